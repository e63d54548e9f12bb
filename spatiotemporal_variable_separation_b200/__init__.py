"""B200-native training step of the PDE-driven S/T separation model (drop-in for var_sep.networks / var_sep.train)."""
__version__ = '0.1.0'
