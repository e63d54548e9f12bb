"""Drop-in counterparts of var_sep.networks (same module, class and function names)."""
