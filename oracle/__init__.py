"""CPU oracle for the S/T-separation training step.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain functional PyTorch running on the CPU, the
algorithm of the reference hot path (``var_sep/networks/*.py`` and
``var_sep/train.py`` of JeremieDona/spatiotemporal_variable_separation).  It
is the checker for the CUDA path, never the thing measured or shipped:

* only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
  ``--impl reference`` legs of ``bench.py`` may import it;
* the product package (``spatiotemporal_variable_separation_b200``) must never
  import it and has no CPU fallback.

Pinning: the reference ships no golden vectors and no unit tests
(SURVEY.md section 4), and its arithmetic is PyTorch's own.  The oracle is therefore
pinned against the *unmodified reference imported in the build container*
(``tests/golden/gen_golden.py`` drives ``/root/reference`` and writes the
committed fixtures under ``tests/golden/``); ``tests/test_oracle_golden.py``
re-checks the oracle against those fixtures on every run.
"""
