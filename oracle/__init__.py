"""CPU oracle for the LNST stylisation hot path -- TEST INFRASTRUCTURE ONLY.

This package is a PyTorch-CPU restatement (fp32 by default, fp64 on request) of the
reference's TensorFlow-1.15 graph and optimisation loop (byungsook/neural-flow-style,
files cited per function as ``file:line`` relative to the reference root).  It exists to
CHECK the CUDA path; it is never imported by the product package (``lnst``).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.

PARITY STATUS: **parity unpinned** except for the one known-answer the reference holds
(the 5x5 bilinear-warp tables in ``transform.py:1865-1884``, checked in
``tests/test_oracle_warp_kat.py``).  The reference has no tests, no golden vectors, and
cannot be executed here (TensorFlow 1.15 has no wheel for this interpreter), so every other
function is pinned only by line-by-line restatement plus brute-force definitions on tiny
grids (``tests/test_oracle_*.py``).
"""
from . import transform, render, vgg, loss, adam, styler  # noqa: F401
