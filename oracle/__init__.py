"""CPU oracle for the LNST stylisation hot path -- TEST INFRASTRUCTURE ONLY.

This package is a PyTorch-CPU restatement (fp32 by default, fp64 on request) of the
reference's TensorFlow-1.15 graph and optimisation loop (byungsook/neural-flow-style,
files cited per function as ``file:line`` relative to the reference root).  It exists to
CHECK the CUDA path; it is never imported by the product package (``lnst``).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.

PARITY STATUS: **pinned against the reference's own code executed in this container**, to the
extent that is possible without TensorFlow:
  * ``tests/golden/ref_*.npz`` are outputs of the UNMODIFIED reference modules
    (``styler_3p.Styler.run``, ``styler_2p.Styler.run``, ``transform.py`` operators, ``vgg.py``,
    ``styler_base.py``) run by ``tests/golden/make_reference_golden.py`` on top of ``oracle/tfshim`` --
    a stand-in for the absent third-party TensorFlow 1.15 that restates the published semantics of
    each op the path calls.  ``tests/test_reference_golden.py`` holds this oracle to them (loss
    2e-5, field 1e-4, operators 2e-6 with identical NaN patterns, view matrices bit-equal) on 23
    loop-level cases (density / position / colour modes, views, octaves, sequences, resize, TV,
    content, style mask) and 20 operator-level vectors;
  * the one known-answer the reference itself holds (the 5x5 bilinear-warp tables in
    ``transform.py:1865-1884``) is checked in ``tests/test_golden.py``.
  * the resimulation path (``oracle/resim.py``, ``transform.g2p*``) is pinned the same way by
    ``tests/golden/ref_resim.npz`` (the reference's ``test_smokegun_resim.SimG2P`` run here), the inception path
    (``oracle/graphnet.py``) by ``ref_density_inception*.npz`` (the reference's own GraphDef parsing / import /
    tensor naming / losses on a seeded synthetic graph; the network's op numerics themselves are UNPINNED --
    see that module's header).
Not pinned (no TensorFlow binary): the floating-point summation order inside TF's kernels and
TF's platform-dependent NaN handling on CPU (the GPU rule is the one restated).
"""
from . import transform, render, vgg, loss, adam, styler, resim, graphnet  # noqa: F401
