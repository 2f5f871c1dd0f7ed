#!/usr/bin/env python
"""Scene constants of the reference's four driver scripts, captured from the scripts themselves.

    python tests/golden/make_driver_config_golden.py            (needs /root/reference)

Imports the UNMODIFIED ``test_smokegun.py``, ``test_chocolate.py``, ``test_dambreak2d.py`` and
``test_smokegun_resim.py`` (on the TensorFlow / matplotlib stand-ins of ``oracle/tfshim``, with a stub for the absent
``partio`` module), replaces each module's ``run`` by a recorder and calls its ``main(config)`` with the reference's
own ``get_config()`` defaults.  What ``main`` left on the config namespace goes to
``tests/golden/ref_driver_configs.json``; ``tests/test_widen_io_drivers.py`` holds ``lnst.drivers.*.main`` to it.
"""
import copy
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M  # noqa: E402

DRIVERS = ['test_smokegun', 'test_chocolate', 'test_dambreak2d', 'test_smokegun_resim']


def simple(v):
    return isinstance(v, (int, float, str, bool, type(None))) or (isinstance(v, (list, tuple)) and all(simple(e) for e in v))


def capture():
    M._setup_paths()
    sys.modules.setdefault('partio', types.ModuleType('partio'))
    import config as ref_config
    out = {}
    for name in DRIVERS:
        mod = __import__(name)
        got = {}
        mod.run = lambda cfg, got=got: got.update({k: v for k, v in vars(cfg).items() if simple(v)})
        argv, sys.argv = sys.argv, sys.argv[:1]
        try:
            cfg, _ = ref_config.get_config()
        finally:
            sys.argv = argv
        mod.main(copy.deepcopy(cfg))
        out[name] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in sorted(got.items())}
    return out


if __name__ == '__main__':
    o = capture()
    with open(os.path.join(HERE, 'ref_driver_configs.json'), 'w') as f:
        json.dump(o, f, indent=1, sort_keys=True)
    for k, v in o.items():
        print(k, len(v), 'attributes')
