"""Empty stand-in: the reference imports matplotlib at module level for its plotting helpers, which
the hot path never calls.  TEST INFRASTRUCTURE (see oracle/tfshim/tensorflow/__init__.py)."""
