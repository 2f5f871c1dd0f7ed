"""Stand-in for scikit-image: only ``transform.resize`` is reachable from the hot path
(``util.py:187-207``, the style/content TARGET resampling).  TEST INFRASTRUCTURE."""
from . import transform  # noqa: F401
