"""Drop-in module name of the reference extension (`pycudwt`, setup.py:105)."""
from pypwt_b200 import Wavelets, pinned_empty, pinned_zeros, device_count, lookup_filters  # noqa: F401
from pypwt_b200 import __version__  # noqa: F401
