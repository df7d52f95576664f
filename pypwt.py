"""Legacy module name still imported by the reference's tests and docs (test/test_wavelets.py:23)."""
from pycudwt import *  # noqa: F401,F403
from pycudwt import Wavelets, __version__  # noqa: F401
