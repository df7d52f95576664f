"""pypwt_b200 -- B200-native wavelet hot path behind the `pycudwt.Wavelets` API.

The package holds the CUDA/C-ABI library (`csrc/`, built to `libpwt_b200.so`), the Cython wrapper
(`pycudwt.pyx`) and the multi-GPU stack front-end (`sharded.py`).  Everything numerical runs on the
GPU; importing this package on a machine without the built extension raises ImportError with the
build command (there is deliberately no CPU fallback).
"""
import os as _os

try:
    from . import pycudwt as _ext
except ImportError as _e:  # pragma: no cover - exercised only on unbuilt trees
    raise ImportError(
        "pypwt_b200: the native extension is not built (%s). Run `python pypwt_b200/_build.py` "
        "(needs nvcc + cython); there is no CPU fallback." % (_e,)) from _e

# let the C side find the NCCL the Python environment ships, without importing torch
if "PWT_NCCL_LIB" not in _os.environ:
    try:
        import importlib.util as _ilu
        _spec = _ilu.find_spec("nvidia.nccl")
        if _spec and _spec.submodule_search_locations:
            _cand = _os.path.join(list(_spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if _os.path.exists(_cand):
                _os.environ["PWT_NCCL_LIB"] = _cand
    except Exception:
        pass

Wavelets = _ext.Wavelets
Wavelets64 = _ext.Wavelets64
Wavelets3D = _ext.Wavelets3D
VOLUME_BAND_KEYS = _ext.VOLUME_BAND_KEYS
lookup_filters64 = _ext.lookup_filters64
pinned_empty = _ext.pinned_empty
pinned_zeros = _ext.pinned_zeros
device_count = _ext.device_count
set_device = _ext.set_device
lookup_filters = _ext.lookup_filters
comm_unique_id = _ext.comm_unique_id
comm_init_all = _ext.comm_init_all
norms_allreduce_group = _ext.norms_allreduce_group
DeviceArray = _ext.DeviceArray
LIBRARY_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "libpwt_b200.so")
__version__ = "1.0.3"
__all__ = ["Wavelets", "Wavelets64", "Wavelets3D", "VOLUME_BAND_KEYS", "lookup_filters64", "pinned_empty", "pinned_zeros", "device_count", "set_device", "lookup_filters",
           "comm_unique_id", "comm_init_all", "norms_allreduce_group", "DeviceArray", "LIBRARY_PATH"]
