"""ctypes loader of oracle/dwt_cpu.c (plain-C, OpenMP) -- TEST / BASELINE INFRASTRUCTURE ONLY.

A second, independent CPU restatement of the separable DWT (checked against the numpy oracle and the
reference goldens in tests/test_oracle.py) and the `cpu_baseline` / `--impl reference` arm of bench.py:
it plays the role of the reference workflow's CPU path (pywt.wavedec2 / waverec2, periodization),
which cannot be installed here.  Build: `make -C oracle cpu` (gcc -O3 -fopenmp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libdwt_cpu.so")


def build():
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    src = os.path.join(_HERE, "dwt_cpu.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", _LIB, src], check=True)
    return _LIB


def _lib():
    lib = ctypes.CDLL(build())
    fp = ctypes.POINTER(ctypes.c_float)
    lib.dwt_cpu_forward2d.argtypes = [fp, ctypes.POINTER(fp), fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, ctypes.c_int]
    lib.dwt_cpu_inverse2d.argtypes = [fp, ctypes.POINTER(fp), fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, ctypes.c_int]
    return lib


def threads():
    return _lib().dwt_cpu_threads()


def set_threads(n=None):
    """Use n OpenMP threads (default: every core this process may run on), whatever OMP_NUM_THREADS says --
    torchrun exports OMP_NUM_THREADS=1 to its workers.  Returns the thread count now in effect."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib = _lib()
    lib.dwt_cpu_set_threads(int(n))
    return lib.dwt_cpu_threads()


class CpuDwt2:
    """Multi-level separable 2D DWT plan on the CPU (fp32), band layout of the reference."""

    def __init__(self, shape, wname, levels):
        from . import pdwt_oracle as O
        self.lib = _lib()
        self.Nr, self.Nc = shape
        self.L, self.H, self.IL, self.IH = (np.ascontiguousarray(f, np.float32) for f in O.filters(wname, np.float32))
        self.F = self.L.size
        self.levels = min(max(levels, 1), O.max_level(self.Nr, self.Nc, self.F, 2))
        sizes = O.band_sizes(self.Nr, self.Nc, self.levels, 0, 2)
        self.bands = [np.zeros(sizes[-1], np.float32)]
        for s in sizes:
            self.bands += [np.zeros(s, np.float32) for _ in range(3)]
        self.tmp = np.zeros(2 * self.Nr * self.Nc, np.float32)
        fp = ctypes.POINTER(ctypes.c_float)
        self._bp = (fp * len(self.bands))(*[b.ctypes.data_as(fp) for b in self.bands])
        self._fp = fp

    def _p(self, a):
        return a.ctypes.data_as(self._fp)

    def forward(self, img):
        img = np.ascontiguousarray(img, np.float32)
        self.lib.dwt_cpu_forward2d(self._p(img), self._bp, self._p(self.tmp), self.Nr, self.Nc, self.levels,
                                   self._p(self.L), self._p(self.H), self.F)
        return self.bands

    def inverse(self):
        out = np.empty((self.Nr, self.Nc), np.float32)
        self.lib.dwt_cpu_inverse2d(self._p(out), self._bp, self._p(self.tmp), self.Nr, self.Nc, self.levels,
                                   self._p(self.IL), self._p(self.IH), self.F)
        return out
