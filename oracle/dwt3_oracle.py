"""CPU oracle of the volumetric (3D) separable DWT -- TEST INFRASTRUCTURE ONLY (tests/ and smoke() may import it).

The reference has no 3D transform ("3D is not handled", pdwt/README.md:29), so parity is UNPINNED by any reference
output: this oracle composes the reference's own, pinned, 1D closed forms (`pdwt_oracle.analysis` / `synthesis`,
separable.cu:91-131 and :293-328) along x, then y, then z -- which is the definition of the separable n-D transform
(pywt.wavedecn(mode="periodization"); cross-checked against pywt when it is importable, tests/test_oracle.py).
Band keys follow pywt.wavedecn: one letter per axis in (z, y, x) order, 'a' low-pass, 'd' high-pass."""
import numpy as np

from . import pdwt_oracle as O

KEYS = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")


def _along(x, axis, fn):
    return np.moveaxis(fn(np.moveaxis(x, axis, -1)), -1, axis)


def dwt3_level(v, L, H):
    """One level: returns {'aaa': ..., 'aad': ..., ...} (keys in (z, y, x) order)."""
    out = {"": v}
    for axis in (2, 1, 0):                      # x first, then y, then z
        nxt = {}
        for key, arr in out.items():
            nxt["a" + key] = _along(arr, axis, lambda t: O.analysis(t, L))
            nxt["d" + key] = _along(arr, axis, lambda t: O.analysis(t, H))
        out = nxt
    return out


def idwt3_level(bands, IL, IH, shape):
    """Inverse of dwt3_level: z first, then y, then x; `shape` = output (nz, ny, nx)."""
    cur = dict(bands)
    for axis in (0, 1, 2):
        nxt = {}
        for key in {k[1:] for k in cur}:
            a, d = cur["a" + key], cur["d" + key]
            am, dm = np.moveaxis(a, axis, -1), np.moveaxis(d, axis, -1)
            nxt[key] = np.moveaxis(O.synthesis(am, dm, IL, IH, shape[axis]), -1, axis)
        cur = nxt
    return cur[""]


def max_level3(shape, hlen):
    n = min(shape) // (hlen - 1)
    lev = 0
    while n > 1:
        n >>= 1
        lev += 1
    return lev


class OracleWavelets3D:
    def __init__(self, vol, wname, levels, dtype=np.float64):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        assert vol.ndim == 3
        self.shape = vol.shape
        self.L, self.H, self.IL, self.IH = O.filters(wname, dtype)
        self.levels = min(max(int(levels), 1), max_level3(vol.shape, self.L.size))
        if self.levels < 1:
            raise ValueError("volume too small for this wavelet")
        self._vol = vol.astype(dtype)
        self._c = None
        self._shapes = []

    def forward(self):
        a = self._vol
        self._c, self._shapes = [None], []
        for _ in range(self.levels):
            self._shapes.append(a.shape)
            b = dwt3_level(a, self.L, self.H)
            a = b.pop("aaa")
            self._c.append(b)
        self._c[0] = a

    def inverse(self):
        a = self._c[0]
        for lev in range(self.levels, 0, -1):
            bands = dict(self._c[lev])
            bands["aaa"] = a
            a = idwt3_level(bands, self.IL, self.IH, self._shapes[lev - 1])
        self._vol = a

    def soft_threshold(self, beta, app=0):
        b = np.float32(beta)
        if app:
            self._c[0] = O.soft_thresh(self._c[0], self._c[0].dtype.type(b))
        for lev in range(1, self.levels + 1):
            self._c[lev] = {k: O.soft_thresh(v, v.dtype.type(b)) for k, v in self._c[lev].items()}

    def hard_threshold(self, beta, app=0):
        b = np.float32(beta)
        if app:
            self._c[0] = O.hard_thresh(self._c[0], self._c[0].dtype.type(b))
        for lev in range(1, self.levels + 1):
            self._c[lev] = {k: O.hard_thresh(v, v.dtype.type(b)) for k, v in self._c[lev].items()}

    def norm1(self):
        return float(np.abs(self._c[0]).sum() + sum(np.abs(v).sum() for d in self._c[1:] for v in d.values()))

    def norm2sq(self):
        return float((self._c[0] ** 2).sum() + sum((v ** 2).sum() for d in self._c[1:] for v in d.values()))

    @property
    def coeffs(self):
        return [np.asarray(self._c[0], np.float32)] + [{k: np.asarray(v, np.float32) for k, v in d.items()} for d in self._c[1:]]

    @property
    def image(self):
        return np.asarray(self._vol, np.float32)
