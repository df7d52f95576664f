"""CPU oracle for the pycudwt / PDWT wavelet hot path -- TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the arithmetic of the reference's CUDA
kernels.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; the product
(`pypwt_b200/`) never does and has no CPU fallback.

Parity pinning
--------------
* `oracle/ref_emulation.py` re-executes the reference kernels' per-thread index
  arithmetic literally (pure-Python loops, small sizes) and `tests/test_oracle.py`
  checks every closed form below against it.
* `tests/golden/*.npz` hold outputs of the reference's own CUDA build
  (`oracle/_ref`, built by `oracle/Makefile` from /root/reference) executed on a
  B200 through `gpurun` by `tests/golden/make_golden.py`; the oracle is checked
  against them in the CPU suite.
* PyWavelets (the reference tests' ground truth, test/test_wavelets.py:230,301,372,438)
  is not installable here (no network); if `import pywt` ever succeeds,
  `tests/test_oracle.py` cross-checks against `mode="periodization"` as well.

All references below are relative to /root/reference/.

Conventions (pdwt/src/filters.cpp:5919-6002): L=dec_lo, H=dec_hi, IL=rec_lo, IH=rec_hi.
Band slots (separable.cu:165-174,197,206): coeffs[0]=A_L, coeffs[3i+1]=H_{i+1}=(Lx,Hy),
coeffs[3i+2]=V_{i+1}=(Hx,Ly), coeffs[3i+3]=D_{i+1}; level 1 = finest.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SQRT_2 = 1.4142135623730951          # common.cu:8
HAAR_ALIASES = ("haar", "db1", "bior1.1", "rbior1.1")   # separable.cu:25 ("rbior1.1" sic)

with open(os.path.join(_HERE, "filter_table.json")) as _f:
    _TABLE = json.load(_f)

WAVELET_NAMES = list(_TABLE.keys())


# ----------------------------------------------------------------------------
# integer geometry (utils.cu:14-27, wt.cu:156-165, pypwt.pyx:247-258)
# ----------------------------------------------------------------------------
def div2(n):
    """ceil(n/2) -- utils.cu:24-27."""
    return (n + (n & 1)) // 2


def ilog2(i):
    """floor(log2(i)) for i>=1, 0 for i in {0,1} -- utils.cu:14-20."""
    l = 0
    while i > 1:
        i >>= 1
        l += 1
    return l


def max_level(Nr, Nc, hlen, ndim):
    """wt.cu:156-159: ilog2(N/(hlen-1)), N=min(Nr,Nc) in 2D, Nc in 1D."""
    N = min(Nr, Nc) if ndim == 2 else Nc
    return ilog2(N // (hlen - 1))


def band_sizes(Nr, Nc, levels, do_swt, ndim):
    """[(Nr_l, Nc_l)] for l=1..levels -- pypwt.pyx:247-258 / common.cu:400-445."""
    if do_swt:
        return [(Nr, Nc)] * levels
    res = []
    for _ in range(levels):
        Nc = div2(Nc)
        if ndim == 2:
            Nr = div2(Nr)
        res.append((Nr, Nc))
    return res


# ----------------------------------------------------------------------------
# filters (filters.cpp; relations checked bit-exactly by tools/gen_filter_table.py)
# ----------------------------------------------------------------------------
def filters(wname, dtype=np.float64, table_dtype=np.float32):
    """Return (L, H, IL, IH), each first rounded to fp32 like the reference's DTYPE table
    (table_dtype=np.float64: the DOUBLEPRECISION build's table, filters.h:16-30)."""
    key = wname.lower()
    if key in HAAR_ALIASES:
        key = "haar"
    if key not in _TABLE:
        raise ValueError("unknown wavelet %r" % wname)
    e = _TABLE[key]
    L = np.asarray(e["dec_lo"], np.float64).astype(table_dtype)
    IL = L[::-1].copy() if e["orthogonal"] else np.asarray(e["rec_lo"], np.float64).astype(table_dtype)
    sign = np.where(np.arange(L.size) % 2 == 0, 1.0, -1.0).astype(table_dtype)
    H = -sign * IL
    IH = sign * L
    return tuple(np.asarray(f, dtype) for f in (L, H, IL, IH))


# ----------------------------------------------------------------------------
# 1D building blocks along the LAST axis
# ----------------------------------------------------------------------------
def _ext_odd(x):
    """Odd length: repeat the last sample (separable.cu:116-121)."""
    if x.shape[-1] & 1:
        return np.concatenate([x, x[..., -1:]], axis=-1)
    return x


def analysis(x, f):
    """separable.cu:91-131: out[k] = sum_j f[F-1-j] * xe[(2k - c + j) mod Ne], c=(F-1)//2."""
    xe = _ext_odd(x)
    Ne = xe.shape[-1]
    F = f.size
    c = (F - 1) // 2
    k2 = 2 * np.arange(Ne // 2)
    out = np.zeros(x.shape[:-1] + (Ne // 2,), dtype=np.result_type(x, f))
    for j in range(F):
        out += f[F - 1 - j] * xe[..., (k2 - c + j) % Ne]
    return out


def synthesis(a, d, fl, fh, n_out):
    """separable.cu:293-328 (closed form, even F): for output n,
    x[n] = sum_{t: (n+F/2-1-t) even} fl[t]*a[k] + fh[t]*d[k], k=((n+F/2-1-t)/2) mod n2."""
    F = fl.size
    assert F % 2 == 0, "closed form holds for even filter lengths"
    n2 = a.shape[-1]
    n = np.arange(n_out)
    p = F // 2 - 1
    out = np.zeros(a.shape[:-1] + (n_out,), dtype=np.result_type(a, fl))
    for t in range(F):
        sel = ((n + p - t) % 2) == 0
        k = ((n[sel] + p - t) // 2) % n2
        out[..., sel] += fl[t] * a[..., k] + fh[t] * d[..., k]
    return out


def swt_analysis(x, f, level):
    """separable.cu:409-449: out[g] = sum_j f[F-1-j] * x[(g + (j-c)*s) mod N], s=2^(level-1), c=(F-1)//2."""
    N = x.shape[-1]
    F = f.size
    s = 1 << (level - 1)
    c = (F - 1) // 2
    g = np.arange(N)
    out = np.zeros(x.shape, dtype=np.result_type(x, f))
    for j in range(F):
        out += f[F - 1 - j] * x[..., (g + (j - c) * s) % N]
    return out


def swt_synthesis(a, d, fl, fh, level):
    """separable.cu:593-626: x[g] = sum_j (fl[F-1-j]/2)*a[(g+(j-c)s) mod N] + (fh[F-1-j]/2)*d[...], c=F//2."""
    N = a.shape[-1]
    F = fl.size
    s = 1 << (level - 1)
    c = F // 2
    g = np.arange(N)
    out = np.zeros(a.shape, dtype=np.result_type(a, fl))
    for j in range(F):
        idx = (g + (j - c) * s) % N
        out += (fl[F - 1 - j] / 2) * a[..., idx] + (fh[F - 1 - j] / 2) * d[..., idx]
    return out


def _cols(fn, x, *args):
    """Apply a last-axis operator along axis 0 of 2D arrays."""
    xs = [np.swapaxes(v, 0, 1) if isinstance(v, np.ndarray) and v.ndim == 2 else v for v in (x,) + args]
    return np.swapaxes(fn(*xs), 0, 1)


# ----------------------------------------------------------------------------
# one level, 2D
# ----------------------------------------------------------------------------
def dwt2_level(x, L, H, haar=False):
    """Rows then columns (separable.cu:179-209).  Returns A, H(Lx,Hy), V(Hx,Ly), D.
    haar=True follows haar.cu:10-38 (exact 1/2 factor butterfly)."""
    if haar:
        xe = _ext_odd(x)
        xe = np.swapaxes(_ext_odd(np.swapaxes(xe, 0, 1)), 0, 1)
        a, b = xe[0::2, 0::2], xe[0::2, 1::2]
        c, d = xe[1::2, 0::2], xe[1::2, 1::2]
        A = 0.5 * ((a + c) + (b + d))
        V = 0.5 * ((a + c) - (b + d))
        Hb = 0.5 * ((a - c) + (b - d))
        D = 0.5 * ((a - c) - (b - d))
        return A, Hb, V, D
    lo, hi = analysis(x, L), analysis(x, H)
    A = _cols(analysis, lo, L)
    Hb = _cols(analysis, lo, H)
    V = _cols(analysis, hi, L)
    D = _cols(analysis, hi, H)
    return A, Hb, V, D


def idwt2_level(A, Hb, V, D, IL, IH, shape, haar=False):
    """Columns then rows (separable.cu:332-364).  shape = (Nr_out, Nc_out).
    haar=True follows haar.cu:41-58."""
    Nr, Nc = shape
    if haar:
        out = np.zeros((2 * A.shape[0], 2 * A.shape[1]), dtype=A.dtype)
        a, b, c, d = A, V, Hb, D
        out[0::2, 0::2] = 0.5 * ((a + c) + (b + d))
        out[0::2, 1::2] = 0.5 * ((a + c) - (b + d))
        out[1::2, 0::2] = 0.5 * ((a - c) + (b - d))
        out[1::2, 1::2] = 0.5 * ((a - c) - (b - d))
        return out[:Nr, :Nc]
    t1 = _cols(lambda a_, d_: synthesis(a_, d_, IL, IH, Nr), A, Hb)
    t2 = _cols(lambda a_, d_: synthesis(a_, d_, IL, IH, Nr), V, D)
    return synthesis(t1, t2, IL, IH, Nc)


def swt2_level(x, L, H, level):
    """separable.cu:496-515."""
    lo, hi = swt_analysis(x, L, level), swt_analysis(x, H, level)
    return (_cols(swt_analysis, lo, L, level), _cols(swt_analysis, lo, H, level),
            _cols(swt_analysis, hi, L, level), _cols(swt_analysis, hi, H, level))


def iswt2_level(A, Hb, V, D, IL, IH, level):
    """separable.cu:629-649 (each 1D pass carries a factor 1/2, :581,:621)."""
    t1 = _cols(lambda a_, d_: swt_synthesis(a_, d_, IL, IH, level), A, Hb)
    t2 = _cols(lambda a_, d_: swt_synthesis(a_, d_, IL, IH, level), V, D)
    return swt_synthesis(t1, t2, IL, IH, level)


# ----------------------------------------------------------------------------
# element-wise ops (common.cu)
# ----------------------------------------------------------------------------
def soft_thresh(v, beta):
    """common.cu:13-27."""
    return np.copysign(np.maximum(np.abs(v) - beta, 0), v)


def hard_thresh(v, beta):
    """common.cu:57-71 with W_SIGN (common.cu:7): keep iff |v|-beta > 0 (strict)."""
    return np.where(np.abs(v) - beta > 0, v, 0 * v)


def proj_linf(v, beta):
    """common.cu:101-115."""
    return np.copysign(np.minimum(np.abs(v), beta), v)


def beta_schedule(beta, levels, normalize, ft=np.float32):
    """Per-level thresholds for the detail bands (common.cu:239-247):
    cumulative DTYPE `beta /= SQRT_2` (fp32 / double -> fp32 in the default build)."""
    b = ft(beta)
    out = []
    for _ in range(levels):
        if normalize > 0:
            b = ft(np.float64(b) / SQRT_2)
        out.append(b)
    return out


def beta_appcoeffs(beta, levels, normalize, ft=np.float32):
    """common.cu:230-235: beta / sqrt(2)^levels computed as /(1<<(L/2)) then /SQRT_2 if L odd."""
    b = ft(beta)
    if normalize > 0:
        n2 = levels // 2
        b = ft(b / ft(1 << n2))
        if n2 * 2 != levels:
            b = ft(np.float64(b) / SQRT_2)
    return b


def circshift(img, sr, sc):
    """common.cu:202-211: out[y,x] = in[(y-sr) mod Nr, (x-sc) mod Nc]."""
    return np.roll(img, (sr, sc), axis=(0, 1))


class GlibcRand:
    """glibc `rand()` (TYPE_3 additive feedback, r[i]=r[i-3]+r[i-31]) seeded with 1, i.e. the
    unseeded sequence 1804289383, 846930886, ... that wt.cu:243-244 draws its shifts from."""

    def __init__(self, seed=1):
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            r[i] = w + 2147483647 if w < 0 else w
        for i in range(31, 34):
            r[i] = r[i - 31]
        self._r = r
        for _ in range(310):
            self._step()

    def _step(self):
        r = self._r
        v = (r[-31] + r[-3]) & 0xFFFFFFFF
        r.append(v)
        del r[0]
        return v

    def rand(self):
        return self._step() >> 1


# ----------------------------------------------------------------------------
# full multi-level transforms mirroring the Wavelets class (wt.cu)
# ----------------------------------------------------------------------------
class OracleWavelets:
    """Mirror of `pycudwt.Wavelets` (src/pypwt.pyx:64-616 over pdwt/src/wt.cu) on the CPU.

    dtype=np.float64 (default) evaluates the closed forms exactly on fp32-rounded filters;
    dtype=np.float32 keeps fp32 intermediates (numpy accumulation order, not the GPU's).
    Documented divergences from the reference: Q3 (norm2sq in 1D is the
    true sum of squares), Q4 (per-instance filters), Q5 (unknown wavelet -> ValueError).
    """

    W_INIT, W_FORWARD, W_INVERSE = 0, 1, 2

    def __init__(self, img, wname, levels, do_separable=1, do_cycle_spinning=0, do_swt=0, ndim=2,
                 dtype=np.float64, rng=None, double_build=False):
        # double_build: the reference compiled with DOUBLEPRECISION (filters.h:16-30): DTYPE = double for the samples,
        # the filter table and the thresholds
        self._ft = np.float64 if double_build else np.float32
        img = np.ascontiguousarray(img, dtype=self._ft)
        ndim = min(ndim, 2)
        self.batched1d = 0
        if img.ndim == 2:
            self.Nr, self.Nc = img.shape
            if ndim != 2:
                self.batched1d = 1
        elif img.ndim == 1:
            self.Nr, self.Nc = 1, img.shape[0]
        else:
            raise NotImplementedError("Only 1D and 2D transforms are supported")
        self.shape = img.shape
        self.ndim = img.ndim
        self.wname = wname
        self.dtype = dtype
        self.do_swt = int(do_swt)
        self.do_cycle_spinning = int(do_cycle_spinning)
        # wt.cu:133-142
        self._ndims = 1 if (self.Nr == 1 or ndim == 1) else 2
        self.do_separable = 1 if self._ndims == 1 else int(do_separable)
        self._haar = (wname.lower() in HAAR_ALIASES) and not self.do_swt      # wt.cu:248,255
        if self.do_swt and wname.lower() != "haar" and wname.lower() in HAAR_ALIASES:
            raise ValueError("unknown wavelet %r for SWT (separable.cu:24-28)" % wname)
        self.L, self.H, self.IL, self.IH = filters(wname, dtype, self._ft)
        self.hlen = 2 if self._haar else self.L.size
        levels = max(int(levels), 1)                                             # wt.cu:111-114
        self.levels = min(levels, max_level(self.Nr, self.Nc, self.hlen, self._ndims))  # wt.cu:156-165
        if self.levels < 1:
            raise ValueError("image too small for this wavelet")
        if self.do_cycle_spinning and self._ndims == 1:
            raise ValueError("cycle spinning is not implemented for 1D (wt.cu:179-183)")
        self.sizes = band_sizes(self.Nr, self.Nc, self.levels, self.do_swt, self._ndims)
        self._image = img.reshape(self.Nr, self.Nc).astype(dtype)
        self._rng = rng if rng is not None else GlibcRand()
        self._shift = (0, 0)
        self.state = self.W_INIT
        nb = (3 * self.levels + 1) if self._ndims == 2 else (self.levels + 1)
        self._c = [None] * nb
        self._c[0] = np.zeros(self.sizes[-1], dtype)
        for i in range(self.levels):
            if self._ndims == 2:
                for j in range(3):
                    self._c[3 * i + 1 + j] = np.zeros(self.sizes[i], dtype)
            else:
                self._c[i + 1] = np.zeros(self.sizes[i], dtype)

    # -- helpers -------------------------------------------------------------
    def _swap_ns(self, h, v):
        """Non-separable mode stores (Ly,Hx) in slot 1 and (Hy,Lx) in slot 2
        (nonseparable.cu:71-74,159-162); Haar uses the dedicated kernels (wt.cu:255)."""
        if self._ndims == 2 and not self.do_separable and not self._haar:
            return v, h
        return h, v

    # -- transforms ----------------------------------------------------------
    def set_image(self, img):
        img = np.ascontiguousarray(img, dtype=self._ft)
        if img.shape != (self.Nr, self.Nc):
            raise ValueError("wrong shape")
        self._image = img.astype(self.dtype)
        self.state = self.W_INIT

    def forward(self, img=None):
        if img is not None:
            img = np.ascontiguousarray(img, dtype=self._ft)
            if img.shape != self.shape:
                raise ValueError("wrong shape")
            self._image = img.reshape(self.Nr, self.Nc).astype(self.dtype)
        if self.do_cycle_spinning:                                               # wt.cu:242-246
            sr = self._rng.rand() % self.Nr
            sc = self._rng.rand() % self.Nc
            self._shift = (sr, sc)
            self._image = circshift(self._image, sr, sc)
        a = self._image
        L, H = self.L, self.H
        for i in range(self.levels):
            if self._ndims == 1:
                if self.do_swt:
                    a, d = swt_analysis(a, L, i + 1), swt_analysis(a, H, i + 1)
                elif self._haar:                                                 # haar.cu:132-146
                    xe = _ext_odd(a)
                    s = self.dtype(0.70710678118654746)
                    a, d = s * (xe[..., 0::2] + xe[..., 1::2]), s * (xe[..., 0::2] - xe[..., 1::2])
                else:
                    a, d = analysis(a, L), analysis(a, H)
                self._c[i + 1] = d
            else:
                if self.do_swt:
                    a, h, v, d = swt2_level(a, L, H, i + 1)
                else:
                    a, h, v, d = dwt2_level(a, L, H, haar=self._haar)
                h, v = self._swap_ns(h, v)
                self._c[3 * i + 1], self._c[3 * i + 2], self._c[3 * i + 3] = h, v, d
        self._c[0] = a
        self.state = self.W_FORWARD

    def inverse(self):
        if self.state == self.W_INVERSE:                                         # wt.cu:272-275
            return
        a = self._c[0]
        IL, IH = self.IL, self.IH
        shapes = [(self.Nr, self.Nc)] + list(self.sizes)
        for i in range(self.levels - 1, -1, -1):
            if self._ndims == 1:
                d = self._c[i + 1]
                if self.do_swt:
                    a = swt_synthesis(a, d, IL, IH, i + 1)
                elif self._haar:                                                 # haar.cu:149-160
                    s = self.dtype(0.70710678118654746)
                    out = np.zeros(a.shape[:-1] + (2 * a.shape[-1],), a.dtype)
                    out[..., 0::2], out[..., 1::2] = s * (a + d), s * (a - d)
                    a = out[..., :shapes[i][1]]
                else:
                    a = synthesis(a, d, IL, IH, shapes[i][1])
            else:
                h, v, d = self._c[3 * i + 1], self._c[3 * i + 2], self._c[3 * i + 3]
                h, v = self._swap_ns(h, v)
                if self.do_swt:
                    a = iswt2_level(a, h, v, d, IL, IH, i + 1)
                else:
                    a = idwt2_level(a, h, v, d, IL, IH, shapes[i], haar=self._haar)
        self._image = a
        if self.do_cycle_spinning:                                               # wt.cu:303
            self._image = circshift(self._image, -self._shift[0], -self._shift[1])
        self.state = self.W_INVERSE

    # -- coefficient ops -----------------------------------------------------
    def _detail_slots(self, i):
        return [3 * i + 1, 3 * i + 2, 3 * i + 3] if self._ndims == 2 else [i + 1]

    def _apply(self, fn, beta, app, normalize, app_unscaled=False):
        if self.state == self.W_INVERSE:                                         # wt.cu:309,319
            return
        if app:
            b = self._ft(beta) if app_unscaled else beta_appcoeffs(beta, self.levels, normalize, self._ft)
            self._c[0] = fn(self._c[0], self.dtype(b))
        for i, b in enumerate(beta_schedule(beta, self.levels, normalize, self._ft)):
            for s in self._detail_slots(i):
                self._c[s] = fn(self._c[s], self.dtype(b))

    def soft_threshold(self, beta, do_threshold_appcoeffs=0, normalize=0):
        self._apply(soft_thresh, beta, do_threshold_appcoeffs, normalize)

    def hard_threshold(self, beta, do_threshold_appcoeffs=0, normalize=0):
        # common.cu:264-270: A is thresholded with the UNscaled beta (reference quirk Q2)
        self._apply(hard_thresh, beta, do_threshold_appcoeffs, normalize, app_unscaled=True)

    def proj_linf(self, beta, do_threshold_appcoeffs=1):
        self._apply(proj_linf, beta, do_threshold_appcoeffs, 0)

    def shrink(self, beta, do_threshold_appcoeffs=1):
        if self.state == self.W_INVERSE:
            return
        f = self.dtype(self._ft(1.0) / (self._ft(1.0) + self._ft(beta)))  # common.cu:355
        start = 0 if do_threshold_appcoeffs else 1
        for s in range(start, len(self._c)):
            self._c[s] = self._c[s] * f

    def group_soft_threshold(self, beta, do_threshold_appcoeffs=0, normalize=0):
        """common.cu:145-198,311-341."""
        if self.state == self.W_INVERSE:
            return
        for i, b in enumerate(beta_schedule(beta, self.levels, normalize)):
            slots = self._detail_slots(i)
            if do_threshold_appcoeffs and i == self.levels - 1:
                slots = slots + [0]
            nrm = np.sqrt(sum(self._c[s] ** 2 for s in slots))
            with np.errstate(divide="ignore", invalid="ignore"):
                res = np.where(nrm == 0, 0, np.maximum(1 - self.dtype(b) / nrm, 0))
            for s in slots:
                self._c[s] = self._c[s] * res

    def norm1(self):
        return float(sum(np.abs(c, dtype=np.float64).sum() for c in self._c))   # wt.cu:396-416

    def norm2sq(self):
        return float(sum((c.astype(np.float64) ** 2).sum() for c in self._c))   # wt.cu:368-393 (Q3)

    def add_wavelet(self, W, alpha=1.0):
        """wt.cu:622-655."""
        if self.levels != W.levels or self.wname.lower() != W.wname.lower():
            return -1
        if self.state == self.W_INVERSE or W.state == self.W_INVERSE:
            return 1
        if (self.Nr, self.Nc, self._ndims) != (W.Nr, W.Nc, W._ndims):
            return -2
        if bool(self.do_swt) != bool(W.do_swt):
            return -3
        if self.do_cycle_spinning and W.do_cycle_spinning and self._shift != W._shift:
            return -4
        a = self.dtype(np.float32(alpha))
        for s in range(len(self._c)):
            self._c[s] = self._c[s] + a * W._c[s]
        return 0

    def set_coeff(self, coeff, num):
        self._c[num] = np.ascontiguousarray(coeff, self._ft).reshape(self._c[num].shape).astype(self.dtype)

    # -- read-back -----------------------------------------------------------
    @property
    def image(self):
        return np.asarray(self._image, self._ft).reshape(self.Nr, self.Nc)

    def coeff_only(self, num):
        if self.state == self.W_INVERSE:                                         # wt.cu:474-477 + pyx:284
            raise RuntimeError("coefficients were consumed by inverse()")
        return np.asarray(self._c[num], self._ft)

    @property
    def coeffs(self):
        if self.state == self.W_INVERSE:
            raise RuntimeError("coefficients were consumed by inverse()")
        out = [np.asarray(self._c[0], self._ft)]
        for i in range(self.levels):
            if self._ndims == 2:
                out.append([np.asarray(self._c[3 * i + 1 + j], self._ft) for j in range(3)])
            else:
                out.append(np.asarray(self._c[i + 1], self._ft))
        return out
