"""GPU parity AT THE BENCHMARK SIZES against the reference itself (oracle/_ref: the unmodified PDWT +
pypwt.pyx recompiled for sm_100a), band by band and on the reconstruction.

The small-image tests of test_gpu_vs_pdwt.py stay below the size thresholds of the kernels that produce the
benchmark numbers (fused cascade: >= 512 columns, strip kernels: >= 64 x 256, multi-wave task queues, 32-bit
offset guards, batch strides).  Here the reference meets exactly those kernels: BASELINE.json configs C2 / M
(4096^2, 8192^2 haar + db2, 3 levels), C3 (2048^2 sym8 stack, per slice), C4 (SWT db4 4 levels with cycle
spinning + hard threshold), C5 (8192^2, 5 levels, long filters) and the batched 1D transform of 8192 rows.

Band mapping is the reference's own (test/test_wavelets.py:245-255): coeffs[0] = A, coeffs[i+1] = [H, V, D] of
level i+1.  Tolerance: 1e-5 * max(|x|max, |band|max) -- the reference's tests scale their tolerance with the
level the same way (test_wavelets.py:236,246: tol * 2**level) because the approximation grows by 2 per level.
The worst err / (1e-5 * max|x|) of every case is written to gpurun_out/parity_bench_sizes.json.
"""
import ctypes
import gc
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, synth_image

pytestmark = pytest.mark.gpu
RTOL = 1e-5
_REPORT = {}


def _ref():
    p = os.path.join(ROOT, "oracle", "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    try:
        import pycudwt_ref
    except ImportError as e:
        pytest.skip("reference build oracle/_ref not available: %s" % e)
    return pycudwt_ref


def _mine():
    import pycudwt
    return pycudwt


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_bench_sizes.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _bands(W, is1d):
    """yield (label, level, host array) band by band (one D2H at a time, nothing kept)."""
    yield "A", W.levels, W.coeff_only(0)
    for l in range(1, W.levels + 1):
        if is1d:
            yield "D%d" % l, l, W.coeff_only(l)
        else:
            for j, nm in enumerate("HVD"):
                yield "%s%d" % (nm, l), l, W.coeff_only(3 * (l - 1) + j + 1)


def _cmp(g, r, xmax, what, rec, key):
    assert g.shape == r.shape, "%s: shape %s vs %s" % (what, g.shape, r.shape)
    err = float(np.abs(g - r).max())
    bmax = float(np.abs(r).max())
    tol = RTOL * max(xmax, bmax)
    rec[key] = max(rec.get(key, 0.0), err / (RTOL * xmax))
    rec["worst_err_over_tol"] = max(rec.get("worst_err_over_tol", 0.0), err / tol)
    assert err <= tol, "%s: err %.3e > tol %.3e (|x|max %.1f, |band|max %.1f)" % (what, err, tol, xmax, bmax)


def _cmp_hard(W, R, b, xmax, what, rec):
    """Hard-thresholded band: a coefficient within rounding distance of beta may be kept by one implementation and
    zeroed by the other (|c| - beta > 0 on values that differ in the last bits).  Such flips are counted (they must
    stay a ~1e-6 fraction), everything else must agree; flipped bands are then overwritten with the reference's
    values so that the two inverses start from identical coefficients."""
    g, r = W.coeff_only(b), R.coeff_only(b)
    flip = (g == 0) != (r == 0)
    nflip = int(flip.sum())
    rec["threshold_flips"] = rec.get("threshold_flips", 0) + nflip
    assert nflip <= 4 + g.size * 5e-6, "%s: %d flips in %d coefficients" % (what, nflip, g.size)
    d = np.abs(g - r)
    if nflip:
        d[flip] = 0
    tol = RTOL * max(xmax, float(np.abs(r).max()))
    err = float(d.max())
    rec["worst_err_over_tol"] = max(rec.get("worst_err_over_tol", 0.0), err / tol)
    assert err <= tol, "%s: err %.3e > tol %.3e" % (what, err, tol)
    if nflip:
        W.set_coeff(r, b)


def _side_by_side(name, img, wname, levels, kw=None, thresh=None, srand=None):
    """Reference and ours on the same input: every band after forward (and after the optional threshold),
    then the reconstruction.  thresh = ("soft"|"hard", beta).  srand: seed libc rand() before each forward
    so that both cycle-spinning instances draw the same shift."""
    ref, mine = _ref(), _mine()
    kw = dict(kw or {})
    is1d = kw.get("ndim", 2) == 1 or img.ndim == 1
    xmax = float(np.abs(img).max())
    rec = _REPORT.setdefault(name, {"shape": list(img.shape), "wname": wname, "levels": levels, "xmax": xmax})
    libc = ctypes.CDLL("libc.so.6") if srand is not None else None

    def fwd(X):
        if libc:
            libc.srand(srand)
        X.forward(img)          # (re)loads the input: earlier inverses left a reconstruction in the image plane

    R = ref.Wavelets(img, wname, levels, **kw)
    W = mine.Wavelets(img, wname, levels, **kw)
    assert W.levels == R.levels
    fwd(R)
    fwd(W)
    for (lab, l, g), (_, _, r) in zip(_bands(W, is1d), _bands(R, is1d)):
        _cmp(g, r, xmax, "%s fwd %s" % (name, lab), rec, "fwd_L%d" % l)
    if thresh and thresh[0] == "soft":
        R.soft_threshold(thresh[1], 0, 1)
        W.soft_threshold(thresh[1], 0, 1)
        for (lab, l, g), (_, _, r) in zip(_bands(W, is1d), _bands(R, is1d)):     # the read flushes ours to memory
            _cmp(g, r, xmax, "%s soft %s" % (name, lab), rec, "thr_L%d" % l)
        fwd(R)
        fwd(W)
        R.soft_threshold(thresh[1], 0, 1)
        W.soft_threshold(thresh[1], 0, 1)       # not read back: stays deferred, applied on load by our inverse
    elif thresh:
        # deferred (applied on load by the inverse) against applied in memory, ours against ours: bit-identical
        W.hard_threshold(thresh[1], 0, 1)
        W.inverse()
        deferred = np.array(W.image)
        fwd(W)
        W.hard_threshold(thresh[1], 0, 1)
        W.coeff_only(0)                         # any observer flushes the pending threshold to memory
        W.inverse()
        assert np.array_equal(deferred, W.image), "%s: deferred hard threshold differs from the flushed one" % name
        del deferred
        fwd(W)
        R.hard_threshold(thresh[1], 0, 1)
        W.hard_threshold(thresh[1], 0, 1)
        nb = (levels if is1d else 3 * levels) + 1
        for b in range(nb):
            _cmp_hard(W, R, b, xmax, "%s hard band %d" % (name, b), rec)
    R.inverse()
    W.inverse()
    rimg = np.array(R.image)
    gimg = np.array(W.image)
    _cmp(gimg, rimg, xmax, "%s inverse" % name, rec, "inverse")
    if not thresh:
        rec["reconstruction_err_over_1e-5xmax"] = float(np.abs(gimg - img).max()) / (RTOL * xmax)
    del R, W
    gc.collect()
    return rec


# ---- C2 / M: the headline kernels (fused 3-level cascade) ---------------------------------------------------
@pytest.mark.parametrize("side", [4096, 8192])
@pytest.mark.parametrize("wname", ["haar", "db2"])
def test_c2_m_fused_cascade_vs_pdwt(wname, side):
    img = synth_image((side, side), seed=side + len(wname), kind="smooth")
    _side_by_side("C2/M %s %d^2 L3" % (wname, side), img, wname, 3)


def test_m_db2_soft_threshold_deferred_vs_pdwt():
    """C1's step (forward + soft_threshold + inverse) at the metric's size: the threshold is deferred into the
    fused inverse; both the thresholded coefficients (flushed by the read) and the fused-on-load path are compared."""
    img = synth_image((8192, 8192), seed=77, kind="smooth")
    _side_by_side("M db2 8192^2 L3 soft(10)", img, "db2", 3, thresh=("soft", 10.0))


def test_db3_five_levels_vs_pdwt():
    """F = 6 cascade (levels 1-3 fused) followed by the per-level kernels for levels 4, 5."""
    img = synth_image((4096, 4096), seed=78, kind="smooth")
    _side_by_side("db3 4096^2 L5", img, "db3", 5)


# ---- C3: sym8 stack, per slice ---------------------------------------------------------------------------------
def test_c3_sym8_stack_vs_pdwt():
    ref, mine = _ref(), _mine()
    S, side = 4, 2048
    stack = synth_image((S, side, side), seed=31, kind="smooth")
    xmax = float(np.abs(stack).max())
    rec = _REPORT.setdefault("C3 sym8 %dx%d^2 L3 (stack, per slice)" % (S, side),
                             {"shape": list(stack.shape), "wname": "sym8", "levels": 3, "xmax": xmax})
    W = mine.Wavelets(stack, "sym8", 3)
    W.forward()
    n1, n2 = W.norms()
    mine_c = [np.array(W.coeff_only(b)) for b in range(10)]
    W.inverse()
    mine_img = np.array(W.image)
    r1 = r2 = 0.0
    for s in range(S):
        R = ref.Wavelets(stack[s], "sym8", 3)
        R.forward()
        r1 += float(R.norm1())
        r2 += float(R.norm2sq())
        for b in range(10):
            _cmp(mine_c[b][s], R.coeff_only(b), xmax, "C3 slice %d band %d" % (s, b), rec, "fwd_L%d" % (3 if b == 0 else (b - 1) // 3 + 1))
        R.inverse()
        _cmp(mine_img[s], np.array(R.image), xmax, "C3 slice %d inverse" % s, rec, "inverse")
        del R
    # global norms of the stack = sum of the reference's per-slice norms (fp32 cuBLAS sums on its side)
    assert abs(n1 - r1) <= 2e-5 * r1
    assert abs(n2 - r2) <= 2e-5 * r2
    rec["norm1_rel_diff"] = abs(n1 - r1) / r1
    rec["norm2sq_rel_diff"] = abs(n2 - r2) / r2


# ---- C4: stationary transform, cycle spinning, hard threshold ---------------------------------------------------
@pytest.mark.parametrize("side", [4096, 8192])
def test_c4_swt_db4_cycle_spinning_hard_threshold_vs_pdwt(side):
    img = synth_image((side, side), seed=41, kind="smooth")
    _side_by_side("C4 swt db4 %d^2 L4 cs + hard(20)" % side, img, "db4", 4,
                  kw=dict(do_swt=1, do_cycle_spinning=1), thresh=("hard", 20.0), srand=4242 + side)


def test_c4_swt_db4_plain_vs_pdwt():
    img = synth_image((4096, 4096), seed=42, kind="smooth")
    _side_by_side("C4 swt db4 4096^2 L4", img, "db4", 4, kw=dict(do_swt=1))


# ---- C5: five levels, mid-length and long filters -----------------------------------------------------------------
@pytest.mark.parametrize("wname", ["db4", "sym8", "db12", "db20", "coif5"])
def test_c5_long_filters_vs_pdwt(wname):
    img = synth_image((8192, 8192), seed=50 + len(wname), kind="smooth")
    _side_by_side("C5 %s 8192^2 L5" % wname, img, wname, 5)


def test_c5_nonseparable_vs_pdwt():
    """do_separable=0 at a size where the reference's F^2 stencil is still quick (db4: 64 taps)."""
    img = synth_image((2048, 2048), seed=59, kind="smooth")
    _side_by_side("C5 nonsep db4 2048^2 L5", img, "db4", 5, kw=dict(do_separable=0))


# ---- batched 1D over 8192 rows --------------------------------------------------------------------------------------
@pytest.mark.parametrize("wname", ["haar", "db2", "sym8"])
def test_batched_1d_8192_rows_vs_pdwt(wname):
    img = synth_image((8192, 8192), seed=60 + len(wname), kind="smooth")
    _side_by_side("1D %s 8192x8192 L3" % wname, img, wname, 3, kw=dict(ndim=1))


def test_batched_1d_swt_vs_pdwt():
    img = synth_image((2048, 8192), seed=66, kind="smooth")
    _side_by_side("1D swt db4 2048x8192 L3", img, "db4", 3, kw=dict(ndim=1, do_swt=1))


# ---- maximum sizes ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wname,levels", [("db2", 3), ("sym8", 2)])
def test_largest_image_vs_pdwt(wname, levels):
    """2^29 samples (16384 x 32768, 2 GiB): element offsets beyond 2^28 in every kernel family that serves large images
    (fused cascade, strip kernels), beside the reference run on the same array."""
    tile = synth_image((2048, 4096), seed=70 + levels, kind="smooth")
    img = np.ascontiguousarray(np.tile(tile, (8, 8)))
    assert img.shape == (16384, 32768)
    _side_by_side("MAX %s 16384x32768 L%d" % (wname, levels), img, wname, levels)


def test_image_size_limit_is_refused_cleanly():
    """One image must hold fewer than 2^31 samples (32-bit element offsets inside an image): the C ABI refuses larger
    ones with PWT_ERR_ARG and a message before anything is allocated; the largest accepted width still works in 1D."""
    from pypwt_b200 import LIBRARY_PATH
    lib = ctypes.CDLL(LIBRARY_PATH)
    lib.pwt_last_error.restype = ctypes.c_char_p
    for nr, nc in ((46341, 46341), (32768, 65536), (65536, 32768)):
        h = ctypes.c_void_p()
        rc = lib.pwt_create(ctypes.byref(h), None, nr, nc, b"db2", 1, 1, 1, 0, 0, 2)
        assert rc != 0 and not h.value, (nr, nc)
        assert b"2^31" in lib.pwt_last_error()
    # a long 1D signal: 2^27 + 3 samples, odd length, 4 levels, against perfect reconstruction and energy conservation
    n = 2 ** 27 + 3
    x = np.sin(np.arange(n, dtype=np.float64) * 1e-3).astype(np.float32) * 100 + np.random.default_rng(3).standard_normal(n).astype(np.float32)
    W = _mine().Wavelets(x, "db3", 4, ndim=1)
    W.forward()
    e = sum(float(np.square(np.asarray(c, np.float64)).sum()) for c in W.coeffs)
    ex = float(np.square(x.astype(np.float64)).sum())
    assert abs(e - ex) <= 2e-4 * ex                          # orthogonal bank; the odd-length extension adds one sample per level
    W.inverse()
    assert np.abs(W.image.ravel() - x).max() <= 1e-5 * 8 * np.abs(x).max()
