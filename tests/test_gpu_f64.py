"""Double-precision plans (`Wavelets64` -> pwt64_* C ABI; the reference's DOUBLEPRECISION build, pdwt/src/filters.h:16-30,
SURVEY 8f rank 4) against the oracle evaluated with DTYPE = double (`double_build=True`: samples, filter table and
thresholds in float64).  Tolerance: 1e-12 * max|x| (fp64 accumulation of <= 40 taps per pass over <= 10 levels)."""
import numpy as np
import pytest

from conftest import synth_image
from oracle import pdwt_oracle as O

pytestmark = pytest.mark.gpu

RTOL64 = 1e-12
ALL = O.WAVELET_NAMES


def _W64(*a, **k):
    import pypwt_b200
    return pypwt_b200.Wavelets64(*a, **k)


def _img(shape, seed):
    return synth_image(shape, seed=seed).astype(np.float64) + np.random.default_rng(seed).standard_normal(shape) * 1e-3


def close(got, ref, what, scale=255.0, wname=""):
    got, ref = np.asarray(got), np.asarray(ref, np.float64)
    assert got.dtype == np.float64 and got.shape == ref.shape, (what, got.dtype, got.shape, ref.shape)
    tol = RTOL64 * max(scale, float(np.abs(ref).max())) * (50 if wname in ("bior3.1", "rbio3.1") else 1)
    err = float(np.abs(got - ref).max())
    assert err <= tol, "%s: max err %.3e > %.3e" % (what, err, tol)


def compare(W, Wo, what, wname=""):
    c, co = W.coeffs, Wo.coeffs
    assert len(c) == len(co)
    close(c[0], co[0], what + " A", wname=wname)
    for i in range(1, len(c)):
        if isinstance(co[i], list):
            for j in range(3):
                close(c[i][j], co[i][j], what + " L%d b%d" % (i, j), wname=wname)
        else:
            close(c[i], co[i], what + " D%d" % i, wname=wname)


@pytest.mark.parametrize("wname", ALL)
def test_dwt2_idwt2_f64(wname):
    """Every built-in bank at the maximum depth on an odd-sized image: bands, then reconstruction."""
    img = _img((203, 177), 3)
    W = _W64(img, wname, 99)
    Wo = O.OracleWavelets(img, wname, 99, double_build=True)
    assert W.levels == Wo.levels and [tuple(s) for s in W.sizes] == [tuple(s) for s in Wo.sizes]
    W.forward(); Wo.forward()
    compare(W, Wo, "dwt2 f64 " + wname, wname)
    W.inverse(); Wo.inverse()
    close(W.image, Wo.image, "idwt2 f64 " + wname, wname=wname)
    # perfect reconstruction is limited by the precision of the table itself (~2.5e-8 for coif5)
    assert np.abs(W.image - img).max() <= 1e-6 * np.abs(img).max()


@pytest.mark.parametrize("wname", ["haar", "db2", "sym8", "db20", "bior2.4", "rbio6.8", "coif5"])
@pytest.mark.parametrize("kind", ["1d", "batched", "swt2", "swt1d", "nonsep", "stack"])
def test_other_transforms_f64(wname, kind):
    kw, shape = {}, (96, 160)
    if kind == "1d":
        shape, kw = (4099,), dict(ndim=1)
    elif kind == "batched":
        shape, kw = (37, 1000), dict(ndim=1)
    elif kind == "swt2":
        shape, kw = (96, 132), dict(do_swt=1)
    elif kind == "swt1d":
        shape, kw = (9, 700), dict(do_swt=1, ndim=1)
    elif kind == "nonsep":
        kw = dict(do_separable=0)
    img = _img(shape, 5)
    if kind == "stack":
        stack = np.stack([img, img[::-1], img * 0.5])
        W = _W64(stack, wname, 3)
        W.forward()
        for k in range(3):
            Wo = O.OracleWavelets(stack[k], wname, 3, double_build=True)
            Wo.forward()
            close(W.coeffs[0][k], Wo.coeffs[0], "stack A")
            close(W.coeffs[-1][1][k], Wo.coeffs[-1][1], "stack V of the last level")
        W.inverse()
        assert np.abs(W.image - stack).max() <= 1e-6 * np.abs(stack).max()
        return
    try:
        Wo = O.OracleWavelets(img, wname, 4, double_build=True, **kw)
    except ValueError:
        pytest.skip("not a valid configuration")
    W = _W64(img, wname, 4, **kw)
    assert W.levels == Wo.levels
    W.forward(); Wo.forward()
    compare(W, Wo, kind + " f64 " + wname)
    W.inverse(); Wo.inverse()
    close(W.image.reshape(Wo.image.shape), Wo.image, kind + " inverse f64 " + wname)


def test_thresholds_norms_shrink_f64():
    img = _img((128, 192), 7)
    for op, kw in (("soft_threshold", dict(do_threshold_appcoeffs=1, normalize=1)), ("hard_threshold", dict(do_threshold_appcoeffs=1, normalize=1)),
                   ("soft_threshold", {}), ("hard_threshold", {}), ("shrink", {}), ("shrink", dict(do_threshold_appcoeffs=0))):
        W = _W64(img, "db3", 3)
        Wo = O.OracleWavelets(img, "db3", 3, double_build=True)
        W.forward(); Wo.forward()
        getattr(W, op)(7.3, **kw); getattr(Wo, op)(7.3, **kw)
        compare(W, Wo, op + " f64")
        n1, n2 = W.norms()
        assert abs(n1 - Wo.norm1()) <= 1e-12 * Wo.norm1() and abs(n2 - Wo.norm2sq()) <= 1e-12 * Wo.norm2sq()
        W.inverse(); Wo.inverse()
        close(W.image, Wo.image, op + " inverse f64")
        W.soft_threshold(1.0)                      # refused after inverse (wt.cu:309-312), no exception
        with pytest.raises(RuntimeError):
            W.coeffs


class _FixedRand:
    def __init__(self, vals):
        self.vals = list(vals)

    def rand(self):
        return self.vals.pop(0)


def test_cycle_spinning_f64():
    img = _img((96, 128), 11)
    W = _W64(img, "db2", 2, do_cycle_spinning=1)
    for it in range(2):
        W.forward(img)
        sr, sc = W.current_shift
        Wo = O.OracleWavelets(img, "db2", 2, do_cycle_spinning=1, rng=_FixedRand([sr, sc]), double_build=True)
        Wo.forward()
        assert np.array_equal(W.image, np.roll(img, (sr, sc), axis=(0, 1)))
        compare(W, Wo, "cycle spinning f64")
        W.inverse()
        assert np.abs(W.image - img).max() <= 1e-6 * np.abs(img).max()


def test_f64_is_more_accurate_than_f32():
    """The point of the double build: reconstruction error ~1e-13 relative instead of ~1e-6."""
    import pycudwt
    img = _img((512, 512), 13)
    W = _W64(img, "db8", 5); W.forward(); W.inverse()
    e64 = np.abs(W.image - img).max() / np.abs(img).max()
    V = pycudwt.Wavelets(img.astype(np.float32), "db8", 5); V.forward(); V.inverse()
    e32 = np.abs(V.image - img.astype(np.float32)).max() / np.abs(img).max()
    assert e64 < 1e-9 and e32 > 20 * e64, (e64, e32)


def test_errors_f64():
    img = _img((64, 64), 1)
    with pytest.raises(ValueError):
        _W64(img, "nope", 2)
    with pytest.raises(ValueError):
        _W64(img[:4, :4], "db20", 1)
    with pytest.raises(ValueError):
        _W64(img[0], "db2", 2, do_cycle_spinning=1, ndim=1)
    W = _W64(img, "db2", 2)
    with pytest.raises(ValueError):
        W.set_image(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        W.set_coeff(np.zeros(5), 1)


EVEN_BANKS_4_20 = ["haar", "db2", "db3", "db4", "db5", "sym6", "db7", "sym8", "db9", "db10", "bior2.2", "bior3.1", "rbio2.8", "bior6.8", "coif2", "coif3", "db12", "coif5", "db20"]


@pytest.mark.parametrize("shape", [(300, 520), (257, 1031), (2, 131, 258), (64, 64), (5, 7)])
@pytest.mark.parametrize("wname", EVEN_BANKS_4_20)
def test_fused_level_kernels_f64(wname, shape):
    """kernels_f64_fused.cu (row + column pass of a level in one launch, F = 4 .. 40 and the Haar butterfly): several 128-column strips and row
    segments, odd sizes in both directions (the repeated last sample of the analysis, the clipped last row / column of the
    synthesis), a stack, a size smaller than the filter -- against the double-build oracle."""
    img = _img(shape, 11)
    imgs = img if img.ndim == 3 else img[None]
    try:
        Wos = [O.OracleWavelets(x, wname, 4, double_build=True) for x in imgs]
    except ValueError:
        pytest.skip("not a valid configuration")
    W = _W64(img, wname, 4)
    l0 = W.launch_count
    W.forward()
    assert W.levels == Wos[0].levels
    if True:
        assert W.launch_count - l0 == W.levels, "one launch per level expected (fused row + column pass)"
    for Wo in Wos:
        Wo.forward()
    c = W.coeffs
    for k, Wo in enumerate(Wos):
        pick = (lambda a: a[k]) if img.ndim == 3 else (lambda a: a)
        close(pick(c[0]), Wo.coeffs[0], "fused f64 A", wname=wname)
        for i in range(1, len(c)):
            for j in range(3):
                close(pick(c[i][j]), Wo.coeffs[i][j], "fused f64 L%d b%d %s" % (i, j, wname), wname=wname)
    W.inverse()
    for k, Wo in enumerate(Wos):
        Wo.inverse()
        close(W.image[k] if img.ndim == 3 else W.image, Wo.image, "fused f64 inverse " + wname, wname=wname)


def test_fused_level_kernels_f64_match_the_two_pass_kernels():
    """Same process image, two library instances: PWT_F64_FUSED=0 selects the two-pass kernels.  The analysis keeps their
    summation order (bit-identical bands); the synthesis runs rows before columns (fp64 rounding apart)."""
    import os, subprocess, sys, tempfile
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import pypwt_b200
rng = np.random.default_rng(5)
out = {}
for wn, shape in (("db2", (515, 770)), ("sym8", (384, 640)), ("db10", (3, 200, 300))):
    img = rng.standard_normal(shape) * 40 + 100
    W = pypwt_b200.Wavelets64(img, wn, 3)
    W.forward()
    c = W.coeffs
    out[wn + "_A"] = c[0]
    for i in range(1, len(c)):
        for j in range(3): out["%%s_%%d_%%d" %% (wn, i, j)] = c[i][j]
    W.inverse()
    out[wn + "_img"] = W.image
np.savez(sys.argv[1], **out)
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for mode in ("1", "0"):
            path = os.path.join(d, "m%s.npz" % mode)
            env = dict(os.environ, PWT_F64_FUSED=mode)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            with np.load(path) as z:
                res[mode] = {k: z[k] for k in z.files}
    assert set(res["0"]) == set(res["1"]) and len(res["1"]) > 20
    for k in res["1"]:
        if k.endswith("_img"):
            assert np.abs(res["1"][k] - res["0"][k]).max() <= 1e-12 * 255, k
        else:
            assert np.array_equal(res["1"][k], res["0"][k]), k


def test_custom_banks_f64():
    """`Wavelets64.set_wavelets_filters` (wt.cu:558-600 in the DOUBLEPRECISION build, separable banks): a built-in bank loaded as
    a custom one is bit-identical to the built-in plan; odd-length banks (CDF 9/7, LeGall 5/3, random 7 taps: the `hlen & 1`
    branches of separable.cu:98-102, 251-264) agree with the fp32 plans, which are pinned to the reference's CUDA build for
    exactly these banks (test_gpu_api.py); a custom 2-tap bank is a filter bank, not the Haar butterfly; non-separable plans refuse."""
    import pycudwt
    import pypwt_b200
    from test_gpu_api import BANKS
    img = _img((97, 150), 21)
    for wn, kw in (("db3", {}), ("sym8", {}), ("db2", dict(do_swt=1)), ("db4", dict(ndim=1))):
        L, H, IL, IH = pypwt_b200.lookup_filters64(wn)
        A = _W64(img, wn, 3, **kw); B = _W64(img, "db2" if wn != "db2" else "db3", 3, **kw)
        B.set_wavelets_filters("mine", L, H, IL, IH)
        assert B.hlen == A.hlen and B.wname == "mine"
        if B.levels != A.levels:                             # the level count was fixed at construction (wt.cu:156-165)
            continue
        A.forward(); B.forward()
        for a, b in zip(_flat(A.coeffs), _flat(B.coeffs)):
            assert np.array_equal(a, b)
        A.inverse(); B.inverse()
        assert np.array_equal(A.image, B.image)
    for bank, taps in BANKS.items():
        for do_swt in (0, 1):
            W = _W64(img, "db3", 2, do_swt=do_swt)
            W.set_wavelets_filters(bank, taps["lo"], taps["hi"], taps["ilo"], taps["ihi"])
            F = pycudwt.Wavelets(img.astype(np.float32), "db3", 2, do_swt=do_swt)
            F.set_wavelets_filters(bank, *[np.asarray(taps[k], np.float32) for k in ("lo", "hi", "ilo", "ihi")])
            assert W.hlen == F.hlen
            W.forward(); F.forward()
            for a, b in zip(_flat(W.coeffs), _flat(F.coeffs)):
                assert a.dtype == np.float64 and np.abs(a - b).max() <= 2e-5 * max(255.0, np.abs(b).max()), (bank, do_swt)
            W.inverse(); F.inverse()
            assert np.abs(W.image - F.image).max() <= 2e-5 * max(255.0, np.abs(F.image).max()), (bank, do_swt)
    s = 2.0 ** -0.5
    W = _W64(img, "haar", 2); Hh = _W64(img, "haar", 2)
    W.set_wavelets_filters("haar2", [s, s], [-s, s], [s, s], [s, -s])
    W.forward(); Hh.forward()
    for a, b in zip(_flat(W.coeffs), _flat(Hh.coeffs)):
        close(a, b, "2-tap custom bank vs the Haar butterfly")
    W.inverse()
    close(W.image, img, "2-tap custom bank round trip")
    with pytest.raises(ValueError):
        _W64(img, "db2", 2, do_separable=0).set_wavelets_filters("x", [1, 2], [1, 2], [1, 2], [1, 2])


def _flat(coeffs):
    out = []
    for c in coeffs:
        out.extend(c if isinstance(c, (list, tuple)) else [c])
    return out
