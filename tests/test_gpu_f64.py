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
