"""Volumetric (3D) separable DWT (`Wavelets3D` -> pwt3_* C ABI; SURVEY 8f rank 4, the reference's stated gap
pdwt/README.md:29) against oracle/dwt3_oracle.py (composition of the pinned 1D closed forms along x, y, z).
Tolerance: 1e-5 * max(|x|max, |band|max), as for the 2D transforms."""
import numpy as np
import pytest

from conftest import synth_image
from oracle import dwt3_oracle as D

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _W3(*a, **k):
    import pypwt_b200
    return pypwt_b200.Wavelets3D(*a, **k)


def _vol(shape, seed):
    rng = np.random.default_rng(seed)
    base = synth_image(shape[1:], seed=seed)
    z = np.linspace(0, 3, shape[0], dtype=np.float32)[:, None, None]
    return (base[None] * (0.6 + 0.4 * np.cos(z)) + rng.standard_normal(shape).astype(np.float32) * 5).astype(np.float32)


def close(got, ref, what, scale=255.0):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    tol = RTOL * max(scale, float(np.abs(ref).max()))
    err = float(np.abs(got - ref).max())
    assert err <= tol, "%s: max err %.3e > %.3e" % (what, err, tol)


def compare(W, Wo, what):
    c, co = W.coeffs, Wo.coeffs
    assert len(c) == len(co)
    close(c[0], co[0], what + " aaa")
    for l in range(1, len(c)):
        assert sorted(c[l]) == sorted(co[l]) == sorted(D.KEYS)
        for k in D.KEYS:
            close(c[l][k], co[l][k], "%s L%d %s" % (what, l, k))


@pytest.mark.parametrize("shape", [(64, 96, 128), (33, 47, 59), (40, 256, 512), (130, 64, 72)])
@pytest.mark.parametrize("wname", ["haar", "db2", "db4", "sym8", "bior2.4", "rbio6.8", "db10"])
def test_dwt3_idwt3(wname, shape):
    vol = _vol(shape, 3)
    try:
        Wo = D.OracleWavelets3D(vol, wname, 3)
    except ValueError:
        pytest.skip("volume too small for this filter")
    W = _W3(vol, wname, 3)
    assert W.levels == Wo.levels
    W.forward(); Wo.forward()
    assert [tuple(s) for s in W.sizes] == [tuple(b["ddd"].shape) for b in Wo.coeffs[1:]]
    compare(W, Wo, "dwt3 " + wname)
    W.inverse(); Wo.inverse()
    close(W.image, Wo.image, "idwt3 " + wname)
    close(W.image, vol, "reconstruction " + wname, scale=255.0 * 4)


@pytest.mark.parametrize("shape", [(70, 130, 160), (37, 101, 96), (16, 72, 320), (66, 64, 200), (5, 9, 88)])
@pytest.mark.parametrize("wname", ["haar", "db2", "db3", "bior2.2", "sym2", "coif1", "rbio1.3", "db4", "sym4", "db5", "bior2.4"])
def test_fused_level_kernels_3d(wname, shape):
    """kernels_vol_fused.cu (x + y + z of a level in one launch, F = 2, 4, 6; 8- and 10-tap banks ride along on the two-launch path): several tiles with overhang, several z segments,
    odd heights and depths (the repeated last row / slice of the analysis, the clipped last one of the synthesis), widths where
    only the first level (or only the analysis: Nx % 4 == 0 but Nx % 8 != 0) takes the fused kernels."""
    vol = _vol(shape, 9)
    try:
        Wo = D.OracleWavelets3D(vol, wname, 2)
    except ValueError:
        pytest.skip("volume too small for this filter")
    W = _W3(vol, wname, 2)
    assert W.levels == Wo.levels
    W.forward(); Wo.forward()
    W1 = _W3(vol, wname, 1)                                  # a one-level plan counts the launches of the first level
    W1.forward()
    if Wo.L.size <= 6 and shape[2] % 4 == 0 and shape[2] >= 80 and shape[1] >= 8:
        assert W1.launch_count == 1, "analysis: one fused launch"
        W1.inverse()
        if shape[2] % 8 == 0:
            assert W1.launch_count == 2, "synthesis: one fused launch"
    compare(W, Wo, "fused dwt3 " + wname)
    W.inverse(); Wo.inverse()
    close(W.image, Wo.image, "fused idwt3 " + wname)
    close(W.image, vol, "reconstruction " + wname, scale=255.0 * 4)


def test_thresholds_norms_state_3d():
    vol = _vol((48, 64, 80), 5)
    for op in ("soft_threshold", "hard_threshold"):
        for app in (0, 1):
            W = _W3(vol, "db3", 2); Wo = D.OracleWavelets3D(vol, "db3", 2)
            W.forward(); Wo.forward()
            getattr(W, op)(12.5, app); getattr(Wo, op)(12.5, app)
            compare(W, Wo, op)
            n1, n2 = W.norms()
            assert abs(n1 - Wo.norm1()) <= 1e-5 * Wo.norm1() and abs(n2 - Wo.norm2sq()) <= 1e-5 * Wo.norm2sq()
            W.inverse(); Wo.inverse()
            close(W.image, Wo.image, op + " inverse")
            with pytest.raises(RuntimeError):
                W.coeffs
            W.inverse()                                      # refused with a warning, like the 2D plan (wt.cu:272-275)


def test_set_coeff_and_linearity_3d():
    vol = _vol((32, 48, 64), 7)
    W = _W3(vol, "db2", 2)
    W.forward()
    c = W.coeffs
    V = _W3(np.zeros_like(vol), "db2", 2)
    V.forward()
    V.set_coeff(c[0], 2, "aaa")
    for l in (1, 2):
        for k in D.KEYS:
            V.set_coeff(c[l][k], l, k)
    V.inverse()
    close(V.image, vol, "inverse of copied coefficients", scale=255.0 * 4)
    with pytest.raises(ValueError):
        V.set_coeff(np.zeros((2, 2, 2), np.float32), 1, "ddd")
    with pytest.raises(ValueError):
        _W3(vol[0], "db2", 1)
    with pytest.raises(ValueError):
        _W3(vol, "nope", 1)
    with pytest.raises(ValueError):
        _W3(vol[:4], "db8", 1)


def test_constant_along_z_reduces_to_2d():
    """A volume constant along z: the z high-pass bands vanish and the z low-pass bands are sqrt(2)^L times the 2D
    transform of a slice (ties the 3D path to the 2D path that is pinned against the reference)."""
    import pycudwt
    sl = synth_image((128, 192), seed=9)
    vol = np.repeat(sl[None], 32, axis=0)
    W = _W3(vol, "db2", 2); W.forward()
    W2 = pycudwt.Wavelets(sl, "db2", 2); W2.forward()
    c, c2 = W.coeffs, W2.coeffs
    close(c[0][0], 2.0 * c2[0], "aaa vs 2D A")
    for l, s in ((1, np.sqrt(2.0)), (2, 2.0)):
        close(c[l]["ada"][0], s * c2[l][0], "ada vs H")
        close(c[l]["aad"][0], s * c2[l][1], "aad vs V")
        close(c[l]["add"][0], s * c2[l][2], "add vs D")
        for k in ("daa", "dad", "dda", "ddd"):
            assert np.abs(c[l][k]).max() <= 1e-4 * np.abs(vol).max()
