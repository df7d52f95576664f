"""CPU tests of the oracle: closed forms vs the literal emulation of the reference kernels,
known-answer vectors, golden vectors produced by the reference's own CUDA build
(tests/golden/pdwt_golden.npz, generated on a B200 by tests/golden/make_golden.py), and
PyWavelets when it happens to be importable."""
import os

import numpy as np
import pytest

from conftest import ROOT, synth_image
from oracle import pdwt_oracle as O
from oracle import ref_emulation as E

GOLDEN = os.path.join(ROOT, "tests", "golden", "pdwt_golden.npz")


def test_glibc_rand_sequence():
    g = O.GlibcRand()
    assert [g.rand() for _ in range(4)] == [1804289383, 846930886, 1681692777, 1714636915]


def test_geometry():
    assert [O.div2(n) for n in (1, 2, 3, 8, 9)] == [1, 1, 2, 4, 5]
    assert O.max_level(512, 512, 4, 2) == 7 and O.max_level(8192, 8192, 40, 2) == 7
    assert O.max_level(2048, 2048, 16, 2) == 7 and O.max_level(1, 1000, 2, 1) == 9
    assert O.band_sizes(511, 509, 3, 0, 2) == [(256, 255), (128, 128), (64, 64)]
    assert O.band_sizes(4, 1001, 2, 0, 1) == [(4, 501), (4, 251)]
    assert O.band_sizes(5, 7, 2, 1, 2) == [(5, 7)] * 2


def test_filter_relations_all_72():
    assert len(O.WAVELET_NAMES) == 72
    for w in O.WAVELET_NAMES:
        L, H, IL, IH = O.filters(w)
        F = L.size
        assert F % 2 == 0 and F <= 40
        # two-channel perfect reconstruction conditions of the bank (double shift orthogonality)
        for k in range(-(F // 2) + 1, F // 2):
            s = sum(L[n] * IL[F - 1 - n + 2 * k] for n in range(F) if 0 <= F - 1 - n + 2 * k < F)
            assert abs(s - (1.0 if k == 0 else 0.0)) < 5e-6, (w, k, s)


def test_known_answers():
    """SURVEY.md 8c small vectors (fp64 closed forms)."""
    x = np.arange(8.0)
    L, H, IL, IH = O.filters("haar")
    np.testing.assert_allclose(O.analysis(x, L), [0.7071067812, 3.5355339059, 6.3639610307, 9.1923881554], atol=2e-7)
    np.testing.assert_allclose(O.analysis(x, H), [-0.7071067812] * 4, atol=2e-7)
    L, H, IL, IH = O.filters("db2")
    np.testing.assert_allclose(O.analysis(x, L), [3.3460652150, 2.3107890345, 5.1392161593, 9.0029194644], atol=1e-6)
    np.testing.assert_allclose(O.analysis(x, H), [-1.0352761804, 0, 0, 3.8637033052], atol=1e-6)
    xo = np.arange(7.0)
    np.testing.assert_allclose(O.analysis(xo, L), [2.8631023018, 2.3107890345, 5.1392161593, 8.7787755964], atol=1e-6)
    np.testing.assert_allclose(O.analysis(xo, H), [-0.9058666579, 0, 0, 3.0271870014], atol=1e-6)
    A, Hb, V, D = O.dwt2_level(np.arange(16.0).reshape(4, 4), L, H)
    np.testing.assert_allclose(A, [[10, 12], [18, 20]], atol=1e-5)
    np.testing.assert_allclose(Hb, [[-2.9282032303] * 2, [10.9282032303] * 2], atol=1e-5)
    np.testing.assert_allclose(V, [[-0.7320508076, 2.7320508076]] * 2, atol=1e-5)
    np.testing.assert_allclose(D, 0, atol=1e-5)
    np.testing.assert_allclose(O.swt_analysis(x, L, 1), [3.3460652150, 0.8965754722, 2.3107890345, 3.7250025969,
                                                          5.1392161593, 6.5534297217, 9.0029194644, 8.6239820825], atol=1e-6)
    np.testing.assert_allclose(O.swt_analysis(x, H, 1), [-1.0352761804, 0, 0, 0, 0, 0, 3.8637033052, -2.8284271247], atol=1e-6)
    a1 = O.swt_analysis(x, L, 1)
    np.testing.assert_allclose(O.swt_analysis(a1, L, 2), [7, 4.9019237886, 3.5358983849, 3.9019237886, 7, 9.0980762114,
                                                           10.4641016151, 10.0980762114], atol=2e-6)


WN = ["haar", "db2", "db3", "db4", "sym5", "coif2", "bior2.2", "bior3.1", "rbio3.1", "bior6.8", "db9"]


@pytest.mark.parametrize("shape", [(16, 24), (15, 21), (18, 17)])
@pytest.mark.parametrize("wname", WN)
def test_closed_forms_match_kernel_emulation(wname, shape):
    L, H, IL, IH = O.filters(wname)
    if min(shape) < 2 * (L.size - 1):
        pytest.skip("image smaller than the level-1 support")
    x = np.random.default_rng(0).standard_normal(shape)
    lo, hi = E.fwd_rows(x, L, H)
    np.testing.assert_allclose(lo, O.analysis(x, L), atol=1e-12)
    np.testing.assert_allclose(hi, O.analysis(x, H), atol=1e-12)
    bands = O.dwt2_level(x, L, H)
    for u, v in zip(E.fwd_cols(lo, hi, L, H), bands):
        np.testing.assert_allclose(u, v, atol=1e-12)
    t1, t2 = E.inv_cols(*bands, IL, IH, shape[0])
    np.testing.assert_allclose(E.inv_rows(t1, t2, IL, IH, shape[1]), O.idwt2_level(*bands, IL, IH, shape), atol=1e-12)
    np.testing.assert_allclose(O.idwt2_level(*bands, IL, IH, shape), x, atol=2e-6)
    for lev in (1, 2):
        if (L.size - 1) * (1 << lev) > min(shape):
            continue
        slo, shi = E.swt_rows(x, L, H, lev)
        np.testing.assert_allclose(slo, O.swt_analysis(x, L, lev), atol=1e-12)
        np.testing.assert_allclose(shi, O.swt_analysis(x, H, lev), atol=1e-12)
        np.testing.assert_allclose(E.iswt_rows(slo, shi, IL, IH, lev), O.swt_synthesis(slo, shi, IL, IH, lev), atol=1e-12)
        np.testing.assert_allclose(O.swt_synthesis(slo, shi, IL, IH, lev), x, atol=2e-6)


@pytest.mark.parametrize("shape", [(12, 16), (11, 13)])
@pytest.mark.parametrize("wname", ["db2", "db3", "bior2.2", "sym4"])
def test_nonseparable_slot_swap(wname, shape):
    """The reference's 2D-stencil kernels (nonseparable.cu) equal the separable transform with the
    H and V slots exchanged (quirk Q1) -- for forward, inverse, and the a-trous variants."""
    L, H, IL, IH = O.filters(wname)
    if min(shape) < 2 * (L.size - 1):
        pytest.skip("too small")
    x = np.random.default_rng(1).standard_normal(shape)
    A, Hb, V, D = O.dwt2_level(x, L, H)
    a, h, v, d = E.ns_forward(x, E.ns_filters(L, H))
    for u, w in ((a, A), (h, V), (v, Hb), (d, D)):
        np.testing.assert_allclose(u, w, atol=1e-12)
    rec = E.ns_inverse(a, h, v, d, E.ns_filters(IL, IH), shape)
    np.testing.assert_allclose(rec, O.idwt2_level(A, Hb, V, D, IL, IH, shape), atol=1e-12)
    A, Hb, V, D = O.swt2_level(x, L, H, 1)
    a, h, v, d = E.ns_forward_swt(x, E.ns_filters(L, H), 1)
    for u, w in ((a, A), (h, V), (v, Hb), (d, D)):
        np.testing.assert_allclose(u, w, atol=1e-12)
    rec = E.ns_inverse_swt(a, h, v, d, E.ns_filters(IL, IH), 1)
    np.testing.assert_allclose(rec, O.iswt2_level(A, Hb, V, D, IL, IH, 1), atol=1e-12)


@pytest.mark.parametrize("shape", [(8, 12), (7, 9)])
def test_haar_kernels(shape):
    x = np.random.default_rng(2).standard_normal(shape)
    L, H, IL, IH = O.filters("haar")
    bands = O.dwt2_level(x, L, H, haar=True)
    for u, v in zip(E.haar2d_fwd(x), bands):
        np.testing.assert_allclose(u, v, atol=1e-14)
    # the dedicated butterfly equals the generic filter bank up to the fp32 rounding of 1/sqrt(2)
    for u, v in zip(O.dwt2_level(x, L, H), bands):
        np.testing.assert_allclose(u, v, atol=1e-6)
    np.testing.assert_allclose(E.haar2d_inv(*bands, shape), O.idwt2_level(*bands, IL, IH, shape, haar=True), atol=1e-14)
    np.testing.assert_allclose(O.idwt2_level(*bands, IL, IH, shape, haar=True), x, atol=1e-14)


@pytest.mark.parametrize("cfg", [dict(), dict(do_swt=1), dict(ndim=1), dict(do_separable=0), dict(do_swt=1, ndim=1)])
@pytest.mark.parametrize("wname", ["haar", "db2", "sym8", "bior4.4"])
def test_oracle_class_roundtrip_and_state(wname, cfg):
    img = synth_image((96, 80), seed=3)
    W = O.OracleWavelets(img, wname, 3, **cfg)
    W.forward()
    n2 = W.norm2sq()
    if wname != "bior4.4" and not cfg.get("do_swt"):
        assert abs(n2 - float((img.astype(np.float64) ** 2).sum())) < 1e-5 * n2   # orthogonal: Parseval
    W.inverse()
    assert np.abs(W.image - img).max() < 1e-3
    with pytest.raises(RuntimeError):
        W.coeffs
    W.forward()
    assert len(W.coeffs) == W.levels + 1


def test_threshold_schedule():
    b = O.beta_schedule(10.0, 3, 1)
    assert np.allclose(b, [10 / 2 ** 0.5, 5.0, 10 / 2 ** 1.5], rtol=1e-6)
    assert O.beta_schedule(10.0, 2, 0) == [np.float32(10.0)] * 2
    assert np.isclose(O.beta_appcoeffs(10.0, 3, 1), 10 / 2 ** 1.5, rtol=1e-6)
    assert O.beta_appcoeffs(10.0, 4, 1) == np.float32(2.5)
    v = np.array([-3.0, -1.0, 0.0, 1.0, 3.0])
    assert np.array_equal(O.soft_thresh(v, 1.0), [-2, -0.0, 0, 0, 2])
    assert np.array_equal(O.hard_thresh(v, 1.0), [-3, 0, 0, 0, 3])       # strict >
    assert np.array_equal(O.proj_linf(v, 2.0), [-2, -1, 0, 1, 2])
    assert np.array_equal(O.circshift(np.arange(6).reshape(2, 3), 1, -1), np.roll(np.arange(6).reshape(2, 3), (1, -1), (0, 1)))


# ---- golden vectors from the reference's own CUDA build ---------------------------------------
def _golden():
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/pdwt_golden.npz not generated yet (needs a GPU run of make_golden.py)")
    return np.load(GOLDEN)


def _flat(c):
    out = []
    for b in c:
        out += list(b) if isinstance(b, list) else [b]
    return out


GOLD_MODES = {"dwt2": dict(), "swt2": dict(do_swt=1), "dwt1": dict(ndim=1), "swt1": dict(do_swt=1, ndim=1),
              "ns_dwt2": dict(do_separable=0), "ns_swt2": dict(do_separable=0, do_swt=1)}


def test_oracle_matches_reference_cuda_golden():
    """Every stored output of PDWT's CUDA kernels is reproduced by the oracle to fp32 accuracy
    (tolerance 1e-5 * max(255, |band|max): the reference computes in fp32, the oracle in fp64)."""
    g = _golden()
    keys = sorted(k for k in g.files if k.endswith("/levels"))
    assert len(keys) > 100
    worst = 0.0
    for k in keys:
        iname, wname, tag = k.split("/")[:3]
        base = "%s/%s/%s" % (iname, wname, tag)
        img = g["in/" + iname]
        W = O.OracleWavelets(img, wname, 2, **GOLD_MODES[tag])
        assert W.levels == int(g[k][0]), base
        W.forward()

        def cmp(prefix, coeffs):
            nonlocal worst
            for i, b in enumerate(coeffs):
                r = g["%s/%s%d" % (base, prefix, i)]
                assert r.shape == b.shape, (base, prefix, i)
                tol = 1e-5 * max(255.0, np.abs(r).max())
                err = np.abs(b.astype(np.float64) - r).max()
                worst = max(worst, err / tol)
                assert err <= tol, "%s %s%d err %.3e tol %.3e" % (base, prefix, i, err, tol)

        cmp("c", _flat(W.coeffs))
        if base + "/norm1" in g.files:
            assert abs(W.norm1() - g[base + "/norm1"][0]) <= 2e-5 * W.norm1()
            if base + "/norm2sq" in g.files:
                assert abs(W.norm2sq() - g[base + "/norm2sq"][0]) <= 2e-5 * W.norm2sq()
            W.soft_threshold(10.0, 1, 1)
            cmp("soft", _flat(W.coeffs))
            W.forward()
            W.hard_threshold(10.0, 1, 1)
            nbad = 0
            for i, b in enumerate(_flat(W.coeffs)):
                r = g["%s/hard%d" % (base, i)]
                nbad += int((np.abs(b - r) > 1e-5 * max(255.0, np.abs(r).max())).sum())
            assert nbad <= 2, base
            W.forward()
            W.shrink(0.5, 1)
            r = g[base + "/shrinkA"]
            assert np.abs(W.coeff_only(0) - r).max() <= 1e-5 * max(255.0, np.abs(r).max())
            W.forward()
        W.inverse()
        r = g[base + "/inv"]
        assert r.shape == W.image.shape
        assert np.abs(W.image - r).max() <= 1e-5 * max(255.0, np.abs(r).max()), base
    print("worst err/tol over golden set: %.3f" % worst)


def test_cycle_spinning_golden():
    g = _golden()
    img = g["in/even"]
    W = O.OracleWavelets(img, "db2", 2, do_cycle_spinning=1)     # fresh GlibcRand
    W.forward()
    assert np.array_equal(W.image, g["cs/shifted_image"])
    assert np.abs(W.coeff_only(0) - g["cs/A"]).max() <= 1e-5 * np.abs(g["cs/A"]).max()
    W.inverse()
    assert np.abs(W.image - g["cs/inv"]).max() <= 1e-5 * 255


def test_against_pywt_if_available():
    pywt = pytest.importorskip("pywt")
    img = synth_image((64, 64), seed=5).astype(np.float64)
    for wname in ["haar", "db2", "db4", "sym8", "coif2", "bior2.2"]:
        W = O.OracleWavelets(img, wname, 3)
        W.forward()
        ref = pywt.wavedec2(img, wname, mode="periodization", level=3)
        c = W.coeffs
        assert np.abs(c[0] - ref[0]).max() < 1e-3
        for i in range(3):
            for j in range(3):
                assert np.abs(c[i + 1][j] - ref[3 - i][j]).max() < 1e-3


@pytest.mark.parametrize("shape", [(64, 96), (63, 97), (100, 77)])
@pytest.mark.parametrize("wname", ["haar", "db2", "sym4", "bior2.2", "db8"])
def test_c_port_matches_numpy_oracle(wname, shape):
    """oracle/dwt_cpu.c (the CPU baseline of bench.py) computes the same transform as the numpy oracle."""
    from oracle import dwt_cpu
    x = synth_image(shape, seed=6)
    P = dwt_cpu.CpuDwt2(shape, wname, 3)
    b = P.forward(x)
    W = O.OracleWavelets(x, wname, 3)
    assert P.levels == W.levels
    W.forward()
    c = W.coeffs
    tol = 1e-5 * max(255.0, np.abs(c[0]).max())
    assert np.abs(b[0] - c[0]).max() <= tol
    for i in range(P.levels):
        for j in range(3):
            assert np.abs(b[3 * i + 1 + j] - c[i + 1][j]).max() <= tol
    assert np.abs(P.inverse() - x).max() <= 1e-5 * 255 * (20 if wname in ("bior3.1", "rbio3.1") else 1) * 4


# ---- odd-length custom banks ----------------------------------------------------------------
def _pad(k, mode):
    """The even-length bank (len + 1 taps) that reproduces the reference's odd-length windows; the product applies the
    same paddings in pwt_set_filters_forward / _inverse (pwt_plan.cu, "custom filters")."""
    k = np.asarray(k, dtype=np.float64)
    z = np.zeros(1)
    if mode == "front":
        return np.concatenate([z, k])
    if mode == "drop0":
        return np.concatenate([z, k[1:], z])
    return np.concatenate([k, z])          # "back"


@pytest.mark.parametrize("hlen", [3, 5, 7, 9, 11])
@pytest.mark.parametrize("shape", [(12, 16), (11, 13)])
def test_odd_length_banks_equal_padded_even_banks(hlen, shape):
    """Literal emulation of the reference kernels with an ODD number of taps (separable.cu:98-102, 251-264, 416-420,
    559-568; nonseparable.cu:125, 182-190, 311, 367) against the same kernels fed the padded even-length bank."""
    rng = np.random.default_rng(100 + hlen)
    img = rng.standard_normal(shape)
    fL, fH, iL, iH = (rng.standard_normal(hlen) for _ in range(4))
    # analysis (DWT)
    a = E.fwd_cols(*E.fwd_rows(img, fL, fH), fL, fH)
    b = E.fwd_cols(*E.fwd_rows(img, _pad(fL, "front"), _pad(fH, "front")), _pad(fL, "front"), _pad(fH, "front"))
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, atol=1e-12)
    # synthesis (DWT): tap 0 of an odd-length synthesis filter is never read by the reference
    cA, cH, cV, cD = a
    t = E.inv_cols(cA, cH, cV, cD, iL, iH, shape[0])
    r1 = E.inv_rows(t[0], t[1], iL, iH, shape[1])
    pl, ph = _pad(iL, "drop0"), _pad(iH, "drop0")
    t = E.inv_cols(cA, cH, cV, cD, pl, ph, shape[0])
    r2 = E.inv_rows(t[0], t[1], pl, ph, shape[1])
    np.testing.assert_allclose(r1, r2, atol=1e-12)
    # SWT analysis / synthesis rows (the column passes use the same index arithmetic)
    for level in (1, 2):
        if (hlen - 1) * (1 << (level - 1)) >= min(shape):
            continue
        lo1, hi1 = E.swt_rows(img, fL, fH, level)
        lo2, hi2 = E.swt_rows(img, _pad(fL, "front"), _pad(fH, "front"), level)
        np.testing.assert_allclose(lo1, lo2, atol=1e-12)
        np.testing.assert_allclose(hi1, hi2, atol=1e-12)
        np.testing.assert_allclose(E.iswt_rows(lo1, hi1, iL, iH, level),
                                   E.iswt_rows(lo1, hi1, _pad(iL, "back"), _pad(iH, "back"), level), atol=1e-12)


def _pad2(K, mode):
    n = K.shape[-1]
    out = np.zeros((4, n + 1, n + 1))
    if mode == "front":
        out[:, 1:, 1:] = K
    elif mode == "drop0":
        out[:, 1:n, 1:n] = K[:, 1:, 1:]
    else:
        out[:, :n, :n] = K
    return out


@pytest.mark.parametrize("hlen", [3, 5, 7])
def test_odd_length_nonseparable_banks_equal_padded_even_banks(hlen):
    rng = np.random.default_rng(200 + hlen)
    shape = (12, 14)
    img = rng.standard_normal(shape)
    K = rng.standard_normal((4, hlen, hlen))
    IK = rng.standard_normal((4, hlen, hlen))
    a = E.ns_forward(img, K)
    b = E.ns_forward(img, _pad2(K, "front"))
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, atol=1e-12)
    np.testing.assert_allclose(E.ns_inverse(*a, IK, shape), E.ns_inverse(*a, _pad2(IK, "drop0"), shape), atol=1e-12)
    a = E.ns_forward_swt(img, K, 1)
    b = E.ns_forward_swt(img, _pad2(K, "front"), 1)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, atol=1e-12)
    np.testing.assert_allclose(E.ns_inverse_swt(*a, IK, 1), E.ns_inverse_swt(*a, _pad2(IK, "back"), 1), atol=1e-12)


def test_volume_oracle_properties():
    """oracle/dwt3_oracle.py (composition of the pinned 1D closed forms): perfect reconstruction for even and odd sizes,
    energy preservation for orthogonal banks on even sizes, and consistency with the 2D oracle on a volume that is
    constant along z."""
    from oracle import dwt3_oracle as D
    rng = np.random.default_rng(0)
    for shp in ((40, 44, 48), (31, 37, 43)):
        v = rng.standard_normal(shp).astype(np.float32)
        for wn in ("haar", "db2", "bior2.4"):
            W = D.OracleWavelets3D(v, wn, 2)
            W.forward()
            if shp[0] % 4 == 0 and wn != "bior2.4":
                assert abs(W.norm2sq() / (v.astype(np.float64) ** 2).sum() - 1) < 1e-6
            assert set(W.coeffs[1]) == set(D.KEYS)
            W.inverse()
            assert np.abs(W.image - v).max() < 2e-6
    sl = rng.standard_normal((16, 24)).astype(np.float32)
    W = D.OracleWavelets3D(np.repeat(sl[None], 8, axis=0), "db2", 1)
    W.forward()
    W2 = O.OracleWavelets(sl, "db2", 1)
    W2.forward()
    assert np.abs(W.coeffs[0][0] - np.sqrt(2) * W2.coeffs[0]).max() < 1e-5
    assert np.abs(W.coeffs[1]["ada"][0] - np.sqrt(2) * W2.coeffs[1][0]).max() < 1e-5
    assert np.abs(W.coeffs[1]["aad"][0] - np.sqrt(2) * W2.coeffs[1][1]).max() < 1e-5
    assert np.abs(W.coeffs[1]["daa"]).max() < 1e-5
    try:
        import pywt
    except ImportError:
        return
    v = rng.standard_normal((32, 40, 48)).astype(np.float32)
    W = D.OracleWavelets3D(v, "db3", 2)
    W.forward()
    ref = pywt.wavedecn(v.astype(np.float64), "db3", mode="periodization", level=2)
    assert np.abs(W.coeffs[0] - ref[0]).max() < 1e-5
    for lev in (1, 2):
        for k in D.KEYS:
            assert np.abs(W.coeffs[lev][k] - ref[-lev][k]).max() < 1e-5


def test_double_build_oracle():
    """double_build=True (the DOUBLEPRECISION build): samples, table and thresholds stay float64."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((64, 80))
    W = O.OracleWavelets(x, "db4", 3, double_build=True)
    W.forward()
    assert W.coeffs[0].dtype == np.float64 and W.coeffs[1][0].dtype == np.float64
    W.soft_threshold(0.0)
    W.inverse()
    assert np.abs(W.image - x).max() < 1e-9            # limited by the table's own precision
    W32 = O.OracleWavelets(x, "db4", 3)
    W32.forward(); W32.inverse()
    assert np.abs(W32.image - x).max() > 1e-9          # fp32-rounded samples and taps
