"""GPU tests of the data path and API surface around the kernels:
odd-length custom banks against the reference's CUDA build, single-copy `coeffs`, device-array interop
(`__cuda_array_interface__`, DLPack, device inputs = the reference's memisonhost=0 / mem_is_on_device=1 paths),
copy() of a custom-bank plan, thread safety of the launchers, and one forced-mode test per fallback kernel family.
"""
import os
import sys
import threading

import numpy as np
import pytest

from conftest import ROOT, synth_image
from oracle import pdwt_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _mine():
    import pycudwt
    return pycudwt


def _ref():
    p = os.path.join(ROOT, "oracle", "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    try:
        import pycudwt_ref
    except ImportError as e:
        pytest.skip("reference build oracle/_ref not available: %s" % e)
    return pycudwt_ref


def flat(c):
    out = []
    for b in c:
        out += [np.array(x) for x in b] if isinstance(b, list) else [np.array(b)]
    return out


def close(g, r, scale, what, k=1.0):
    assert g.shape == r.shape, what
    tol = k * RTOL * max(scale, float(np.abs(r).max()))
    err = float(np.abs(g.astype(np.float64) - r).max())
    assert err <= tol, "%s: err %.3e > %.3e" % (what, err, tol)


# ---- odd-length custom banks (demo.cpp:83-179 without its zero padding) ---------------------------------------------
CDF97 = dict(
    lo=[0.026748757411, -0.016864118443, -0.078223266529, 0.266864118443, 0.602949018236, 0.266864118443,
        -0.078223266529, -0.016864118443, 0.026748757411],
    hi=[0.0, 0.091271763114, -0.057543526229, -0.591271763114, 1.11508705, -0.591271763114, -0.057543526229,
        0.091271763114, 0.0],
    ilo=[0.0, -0.091271763114, -0.057543526229, 0.591271763114, 1.11508705, 0.591271763114, -0.057543526229,
         -0.091271763114, 0.0],
    ihi=[0.026748757411, 0.016864118443, -0.078223266529, -0.266864118443, 0.602949018236, -0.266864118443,
         -0.078223266529, 0.016864118443, 0.026748757411])
LEGALL53 = dict(lo=[-1 / 8, 2 / 8, 6 / 8, 2 / 8, -1 / 8], hi=[-0.5, 1.0, -0.5, 0.0, 0.0],
                ilo=[0.5, 1.0, 0.5, 0.0, 0.0], ihi=[-1 / 8, -2 / 8, 6 / 8, -2 / 8, -1 / 8])
RAND7 = {k: list(np.random.default_rng(5).standard_normal(7) * 0.4) for k in ("lo", "hi", "ilo", "ihi")}
BANKS = {"cdf97": CDF97, "legall53": LEGALL53, "rand7": RAND7}


@pytest.mark.parametrize("shape", [(128, 192), (97, 75)])
@pytest.mark.parametrize("do_swt", [0, 1])
@pytest.mark.parametrize("bank", list(BANKS))
def test_odd_length_custom_bank_vs_pdwt(bank, do_swt, shape):
    """An odd number of taps takes the `hlen & 1` branches of the reference kernels (separable.cu:98-102, 251-264,
    416-420, 559-568), which no built-in bank reaches; ours maps them onto even-length banks (pwt_plan.cu)."""
    ref, mine = _ref(), _mine()
    b = {k: np.asarray(v, dtype=np.float32) for k, v in BANKS[bank].items()}
    img = synth_image(shape, seed=90)
    out = {}
    for name, mod in (("ref", ref), ("mine", mine)):
        W = mod.Wavelets(img, "db3", 2, do_swt=do_swt)
        W.set_wavelets_filters(bank, b["lo"], b["hi"], b["ilo"], b["ihi"])
        W.forward()
        c = flat(W.coeffs)
        W.inverse()
        out[name] = (c, np.array(W.image))
        del W
    for i, (g, r) in enumerate(zip(out["mine"][0], out["ref"][0])):
        close(g, r, 255.0, "%s swt=%d band %d" % (bank, do_swt, i))
    close(out["mine"][1], out["ref"][1], 255.0, "%s swt=%d inverse" % (bank, do_swt))


@pytest.mark.parametrize("do_swt", [0, 1])
def test_odd_length_custom_nonseparable_bank_vs_kernel_emulation(do_swt):
    """Non-separable custom banks cannot be loaded through the reference's Python wrapper at all: it types the 2D
    filters as 1D memoryviews and passes len(LL) as the tap count (pypwt.pyx:520-537), so F x F arrays are rejected and
    flattened ones give hlen = F^2.  Ours takes F x F arrays; the check is the literal emulation of the reference's
    kernels (oracle/ref_emulation.py, nonseparable.cu:114-225, 303-449) with an odd number of taps, one level."""
    from oracle import ref_emulation as E
    mine = _mine()
    rng = np.random.default_rng(6)
    K = [(rng.standard_normal((5, 5)) * 0.3).astype(np.float32) for _ in range(8)]
    img = synth_image((48, 40), seed=91)
    W = mine.Wavelets(img, "db2", 1, do_separable=0, do_swt=do_swt)
    W.set_wavelets_filters("rand5x5", K[0], K[3], K[4], K[7], LH=K[1], HL=K[2], i_LH=K[5], i_HL=K[6])
    W.forward()
    c = flat(W.coeffs)
    W.inverse()
    KF = np.stack(K[0:4]).astype(np.float64)
    KI = np.stack(K[4:8]).astype(np.float64)
    x = img.astype(np.float64)
    if do_swt:
        r = E.ns_forward_swt(x, KF, 1)
        rimg = E.ns_inverse_swt(*r, KI, 1)
    else:
        r = E.ns_forward(x, KF)
        rimg = E.ns_inverse(*r, KI, img.shape)
    # slot order of the reference's non-separable transform (quirk Q1): the emulation returns (A, slot1, slot2, slot3)
    for i, (g, rr) in enumerate(zip(c, r)):
        close(g, rr, 255.0, "nonsep 5x5 swt=%d band %d" % (do_swt, i), k=4)
    close(np.array(W.image), rimg, 255.0, "nonsep 5x5 swt=%d inverse" % do_swt, k=4)


def test_copy_of_a_custom_bank_plan():
    """copy() used to rebuild a plan from wname, which is the user's label after set_wavelets_filters."""
    mine = _mine()
    img = synth_image((128, 128), seed=92)
    L, H, IL, IH = mine.lookup_filters("db3")
    W = mine.Wavelets(img, "db4", 2)
    W.set_wavelets_filters("my own bank", L, H, IL, IH)
    W.forward()
    C = W.copy()
    assert C.wname == "my own bank" and C.levels == W.levels
    for a, b in zip(flat(W.coeffs), flat(C.coeffs)):
        assert np.array_equal(a, b)
    C.inverse()
    W.inverse()
    assert np.array_equal(C.image, W.image)
    C.forward(img)      # the copy carries the custom filters, not db4's
    R = mine.Wavelets(img, "db3", 2)
    R.forward()
    for a, b in zip(flat(C.coeffs), flat(R.coeffs)):
        assert np.array_equal(a, b)


# ---- data path --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [dict(), dict(do_swt=1), dict(ndim=1), dict(do_swt=1, ndim=1)])
@pytest.mark.parametrize("shape", [(256, 512), (129, 75), (3, 64, 96)])
def test_coeffs_single_copy_equals_band_copies(cfg, shape):
    """`coeffs` moves every band with one D2H of the contiguous coefficient region; `coeff_only` copies one band."""
    if len(shape) == 3 and cfg.get("ndim") == 1:
        pytest.skip("stacks are a 2D extension")
    mine = _mine()
    img = synth_image(shape, seed=93)
    W = mine.Wavelets(img, "db3", 3, **cfg)
    W.forward()
    W.soft_threshold(5.0, 1)          # pending threshold: both read paths must flush it
    allc = [x.copy() for x in flat(W.coeffs)]
    W.forward()
    W.soft_threshold(5.0, 1)
    nb = len(allc)
    for b in range(nb):
        assert np.array_equal(np.array(W.coeff_only(b)), allc[b]), "band %d" % b
    Wo = O.OracleWavelets(img[0] if len(shape) == 3 else img, "db3", 3, **cfg)
    Wo.forward()
    Wo.soft_threshold(5.0, 1)
    for g, r in zip(allc, flat(Wo.coeffs)):
        close(g[0] if len(shape) == 3 else g, r, 255.0, "coeffs vs oracle")
    W.inverse()
    with pytest.raises(RuntimeError):
        W.coeffs       # refused after inverse(), like coeff_only (wt.cu:474-477)


def test_device_array_interop_with_torch():
    """Zero-copy views (`__cuda_array_interface__`, DLPack) and device inputs (memisonhost=0, wt.cu:145-150;
    mem_is_on_device=1, wt.cu:425-432, 435-465)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("torch has no CUDA device")
    mine = _mine()
    img = synth_image((256, 384), seed=94)
    t_img = torch.from_numpy(img).cuda()
    W = mine.Wavelets(t_img, "db2", 3)                 # constructed from a DEVICE array
    H = mine.Wavelets(img, "db2", 3)
    W.forward(); H.forward()
    for a, b in zip(flat(W.coeffs), flat(H.coeffs)):
        assert np.array_equal(a, b)
    # views of the bands: CUDA array interface and DLPack see the same memory
    W.sync()
    for num in (0, 1, 5, 9):
        v = W.coeff_device(num)
        a = torch.as_tensor(v, device="cuda")
        d = torch.from_dlpack(v)
        assert a.data_ptr() == d.data_ptr() == W.coeff_int_ptr(num)
        assert np.array_equal(a.cpu().numpy(), np.array(H.coeff_only(num)))
        assert np.array_equal(d.cpu().numpy(), np.array(H.coeff_only(num)))
    # in-place edit through the view is seen by the inverse and by the norms
    a = torch.as_tensor(W.coeff_device(1), device="cuda")
    a.mul_(0.5)
    torch.cuda.synchronize()
    h1 = np.array(H.coeff_only(1)) * 0.5
    H.set_coeff(h1, 1)
    assert abs(W.norm1() - H.norm1()) <= 1e-6 * H.norm1()
    W.inverse(); H.inverse()
    t_out = torch.as_tensor(W.image_device, device="cuda")
    W.sync()
    assert np.array_equal(t_out.cpu().numpy(), H.image)
    # device-side set_image / forward(img) / set_coeff
    t2 = torch.from_numpy(synth_image((256, 384), seed=95)).cuda()
    torch.cuda.synchronize()
    W.forward(t2)
    H.forward(t2.cpu().numpy())
    for x, y in zip(flat(W.coeffs), flat(H.coeffs)):
        assert np.array_equal(x, y)
    band = torch.zeros((32, 48), device="cuda")
    torch.cuda.synchronize()
    W.set_coeff(band, 0)
    assert not np.array(W.coeff_only(0)).any()
    with pytest.raises(ValueError):
        W.set_image(torch.zeros((8, 8), device="cuda"))
    with pytest.raises(ValueError):
        W.set_image(torch.zeros((256, 384), device="cuda", dtype=torch.float64))


# ---- thread safety of the launchers ---------------------------------------------------
@pytest.mark.parametrize("wname,shape", [("db2", (1024, 1024)), ("sym8", (768, 1024)), ("haar", (512, 2048))])
def test_two_threads_two_plans_with_fused_norms(wname, shape):
    """Two host threads drive two plans (own streams) through forward + norms + soft threshold + inverse with the GIL
    released inside the library calls.  The launchers used to share a process-global norm sink and function-static
    occupancy caches; every iteration must reproduce the single-threaded results bit for bit."""
    mine = _mine()
    imgs = [synth_image(shape, seed=300 + k, kind="smooth") for k in range(2)]

    def run(k, W, n, res):
        for _ in range(n):
            W.forward(imgs[k])
            n1, n2 = W.norms()          # arms the fused norm reduction of the following forwards
            W.soft_threshold(4.0, 0, 1)
            W.inverse()
            res.append((n1, n2, np.array(W.image)))

    single = []
    for k in range(2):
        W = mine.Wavelets(imgs[k], wname, 3)
        r = []
        run(k, W, 3, r)
        single.append(r[-1])
        assert r[0][0] == pytest.approx(r[-1][0], rel=1e-5)       # generic reduction (1st pass) vs fused partial sums
        del W
    plans = [mine.Wavelets(imgs[k], wname, 3) for k in range(2)]
    results = [[], []]
    threads = [threading.Thread(target=run, args=(k, plans[k], 100, results[k])) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(2):
        assert len(results[k]) == 100
        for i, (n1, n2, im) in enumerate(results[k]):
            if i >= 1:       # iteration 0 uses the un-fused reduction (different summation order)
                assert n1 == single[k][0] and n2 == single[k][1], "thread %d iteration %d: norms differ" % (k, i)
            assert np.array_equal(im, single[k][2]), "thread %d iteration %d: image differs" % (k, i)


def test_first_use_from_many_threads():
    """Several threads trigger the one-time per-device kernel set-up (dynamic shared memory opt-in) at the same time."""
    mine = _mine()
    img = synth_image((256, 512), seed=310)
    errs, outs = [], [None] * 6

    def work(i):
        try:
            W = mine.Wavelets(img, "db7", 2, do_swt=i & 1)      # F = 14: strip / SWT strip kernels with > 48 KB of smem
            W.forward()
            W.inverse()
            outs[i] = np.array(W.image)
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for o in outs:
        assert np.abs(o - img).max() < 2e-3


# ---- one forced-mode test per retained fallback family ----------------------------------------------------------------
@pytest.mark.parametrize("wname", ["haar", "db2", "db5", "sym8", "db10", "db12", "db16", "db20"])
@pytest.mark.parametrize("shape", [(256, 384), (130, 257)])
def test_mode2_shared_memory_families_against_oracle(wname, shape):
    """kernel mode 2 = kernels_fast.cu (F <= 20) and kernels_tile.cu (F >= 22; here forced from F >= 10 by the size
    rule) + kernels_swt.cu for the stationary transform: never the auto choice any more, kept as fallbacks."""
    mine = _mine()
    img = synth_image(shape, seed=95)
    for do_swt in (0, 1):
        if do_swt and wname == "haar":
            continue
        lev = 2
        W = mine.Wavelets(img, wname, lev, do_swt=do_swt)
        W.set_kernel_mode(2)
        Wo = O.OracleWavelets(img, wname, lev, do_swt=do_swt)
        W.forward(); Wo.forward()
        for g, r in zip(flat(W.coeffs), flat(Wo.coeffs)):
            close(g, r, 255.0, "mode 2 %s swt=%d" % (wname, do_swt))
        W.inverse(); Wo.inverse()
        close(W.image, Wo.image, 255.0, "mode 2 %s swt=%d inverse" % (wname, do_swt))
