"""GPU parity against the REFERENCE ITSELF: the unmodified PDWT + pypwt.pyx, recompiled for sm_100a
into oracle/_ref/ (oracle/Makefile), run side by side with our kernels on the same inputs.

One reference instance at a time: its filters live in process-global __constant__ memory (quirk Q4).
Skipped when oracle/_ref is not present (it is built in the build container and shipped by gpurun).
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, synth_image
from oracle import pdwt_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-5


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


def flat(c):
    out = []
    for b in c:
        out += [np.array(x) for x in b] if isinstance(b, list) else [np.array(b)]
    return out


def check(got, ref, scale, what):
    assert len(got) == len(ref), what
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.shape == r.shape, "%s band %d: %s vs %s" % (what, i, g.shape, r.shape)
        tol = RTOL * max(scale, float(np.abs(r).max()))
        if "bior3.1" in what or "rbio3.1" in what:
            tol *= 20    # ill-conditioned banks: the reference skips them itself (test_wavelets.py:176-181)
        err = float(np.abs(g.astype(np.float64) - r).max())
        assert err <= tol, "%s band %d: err %.3e > %.3e" % (what, i, err, tol)


IMG = synth_image((128, 128), seed=21)
MODES = {
    "dwt2": dict(),
    "swt2": dict(do_swt=1),
    "dwt_batched": dict(ndim=1),
    "swt_batched": dict(do_swt=1, ndim=1),
}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("wname", O.WAVELET_NAMES)
def test_forward_inverse_vs_pdwt(wname, mode):
    ref, mine = _ref(), _mine()
    kw = MODES[mode]
    lev = 999 if "swt" not in mode else 2
    R = ref.Wavelets(IMG, wname, lev, **kw)
    R.forward()
    rc = flat(R.coeffs)
    R.inverse()
    rimg = np.array(R.image)
    rlev = R.levels
    del R
    W = mine.Wavelets(IMG, wname, lev, **kw)
    assert W.levels == rlev
    W.forward()
    check(flat(W.coeffs), rc, 255.0, "%s %s fwd" % (mode, wname))
    W.inverse()
    check([W.image], [rimg], 255.0, "%s %s inv" % (mode, wname))


@pytest.mark.parametrize("shape", [(128, 128), (97, 75)])
@pytest.mark.parametrize("do_swt", [0, 1])
@pytest.mark.parametrize("wname", ["haar", "db2", "db4", "sym5", "bior2.2", "rbio3.1", "coif2"])
def test_nonseparable_vs_pdwt(wname, do_swt, shape):
    ref, mine = _ref(), _mine()
    img = synth_image(shape, seed=22)
    R = ref.Wavelets(img, wname, 2, do_separable=0, do_swt=do_swt)
    R.forward()
    rc = flat(R.coeffs)
    R.inverse()
    rimg = np.array(R.image)
    del R
    W = mine.Wavelets(img, wname, 2, do_separable=0, do_swt=do_swt)
    W.forward()
    check(flat(W.coeffs), rc, 255.0, "nonsep %s fwd" % wname)
    W.inverse()
    check([W.image], [rimg], 255.0, "nonsep %s inv" % wname)


@pytest.mark.parametrize("shape", [(255, 253), (65, 200), (999,)])
@pytest.mark.parametrize("wname", ["haar", "db2", "db5", "sym8", "bior3.1"])
def test_odd_sizes_vs_pdwt(wname, shape):
    ref, mine = _ref(), _mine()
    img = synth_image(shape, seed=23)
    # (a 2D array with a single row crashes the reference's wrapper: C++ switches to 1D, pypwt.pyx
    # keeps the 2D coefficient list -- so the 1D case is given as a true 1D array)
    R = ref.Wavelets(img, wname, 3, ndim=img.ndim)
    R.forward()
    rc = flat(R.coeffs)
    R.inverse()
    rimg = np.array(R.image)
    del R
    W = mine.Wavelets(img, wname, 3, ndim=img.ndim)
    W.forward()
    check(flat(W.coeffs), rc, 255.0, "odd %s fwd" % wname)
    W.inverse()
    check([W.image], [rimg], 255.0, "odd %s inv" % wname)


@pytest.mark.parametrize("cfg", [dict(), dict(do_swt=1), dict(ndim=1)])
def test_thresholds_norms_vs_pdwt(cfg):
    ref, mine = _ref(), _mine()
    img = synth_image((96, 160), seed=24, kind="smooth")
    res = {}
    for name, mod in (("ref", ref), ("mine", mine)):
        W = mod.Wavelets(img, "db3", 3, **cfg)
        W.forward()
        out = {"n1": W.norm1()}
        if "ndim" not in cfg:
            out["n2"] = W.norm2sq()      # the reference's 1D norm2sq is wrong (quirk Q3)
        W.soft_threshold(8.0, 1, 1)
        out["soft"] = flat(W.coeffs)
        W.forward()
        W.shrink(0.3, 1)
        out["shrink"] = flat(W.coeffs)
        W.forward()
        W.hard_threshold(8.0, 1, 1)
        out["hard"] = flat(W.coeffs)
        W.inverse()
        out["img"] = np.array(W.image)
        res[name] = out
        del W
    r, m = res["ref"], res["mine"]
    assert abs(m["n1"] - r["n1"]) <= 2e-5 * r["n1"]
    if "n2" in r:
        assert abs(m["n2"] - r["n2"]) <= 2e-5 * r["n2"]
    check(m["soft"], r["soft"], 255.0, "soft")
    check(m["shrink"], r["shrink"], 255.0, "shrink")
    # hard threshold: a coefficient within rounding distance of beta may flip
    nbad = 0
    for g, rr in zip(m["hard"], r["hard"]):
        tol = RTOL * max(255.0, float(np.abs(rr).max()))
        nbad += int((np.abs(g - rr) > tol).sum())
    assert nbad <= 4
    if nbad == 0:
        check([m["img"]], [r["img"]], 255.0, "hard inverse")


def test_cycle_spinning_vs_pdwt():
    """Both libraries draw shifts from the same libc rand() stream of this process, so we only check
    that the image seen after forward() is a circular shift of the input by the reported shift and
    that the reference reconstructs the input as we do."""
    ref, mine = _ref(), _mine()
    img = synth_image((64, 96), seed=25)
    R = ref.Wavelets(img, "db2", 2, do_cycle_spinning=1)
    R.forward()
    shifted = np.array(R.image)
    rA = np.array(R.coeff_only(0))
    R.inverse()
    assert np.abs(R.image - img).max() < 1e-3
    del R
    # find the shift the reference used
    found = None
    for sr in range(64):
        for sc in range(96):
            if shifted[0, 0] == img[(-sr) % 64, (-sc) % 96] and np.array_equal(np.roll(img, (sr, sc), axis=(0, 1)), shifted):
                found = (sr, sc)
                break
        if found:
            break
    assert found is not None
    W = mine.Wavelets(np.roll(img, found, axis=(0, 1)), "db2", 2)
    W.forward()
    check([W.coeff_only(0)], [rA], 255.0, "cycle-spun A")


@pytest.mark.parametrize("mode", [0, 1, 3])
@pytest.mark.parametrize("shape", [(96, 96), (33, 70), (32, 64), (7, 9), (500, 500), (1001, 777), (512, 1024), (3, 130, 258)])
def test_haar_is_bit_identical_to_pdwt(shape, mode):
    """Haar is additions and an exact factor 1/2 (haar.cu:10-58): every kernel family that serves it (fused cascade, register
    kernels, the flat butterfly for the sizes those do not take, the generic tile kernel) must reproduce the reference's bits,
    odd sizes included.  (The generic tile kernel used to associate along the rows first: 3e-5 away on 8-bit data, within the
    tolerance but not exact -- found when the flat kernels, which follow the reference's order, were compared with it.)"""
    ref, mine = _ref(), _mine()
    img = synth_image(shape, seed=33)
    imgs = img if img.ndim == 3 else img[None]
    W = mine.Wavelets(img, "haar", 4)
    W.set_kernel_mode(mode)
    W.forward()
    mc = flat(W.coeffs)
    W.inverse()
    mi = W.image if img.ndim == 3 else W.image[None]
    for k in range(imgs.shape[0]):
        R = ref.Wavelets(imgs[k], "haar", 4)
        assert R.levels == W.levels
        R.forward()
        for b, (g, r) in enumerate(zip(mc, flat(R.coeffs))):
            assert np.array_equal(g[k] if img.ndim == 3 else g, r), "band %d of image %d differs" % (b, k)
        R.inverse()
        assert np.array_equal(mi[k], np.array(R.image)), "reconstruction of image %d differs" % k
        del R
