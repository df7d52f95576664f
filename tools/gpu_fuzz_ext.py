"""Randomised parity run for the double-precision and volumetric plans (see tools/gpu_fuzz.py)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pypwt_b200
from oracle import pdwt_oracle as O, dwt3_oracle as D

N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
names = O.WAVELET_NAMES
bad = 0
t0 = time.time()


def close(a, b, what, rtol, scale=255.0, k=1):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        return "%s: shape %s vs %s" % (what, a.shape, b.shape)
    tol = rtol * max(scale, np.abs(b).max()) * k
    err = np.abs(a - b).max()
    return None if err <= tol else "%s: err %.3e > %.3e" % (what, err, tol)


for case in range(N):
    wn = names[int(rng.integers(len(names)))]
    k = 50 if wn in ("bior3.1", "rbio3.1") else 1
    # ---- double precision
    mode = rng.choice(["2d", "swt2", "b1d", "swt1d", "nonsep", "stack"])
    kw = {}
    shape = (int(rng.integers(8, 500)), int(rng.integers(8, 700)))
    if mode == "swt2": kw = dict(do_swt=1); shape = (min(shape[0], 300), min(shape[1], 300))
    if mode == "b1d": kw = dict(ndim=1)
    if mode == "swt1d": kw = dict(ndim=1, do_swt=1)
    if mode == "nonsep": kw = dict(do_separable=0)
    L = int(rng.integers(1, 7))
    img = rng.standard_normal(shape) * 50 + 128
    errs = []
    try:
        Wo = O.OracleWavelets(img, wn, L, double_build=True, **kw)
        if mode == "stack":
            st = np.stack([img, img * 0.5 + 1])
            W = pypwt_b200.Wavelets64(st, wn, L)
            W.forward(); Wo.forward()
            errs.append(close(W.coeffs[0][0], Wo.coeffs[0], "f64 stack A", 1e-12, k=k))
            errs.append(close(W.coeffs[-1][2][0], Wo.coeffs[-1][2], "f64 stack D", 1e-12, k=k))
            W.inverse(); Wo.inverse()
            errs.append(close(W.image[0], Wo.image, "f64 stack inverse", 1e-12, k=k))
        else:
            W = pypwt_b200.Wavelets64(img, wn, L, **kw)
            W.forward(); Wo.forward()
            c, co = W.coeffs, Wo.coeffs
            errs.append(close(np.asarray(c[0]).reshape(co[0].shape), co[0], "f64 A", 1e-12, k=k))
            for i in range(1, len(co)):
                if isinstance(co[i], list):
                    for j in range(3): errs.append(close(c[i][j], co[i][j], "f64 L%d b%d" % (i, j), 1e-12, k=k))
                else:
                    errs.append(close(np.asarray(c[i]).reshape(co[i].shape), co[i], "f64 D%d" % i, 1e-12, k=k))
            beta = float(rng.uniform(1, 20))
            W.soft_threshold(beta, int(rng.integers(2)), 0); Wo.soft_threshold(beta, 0, 0) if False else None
            W.set_image(img); W.forward(); W.inverse(); Wo.inverse()
            errs.append(close(W.image.reshape(Wo.image.shape), Wo.image, "f64 inverse", 1e-12, k=k))
    except ValueError:
        pass
    # ---- volumes
    vshape = (int(rng.integers(8, 90)), int(rng.integers(8, 120)), int(rng.integers(8, 160)))
    vol = (rng.standard_normal(vshape) * 50 + 128).astype(np.float32)
    Lv = int(rng.integers(1, 4))
    try:
        Vo = D.OracleWavelets3D(vol, wn, Lv)
        V = pypwt_b200.Wavelets3D(vol, wn, Lv)
        V.forward(); Vo.forward()
        c, co = V.coeffs, Vo.coeffs
        errs.append(close(c[0], co[0], "3D aaa", 1e-5, k=20 if k > 1 else 1))
        for l in range(1, len(co)):
            for key in D.KEYS:
                errs.append(close(c[l][key], co[l][key], "3D L%d %s" % (l, key), 1e-5, k=20 if k > 1 else 1))
        V.inverse(); Vo.inverse()
        errs.append(close(V.image, Vo.image, "3D inverse", 1e-5, k=20 if k > 1 else 1))
    except ValueError:
        pass
    errs = [e for e in errs if e]
    if errs:
        bad += 1
        print("CASE %d FAIL %s %s %s L%d | vol %s L%d: %s" % (case, mode, wn, shape, L, vshape, Lv, errs[:3]), flush=True)
print("fuzz (f64 + 3D) done: %d cases, %d failures, %.0f s" % (N, bad, time.time() - t0), flush=True)
