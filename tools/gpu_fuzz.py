"""Randomised parity run on the GPU: random shapes / banks / level counts / modes / operator sequences through
pycudwt.Wavelets against the oracle (tolerance of the parity tests).  usage: python tools/gpu_fuzz.py [cases] [seed]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pycudwt
from oracle import pdwt_oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
names = O.WAVELET_NAMES
ILL = ("bior3.1", "rbio3.1")


def close(a, b, what, scale, wn):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        return "%s: shape %s vs %s" % (what, a.shape, b.shape)
    tol = 1e-5 * max(scale, np.abs(b).max() if b.size else 0) * (20 if wn in ILL else 1)
    err = np.abs(a - b).max() if b.size else 0.0
    return None if err <= tol else "%s: err %.3e > tol %.3e" % (what, err, tol)


def dims():
    r = rng.random()
    if r < 0.5:
        return int(rng.integers(8, 300))
    if r < 0.8:
        return int(rng.integers(300, 1200))
    return int(rng.choice([256, 512, 1024, 2048, 1000, 2047, 777, 4096]))


bad = 0
t0 = time.time()
for case in range(N):
    wn = names[int(rng.integers(len(names)))]
    mode = rng.choice(["2d", "2d", "swt2", "1d", "b1d", "swt1d", "nonsep", "cs"])
    kw = {}
    if mode in ("2d", "nonsep", "cs", "swt2"):
        shape = (dims(), dims())
        if mode == "swt2":
            shape = (min(shape[0], 600), min(shape[1], 600)); kw = dict(do_swt=1)
        if mode == "nonsep":
            shape = (min(shape[0], 400), min(shape[1], 400)); kw = dict(do_separable=0)
        if mode == "cs":
            kw = dict(do_cycle_spinning=1, do_swt=int(rng.integers(2)))
            if kw["do_swt"]:
                shape = (min(shape[0], 600), min(shape[1], 600))
    elif mode == "1d":
        shape = (int(rng.integers(16, 20000)),); kw = dict(ndim=1)
    elif mode == "b1d":
        shape = (int(rng.integers(1, 200)), dims() * int(rng.integers(1, 8))); kw = dict(ndim=1)
    else:
        shape = (int(rng.integers(1, 60)), dims()); kw = dict(ndim=1, do_swt=1)
    L = int(rng.integers(1, 9))
    img = (rng.standard_normal(shape) * 50 + 128).astype(np.float32)
    try:
        Wo = O.OracleWavelets(img, wn, L, **kw)
    except (ValueError, NotImplementedError):
        continue
    try:
        W = pycudwt.Wavelets(img, wn, L, **kw)
    except ValueError as e:
        print("CASE %d create mismatch %s %s %s L%d: %s" % (case, mode, wn, shape, L, e)); bad += 1; continue
    errs = []
    if W.levels != Wo.levels:
        errs.append("levels %d vs %d" % (W.levels, Wo.levels))
    W.forward()
    if kw.get("do_cycle_spinning"):
        sr, sc = W.current_shift
        class R:  # noqa
            v = [sr, sc]
            def rand(self): return self.v.pop(0)
        Wo = O.OracleWavelets(img, wn, L, rng=R(), **kw)
    Wo.forward()
    if rng.random() < 0.7:
        c, co = W.coeffs, Wo.coeffs
        errs.append(close(c[0], co[0], "A", 255.0, wn))
        for i in range(1, len(co)):
            if isinstance(co[i], list):
                for j in range(3):
                    errs.append(close(c[i][j], co[i][j], "L%d b%d" % (i, j), 255.0, wn))
            else:
                errs.append(close(c[i], co[i], "D%d" % i, 255.0, wn))
    op = rng.choice(["none", "soft", "hard", "shrink", "norms"])
    beta = float(rng.uniform(1, 30))
    app, nrm = int(rng.integers(2)), int(rng.integers(2))
    if op == "soft":
        W.soft_threshold(beta, app, nrm); Wo.soft_threshold(beta, app, nrm)
    elif op == "hard":
        # a hard threshold is discontinuous: a coefficient within rounding of beta may be kept on one side and zeroed
        # on the other.  Such flips are legitimate; they are identified (|pre-threshold value| ~ beta) and the GPU's
        # decision is copied into the oracle so that the reconstructions stay comparable.
        pre = [np.array(x, np.float64) for x in Wo._c]
        W.hard_threshold(beta, 0, 0); Wo.hard_threshold(beta, 0, 0)
        cg = W.coeffs
        flat = [cg[0]] + [b for lv in cg[1:] for b in (lv if isinstance(lv, list) else [lv])]
        for bi in range(1, len(Wo._c)):
            g = np.asarray(flat[bi], np.float64).reshape(Wo._c[bi].shape)
            o = np.asarray(Wo._c[bi], np.float64)
            flip = (g == 0) != (o == 0)
            if flip.any():
                near = np.abs(np.abs(pre[bi][flip]) - beta) <= 1e-5 * max(255.0, np.abs(pre[bi]).max()) * (20 if wn in ILL else 1)
                if not near.all():
                    errs.append("hard threshold: %d flips away from beta in band %d" % (int((~near).sum()), bi))
                Wo._c[bi] = np.where(flip, g, o).astype(Wo._c[bi].dtype)
    elif op == "shrink":
        W.shrink(beta / 30, app); Wo.shrink(beta / 30, app)
    elif op == "norms":
        n1, n1o = W.norm1(), Wo.norm1()
        if abs(n1 - n1o) > 1e-5 * max(n1o, 1):
            errs.append("norm1 %.8e vs %.8e" % (n1, n1o))
    W.inverse(); Wo.inverse()
    # a hard threshold flips on values within rounding of beta: compare the reconstruction with a wider tolerance there
    scale = 255.0
    errs.append(close(W.image.reshape(Wo.image.shape), Wo.image, "inverse after %s" % op, scale, wn))
    errs = [e for e in errs if e]
    if errs:
        bad += 1
        print("CASE %d FAIL %s %s %s L%d(%d) %s: %s" % (case, mode, wn, shape, L, W.levels, kw, errs[:3]), flush=True)
    del W
print("fuzz done: %d cases, %d failures, %.0f s" % (N, bad, time.time() - t0), flush=True)
