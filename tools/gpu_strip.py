"""Strip kernels (kernel mode 4) against the generic kernels (mode 1) on many shapes / filters, then timings.
usage: python tools/gpu_strip.py [check] [time] [wname ...]"""
import os, sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
SMALL = bool(os.environ.get("SMALL"))     # compute-sanitizer runs: small shapes, fewer filters


def _small(shapes, wns):
    if not SMALL:
        return shapes, wns
    return [s for s in shapes if int(np.prod(s)) <= 400000], wns[::3] + wns[-1:]


def check():
    rng = np.random.default_rng(3)
    bad = 0
    shapes = [(512, 768), (511, 509), (64, 1000), (1001, 777), (40, 36), (300, 2048), (2048, 300), (3, 260, 516), (17, 23), (2048, 2048)]
    wns = ["db3", "db4", "sym5", "db6", "db7", "sym8", "coif3", "db10", "db12", "coif5", "db16", "db20", "bior6.8", "rbio2.8", "bior3.9"]
    shapes, wns = _small(shapes, wns)
    for shp in shapes:
        img = (rng.standard_normal(shp) * 50 + 128).astype(np.float32)
        for wn in wns:
            for L in (1, 3):
                try:
                    S = pycudwt.Wavelets(img, wn, L); G = pycudwt.Wavelets(img, wn, L)
                except ValueError:
                    continue
                S.set_kernel_mode(4); G.set_kernel_mode(1)
                S.forward(); G.forward()
                cs, cg = S.coeffs, G.coeffs
                scale = 1e-5 * max(np.abs(img).max(), 1.0)
                errs = [np.abs(cs[0] - cg[0]).max() / max(scale, 1e-5 * np.abs(cg[0]).max())]
                for i in range(1, len(cs)):
                    for j in range(3):
                        errs.append(np.abs(cs[i][j] - cg[i][j]).max() / max(scale, 1e-5 * np.abs(cg[i][j]).max()))
                S.inverse(); G.inverse()
                ei = np.abs(S.image - G.image).max() / scale
                # strip inverse on the generic coefficients' own forward: also check reconstruction
                er = np.abs(S.image - img).max() / scale
                tol_r = 40 if wn in ("bior3.9",) else 8
                ok = max(errs) < 1.0 and ei < 1.0 and er < tol_r
                if not ok:
                    bad += 1
                    print("FAIL", shp, wn, L, S.levels, "fwd %.3g inv %.3g rec %.3g" % (max(errs), ei, er), flush=True)
    print("check done, failures:", bad)
    return bad

def check1d():
    rng = np.random.default_rng(5)
    bad = 0
    shapes = [(512, 768), (511, 509), (64, 1000), (1001, 777), (40, 36), (300, 2048), (7, 8192), (1000003,), (4099,), (33, 130)]
    wns = ["db2", "db3", "db4", "sym5", "db6", "sym8", "db10", "coif5", "db20", "bior6.8", "bior3.9", "rbio2.8"]
    shapes, wns = _small(shapes, wns)
    for shp in shapes:
        img = (rng.standard_normal(shp) * 50 + 128).astype(np.float32)
        for wn in wns:
            for L in (1, 4):
                try:
                    S = pycudwt.Wavelets(img, wn, L, ndim=1); G = pycudwt.Wavelets(img, wn, L, ndim=1)
                except ValueError:
                    continue
                S.set_kernel_mode(4); G.set_kernel_mode(1)
                S.forward(); G.forward()
                cs, cg = S.coeffs, G.coeffs
                scale = 1e-5 * max(np.abs(img).max(), 1.0)
                errs = [np.abs(np.asarray(a) - np.asarray(b)).max() / max(scale, 1e-5 * np.abs(b).max()) for a, b in zip(cs, cg)]
                shp_ok = all(np.asarray(a).shape == np.asarray(b).shape for a, b in zip(cs, cg))
                S.inverse(); G.inverse()
                ei = np.abs(S.image - G.image).max() / scale
                if not (shp_ok and max(errs) < 1.0 and ei < 1.0):
                    bad += 1
                    print("FAIL1D", shp, wn, L, S.levels, "fwd %.3g inv %.3g" % (max(errs), ei), shp_ok, flush=True)
    # stationary transform, 1D: auto (staged-row kernels) against generic
    for shp in [sh for sh in [(512, 768), (64, 1000), (300, 2048), (7, 8192), (4100,), (33, 132), (5, 20)] if not SMALL or int(np.prod(sh)) <= 100000]:
        img = (rng.standard_normal(shp) * 50 + 128).astype(np.float32)
        for wn in ["haar", "db2", "db4", "sym8", "db10", "coif5", "db20", "bior3.9"]:
            for L in (1, 2, 3, 5):
                try:
                    S = pycudwt.Wavelets(img, wn, L, ndim=1, do_swt=1); G = pycudwt.Wavelets(img, wn, L, ndim=1, do_swt=1)
                except ValueError:
                    continue
                G.set_kernel_mode(1)
                S.forward(); G.forward()
                cs, cg = S.coeffs, G.coeffs
                scale = 1e-5 * max(np.abs(img).max(), 1.0)
                errs = [np.abs(np.asarray(a) - np.asarray(b)).max() / max(scale, 1e-5 * np.abs(b).max()) for a, b in zip(cs, cg)]
                S.inverse(); G.inverse()
                ei = np.abs(S.image - G.image).max() / scale
                if not (max(errs) < 1.0 and ei < 1.0):
                    bad += 1
                    print("FAILSWT1D", shp, wn, L, S.levels, "fwd %.3g inv %.3g" % (max(errs), ei), flush=True)
    print("check1d done, failures:", bad)
    return bad

def checkswt():
    rng = np.random.default_rng(7)
    bad = 0
    for shp in [sh for sh in [(512, 768), (300, 520), (64, 1000), (1024, 1024), (2, 260, 516), (129, 260), (2048, 2048)] if not SMALL or int(np.prod(sh)) <= 300000]:
        img = (rng.standard_normal(shp) * 50 + 128).astype(np.float32)
        for wn in ["haar", "db2", "db3", "db4", "sym5", "db6", "sym7", "sym8"]:
            for L in (1, 2, 4):
                try:
                    S = pycudwt.Wavelets(img, wn, L, do_swt=1); G = pycudwt.Wavelets(img, wn, L, do_swt=1)
                except ValueError:
                    continue
                G.set_kernel_mode(1)
                S.forward(); G.forward()
                cs, cg = S.coeffs, G.coeffs
                scale = 1e-5 * max(np.abs(img).max(), 1.0)
                errs = [np.abs(cs[0] - cg[0]).max() / max(scale, 1e-5 * np.abs(cg[0]).max())]
                for i in range(1, len(cs)):
                    for j in range(3):
                        errs.append(np.abs(cs[i][j] - cg[i][j]).max() / max(scale, 1e-5 * np.abs(cg[i][j]).max()))
                S.inverse(); G.inverse()
                ei = np.abs(S.image - G.image).max() / scale
                if not (max(errs) < 1.0 and ei < 1.0):
                    bad += 1
                    print("FAILSWT2D", shp, wn, L, S.levels, "fwd %.3g inv %.3g" % (max(errs), ei), flush=True)
    print("checkswt done, failures:", bad)
    return bad

def timeit(wn, shape=(8192, 8192), L=1, mode=0, n=20):
    img = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    W = pycudwt.Wavelets(img, wn, L)
    W.set_kernel_mode(mode)
    out = []
    for what in ("f", "fi"):
        for _ in range(3):
            W.forward(); W.inverse()
        W.timer_start()
        for _ in range(n):
            W.forward()
            if what == "fi": W.inverse()
        out.append(W.timer_stop() / n)
    return out[0], out[1] - out[0]

if __name__ == "__main__":
    args = sys.argv[1:]
    if "check" in args:
        if check(): sys.exit(1)
    if "checkswt" in args:
        if checkswt(): sys.exit(1)
    if "check1d" in args:
        if check1d(): sys.exit(1)
    if "time5" in args:
        wns = [a for a in args if a not in ("check", "check1d", "checkswt", "time", "time5", "nostack")]
        for wn in wns:
            a = timeit(wn, L=5, mode=0); b = timeit(wn, L=5, mode=4)
            print("%-6s 8192^2 L5  auto fwd %.4f inv %.4f sum %.4f | strip fwd %.4f inv %.4f sum %.4f ms" % (wn, a[0], a[1], a[0] + a[1], b[0], b[1], b[0] + b[1]), flush=True)
    if "time" in args:
        wns = [a for a in args if a not in ("check", "check1d", "checkswt", "time", "time5", "nostack")] or ["db3", "db4", "db6", "sym8", "db10", "db12", "coif5", "db20"]
        for wn in wns:
            a = timeit(wn, mode=0); b = timeit(wn, mode=4)
            print("%-6s 8192^2 L1  auto fwd %.4f inv %.4f | strip fwd %.4f inv %.4f ms" % (wn, a[0], a[1], b[0], b[1]), flush=True)
        for wn in ("sym8",):
            if "nostack" in args: break
            for mode in (0, 4):
                f, i = timeit(wn, shape=(64, 2048, 2048), L=3, mode=mode, n=5)
                print("%-6s 64x2048^2 L3 mode %d fwd %.4f inv %.4f ms" % (wn, mode, f, i), flush=True)
