"""Batched 1D transforms, every level in one launch (kernels_row1d.cu): auto mode against the generic kernels
(mode 1) on many shapes / filters / level counts, then timings.   usage: python tools/gpu_row1d.py [check] [time]"""
import os, sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
SMALL = bool(os.environ.get("SMALL"))


def check(swt=0):
    rng = np.random.default_rng(11)
    bad = n = 0
    shapes = [(512, 768), (511, 509), (64, 1000), (1001, 777), (40, 36), (300, 2048), (7, 8192), (3, 16384), (2, 40000), (100003,), (4099,), (33, 130), (5, 24), (9, 6)]
    wns = ["haar", "db2", "db3", "db4", "sym5", "db6", "sym8", "db10", "coif5", "db20", "bior6.8", "bior3.9", "rbio2.8", "bior1.3"]
    if SMALL:
        shapes = [s for s in shapes if int(np.prod(s)) <= 100000]; wns = wns[::3] + wns[-1:]
    for shp in shapes:
        img = (rng.standard_normal(shp) * 50 + 128).astype(np.float32)
        for wn in wns:
            for L in (1, 2, 3, 4, 7, 20):
                try:
                    S = pycudwt.Wavelets(img, wn, L, ndim=1, do_swt=swt); G = pycudwt.Wavelets(img, wn, L, ndim=1, do_swt=swt)
                except ValueError:
                    continue
                G.set_kernel_mode(1)
                l0 = S.launch_count
                S.forward(); G.forward()
                lf = S.launch_count - l0
                cs, cg = S.coeffs, G.coeffs
                scale = 1e-5 * max(np.abs(img).max(), 1.0)
                errs = [np.abs(np.asarray(a) - np.asarray(b)).max() / max(scale, 1e-5 * np.abs(b).max()) for a, b in zip(cs, cg)]
                shp_ok = all(np.asarray(a).shape == np.asarray(b).shape for a, b in zip(cs, cg))
                S.soft_threshold(5.0); G.soft_threshold(5.0)
                S.inverse(); G.inverse()
                ei = np.abs(S.image - G.image).max() / scale
                # inverse of the other path's coefficients too (exercises the fused inverse on exact inputs)
                n += 1
                if not (shp_ok and max(errs) < 1.0 and ei < 1.0):
                    bad += 1
                    print("FAIL1D swt=%d" % swt, shp, wn, L, S.levels, "fwd %.3g inv %.3g launches %d" % (max(errs), ei, lf), shp_ok, flush=True)
    print("check swt=%d done: %d cases, failures: %d" % (swt, n, bad), flush=True)
    return bad


def timeit(wn, shape, L, swt=0, mode=0, n=20):
    img = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    W = pycudwt.Wavelets(img, wn, L, ndim=1, do_swt=swt)
    W.set_kernel_mode(mode)
    for _ in range(5):
        W.forward(); W.inverse()
    out = []
    for what in ("f", "fi"):
        ts = []
        for rep in range(5):
            W.timer_start()
            for _ in range(n):
                W.forward()
                if what == "fi": W.inverse()
            ts.append(W.timer_stop() / n)
        out.append(sorted(ts)[2])
    return out[0], out[1]


if __name__ == "__main__":
    args = sys.argv[1:]
    rc = 0
    if "check" in args:
        rc |= check(0)
    if "checkswt" in args:
        rc |= check(1)
    if "time" in args:
        for swt in (0, 1):
            for shape, L in (((8192, 8192), 3), ((8192, 8192), 6), ((16384, 4096), 3), ((2048, 16384), 3), ((65536, 512), 3)):
                for wn in ("haar", "db2", "db4", "sym8", "db20"):
                    f, fi = timeit(wn, shape, L, swt)
                    px = shape[0] * shape[1]
                    bpp = 16 if not swt else 8 * (L + 2)
                    print("1D swt=%d %-5s %5dx%-5d L%d fwd %.4f fwd+inv %.4f ms  %.1f Gpx/s  frac %.3f" % (swt, wn, shape[0], shape[1], L, f, fi, px / fi / 1e6, bpp * px / fi / 1e6 / 6549.4), flush=True)
    sys.exit(1 if rc else 0)
