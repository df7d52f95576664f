"""Double-precision plans: timings (CUDA events) against the 32 B/px fp64 roofline."""
import sys; sys.path.insert(0, ".")
import numpy as np, pypwt_b200
for wn, N, L, kw in (("haar", 8192, 3, {}), ("db2", 8192, 3, {}), ("sym8", 8192, 3, {}), ("db20", 8192, 3, {}), ("db2", 4096, 3, {}),
                     ("db4", 4096, 4, dict(do_swt=1)), ("db2", 8192, 3, dict(ndim=1))):
    img = np.random.default_rng(0).standard_normal((N, N))
    W = pypwt_b200.Wavelets64(img, wn, L, **kw)
    for _ in range(3): W.forward(); W.inverse()
    W.sync()
    ts = []
    for rep in range(3):
        W.timer_start()
        for _ in range(10): W.forward(); W.inverse()
        ts.append(W.timer_stop() / 10)
    t = sorted(ts)[1]
    bpp = 32 if not kw.get("do_swt") else 2 * (3 * L + 2) * 8
    print("f64 %-5s %d^2 L%d %s fwd+inv %.4f ms  %.1f Gpx/s  frac of %d B/px roofline %.3f" % (wn, N, L, kw, t, N * N / t / 1e6, bpp, bpp * N * N / t / 1e6 / 6549.4), flush=True)
