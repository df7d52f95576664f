"""1- and 2-level transforms (no 3-level cascade): per-level kernels against the 16 B/px roofline."""
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
for shape in ((8192, 8192), (4096, 4096), (2048, 2048)):
    img = rng.standard_normal(shape).astype(np.float32)
    for wn in ("haar", "db2", "db3", "sym8"):
        for L in (1, 2, 3, 4):
            W = pycudwt.Wavelets(img, wn, L)
            for _ in range(3): W.forward(); W.inverse()
            W.sync(); ts = []
            for r in range(3):
                W.timer_start()
                for _ in range(10): W.forward(); W.inverse()
                ts.append(W.timer_stop() / 10)
            t = sorted(ts)[1]
            l0 = W.launch_count; W.forward(); W.inverse()
            print("%dx%d %-5s L%d %.4f ms frac %.3f launches %d" % (shape[0], shape[1], wn, L, t, 16 * img.size / t / 1e6 / 6549.4, W.launch_count - l0), flush=True)
