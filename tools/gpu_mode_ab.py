"""Per-level kernels for short filters on sizes the fused cascade / register kernels do not take: default choice against the strip kernels (mode 4)."""
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
for shape in ((8188, 8188), (8184, 8184), (4092, 4092), (2000, 3000), (1500, 1500), (1000, 1000), (8192, 8200), (520, 520)):
    img = rng.standard_normal(shape).astype(np.float32)
    for wn in ("haar", "db2", "db3"):
        res = []
        for mode in (0, 4):
            W = pycudwt.Wavelets(img, wn, 3)
            W.set_kernel_mode(mode)
            for _ in range(3): W.forward(); W.inverse()
            W.sync(); ts = []
            for r in range(3):
                W.timer_start()
                for _ in range(10): W.forward(); W.inverse()
                ts.append(W.timer_stop() / 10)
            l0 = W.launch_count; W.forward(); W.inverse()
            res.append("%.4f ms (%d launches)" % (sorted(ts)[1], W.launch_count - l0))
        print("%-12s %-5s default %s   strip %s" % ("%dx%d" % shape, wn, res[0], res[1]), flush=True)
