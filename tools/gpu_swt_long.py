import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
for shape in ((4096, 4096), (4095, 4097)):
    img = rng.standard_normal(shape).astype(np.float32)
    for wn in ("db4", "sym8", "db10", "db20"):
        try:
            W = pycudwt.Wavelets(img, wn, 3, do_swt=1)
        except ValueError as e:
            print(wn, e); continue
        for _ in range(2): W.forward(); W.inverse()
        W.sync(); ts = []
        for r in range(3):
            W.timer_start()
            for _ in range(5): W.forward(); W.inverse()
            ts.append(W.timer_stop() / 5)
        t = sorted(ts)[1]
        l0 = W.launch_count; W.forward(); W.inverse(); nl = W.launch_count - l0
        bpp = 2 * (3 * 3 + 2) * 4
        print("swt2 %-5s %s L3 %.4f ms frac %.3f launches %d" % (wn, shape, t, bpp * img.size / t / 1e6 / 6549.4, nl), flush=True)
