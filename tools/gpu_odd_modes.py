import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
for shape in ((8191, 8191), (8190, 8190), (8188, 8188), (4095, 4097), (2047, 2049), (1001, 777)):
    img = rng.standard_normal(shape).astype(np.float32)
    for wn in ("db2", "db3", "sym8"):
        for mode in (0, 4):
            W = pycudwt.Wavelets(img, wn, 3)
            W.set_kernel_mode(mode)
            for _ in range(3): W.forward(); W.inverse()
            W.sync(); ts = []
            for r in range(3):
                W.timer_start()
                for _ in range(10): W.forward(); W.inverse()
                ts.append(W.timer_stop() / 10)
            t = sorted(ts)[1]
            print("%-12s %-5s mode %d %.4f ms  %.1f Gpx/s" % ("%dx%d" % shape, wn, mode, t, img.size / t / 1e6), flush=True)
