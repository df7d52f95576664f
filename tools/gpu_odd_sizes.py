"""Throughput on sizes that are not multiples of 4 / 8 against the aligned neighbours (forward + inverse, CUDA events)."""
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
for shape in ((8192, 8192), (8191, 8191), (8190, 8190), (8188, 8188), (8192, 8190), (4096, 4096), (4095, 4097), (4094, 4094), (2048, 2048), (2047, 2049), (1001, 777)):
    img = rng.standard_normal(shape).astype(np.float32)
    for wn, kw in (("haar", {}), ("db2", {}), ("sym8", {}), ("db2", dict(ndim=1)), ("db4", dict(do_swt=1))):
        if kw.get("do_swt") and shape[0] > 4100:
            continue
        W = pycudwt.Wavelets(img, wn, 3, **kw)
        for _ in range(3): W.forward(); W.inverse()
        W.sync(); ts = []
        for r in range(3):
            W.timer_start()
            for _ in range(10): W.forward(); W.inverse()
            ts.append(W.timer_stop() / 10)
        t = sorted(ts)[1]
        l0 = W.launch_count; W.forward(); W.inverse(); nl = W.launch_count - l0
        print("%-12s %-5s %-22s %.4f ms  %.1f Gpx/s  launches %d" % ("%dx%d" % shape, wn, kw, t, img.size / t / 1e6, nl), flush=True)
        del W
