import sys, os; sys.path.insert(0, ".")
import numpy as np, pycudwt
img = (np.random.default_rng(0).standard_normal((64, 2048, 2048)) * 50 + 128).astype(np.float32)
for wn in ("sym8", "db7"):
    W = pycudwt.Wavelets(img, wn, 3)
    def step(): W.forward(); W.soft_threshold(5.0, 0, 1); W.inverse()
    for _ in range(5): step()
    W.sync(); ts = []
    for r in range(5):
        W.timer_start()
        for _ in range(10): step()
        ts.append(W.timer_stop() / 10)
    print("OCC3=%s %s 64x2048^2 fwd+soft+inv %.4f ms" % (os.environ.get("PWT_STRIP_THR_OCC3", "1"), wn, sorted(ts)[2]), flush=True)
