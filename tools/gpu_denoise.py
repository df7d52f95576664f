"""Scratch: forward + soft threshold + inverse at 8192^2, 5 levels: deferred (default) vs PWT_NO_DEFER=1."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
img = (np.random.default_rng(0).standard_normal((8192, 8192)) * 50 + 128).astype(np.float32)
for wn in sys.argv[1:] or ("db2", "db4", "sym8"):
    W = pycudwt.Wavelets(img, wn, 5)
    def step():
        W.forward(); W.soft_threshold(10.0); W.inverse()
    for _ in range(3): step()
    W.timer_start()
    for _ in range(20): step()
    print("%s L5 fwd+soft+inv %.4f ms (PWT_NO_DEFER=%s)" % (wn, W.timer_stop() / 20, os.environ.get("PWT_NO_DEFER")), flush=True)
