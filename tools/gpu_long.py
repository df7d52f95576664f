"""Scratch driver for ncu: one forward + inverse at 8192^2 with a long filter (default db20, 5 levels)."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
wn = sys.argv[1] if len(sys.argv) > 1 else "db20"
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
W = pycudwt.Wavelets(img, wn, 5)
for _ in range(3):
    W.forward(); W.inverse()
W.timer_start()
for _ in range(20):
    W.forward(); W.inverse()
print(wn, "fwd+inv ms", W.timer_stop() / 20)
