"""Scratch: per-launch durations (events) of a deep transform, to see what the small levels cost."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
for wn, L in (("haar", 7), ("db2", 7), ("db4", 7), ("sym8", 6)):
    W = pycudwt.Wavelets(img, wn, L)
    for _ in range(3): W.forward(); W.inverse()
    W.profile_enable(1); W.forward(); W.inverse(); prof = W.profile_read(); W.profile_enable(0)
    W.timer_start()
    for _ in range(20): W.forward(); W.inverse()
    print(wn, L, "fwd+inv %.4f ms" % (W.timer_stop() / 20), [(t, round(m * 1000, 1)) for t, m in prof], flush=True)
