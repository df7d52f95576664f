"""Scratch: config C4 (SWT db4 4 levels, 8192^2): per-kernel times and the denoising loop."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
wn = sys.argv[1] if len(sys.argv) > 1 else "db4"
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
W = pycudwt.Wavelets(img, wn, 4, do_swt=1)
for _ in range(2):
    W.forward(); W.hard_threshold(0.5); W.inverse()
W.profile_enable(1)
W.forward(); W.hard_threshold(0.5); W.inverse()
print("profile:", [(t, round(ms, 4)) for t, ms in W.profile_read()])
W.profile_enable(0)
for what in ("fwd", "fwd+inv", "fwd+hard+inv"):
    def step():
        W.forward()
        if what == "fwd+hard+inv": W.hard_threshold(0.5)
        if what != "fwd": W.inverse()
    for _ in range(3): step()
    W.timer_start()
    for _ in range(20): step()
    print(f"swt {wn} L4 {what}: {W.timer_stop()/20:.4f} ms", flush=True)
