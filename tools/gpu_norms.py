"""Scratch timing: forward + norms (C3-like flow) and forward alone, 8192^2 db2 L3."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
from pypwt_b200 import pycudwt as X
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
for wname in ("db2", "haar"):
    for mode in (0, 3):
        W = pycudwt.Wavelets(img, wname, 3)
        W.set_kernel_mode(mode)
        for what in ("fwd", "fwd+norms", "fwd+inv"):
            def step():
                W.forward()
                if what == "fwd+norms":
                    W.norms_async() if hasattr(W, "norms_async") else W.norms()
                if what == "fwd+inv":
                    W.inverse()
            for _ in range(5): step()
            W.timer_start()
            for _ in range(200): step()
            ms = W.timer_stop() / 200
            print(f"{wname} mode{mode} {what}: {ms:.4f} ms", flush=True)
