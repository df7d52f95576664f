"""Scratch: small and medium images, fwd+inv, several filters / depths (A/B of launch-related changes)."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
rng = np.random.default_rng(0)
out = []
for N in (512, 1024, 2048, 4096):
    img = rng.standard_normal((N, N)).astype(np.float32)
    for wn in ("haar", "db2", "db4", "sym8"):
        for L in (3, 5):
            W = pycudwt.Wavelets(img, wn, L)
            for _ in range(10): W.forward(); W.inverse()
            best = 1e9
            for rep in range(3):
                W.timer_start()
                for _ in range(200): W.forward(); W.inverse()
                best = min(best, W.timer_stop() / 200)
            out.append("%d %s L%d %.4f" % (N, wn, L, best))
print("PDL off" if os.environ.get("PWT_NO_PDL") else "PDL on", " | ".join(out), flush=True)
