"""Scratch: C3 shard (64 x 2048^2 sym8 L3): forward + norms, fused vs plain reduction."""
import os, sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
img = np.random.default_rng(0).standard_normal((64, 2048, 2048)).astype(np.float32)
W = pycudwt.Wavelets(img, "sym8", 3)
def t(fn, n=10):
    for _ in range(3): fn()
    W.timer_start()
    for _ in range(n): fn()
    return W.timer_stop() / n
print("fwd only %.4f ms" % t(lambda: W.forward()))
def fn():
    W.forward(); W.norm1()
print("fwd+norm1 %.4f ms (PWT_NO_FUSED_NORMS=%s)" % (t(fn), os.environ.get("PWT_NO_FUSED_NORMS")))
W.forward(); a = W.norms()
c = W.coeffs
flat = np.concatenate([c[0].ravel().astype(np.float64)] + [b.ravel().astype(np.float64) for lvl in c[1:] for b in lvl])
print("rel err", abs(a[0] - np.abs(flat).sum()) / a[0], abs(a[1] - (flat * flat).sum()) / a[1])
