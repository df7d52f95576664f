"""Slice groups of a batched per-level 2D DWT (PWT_GROUP_MB): digest of the results + timings, one process per setting.
usage: PWT_GROUP_MB=<mb> python tools/gpu_group.py [wname] [slices] [N] [levels]"""
import os, sys, hashlib, numpy as np
sys.path.insert(0, ".")
import pycudwt
wn = sys.argv[1] if len(sys.argv) > 1 else "sym8"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
L = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rng = np.random.default_rng(0)
img = (rng.standard_normal((S, N, N)) * 50 + 128).astype(np.float32)
W = pycudwt.Wavelets(img, wn, L)
W.forward()
n1a = W.norm1()                       # full reduction
W.forward()
n1b, n2b = W.norm1(), W.norm2sq()     # fused partial sums
h = hashlib.sha1()
c = W.coeffs
h.update(c[0].tobytes())
for lv in c[1:]:
    for b in lv:
        h.update(b.tobytes())
W.soft_threshold(10.0)
W.inverse()
h2 = hashlib.sha1(W.image.tobytes()).hexdigest()[:12]
W.set_image(img)
W.forward(); W.inverse()
rec = float(np.abs(W.image - img).max())
for _ in range(20): W.forward(); W.inverse()
W.sync()
ts = []
for rep in range(5):
    W.timer_start()
    for _ in range(20): W.forward(); W.inverse()
    ts.append(W.timer_stop() / 20)
tf = []
for rep in range(3):
    W.timer_start()
    for _ in range(20): W.forward()
    tf.append(W.timer_stop() / 20)
px = S * N * N
t = sorted(ts)[2]
print(f"{wn} {S}x{N}^2 L{W.levels} group_mb={os.environ.get('PWT_GROUP_MB','dflt')} coeffs {h.hexdigest()[:12]} thr+inv {h2} rec {rec:.2e} "
      f"n1 {n1a:.6e} {n1b:.6e} n2 {n2b:.6e} | fwd {sorted(tf)[1]:.4f} fwd+inv {t:.4f} ms  {px / t / 1e6:.1f} Gpx/s  frac16 {16 * px / t / 1e6 / 6549.4:.3f}", flush=True)
