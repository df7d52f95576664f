"""A/B of the fused forward variants: forward alone, best and median of several event-timed batches."""
import sys, os, numpy as np
sys.path.insert(0, ".")
import pycudwt
wn = sys.argv[1] if len(sys.argv) > 1 else "db2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(0)
img = rng.standard_normal((N, N)).astype(np.float32)
W = pycudwt.Wavelets(img, wn, L)
for _ in range(300): W.forward()
W.sync()
ts = []
for rep in range(7):
    W.timer_start()
    for _ in range(200): W.forward()
    ts.append(W.timer_stop() / 200)
f = sorted(ts)
for _ in range(100): W.forward(); W.inverse()
ts = []
for rep in range(7):
    W.timer_start()
    for _ in range(200): W.forward(); W.inverse()
    ts.append(W.timer_stop() / 200)
g = sorted(ts)
print(f"{wn} {N} L{L} variant={os.environ.get('PWT_FUSED_VARIANT','0')} fwd best {f[0]:.4f} med {f[3]:.4f} | fwd+inv best {g[0]:.4f} med {g[3]:.4f} ms", flush=True)
