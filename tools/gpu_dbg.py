import sys, os, numpy as np
sys.path.insert(0, ".")
import pycudwt
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
W = pycudwt.Wavelets(img, sys.argv[1] if len(sys.argv) > 1 else "db2", 3)
for _ in range(10): W.forward()
best = 1e9
for rep in range(3):
    W.timer_start()
    for _ in range(300): W.forward()
    best = min(best, W.timer_stop() / 300)
print(f"fwd debug={os.environ.get('PWT_FUSED_DEBUG','0')} variant={os.environ.get('PWT_FUSED_VARIANT','0')}: {best:.4f} ms", flush=True)
