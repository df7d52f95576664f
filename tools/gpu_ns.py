"""Scratch: non-separable transform timing (rank-1 filters evaluated separably vs direct F x F kernels)."""
import sys, os, numpy as np
sys.path.insert(0, ".")
import pycudwt
img = np.random.default_rng(0).standard_normal((4096, 4096)).astype(np.float32)
for wn in ("db2", "db4", "db8"):
    for swt in (0, 1):
        W = pycudwt.Wavelets(img, wn, 5 if not swt else 3, do_separable=0, do_swt=swt)
        for _ in range(3): W.forward(); W.inverse()
        W.timer_start()
        for _ in range(10): W.forward(); W.inverse()
        print(f"nonsep {wn} swt={swt} 4096^2: {W.timer_stop()/10:.4f} ms  direct={os.environ.get('PWT_NS_DIRECT','0')}", flush=True)
