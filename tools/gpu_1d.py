"""Scratch: batched 1D transforms (ndim=1 over the rows of a 2D array)."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
for wn in ([a for a in sys.argv[1:] if a != "noswt"] or ("haar", "db2", "db4", "sym8")):
    for swt in ((0,) if "noswt" in sys.argv else (0, 1)):
        W = pycudwt.Wavelets(img, wn, 3, ndim=1, do_swt=swt)
        for _ in range(3): W.forward(); W.inverse()
        W.profile_enable(1); W.forward(); W.inverse(); prof = W.profile_read(); W.profile_enable(0)
        W.timer_start()
        for _ in range(10): W.forward(); W.inverse()
        print(f"1d-batched {wn} swt={swt} 8192x8192 L3 fwd+inv: {W.timer_stop()/10:.4f} ms", [(t, round(m, 3)) for t, m in prof], flush=True)
