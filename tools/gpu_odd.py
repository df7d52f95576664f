"""Scratch: sizes the register kernels do not cover (width not a multiple of 128): auto vs strip kernels."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
for shape in ((8000, 8000), (5000, 7001), (64, 2000, 2000)):
    img = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    for wn in ("db2", "db3", "db4", "sym8"):
        res = []
        for mode in (0, 4):
            W = pycudwt.Wavelets(img, wn, 3); W.set_kernel_mode(mode)
            for _ in range(3): W.forward(); W.inverse()
            W.timer_start()
            for _ in range(10): W.forward(); W.inverse()
            res.append(W.timer_stop() / 10)
        px = np.prod(shape)
        print(f"{shape} {wn} L3 fwd+inv: auto {res[0]:.4f} ms ({16*px/res[0]/1e6:.0f} GB/s)  strip {res[1]:.4f} ms ({16*px/res[1]/1e6:.0f} GB/s)", flush=True)
