"""Scratch: does the power-of-two row pitch (8192 floats = 32 KB) cost bandwidth?  Same kernel, neighbouring widths."""
import sys, numpy as np
sys.path.insert(0, ".")
import pycudwt
rng = np.random.default_rng(0)
for wn in ("db2",):
    for nc in (7680, 7936, 8064, 8192, 8320, 8448, 8704, 9216, 4096, 4224, 16384, 16512):
        nr = 8192 if nc < 10000 else 4096
        img = rng.standard_normal((nr, nc)).astype(np.float32)
        W = pycudwt.Wavelets(img, wn, 3)
        res = []
        for what in ("fwd", "fwd+inv"):
            def step():
                W.forward()
                if what == "fwd+inv": W.inverse()
            for _ in range(10): step()
            best = 1e9
            for rep in range(3):
                W.timer_start()
                for _ in range(200): step()
                best = min(best, W.timer_stop() / 200)
            res.append(best)
        px = nr * nc
        print(f"{wn} {nr}x{nc}: fwd {res[0]:.4f} ms ({px*8/res[0]/1e9:.0f} GB/s)  fwd+inv {res[1]:.4f} ms ({px*16/res[1]/1e9:.0f} GB/s, {px/res[1]/1e3:.0f} Mpx/s)", flush=True)
        del W
