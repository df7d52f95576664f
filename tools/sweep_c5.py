#!/usr/bin/env python3
"""BASELINE config 5: filter-length sweep haar -> db20 / coif5 at 8192^2, 5 levels, forward + inverse, separable against
non-separable (both the rank-1 shortcut, which runs the separable kernels with the reference's slot swap, and the TRUE
F x F stencils, `PWT_NS_DIRECT=1`), with the multiply-add count per pixel and the resulting fraction of the fp32 FMA pipe.
One process per mode (the knob is read once):  python tools/sweep_c5.py sep|ns|nsdirect out.json [quick]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode, out = sys.argv[1], sys.argv[2]
quick = len(sys.argv) > 3
if mode == "nsdirect":
    os.environ["PWT_NS_DIRECT"] = "1"
import pycudwt  # noqa: E402

N, L = 8192, 5
img = (np.random.default_rng(1).standard_normal((N, N), dtype=np.float32) * 50 + 128)
wl = ["haar", "db2", "db3", "db4", "db6", "db8", "db10", "db12", "db16", "db20", "coif5", "sym8", "bior6.8"]
if quick:
    wl = ["haar", "db2", "db4", "db8", "db20"]
if mode == "nsdirect":
    wl = [w for w in wl if w not in ("db16", "coif5", "sym8", "bior6.8")]       # F^2 stencils: minutes beyond F = 40
res = []
for w in wl:
    W = pycudwt.Wavelets(img, w, L, do_separable=0 if mode != "sep" else 1)
    F = W.hlen
    reps = 20 if (mode != "nsdirect" or F <= 8) else (5 if F <= 20 else 2)
    for _ in range(2 if mode == "nsdirect" else 5):
        W.forward(); W.inverse()
    W.sync()
    ts = []
    for rep in range(3):
        W.timer_start()
        for _ in range(reps):
            W.forward(); W.inverse()
        ts.append(W.timer_stop() / reps)
    ms = sorted(ts)[1]
    l0 = W.launch_count
    W.forward(); W.inverse()
    nl = W.launch_count - l0
    geo = sum(0.25 ** l for l in range(L))
    haar = F == 2
    if mode == "nsdirect" and not haar:
        fma_px = 2 * F * F * geo                  # 4 bands x F^2 per output quad, forward and inverse (nonseparable.cu:114-225)
    else:
        fma_px = (4 * F if not haar else 4) * geo # row + column pass, analysis and synthesis: 4F per level-input pixel
    px = N * N
    res.append(dict(mode=mode, wname=w, F=F, levels=W.levels, ms=ms, gpx_s=px / ms / 1e6, frac_hbm_16B=16 * px / ms / 1e6 / 6549.4,
                    fma_per_px=fma_px, tfma_s=fma_px * px / ms / 1e9, launches=nl))
    print(res[-1], flush=True)
    del W
json.dump(res, open(out, "w"), indent=1)
