"""Sweep for performance cliffs: stacks of small images, batched 1D shapes, mid-size SWT, against their compulsory traffic."""
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
def run(shape, wn, L, bpp, **kw):
    img = rng.standard_normal(shape).astype(np.float32)
    W = pycudwt.Wavelets(img, wn, L, **kw)
    for _ in range(3): W.forward(); W.inverse()
    W.sync(); ts = []
    for r in range(3):
        W.timer_start()
        for _ in range(10): W.forward(); W.inverse()
        ts.append(W.timer_stop() / 10)
    t = sorted(ts)[1]
    l0 = W.launch_count; W.forward(); W.inverse()
    print("%-18s %-5s L%d %-28s %.4f ms frac %.3f launches %d" % ("x".join(map(str, shape)), wn, W.levels, kw, t, bpp * img.size / t / 1e6 / 6549.4, W.launch_count - l0), flush=True)
for shape in ((64, 512, 512), (16, 1024, 1024), (256, 256, 256)):
    for wn in ("haar", "db2", "sym8"):
        for L in (1, 2, 3): run(shape, wn, L, 16)
for shape in ((8192, 1024), (65536, 256), (1024, 65536), (4096, 4096)):
    for wn in ("haar", "db2", "sym8"):
        for L in (1, 3): run(shape, wn, L, 16, ndim=1)
for shape in ((1024, 1024), (2048, 2048)):
    for wn in ("haar", "db4"):
        run(shape, wn, 2, 2 * 8 * 4, do_swt=1)
