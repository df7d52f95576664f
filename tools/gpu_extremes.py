"""Shape extremes: long single rows, thin and tall images, against the 16 B/px roofline (looking for pathological cases)."""
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
rng = np.random.default_rng(0)
def run(shape, wn, L, bpp=16, **kw):
    img = rng.standard_normal(shape).astype(np.float32)
    W = pycudwt.Wavelets(img, wn, L, **kw)
    for _ in range(2): W.forward(); W.inverse()
    W.sync(); ts = []
    for r in range(3):
        W.timer_start()
        for _ in range(5): W.forward(); W.inverse()
        ts.append(W.timer_stop() / 5)
    t = sorted(ts)[1]
    l0 = W.launch_count; W.forward(); W.inverse()
    print("%-16s %-5s L%d %-24s %.4f ms frac %.3f launches %d" % ("x".join(map(str, shape)), wn, W.levels, kw, t, bpp * img.size / t / 1e6 / 6549.4, W.launch_count - l0), flush=True)
for wn in ("haar", "db2", "sym8"):
    run((16777216,), wn, 4, ndim=1)
    run((1, 16777216), wn, 4, ndim=1)
    run((16, 1048576), wn, 3)
    run((1048576, 16), wn, 1)
    run((64, 262144), wn, 3)
    run((262144, 64), wn, 3)
    run((16777216,), wn, 3, bpp=40, ndim=1, do_swt=1)
    run((16, 1048576), wn, 2, bpp=64, do_swt=1)
