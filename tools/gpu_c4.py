"""BASELINE config 4: SWT db4 4 levels 8192^2, denoising loop with cycle spinning + hard threshold."""
import sys, os; sys.path.insert(0, ".")
import numpy as np, pycudwt
N = 8192
img = (np.random.default_rng(1).standard_normal((N, N), dtype=np.float32) * 50 + 128)
def t(W, fn, reps=10):
    for _ in range(3): fn(W)
    W.sync(); ts = []
    for r in range(3):
        W.timer_start()
        for _ in range(reps): fn(W)
        ts.append(W.timer_stop() / reps)
    return sorted(ts)[1]
def den(W): W.forward(); W.hard_threshold(20.0); W.inverse()
def fi(W): W.forward(); W.inverse()
bpp = 2 * (3 * 4 + 2) * 4
for name, kw, fn in (("fwd+hard+inv, cycle spinning", dict(do_swt=1, do_cycle_spinning=1), den), ("fwd+hard+inv", dict(do_swt=1), den), ("fwd+inv", dict(do_swt=1), fi)):
    W = pycudwt.Wavelets(img, "db4", 4, **kw)
    ms = t(W, fn)
    l0 = W.launch_count; fn(W); nl = W.launch_count - l0
    print("C4 swt db4 L4 8192^2 %-30s NO_FOLD_CS=%s %.4f ms %.1f Gpx/s frac %.3f launches %d" % (name, os.environ.get("PWT_NO_FOLD_CS", "0"), ms, N * N / ms / 1e6, bpp * N * N / ms / 1e6 / 6549.4, nl), flush=True)
    del W
