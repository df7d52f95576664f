for v in 8 6; do for w in db3 coif1 bior2.2 sym3; do PWT_STRIP_MIN_F=$v python - $w <<'PY'
import sys, os; sys.path.insert(0, ".")
import numpy as np, pycudwt
wn = sys.argv[1]
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
for L in (3, 5):
    W = pycudwt.Wavelets(img, wn, L)
    for _ in range(5): W.forward(); W.inverse()
    W.sync(); ts = []
    for r in range(3):
        W.timer_start()
        for _ in range(20): W.forward(); W.inverse()
        ts.append(W.timer_stop() / 20)
    print("STRIP_MIN_F=%s %-7s F=%d L%d %.4f ms frac %.3f" % (os.environ["PWT_STRIP_MIN_F"], wn, W.hlen, L, sorted(ts)[1], 16 * img.size / sorted(ts)[1] / 1e6 / 6549.4), flush=True)
PY
done; done 2>&1 | tee gpurun_out/strip6.txt
