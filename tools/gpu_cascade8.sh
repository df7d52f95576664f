for v in 8 10; do PWT_STRIP_MIN_F=$v python - <<'PY'
import sys, os; sys.path.insert(0, ".")
import numpy as np, pycudwt
for N in (8192, 4096):
  img = np.random.default_rng(0).standard_normal((N, N)).astype(np.float32)
  for wn in ("db4", "sym4"):
    for L in (3, 5):
        W = pycudwt.Wavelets(img, wn, L)
        for _ in range(5): W.forward(); W.inverse()
        W.sync(); ts = []
        for r in range(3):
            W.timer_start()
            for _ in range(20): W.forward(); W.inverse()
            ts.append(W.timer_stop() / 20)
        l0 = W.launch_count; W.forward(); W.inverse(); nl = W.launch_count - l0
        print("STRIP_MIN_F=%s %-5s %d^2 L%d %.4f ms frac %.3f launches %d" % (os.environ["PWT_STRIP_MIN_F"], wn, N, L, sorted(ts)[1], 16 * img.size / sorted(ts)[1] / 1e6 / 6549.4, nl), flush=True)
PY
done
