for v in 0 1; do for w in db2 db3; do PWT_TAIL_STRIP=$v python - $w <<'PY'
import sys, os; sys.path.insert(0, ".")
import numpy as np, pycudwt
wn = sys.argv[1]
for N in (8192, 4096):
  img = np.random.default_rng(0).standard_normal((N, N)).astype(np.float32)
  for L in (3, 4, 5, 7):
    W = pycudwt.Wavelets(img, wn, L)
    for _ in range(5): W.forward(); W.inverse()
    W.sync(); ts = []
    for r in range(3):
        W.timer_start()
        for _ in range(20): W.forward(); W.inverse()
        ts.append(W.timer_stop() / 20)
    print("TAIL_STRIP=%s %-4s %d^2 L%d %.4f ms frac %.3f" % (os.environ["PWT_TAIL_STRIP"], wn, N, L, sorted(ts)[1], 16 * img.size / sorted(ts)[1] / 1e6 / 6549.4), flush=True)
PY
done; done 2>&1 | tee gpurun_out/tail.txt
