for v in 0 1; do for w in db2 db3; do PWT_TAIL_STRIP=$v python - $w <<'PY'
import sys, os; sys.path.insert(0, ".")
import numpy as np, pycudwt
wn = sys.argv[1]
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
W = pycudwt.Wavelets(img, wn, 4)
for _ in range(5): W.forward(); W.inverse()
W.profile_enable(1)
for _ in range(3): W.forward(); W.inverse()
r = W.profile_read()
print("TAIL_STRIP=%s %s L4:" % (os.environ["PWT_TAIL_STRIP"], wn), [(t, round(ms * 1000, 1)) for t, ms in r[-4:]], flush=True)
PY
done; done
