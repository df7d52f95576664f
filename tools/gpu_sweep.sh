python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -k "swt" 2>&1 | grep -v "^Warn\|^Forc" | tail -8
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import pycudwt
img=(np.random.default_rng(1).standard_normal((8192,8192),dtype=np.float32)*50+128)
for cs in (0,1):
    W=pycudwt.Wavelets(img,'db4',4,do_swt=1,do_cycle_spinning=cs)
    def f(): W.forward(); W.hard_threshold(20.0); W.inverse()
    for _ in range(3): f()
    W.sync(); W.timer_start()
    for _ in range(10): f()
    ms=W.timer_stop()/10
    print("C4 swt db4 L4 8192^2 fwd+hard+inv cycle_spinning=%d: %.3f ms  %.0f Mpx/s  %.0f GB/s (112 B/px)"%(cs,ms,img.size/ms/1e3,112*img.size/ms/1e6))
    W.profile_enable(1); W.forward(); W.inverse(); print(W.profile_read())
    del W
PY
