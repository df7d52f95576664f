python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -k "deferred or thresh or state_machine or add_wavelet or golden or vs_pdwt" 2>&1 | grep -v "^Warn\|^Forc" | tail -12
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'.')
import pycudwt
img=(np.random.default_rng(1).standard_normal((8192,8192),dtype=np.float32)*50+128)
W=pycudwt.Wavelets(img,'db2',3)
def f(): W.forward(); W.soft_threshold(10.0); W.inverse()
for _ in range(5): f()
W.sync(); W.timer_start()
for _ in range(50): f()
ms=W.timer_stop()/50
print("denoise fwd+soft+inv 8192^2 db2 L3: %.4f ms  %.0f Mpx/s"%(ms,img.size/ms/1e3))
PY
