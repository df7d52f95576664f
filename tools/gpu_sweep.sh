python bench.py --steps 500 --no-pdwt 2>&1 | grep -o "\"value\": [0-9.]*, \"unit\": \"Mpixel/s\", \"n_gpus\"\|kernel_ms_by_level[^}]*}"
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'.')
import pycudwt
for shape,w in [((4096,4096),'haar'),((4096,4096),'db2'),((8192,8192),'haar'),((2048,2048),'db2'),((16,2048,2048),'db2'),((8192,8192),'db3')]:
    img=(np.random.default_rng(1).standard_normal(shape,dtype=np.float32)*50+128)
    W=pycudwt.Wavelets(img,w,3)
    def f(): W.forward(); W.inverse()
    for _ in range(5): f()
    W.sync(); W.timer_start()
    for _ in range(50): f()
    ms=W.timer_stop()/50
    print("%s %s: %.4f ms  %.0f Mpx/s  frac %.3f"%(shape,w,ms,img.size/ms/1e3, 16*img.size/ms/1e6/6549.4), flush=True)
    del W
PY
