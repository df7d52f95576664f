python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -k "test_dwt2 or test_idwt2 or odd_sizes or vs_pdwt or stack or agree" 2>&1 | grep -v "^Warn\|^Forc" | tail -8
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import pycudwt
img=(np.random.default_rng(1).standard_normal((8192,8192),dtype=np.float32)*50+128)
for w in ['db6','sym8','db10','db12','db20']:
    for minf in (12, 99):
        import os; os.environ['PWT_TILE_MIN_F']=str(minf)
        W=pycudwt.Wavelets(img,w,5)
        def f(): W.forward(); W.inverse()
        for _ in range(3): f()
        W.sync(); W.timer_start()
        for _ in range(10): f()
        ms=W.timer_stop()/10
        print("%-6s F=%2d tile_min_f=%2d: %.3f ms  %.0f Mpx/s"%(w,W.hlen,minf,ms,img.size/ms/1e3), flush=True)
        del W
PY
