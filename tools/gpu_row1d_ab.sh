for v in 000 100 010 001 110; do cp gpurun_variants/lib_$v.so pypwt_b200/libpwt_b200.so; python tools/gpu_row1d_time.py v$v; done 2>&1 | tee gpurun_out/row1d_ab.txt
cp gpurun_variants/lib_000.so pypwt_b200/libpwt_b200.so
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
for wn in ("haar", "sym8"):
    W = pycudwt.Wavelets(img, wn, 3, ndim=1)
    for _ in range(3): W.forward(); W.inverse()
    W.sync()
PY
ncu --set full --clock-control none --import-source on -k regex:k_row -s 8 -c 4 -o gpurun_out/prof_row1d -f python /tmp/one.py > gpurun_out/ncu_row1d.log 2>&1
ncu -i gpurun_out/prof_row1d.ncu-rep --page raw --csv > gpurun_out/ncu_row1d_raw.csv 2>/dev/null
