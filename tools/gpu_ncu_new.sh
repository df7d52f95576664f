# ncu --set full of the kernels added in round 2 (one launch each, warm): 1D rows DWT / SWT, volumetric z pass, fp64 two-pass
cat > /tmp/new1.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt, pypwt_b200
rng = np.random.default_rng(0)
img = rng.standard_normal((8192, 8192)).astype(np.float32)
for swt in (0, 1):
    W = pycudwt.Wavelets(img, "db2", 3, ndim=1, do_swt=swt)
    for _ in range(3): W.forward(); W.inverse()
    W.sync(); del W
vol = rng.standard_normal((512, 512, 512)).astype(np.float32)
V = pypwt_b200.Wavelets3D(vol, "db2", 1)
for _ in range(3): V.forward(); V.inverse()
V.sync(); del V
d = rng.standard_normal((4096, 4096))
D = pypwt_b200.Wavelets64(d, "db2", 1)
for _ in range(3): D.forward(); D.inverse()
D.sync()
PY
ncu --set full --clock-control none -k regex:"k_row_fwd|k_row_inv|k_row_swt|k_vol_z|k64_rows|k64_cols" --launch-skip 0 -c 60 -f -o gpurun_out/prof_new python /tmp/new1.py > gpurun_out/ncu_new.log 2>&1
tail -2 gpurun_out/ncu_new.log
python tools/ncu_summary.py gpurun_out/ncu_new_summary.csv gpurun_out/prof_new.ncu-rep
