# ncu --set full of the fused volumetric level kernels (kernels_vol_fused.cu): 512^3 db2, one level
cat > /tmp/vol1.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pypwt_b200
vol = np.random.default_rng(0).standard_normal((512, 512, 512)).astype(np.float32)
V = pypwt_b200.Wavelets3D(vol, sys.argv[1], 1)
for _ in range(2): V.forward(); V.inverse()
V.sync()
PY
for wn in db2 db3; do
  ncu --set full --clock-control none -k regex:"k_vol3" --launch-skip 1 -c 3 -f -o gpurun_out/prof_vol_$wn python /tmp/vol1.py $wn > gpurun_out/ncu_vol_$wn.log 2>&1
  tail -1 gpurun_out/ncu_vol_$wn.log
done
python tools/ncu_summary.py gpurun_out/ncu_vol_summary.csv gpurun_out/prof_vol_db2.ncu-rep gpurun_out/prof_vol_db3.ncu-rep
rm -f gpurun_out/prof_vol_*.ncu-rep
