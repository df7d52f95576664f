# ncu --set full of the fused double-precision level kernels (kernels_f64_fused.cu): 8192^2, one level, db2 / sym8 / db20
# (second forward + inverse of each plan; the reports are summarised and removed: gpurun_out/ returns at most 64 MiB)
cat > /tmp/f64n.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pypwt_b200
d = np.random.default_rng(0).standard_normal((8192, 8192))
D = pypwt_b200.Wavelets64(d, sys.argv[1], 1)
for _ in range(2): D.forward(); D.inverse()
D.sync()
PY
for wn in db2 sym8 db20; do
  ncu --set full --clock-control none -k regex:"k64_fused" --launch-skip 2 -c 2 -f -o gpurun_out/prof_f64_$wn python /tmp/f64n.py $wn > gpurun_out/ncu_f64_$wn.log 2>&1
  tail -1 gpurun_out/ncu_f64_$wn.log
done
python tools/ncu_summary.py gpurun_out/ncu_f64_summary.csv gpurun_out/prof_f64_db2.ncu-rep gpurun_out/prof_f64_sym8.ncu-rep gpurun_out/prof_f64_db20.ncu-rep
rm -f gpurun_out/prof_f64_*.ncu-rep
