# compute-sanitizer over the kernels added late in round 2 (small shapes): two-pass SWT (float / double), fused SWT strip F = 18 / 20,
# fused double-precision levels (rotating and shifted windows, Haar), fused volumetric levels -- memcheck, then racecheck
cat > /tmp/san2.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt, pypwt_b200
rng = np.random.default_rng(0)
# 2D SWT: widths = 0..3 mod 4 (two-pass kernels with 4 / 2 / 1 columns per thread), F = 18 / 20 on the strip kernels, F >= 22 two-pass
for shp in ((64, 128), (37, 90), (41, 67), (2, 48, 132)):
    img = (rng.standard_normal(shp) * 50).astype(np.float32)
    for wn in ("db2", "sym8", "db9", "db10", "db12", "coif5", "db20"):
        try:
            W = pycudwt.Wavelets(img, wn, 3, do_swt=1)
        except ValueError:
            continue
        W.forward(); c = W.coeffs; W.soft_threshold(2.0); W.inverse(); _ = W.image
# double precision: fused levels (F <= 20 rotating windows, F >= 22 shifted windows / two threads per column), Haar, two-pass SWT
for shp in ((70, 300), (33, 131), (2, 40, 264)):
    img = rng.standard_normal(shp) * 50
    for wn in ("haar", "db2", "db3", "sym8", "db10", "db11", "coif5", "db20"):
        try:
            W = pypwt_b200.Wavelets64(img, wn, 3)
        except ValueError:
            continue
        W.forward(); c = W.coeffs; W.inverse(); _ = W.image
    for wn in ("db2", "sym8", "db12"):
        try:
            W = pypwt_b200.Wavelets64(img, wn, 2, do_swt=1)
        except ValueError:
            continue
        W.forward(); c = W.coeffs; W.inverse(); _ = W.image
# volumes: fused levels (tiles with overhang, odd heights / depths, several z segments)
for shp in ((20, 40, 96), (13, 37, 88), (34, 70, 160)):
    vol = (rng.standard_normal(shp) * 50).astype(np.float32)
    for wn in ("haar", "db2", "db3"):
        try:
            W = pypwt_b200.Wavelets3D(vol, wn, 2)
        except ValueError:
            continue
        W.forward(); c = W.coeffs; W.inverse(); _ = W.image
print("sanitizer workload done")
PY
compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/sanitize2_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize2_memcheck.log; tail -4 gpurun_out/sanitize2_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/sanitize2_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize2_racecheck.log; tail -4 gpurun_out/sanitize2_racecheck.log
