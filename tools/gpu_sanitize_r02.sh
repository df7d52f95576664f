# compute-sanitizer over the round-2 kernels (small shapes): memcheck, then racecheck on the shared-memory kernels
export SMALL=1
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt, pypwt_b200
rng = np.random.default_rng(0)
# fused 1D rows (DWT + SWT), odd and even widths, several rows per CTA and one row per CTA
for shp in ((5, 24), (33, 130), (7, 1000), (3, 4099), (2, 8192), (64, 512)):
    img = (rng.standard_normal(shp) * 50).astype(np.float32)
    for wn in ("haar", "db2", "sym8", "db20", "bior3.9"):
        for swt in (0, 1):
            if swt and shp[1] % 4:
                continue
            try:
                W = pycudwt.Wavelets(img, wn, 5, ndim=1, do_swt=swt)
            except ValueError:
                continue
            W.forward(); c = W.coeffs; W.inverse(); _ = W.image
# volumes (window z pass + generic z pass on odd planes) and double precision
for shp in ((16, 24, 32), (17, 21, 27), (40, 64, 64)):
    vol = (rng.standard_normal(shp) * 50).astype(np.float32)
    for wn in ("haar", "db2", "sym8"):
        try:
            W = pypwt_b200.Wavelets3D(vol, wn, 2)
        except ValueError:
            continue
        W.forward(); c = W.coeffs; W.soft_threshold(1.0); n = W.norms(); W.inverse(); _ = W.image
for kw in (dict(), dict(do_swt=1), dict(ndim=1), dict(do_cycle_spinning=1)):
    img = rng.standard_normal((67, 93))
    W = pypwt_b200.Wavelets64(img, "db3", 2, **kw)
    W.forward(); c = W.coeffs; W.hard_threshold(0.5); n = W.norms(); W.inverse(); _ = W.image
# recorded cycle-spinning shifts of SWT plans
img = (rng.standard_normal((96, 128)) * 50).astype(np.float32)
W = pycudwt.Wavelets(img, "db4", 3, do_swt=1, do_cycle_spinning=1)
for _ in range(2):
    W.forward(); c = W.coeffs; i2 = W.image; W.hard_threshold(3.0); W.inverse(); _ = W.image
print("sanitizer workload done")
PY
compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log; tail -4 gpurun_out/sanitize_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log; tail -4 gpurun_out/sanitize_racecheck.log
