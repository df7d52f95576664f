# A/B of the fused forward's staging variants under ncu (--set full): 0 direct 128-bit loads, 4 bulk-copy (TMA) ring, 5 cp.async ring
cat > /tmp/f3.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
img = np.random.default_rng(0).standard_normal((8192, 8192)).astype(np.float32)
W = pycudwt.Wavelets(img, "db2", 3)
for _ in range(3): W.forward()
W.sync()
PY
for v in 0 4 5; do
  PWT_FUSED_VARIANT=$v ncu --set full --clock-control none -k regex:k_fwd3 -s 2 -c 1 -o gpurun_out/prof_fwd3_v$v -f python /tmp/f3.py > gpurun_out/ncu_fwd3_v$v.log 2>&1
done
for v in 0 4 5; do PWT_FUSED_VARIANT=$v python tools/gpu_fwd_ab.py db2 8192 3; done 2>&1 | tee gpurun_out/fwd3_tma_ab.txt
