for N in 4096 2048; do
for t in 0 2 4 6 8 12 16 24 32; do PWT_FUSED_T3=$t PWT_FUSED_INV_T3=$t python tools/gpu_fwd_ab.py db2 $N 3 2>&1 | sed "s/^/T3=$t /"; done
done 2>&1 | tee gpurun_out/t3.txt
