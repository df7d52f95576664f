python tools/sweep_c5.py sep gpurun_out/c5_sep.json > gpurun_out/c5.log 2>&1
python tools/sweep_c5.py ns gpurun_out/c5_ns.json >> gpurun_out/c5.log 2>&1
timeout 600 python tools/sweep_c5.py nsdirect gpurun_out/c5_nsdirect.json >> gpurun_out/c5.log 2>&1
tail -40 gpurun_out/c5.log
# FMA-pipe utilisation of the level-1 kernels, one forward + inverse per wavelet
cat > /tmp/c5one.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, pycudwt
img = (np.random.default_rng(1).standard_normal((8192, 8192), dtype=np.float32) * 50 + 128)
for w in sys.argv[1:]:
    W = pycudwt.Wavelets(img, w, 1)
    W.forward(); W.inverse(); W.sync()
    W.forward(); W.inverse(); W.sync()
PY
ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --csv --log-file gpurun_out/c5_ncu_fma.csv python /tmp/c5one.py db2 db4 db6 sym8 db10 db12 db16 coif5 db20 > /dev/null 2>&1
tail -3 gpurun_out/c5_ncu_fma.csv | cut -c1-300
