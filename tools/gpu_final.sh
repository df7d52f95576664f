# Round-end artifact capture (run on the GPU box through gpurun): bench line, ncu launch list, ncu --set full
# of the two fused kernels and of the long-filter kernels, configuration sweep.  Outputs under gpurun_out/final/.
set -x
O=gpurun_out/final
mkdir -p $O
for what in "$@"; do
case $what in
bench) python bench.py 2>&1 | grep -v "^Warn\|^Forc" | tail -1 > $O/bench.json; cut -c1-400 $O/bench.json ;;
ref) python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json ;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-pdwt > $O/ncu_bench.log 2>&1; grep -c . $O/launches.csv ;;
ncufused) ncu --set full --clock-control none --import-source on -k regex:"k_fwd3|k_inv3" -s 12 -c 2 -f -o $O/prof_fused python bench.py --steps 2 --warmup 1 --no-pdwt > $O/ncu_fused.log 2>&1; tail -1 $O/ncu_fused.log ;;
nculong) for w in db20 coif5 sym8; do ncu --set full --clock-control none -k regex:"k_tile|k_fwd|k_inv" -c 2 -f -o $O/prof_$w python tools/gpu_long.py $w > $O/ncu_$w.log 2>&1; tail -1 $O/ncu_$w.log; done ;;
ncuswt) ncu --set full --clock-control none -k regex:"k_swt" -s 8 -c 8 -f -o $O/prof_swt python tools/gpu_swt.py db4 > $O/ncu_swt.log 2>&1; tail -1 $O/ncu_swt.log ;;
sweep) python tools/sweep.py $O/sweep.md 2>&1 | grep -v "^Warn\|^Forc" | tail -3 ;;
swt) python tools/gpu_swt.py db4 2>&1 | grep -v "^Warn\|^Forc" > $O/swt_c4.txt; cat $O/swt_c4.txt ;;
esac
done
ls -la $O
