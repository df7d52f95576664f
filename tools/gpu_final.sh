# Round-end artifact capture (run on the GPU box through gpurun): bench line, ncu launch list, ncu --set full
# of the two fused kernels and of the long-filter kernels, configuration sweep.  Outputs under gpurun_out/final/.
set -x
O=${PWT_FINAL_DIR:-gpurun_out/final}
mkdir -p $O
for what in "$@"; do
case $what in
bench) python bench.py 2>&1 | grep -v "^Warn\|^Forc" | tail -1 > $O/bench.json; cut -c1-400 $O/bench.json ;;
ref) python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json ;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-pdwt --no-c3 --no-other > $O/ncu_bench.log 2>&1; grep -c . $O/launches.csv ;;
ncufused) ncu --set full --clock-control none --import-source on -k regex:"k_fwd3|k_inv3" -s 12 -c 2 -f -o $O/prof_fused python bench.py --steps 2 --warmup 1 --no-pdwt --no-c3 --no-other > $O/ncu_fused.log 2>&1; tail -1 $O/ncu_fused.log
  python tools/ncu_summary.py $O/ncu_fused_summary.csv $O/prof_fused.ncu-rep; for i in 0 1; do python tools/ncu_mix.py $O/prof_fused.ncu-rep $i > $O/ncu_fused_mix$i.txt; done; rm -f $O/prof_fused.ncu-rep ;;
nculong) for w in db4 sym8 db20; do ncu --set full --clock-control none --import-source on -k regex:"k_strip" -s 4 -c 2 -f -o $O/prof_$w python tools/gpu_one.py $w > $O/ncu_$w.log 2>&1; tail -1 $O/ncu_$w.log; done
  python tools/ncu_summary.py $O/ncu_strip_summary.csv $O/prof_db4.ncu-rep $O/prof_sym8.ncu-rep $O/prof_db20.ncu-rep
  for w in sym8 db20; do for i in 0 1; do python tools/ncu_mix.py $O/prof_$w.ncu-rep $i --regions > $O/ncu_strip_mix_${w}_$i.txt; done; done; rm -f $O/prof_db4.ncu-rep $O/prof_sym8.ncu-rep $O/prof_db20.ncu-rep ;;
ncuswt) ncu --set full --clock-control none --import-source on -k regex:"k_swt" -s 8 -c 8 -f -o $O/prof_swt python tools/gpu_swt.py db4 > $O/ncu_swt.log 2>&1; tail -1 $O/ncu_swt.log
  python tools/ncu_summary.py $O/ncu_swt_summary.csv $O/prof_swt.ncu-rep; for i in 0 7; do python tools/ncu_mix.py $O/prof_swt.ncu-rep $i > $O/ncu_swt_mix$i.txt; done; rm -f $O/prof_swt.ncu-rep ;;
ncu1d) ncu --set full --clock-control none -k regex:"k_strip_fwd1d|k_strip_inv1d|k_swt1d|k_haar" -s 6 -c 6 -f -o $O/prof_1d python tools/gpu_1d.py db2 > $O/ncu_1d.log 2>&1; tail -1 $O/ncu_1d.log
  ncu --set full --clock-control none -k regex:"k_swt1d" -s 6 -c 6 -f -o $O/prof_1ds python tools/gpu_1d.py db2 > $O/ncu_1ds.log 2>&1; tail -1 $O/ncu_1ds.log
  python tools/ncu_summary.py $O/ncu_1d_summary.csv $O/prof_1d.ncu-rep $O/prof_1ds.ncu-rep; rm -f $O/prof_1d.ncu-rep $O/prof_1ds.ncu-rep ;;
sweep) python tools/sweep.py $O/sweep.md 2>&1 | grep -v "^Warn\|^Forc" | tail -3 ;;
swt) python tools/gpu_swt.py db4 2>&1 | grep -v "^Warn\|^Forc" > $O/swt_c4.txt; cat $O/swt_c4.txt ;;
esac
done
ls -la $O
