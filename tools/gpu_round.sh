# usage: bash tools/gpu_round.sh [tests|bench|ncu ...]   (run on the GPU box through gpurun)
set -x
mkdir -p gpurun_out
for what in "$@"; do
case $what in
golden) python tests/golden/make_golden.py gpurun_out/pdwt_golden.npz 2>&1 | tail -2 ;;
quick) python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "agree or odd_sizes or idwt2 or full_size or stack or vs_pdwt" 2>&1 | grep -v "^Warning\|^Forcing" | tail -15 ;;
tests) python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 -p no:cacheprovider 2>&1 | grep -v "^Warning\|^Forcing" | tail -40 ;;
smoke) python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
bench) python bench.py --steps 20 --warmup 3 2>&1 | grep -v "^Warning\|^Forcing" | tee gpurun_out/bench.json | tail -3 ;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-pdwt > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log; grep -c . gpurun_out/launches.csv ;;
ncufull) ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-k_dwt}" -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-6} -f -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --no-pdwt > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/ ;;
multi) python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -15 ;;
bench2) python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 500 --warmup 5 --no-pdwt 2>&1 | grep -v "^Warn\|^Forc\|^W1\|^\*\*" | tee gpurun_out/bench_n2.json | cut -c1-2500 ;;
esac
done
