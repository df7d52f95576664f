python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gputests.txt
bash tools/gpu_tma_ab.sh
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
