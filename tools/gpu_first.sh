set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python tests/golden/make_golden.py gpurun_out/pdwt_golden.npz 2>&1 | tail -5
python -m pytest tests -m gpu -q --maxfail=60 --timeout=600 -p no:cacheprovider 2>&1 | tail -60
