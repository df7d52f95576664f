python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "fused or agree or stack or idwt2 or full_size or vs_pdwt" 2>&1 | tail -12
for t3 in 8 16 32; do
  echo "INV T3=$t3"; PWT_FUSED_INV_T3=$t3 python bench.py --steps 20 --no-pdwt 2>&1 | grep -o "\"value\": [0-9.]*, \"unit\": \"Mpixel/s\", \"n_gpus\"\|kernel_ms_by_level[^}]*}"
done
