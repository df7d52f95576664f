python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "fused or agree or stack or test_dwt2 or idwt2 or vs_pdwt or odd" 2>&1 | tail -2
for v in 0 2 3; do for t3 in 8 16; do
  echo "FUSED VARIANT=$v T3=$t3"; PWT_FUSED_VARIANT=$v PWT_FUSED_T3=$t3 python bench.py --steps 20 --no-pdwt 2>&1 | grep -o "\"value\": [0-9.]*, \"unit\": \"Mpixel/s\", \"n_gpus\"\|kernel_ms_by_level[^}]*}"
done; done
echo "NO FUSED"; PWT_NO_FUSED=1 python bench.py --steps 20 --no-pdwt 2>&1 | grep -o "\"value\": [0-9.]*, \"unit\": \"Mpixel/s\", \"n_gpus\"\|kernel_ms_by_level[^}]*}"
