for mb in 0 32 64 200; do for nh in 0; do
  echo "L2_PERSIST_MB=$mb"; PWT_VERBOSE=1 PWT_L2_PERSIST_MB=$mb PWT_NO_HINTS=$nh PWT_REG_TILE_ROWS=16 PWT_REG_FWD_VARIANT=2 python bench.py --steps 20 --no-pdwt 2>&1 | grep -o "pwt: pers.*\|\"value\": [0-9.]*, \"unit\": \"Mpixel/s\", \"n_gpus\"\|kernel_ms_by_level[^}]*}"
done; done
