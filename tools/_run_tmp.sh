for b in 16 48 128 512; do
  echo "bps=$b"; BFM_BAND_BLOCKS_PER_SM=$b python tools/stage_bench.py 2>&1 | tail -1
done
