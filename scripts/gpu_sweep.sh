for sl in 65536 100000000; do for ia in 16 32 64; do
  echo "== split_limit $sl inline_area $ia"
  RAD_SPLIT_LIMIT=$sl RAD_INLINE_AREA=$ia python scripts/prof_batches.py --workload config2 --batches 32 | cut -c100-
  RAD_SPLIT_LIMIT=$sl RAD_INLINE_AREA=$ia python scripts/prof_batches.py --workload config2_k1 --batches 256 | cut -c100-
done; done
