timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for ia in 8 16 32 64; do
  echo "== inline_area $ia"
  RAD_INLINE_AREA=$ia python scripts/prof_batches.py --workload config2 --batches 32 | cut -c100-
  RAD_INLINE_AREA=$ia python scripts/prof_batches.py --workload config2_k1 --batches 256 | cut -c100-
done
python scripts/prof_batches.py --workload config3 --batches 2 | cut -c100-
