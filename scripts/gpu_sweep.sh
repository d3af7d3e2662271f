timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for mb in 4 6 8; do
  echo "== setup_minb $mb"
  RAD_SETUP_MINB=$mb python scripts/prof_batches.py --workload config2 --batches 32 | cut -c100-
  RAD_SETUP_MINB=$mb python scripts/prof_batches.py --workload config2_k1 --batches 256 | cut -c100-
done
