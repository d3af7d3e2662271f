for mb in 4 6 8; do
  echo "== setup_minb $mb"
  RAD_SETUP_MINB=$mb python scripts/prof_batches.py --workload config2 --batches 32 | cut -c100-
  RAD_SETUP_MINB=$mb python scripts/prof_batches.py --workload config2_k1 --batches 256 | cut -c100-
done
echo "== config3 (P=250063, N=1024, k=64)"
python scripts/prof_batches.py --workload config3 --batches 2 | cut -c100-
RAD_SPLIT_LIMIT=0 python scripts/prof_batches.py --workload config3 --batches 2 | cut -c100-
