mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for ia in 32 64 128; do
  echo "== inline_area $ia"
  RAD_INLINE_AREA=$ia python scripts/prof_batches.py --workload config2 --batches 32
  RAD_INLINE_AREA=$ia python scripts/prof_batches.py --workload config2_k1 --batches 128
done
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:(round(v['ms_per_launch'],4),round(v['share'],3)) for k,v in d['kernels'].items()}); print(d['process_hemicube']['gpix_per_s'], d['process_hemicube']['frac']); print(d['cpu_baseline'])"
tail -3 gpurun_out/bench5.err
timeout 300 python bench.py --workload config2_k1 --steps 5 --no-cpu-baseline > gpurun_out/bench5_k1.json 2>gpurun_out/bench5_k1.err
python -c "
import json; d=json.load(open('gpurun_out/bench5_k1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:(round(v['ms_per_launch'],4),round(v['share'],3)) for k,v in d['kernels'].items()})"
tail -3 gpurun_out/bench5_k1.err
