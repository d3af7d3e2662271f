# strong scaling of the batched top-k shooting over 1/2/4/8 GPUs (bench.py contract launch)
mkdir -p gpurun_out
for wl in config2 config3 config4; do
  for n in 1 2 4 8; do
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${wl}_$n.json 2> gpurun_out/scale_${wl}_$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --workload $wl --steps 6 --warmup 3 > gpurun_out/scale_${wl}_$n.json 2> gpurun_out/scale_${wl}_$n.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${wl}_$n.json").read().strip().splitlines()[-1]); print("${wl}", $n, "shots/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("${wl}", $n, "ERR", e)
PY
  done
done
