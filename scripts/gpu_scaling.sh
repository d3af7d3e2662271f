# scaling of the batched top-k shooting (bench.py contract launch):  bash scripts/gpu_scaling.sh "<N list>" "<workload[:scaling] list>"
mkdir -p gpurun_out
NS=${1:-"1 2 4 8"}; WLS=${2:-"config2:weak config2:strong config3:strong config4:strong"}
for w in $WLS; do
  wl=${w%%:*}; sc=${w##*:}; [ "$sc" = "$w" ] && sc=weak
  for n in $NS; do
    out=gpurun_out/scale_${wl}_${sc}_$n
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > $out.json 2> $out.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --workload $wl --scaling $sc ${EXCH:+--exchange $EXCH} --steps 6 --warmup 3 > $out.json 2> $out.err || tail -5 $out.err
    fi
    python - <<PY
import json
try:
    d=json.loads(open("$out.json").read().strip().splitlines()[-1]); print("$wl", "$sc", $n, "shots/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "shots/step", d["config"]["shots_per_step"])
except Exception as e: print("$wl", "$sc", $n, "ERR", e)
PY
  done
done
