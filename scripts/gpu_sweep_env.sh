# A/B of tuning knobs: each argument is a comma-separated list of NAME=VALUE settings for one short bench run
mkdir -p gpurun_out
for v in "$@"; do
  tag=$(echo "$v" | tr '=,' '__')
  env $(echo "$v" | tr ',' ' ') timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${WL:+--workload $WL} > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err || tail -3 gpurun_out/q_$tag.err
  python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/q_{t}.json")); print(t, round(d["value"],1), "shots/s e2e", round(d["e2e"]["value"],1))
except Exception as e: print(t,"ERR",e)
PY
done
