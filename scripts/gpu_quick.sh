# quick correctness + A/B pass: GPU tests, then short bench lines for tuning knobs given as "NAME=VALUE" arguments
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${WL:+--workload $WL} > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err || tail -3 gpurun_out/q_$tag.err
python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/q_{t}.json"))
    print(t, round(d["value"],1), "shots/s e2e", round(d["e2e"]["value"],1), {k.split(" ")[0]:round(v["ms_per_batch"],4) for k,v in d["kernels"].items()}, "setup/queue", round(d["kernels"]["raster (K1: raster_setup + raster_queue)"]["setup_ms"],4), round(d["kernels"]["raster (K1: raster_setup + raster_queue)"]["queue_ms"],4))
except Exception as e: print(t,"ERR",e)
PY
}
run base RAD_X=0
for v in "$@"; do run "$(echo $v | tr '=' '_')" "$v"; done
