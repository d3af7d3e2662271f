# round 2 quick pass: GPU tests (optional, TESTS=1), then one short bench line per "NAME=VALUE[,NAME=VALUE...]" argument ("base" = no knob)
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; fi
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline ${EXTRAS:---no-extras} ${WL:+--workload $WL} > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err || tail -3 gpurun_out/q_$tag.err
python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/q_{t}.json"))
    r=[v for k,v in d["kernels"].items() if k.startswith("raster")][0]
    print(t, round(d["value"],1), "shots/s e2e", round(d["e2e"]["value"],1), {k.split(" ")[0]:round(v["ms_per_batch"],4) for k,v in d["kernels"].items()}, "setup/queue", round(r["setup_ms"],4), round(r.get("queue_ms", r.get("ring_ms", 0)),4),
          "K2", round(d["process_hemicube"]["gpix_per_s"],1), "Gpix/s", round(d["process_hemicube"]["ms_per_launch"],4), "ms")
except Exception as e: print(t,"ERR",e)
PY
}
for v in "$@"; do if [ "$v" = "base" ]; then run base RAD_X=0; else run "$(echo $v | tr '=,' '__')" $(echo $v | tr ',' ' '); fi; done
