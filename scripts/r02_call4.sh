mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err || tail -3 gpurun_out/q_$tag.err
python - "$tag" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/q_{t}.json"))
    print(t, round(d["value"],1), "shots/s e2e", round(d["e2e"]["value"],1), {k.split(" ")[0]:round(v["ms_per_batch"],4) for k,v in d["kernels"].items()}, "setup/queue", round(d["kernels"]["raster (K1: raster_setup + raster_queue)"]["setup_ms"],4), round(d["kernels"]["raster (K1: raster_setup + raster_queue)"]["queue_ms"],4))
except Exception as e: print(t,"ERR",e)
PY
}
run l8_g13 RAD_LANES=8 RAD_L2_GROUP_MB=13
run l8_g7 RAD_LANES=8 RAD_L2_GROUP_MB=7
run l4_g26 RAD_LANES=4 RAD_L2_GROUP_MB=26
run l4_g13 RAD_LANES=4 RAD_L2_GROUP_MB=13
run l2_g51 RAD_LANES=2 RAD_L2_GROUP_MB=51
run l1_g101 RAD_LANES=1 RAD_L2_GROUP_MB=101
run l1_g51 RAD_LANES=1 RAD_L2_GROUP_MB=51
