# full measurement pass for profiles/: tests, bench lines, ncu launch lists, ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; tail -2 gpurun_out/bench_config2.err
timeout 300 python bench.py --workload config2_k1 --steps 5 --no-cpu-baseline > gpurun_out/bench_config2_k1.json 2> gpurun_out/bench_config2_k1.err
timeout 300 python bench.py --workload config3 --steps 5 --no-cpu-baseline > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# launch lists: lanes off so that the kernels of one batch appear in pipeline order (ncu serialises them anyway)
RAD_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_config2.csv python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
RAD_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_config2_k1.csv python scripts/prof_batches.py --workload config2_k1 --batches 40 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_config2_lanes.csv python scripts/prof_batches.py --workload config2 --batches 2 2>&1 | tail -1
RAD_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"raster_|process_kernel|topk|apply|camera" -s 7 -c 7 -f -o gpurun_out/prof_config2 python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:"process_kernel" -s 2 -c 2 -f -o gpurun_out/prof_process python scripts/prof_batches.py --workload config2 --batches 1 --process-reps 3 2>&1 | tail -1
python - <<'PY'
import json
for f in ("bench_config2","bench_config2_k1","bench_config3","bench_reference"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("cpu_baseline"))
        if "kernels" in d: print("  ", {k:(round(v["ms_per_batch"],4),round(v["share"],3)) for k,v in d["kernels"].items()}, "K2 Gpix/s", round(d["process_hemicube"]["gpix_per_s"],1), round(d["process_hemicube"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
PY
