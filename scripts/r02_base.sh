# round-2 baseline pass: GPU tests, smoke, the default bench line, launch lists, ncu --set full of one batch
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_config2.json 2> gpurun_out/r02_bench_config2.err; echo bench rc=$?; tail -3 gpurun_out/r02_bench_config2.err
RAD_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config2.csv python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
RAD_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"raster_|process_kernel|topk|apply|camera" -s 7 -c 7 -f -o gpurun_out/r02_prof_config2 python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02_bench_config2.json"))
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],3), d["clocks"])
    print({k.split(" ")[0]:(round(v["ms_per_batch"],4),round(v["share"],3)) for k,v in d["kernels"].items()})
    r=d["kernels"]["raster (K1: raster_setup + raster_queue)"]; print("setup/queue", r["setup_ms"], r["queue_ms"])
    print("K2", {a:b for a,b in d["process_hemicube"].items() if a not in ("note",)})
    for s in ("k1","reference_schedule_k64","config3"): print(s, {a:b for a,b in d.get(s,{}).items() if a!="workload"})
    print("cpu", d.get("cpu_baseline"))
except Exception as e: print("ERR", e)
PY
