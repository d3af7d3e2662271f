# round-2 final measurement pass for profiles/ (one B200): tests, smoke, bench lines, ncu launch lists, ncu --set full of one
# batch, L2 atomic counters of the hot kernels, config-5 resolution sweep, ring / tile A-B lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_config2.json 2> gpurun_out/r02_bench_config2.err; echo bench rc=$?; tail -2 gpurun_out/r02_bench_config2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
RAD_RING=1 timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_config2_ring.json 2>/dev/null
timeout 120 python scripts/fuzz_parity.py --seconds 18 > gpurun_out/r02_fuzz_parity.json 2>/dev/null; tail -c 400 gpurun_out/r02_fuzz_parity.json; echo
timeout 300 python scripts/sweep_config5.py > gpurun_out/r02_config5_resolution_sweep.json 2>/dev/null
RAD_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config2.csv python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config2_lanes.csv python scripts/prof_batches.py --workload config2 --batches 2 2>&1 | tail -1
RAD_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_config2_k1.csv python scripts/prof_batches.py --workload config2_k1 --batches 40 2>&1 | tail -1
RAD_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"raster_|process_kernel|topk|apply|camera" -s 7 -c 7 -f -o gpurun_out/r02_prof_config2 python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
RAD_LANES=1 ncu --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"raster_queue|process_kernel|raster_setup|raster_cull" -s 4 -c 4 --csv --log-file gpurun_out/r02_l2_atomics.csv python scripts/prof_batches.py --workload config2 --batches 2 --process-reps 2 2>&1 | tail -1
python - <<'PY'
import json
def last(p): return [json.loads(l) for l in open(p) if l.startswith('{')][-1]
try:
    d=last("gpurun_out/r02_bench_config2.json")
    print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],3), d["clocks"])
    print({k.split(" ")[0]:(round(v["ms_per_batch"],4),round(v["share"],3)) for k,v in d["kernels"].items()})
    print("roofline", {a:b for a,b in d["roofline"].items() if a not in ("note","traffic_source","peak_source")})
    print("K2", {a:b for a,b in d["process_hemicube"].items() if a not in ("note",)})
    for s in ("k1","reference_schedule_k64","config3"): print(s, {a:b for a,b in d.get(s,{}).items() if a!="workload"})
    print("cpu", d.get("cpu_baseline"))
    r=last("gpurun_out/r02_bench_reference_arm.json"); print("reference arm", round(r["value"],1), r["cpu_baseline"])
    g=last("gpurun_out/r02_bench_config2_ring.json"); print("ring", round(g["value"],1), {k.split(" ")[0]:round(v["ms_per_batch"],4) for k,v in g["kernels"].items()})
except Exception as e: print("ERR", e)
PY
