# N-GPU pass (gpurun --gpus N): multi-GPU tests, then bench.py under torchrun (ours + the reference arm)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BARGS:-} > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo rc=$?; tail -5 gpurun_out/r02_bench_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/r02_bench_n{n}.json"))
    print("N",n,"value",round(d["value"],1),"e2e",round(d["e2e"]["value"],1),"ms/step",round(d["ms_per_step"],3), d["run"]["parallelism"])
    print(" multichip_check", d.get("multichip_check"))
    for k in ("strong_config3","strong_config4"):
        if k in d: print(" ",k,{a:(round(b,4) if isinstance(b,float) else b) for a,b in d[k].items() if a!="workload"})
    print(" kernels", {k.split(" ")[0]:round(v["ms_per_batch"],4) for k,v in d["kernels"].items()})
except Exception as e: print("ERR",e)
PY
