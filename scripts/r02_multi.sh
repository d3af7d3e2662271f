# N-GPU pass (gpurun --gpus N): multi-GPU tests, then bench.py under torchrun (ours + the reference arm)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BARGS:-} > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo rc=$?; tail -5 gpurun_out/r02_bench_n$N.err
python scripts/show_multi.py gpurun_out/r02_bench_n$N.json
