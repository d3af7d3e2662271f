# A/B on one 8-GPU box: bench lines (no extras) for each "NAME=VALUE[,..]" argument ("base" = no knob)
mkdir -p gpurun_out
for v in "$@"; do
  tag=$(echo $v | tr '=,' '__')
  if [ "$v" = "base" ]; then envs="RAD_X=0"; else envs=$(echo $v | tr ',' ' '); fi
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-8} --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus ${N:-8} --steps ${STEPS:-4} --warmup 3 --no-extras ${BARGS:-} 2>/dev/null | grep "^{" > gpurun_out/ab8_$tag.json
  python scripts/show_multi.py gpurun_out/ab8_$tag.json | grep -v multichip
done
