# round 2, GPU call 1: design data for K1 (atomics micro-benchmark), baseline re-measure, existing knobs A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
./scripts/ubench/atomics_ubench > gpurun_out/r02_atomics_ubench.jsonl 2>&1; tail -40 gpurun_out/r02_atomics_ubench.jsonl
bash scripts/gpu_quick.sh RAD_QUEUE_PREFETCH=1 RAD_RASTER=tiles RAD_L2_GROUP_MB=96
