TESTS=1 bash scripts/r02_quick.sh base
RAD_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"process_kernel" -s 1 -c 3 -f -o gpurun_out/r02_prof_process python scripts/prof_batches.py --workload config2 --batches 1 --process-reps 2 2>&1 | tail -1
