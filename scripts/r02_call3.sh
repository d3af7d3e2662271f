mkdir -p gpurun_out
RAD_LANES=1 RAD_RASTER=tiles ncu --set full --clock-control none --import-source on -k regex:"tile_kernel|bin_" -s 4 -c 4 -f -o gpurun_out/r02_prof_tiles python scripts/prof_batches.py --workload config2 --batches 2 2>&1 | tail -1
RAD_LANES=1 ncu --set full --clock-control none --import-source on -k regex:"raster_queue" -s 1 -c 1 -f -o gpurun_out/r02_prof_queue python scripts/prof_batches.py --workload config2 --batches 2 2>&1 | tail -1
