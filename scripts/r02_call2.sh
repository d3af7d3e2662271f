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
run t64x32_128 RAD_RASTER=tiles
run t64x32_256 RAD_RASTER=tiles RAD_TILE_THREADS=256
run t128x32_128 RAD_RASTER=tiles RAD_CUDA_LIB=librad_cuda_t128x32.so
run t128x32_256 RAD_RASTER=tiles RAD_CUDA_LIB=librad_cuda_t128x32.so RAD_TILE_THREADS=256
run t64x64_128 RAD_RASTER=tiles RAD_CUDA_LIB=librad_cuda_t64x64.so
run t64x64_256 RAD_RASTER=tiles RAD_CUDA_LIB=librad_cuda_t64x64.so RAD_TILE_THREADS=256
