# single-GPU diagnostics: k = 1 kernel table, config 3 / config 4 kernel tables (+ set-up knobs)
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith('{')][-1]
r=[v for k,v in d["kernels"].items() if k.startswith("raster")][0]
print(sys.argv[1].split("/")[-1], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k.split(" ")[0]:round(v["ms_per_batch"],4) for k,v in d["kernels"].items()}, "setup/queue", round(r["setup_ms"],4), round(r.get("queue_ms", r.get("ring_ms", 0)),4))
PY
}
timeout 300 python bench.py --workload config2_k1 --steps 3 --no-extras --no-cpu-baseline > gpurun_out/s_k1.json 2>/dev/null; show gpurun_out/s_k1.json
for v in base RAD_SETUP_MINB=4 RAD_SETUP_MINB=5 RAD_SETUP_CTAS=4 RAD_SETUP_CTAS=6; do
  tag=$(echo $v | tr '=,' '__'); if [ "$v" = "base" ]; then envs="RAD_X=0"; else envs=$(echo $v | tr ',' ' '); fi
  env $envs timeout 300 python bench.py --workload config3 --steps 3 --no-extras --no-cpu-baseline > gpurun_out/s_c3_$tag.json 2>/dev/null; show gpurun_out/s_c3_$tag.json
done
timeout 300 python bench.py --workload config4 --steps 2 --no-extras --no-cpu-baseline > gpurun_out/s_c4.json 2>/dev/null; show gpurun_out/s_c4.json
