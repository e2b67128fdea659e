# usage: knob_bench_mc1.sh "VAR=v ..." ... : bench.py (tf32, MC=1 only; the per-GPU load of the 8-GPU line) under each knob set
mkdir -p gpurun_out
i=0
for kv in "$@"; do
  i=$((i+1))
  env $kv timeout 200 python bench.py --steps 40 --no-cpu --no-modes --mc 1 > gpurun_out/knob1_${i}.json 2> gpurun_out/knob1_${i}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/knob1_${i}.json"))
    print("[${kv}] mc=1: %.1f steps/s  %.3f ms" % (d["value"], d["ms_per_step"]))
except Exception as e:
    print("[${kv}] mc=1: FAILED", e)
PY
done
