# usage: knob_bench2.sh "VAR=v [VAR2=v2 ...]" ... : bench.py (tf32, MC=8 and MC=1) under each set of environment knobs
mkdir -p gpurun_out
i=0
for kv in "$@"; do
  i=$((i+1))
  for mc in 8 1; do
    env $kv timeout 200 python bench.py --steps 40 --no-cpu --no-modes --mc $mc > gpurun_out/knob2_${i}_mc${mc}.json 2> gpurun_out/knob2_${i}_mc${mc}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/knob2_${i}_mc${mc}.json"))
    print("[${kv}] mc=${mc}: %.1f steps/s  %.3f ms" % (d["value"], d["ms_per_step"]))
except Exception as e:
    print("[${kv}] mc=${mc}: FAILED", e)
PY
  done
done
