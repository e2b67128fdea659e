# usage: knob_bench.sh VAR v1 v2 ... : bench.py (tf32, MC=8 and MC=1) for each value of an environment knob
mkdir -p gpurun_out
VAR=$1; shift
for v in "$@"; do
  for mc in 8 1; do
    env $VAR=$v timeout 200 python bench.py --steps 40 --no-cpu --no-modes --mc $mc > gpurun_out/knob_${VAR}_${v}_mc${mc}.json 2> gpurun_out/knob_${VAR}_${v}_mc${mc}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/knob_${VAR}_${v}_mc${mc}.json"))
    k=d["kernels"]
    print("${VAR}=${v} mc=${mc}: %.1f steps/s  %.3f ms | eager ms: fwd %.3f dgrad %.3f wgrad %.3f" % (d["value"], d["ms_per_step"], k["mfvi_conv2d_fwd"]["ms"], k["mfvi_conv2d_dgrad"]["ms"], k["mfvi_conv2d_wgrad"]["ms"]))
except Exception as e:
    print("${VAR}=${v} mc=${mc}: FAILED", e)
PY
  done
done
