mkdir -p gpurun_out
for mf in -1 4 3 2; do
  for mc in 8 1; do
    MFVI_MEGA_FROM=$mf timeout 200 python bench.py --steps 30 --no-cpu --no-modes --mc $mc > gpurun_out/r2_mega_bench_${mf}_mc${mc}.json 2> gpurun_out/r2_mega_bench_${mf}_mc${mc}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_mega_bench_${mf}_mc${mc}.json"))
    k=d["kernels"].get("mfvi_mega_run",{})
    print("mega_from=${mf} mc=${mc}: %.1f steps/s  %.3f ms  launches/step %d  mega(eager) %s ms  loss %.4f" % (d["value"], d["ms_per_step"], d["launches_per_step"], k.get("ms"), d["config"]["loss_at_step"]["loss"]))
except Exception as e:
    print("mega_from=${mf} mc=${mc}: FAILED", e)
PY
  done
done
for mf in -1 3; do
  MFVI_MEGA_FROM=$mf timeout 200 python bench.py --steps 20 --no-cpu --no-modes --math fp32 > gpurun_out/r2_mega_bench_fp32_${mf}.json 2> gpurun_out/r2_mega_bench_fp32_${mf}.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_mega_bench_fp32_${mf}.json')); print('fp32 mega_from=${mf}: %.1f steps/s  launches %d' % (d['value'], d['launches_per_step']))"
done
