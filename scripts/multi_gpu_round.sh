# One 8-GPU box: scaling lines of the metric config (N = 1, 2, 4, 8), config 3 (inpainting MC=8) at N = 8, config 5 (BO sweep)
mkdir -p gpurun_out
run() { # N, extra args, tag
  local n=$1; shift; local tag=$1; shift
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 40 --no-cpu --no-modes "$@" > gpurun_out/r02_scale_${tag}_n1.json 2> gpurun_out/r02_scale_${tag}_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 40 "$@" > gpurun_out/r02_scale_${tag}_n$n.json 2> gpurun_out/r02_scale_${tag}_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_scale_${tag}_n$n.json") if l.startswith("{")][-1])
    print("${tag} N=$n: %.1f steps/s  %.3f ms  e2e %.1f  loss@%d %.6f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["loss_at_step"]["step"], d["config"]["loss_at_step"]["loss"]))
except Exception as e:
    print("${tag} N=$n: FAILED", e)
PY
}
if [ "$1" = sweep ]; then run 8 den; else
for n in 1 2 4 8; do run $n den; done
for n in 1 8; do run $n inp --config inp; done
fi
timeout 300 python -m pytest tests/test_gpu_trials.py -q -x 2>&1 | tail -2
timeout 900 python scripts/run_bo_sweep.py --grid 8 --num-iter 300 --size 256 --out gpurun_out/r02_bo_sweep.json > gpurun_out/r02_bo_sweep.log 2>&1; tail -1 gpurun_out/r02_bo_sweep.log | cut -c1-400
