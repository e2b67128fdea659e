# usage: scale_line.sh N [extra bench args] : one line of the strong-scaling table on N GPUs of the box (torchrun, NCCL)
n=$1; shift
mkdir -p gpurun_out
if [ "$n" = 1 ]; then
  timeout 300 python bench.py --gpus 1 --steps 40 --no-cpu --no-modes "$@" > gpurun_out/r02_scale_den_n1.json 2> gpurun_out/r02_scale_den_n1.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 40 "$@" > gpurun_out/r02_scale_den_n$n.json 2> gpurun_out/r02_scale_den_n$n.err
fi
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_scale_den_n$n.json") if l.startswith("{")][-1])
    print("den N=$n: %.1f steps/s  %.3f ms  e2e %.1f  loss@%d %.6f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["loss_at_step"]["step"], d["config"]["loss_at_step"]["loss"]))
except Exception as e:
    print("den N=$n: FAILED", e)
PY
