#!/usr/bin/env bash
# Bring-up of the bf16-operand mode on a B200 box, simplest stage first; every stage writes its log
# under gpurun_out/ and a failing stage does not stop the later ones (they are independent kernels).
#   gpurun --timeout 900 -- 'bash scripts/bringup_bf16.sh'
set -u
mkdir -p gpurun_out
export MFVI_TEST_NEXT=1
# 0. descriptor semantics of kind::f16 / bf16 operands in isolation (seconds): if this fails, fix tc_ptx / the constants first
(cd scripts && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_bf16_test umma_bf16_test.cu -lcuda 2>/dev/null; \
 timeout 60 ./umma_bf16_test > ../gpurun_out/bf16_descriptor_probe.txt 2>&1; echo "descriptor probe rc=$? $(tail -1 ../gpurun_out/bf16_descriptor_probe.txt)")
run() { # name, pytest -k expression
  timeout 300 python -m pytest tests/test_gpu_next_bf16.py -m gpu_next -q -x -k "$2" > "gpurun_out/bf16_$1.txt" 2>&1
  echo "stage $1: rc=$? $(tail -1 gpurun_out/bf16_$1.txt)"
}
run C_elementwise "bn_act_pad or bn_bwd_apply or conversions"
run A_conv "conv_equals or accumulates or rejects"
run B_wgrad "wgrad"
run D_engine "engine_step or trainer"
run G_acceptance "reference_ensemble"      # 32 seeds x 1200 iterations in bf16 against the reference's distribution (~1 minute)
# independent of bf16: the BatchNorm backward without its intermediate buffer (exact fp32 arithmetic) — kernel test, then the
# whole verified suite and the bench with the switch on
run F_fused_bn_bwd "fused_bn_backward"
MFVI_FUSED_BN_BWD=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/fused_bn_bwd_suite.txt 2>&1
echo "verified suite with MFVI_FUSED_BN_BWD=1: rc=$? $(tail -1 gpurun_out/fused_bn_bwd_suite.txt)"
MFVI_FUSED_BN_BWD=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/fused_bn_bwd_bench.json 2> gpurun_out/fused_bn_bwd_bench.err
echo "bench with MFVI_FUSED_BN_BWD=1 rc=$? $(head -c 200 gpurun_out/fused_bn_bwd_bench.json)"
# only meaningful once the four stages are green
timeout 300 python bench.py --math bf16 --steps 30 --warmup 5 --no-cpu > gpurun_out/bf16_bench.json 2> gpurun_out/bf16_bench.err
echo "bench rc=$? $(head -c 300 gpurun_out/bf16_bench.json)"
# issue rate of kind::f16 bf16 MMAs vs N (K-major and MN-major): the numbers the mode's cost models need
(cd scripts && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate_test umma_rate_test.cu -lcuda 2>/dev/null; \
 timeout 60 ./umma_rate_test bf16 > ../gpurun_out/bf16_umma_rate_probe.txt 2>&1; echo "rate probe rc=$?")
