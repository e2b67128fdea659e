#!/usr/bin/env bash
# Builds libmfvidip.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="basic_kernels.cu elementwise.cu conv_simt.cu conv_dispatch.cu radon.cu"
SRCS="$SRCS conv_tc.cu conv_tc2.cu conv_wgrad2.cu bookkeeping.cu lrt.cu conv_pointwise.cu mega.cu"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC -shared -o libmfvidip.so $SRCS "$@"
echo "built $(pwd)/libmfvidip.so"
