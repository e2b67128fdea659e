// Shared device/host helpers for libmfvidip (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mfvi_dip.h"

namespace mfvi {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define MFVI_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      mfvi::set_error(__VA_ARGS__);  \
      return 1;                      \
    }                                \
  } while (0)

static inline cudaStream_t as_stream(mfvi_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: every hot-path kernel is launched with programmatic stream serialisation, calls
// pdl_trigger() first thing (the next kernel of the stream may start its prologue: block scheduling, barrier init, TMEM
// allocation, shared-memory zero fill) and pdl_wait() before it touches global memory (returns once the preceding kernel
// has completed and flushed).  A step is ~170 short kernels, so hiding launch latency / prologues matters.  MFVI_PDL=0
// switches the attribute off (the device-side instructions are then no-ops).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();

// Planning-only mode (mfvi_conv2d_plan): the host side of a conv launch runs unchanged — shape checks, dispatch, tile
// planning, argument setup — but tensor maps are not encoded, function attributes are not set and nothing is launched;
// the launch geometry is recorded instead.  It lets the dispatch table and the tile plans of every layer be inspected and
// tested on a machine without a GPU.  dry_run() is thread-local and null outside mfvi_conv2d_plan.
struct DryRunInfo {
  int launches;
  unsigned grid[3], block;
  size_t smem;
  char detail[200];
};
DryRunInfo* dry_run();
void set_dry_run(DryRunInfo* info);
void dry_note(dim3 grid, dim3 block, size_t smem);
void dry_detail(const char* fmt, ...);

template <typename... P, typename... A>
static inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  if (dry_run() != nullptr) {
    dry_note(grid, block, smem);
    return cudaSuccess;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// cudaFuncSetAttribute applies to the CURRENT device only: `done_mask` (a static at the call site, one per kernel) remembers the
// devices that already have it, so that a process driving several GPUs opts every one of them in.
template <typename K>
static inline cudaError_t allow_dyn_smem(K kernel, int bytes, unsigned long long* done_mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*done_mask & bit) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) *done_mask |= bit;
  return e;
}

constexpr int kNumSMs = 148;  // B200
constexpr float kBnEps = 1e-5f;
constexpr float kLreluSlope = 0.2f;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (bit-exact with oracle/philox.py) + Box-Muller
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
  const float u1 = (static_cast<float>(xa >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
  const float u2 = (static_cast<float>(xb >> 8) + 0.5f) * 5.9604644775390625e-08f;
  // MUFU-based log / sin / cos: drawing the S*P normals of a step is otherwise instruction-bound (~30 us per pass with
  // logf / sincospif vs the ~8 us its HBM traffic takes).  The angle is shifted into [-pi, pi), where the fast sine and
  // cosine are accurate to ~4e-7 absolute: cos(2 pi u) = -cos(2 pi u - pi).  |z - z_fp64| stays below ~1e-5
  // (tests/test_gpu_parity.py::test_philox_bit_exact_and_normals); the uniform bits themselves are bit-exact.
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(fmaf(u2, 6.283185307179586f, -3.141592653589793f), &sn, &cs);
  sn = -sn;
  cs = -cs;
  z0 = r * cs;
  z1 = r * sn;
}

// 4 standard normals of Philox block `blk` of stream (stream_id, sample, step).
__device__ __forceinline__ float4 philox_normal4(uint32_t blk, uint32_t stream_id, uint32_t sample, uint32_t step,
                                                 uint64_t seed) {
  const uint4 x = philox4x32_10(blk, stream_id, sample, step, static_cast<uint32_t>(seed),
                                static_cast<uint32_t>(seed >> 32));
  float4 z;
  box_muller(x.x, x.y, z.x, z.y);
  box_muller(x.z, x.w, z.z, z.w);
  return z;
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reflect index q in [-p, n+p) into [0,n)  (nn.ReflectionPad2d: edge not repeated)
__device__ __forceinline__ int reflect_idx(int q, int n) {
  q = q < 0 ? -q : q;
  return q >= n ? 2 * (n - 1) - q : q;
}

// BatchNorm statistics from (sum, sumsq) accumulated in double.
__device__ __forceinline__ void bn_mean_invstd(const double* __restrict__ sums, double inv_count, float& mean,
                                               float& invstd) {
  const double m = sums[0] * inv_count;
  double var = sums[1] * inv_count - m * m;
  var = var < 0.0 ? 0.0 : var;
  mean = static_cast<float>(m);
  invstd = static_cast<float>(rsqrt(var + static_cast<double>(kBnEps)));
}

__device__ __forceinline__ size_t view_off(const MfviView& v, int s, int h, int w) {
  return static_cast<size_t>(s) * v.sstride + static_cast<size_t>(h) * v.hstride + static_cast<size_t>(w) * v.wstride;
}

}  // namespace mfvi
