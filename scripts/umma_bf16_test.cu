// Probe for the bf16-operand mode: are the kind::f16 / bf16 shared-memory descriptors right?
// Two tiny GEMMs against a host reference, operands written by TMA exactly as the kernels write them:
//   mode 0  K-major A and B, SWIZZLE_128B, 128-byte rows = 64 bf16 channels, four 32-byte k-steps per row
//           (forward / data-gradient activations and forward weights; byte-identical to the tf32 layout)
//           D[128 x 16] = A[shift + m][0..63] * B[n][0..63]^T
//   mode 1  MN-major A and B, plain SWIZZLE_128B with 64-element atoms, 8-row k groups (SBO 1024), 16 k rows per MMA,
//           64-channel blocks LBO apart (data-gradient weights, both weight-gradient operands)
//           D[128 x 64] = sum_{k<128} A[k][m] * B[shift + k][n]
// `shift` also starts the shifted operand at rows that are not multiples of the 8-row swizzle atom (needed only by a future
// tap-aliasing weight gradient; shift 0 is what stages A / B use).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_bf16_test umma_bf16_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../mfvi_dip_mia_b200/csrc/tc_ptx.cuh"

namespace mfvi { DryRunInfo* dry_run() { return nullptr; } }      // tc_ptx.cuh's planning-only hook (not used here)
using namespace mfvi::tc;

constexpr int R = 512;   // rows of the big operand region

static CUtensorMap map2d(const __nv_bfloat16* base, int rows, int box_rows) {
  CUtensorMap m;
  const uint64_t dims[2] = {64, static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {128};
  const uint32_t box[2] = {64, static_cast<uint32_t>(box_rows)};
  if (!tma_encode(&m, base, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) {
    printf("tensor map encode failed\n");
    exit(1);
  }
  return m;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap tmBig, const __grid_constant__ CUtensorMap tmSmall,
                                               int mode, int shift, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* big = smem;                       // R rows x 128 B
  uint8_t* small = smem + R * 128;           // K-major B: 16 rows; MN-major A: 2 blocks x 128 rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(small + 2 * 128 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init(smem_u32(mma_bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int N = mode == 0 ? 16 : 64;
  if (threadIdx.x == 0) {
    const uint32_t small_bytes = mode == 0 ? 16 * 128 : 2 * 128 * 128;
    mbar_expect_tx(smem_u32(bar), R * 128 + small_bytes);
    for (int r0 = 0; r0 < R; r0 += 256) tma_load_2d(smem_u32(big + r0 * 128), &tmBig, smem_u32(bar), 0, r0);
    if (mode == 0) {
      tma_load_2d(smem_u32(small), &tmSmall, smem_u32(bar), 0, 0);
    } else {
      for (int j = 0; j < 2; ++j) tma_load_2d(smem_u32(small + j * 128 * 128), &tmSmall, smem_u32(bar), 0, j * 128);
    }
  }
  __syncthreads();
  mbar_wait(smem_u32(bar), 0);          // every thread observes the TMA completion (whichever lane gets elected issues the MMAs)
  tc_fence_after();
  if (warp == 0) {
    if (mode == 0) {
      const uint32_t idesc = make_idesc(128, N, 0, 0, kFmtBF16);
      for (int k = 0; k < 4; ++k) {          // 4 k-steps of 16 channels = 32 bytes each
        const uint64_t ad = make_desc(smem_u32(big) + shift * 128 + k * 32, 16, 1024);
        const uint64_t bd = make_desc(smem_u32(small) + k * 32, 16, 1024);
        tc_mma_f16_elect(tmem, ad, bd, idesc, k > 0);
      }
    } else {
      const uint32_t idesc = make_idesc(128, N, 1, 1, kFmtBF16);
      for (int k = 0; k < 8; ++k) {          // 8 k-steps of 16 rows = 2048 bytes each
        const uint64_t ad = make_desc(smem_u32(small) + k * 2048, 128 * 128, 1024, kLayoutSw128);
        const uint64_t bd = make_desc(smem_u32(big) + (shift + 16 * k) * 128, 128 * 128, 1024, kLayoutSw128);
        tc_mma_f16_elect(tmem, ad, bd, idesc, k > 0);
      }
    }
    tc_commit_elect(smem_u32(mma_bar));
  }
  __syncthreads();
  mbar_wait(smem_u32(mma_bar), 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16) + c, v);
    for (int j = 0; j < 16; ++j) out[((threadIdx.x >> 5) * 32 + lane) * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  // small integers: exact in bf16, products and sums exact in fp32
  std::vector<float> big(R * 64), smallK(16 * 64), smallMN(2 * 128 * 64);
  for (int r = 0; r < R; ++r) for (int k = 0; k < 64; ++k) big[r * 64 + k] = static_cast<float>((r * 7 + k * 3) % 15) - 7.f;
  for (int n = 0; n < 16; ++n) for (int k = 0; k < 64; ++k) smallK[n * 64 + k] = static_cast<float>((n * 5 + k * 11) % 13) - 6.f;
  for (int i = 0; i < 2 * 128 * 64; ++i) smallMN[i] = static_cast<float>((i * 13 + (i >> 6) * 3) % 11) - 5.f;
  auto upload = [](const std::vector<float>& h) {
    std::vector<__nv_bfloat16> b(h.size());
    for (size_t i = 0; i < h.size(); ++i) b[i] = __float2bfloat16(h[i]);
    __nv_bfloat16* d;
    cudaMalloc(&d, b.size() * 2);
    cudaMemcpy(d, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    return d;
  };
  __nv_bfloat16 *dBig = upload(big), *dK = upload(smallK), *dMN = upload(smallMN);
  float* dOut;
  cudaMalloc(&dOut, 128 * 64 * 4);
  const size_t smem = 1024 + R * 128 + 2 * 128 * 128 + 64;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shifts[] = {0, 8, 16, 1, 2, 3, 5, 9, 66, 131, 258};
  int bad = 0;
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tmBig = map2d(dBig, R, 256);
    CUtensorMap tmSmall = mode == 0 ? map2d(dK, 16, 16) : map2d(dMN, 2 * 128, 128);
    const int N = mode == 0 ? 16 : 64;
    for (int shift : shifts) {
      cudaMemset(dOut, 0, 128 * 64 * 4);
      k_probe<<<1, 128, smem>>>(tmBig, tmSmall, mode, shift, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      std::vector<float> out(128 * N);
      cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          if (mode == 0) {
            for (int k = 0; k < 64; ++k) ref += (double)big[(shift + m) * 64 + k] * smallK[n * 64 + k];
          } else {      // A[k][m]: block m / 64, row k, channel m % 64
            for (int k = 0; k < 128; ++k) ref += (double)smallMN[((m / 64) * 128 + k) * 64 + (m % 64)] * big[(shift + k) * 64 + n];
          }
          const double err = fabs(ref - out[m * N + n]);
          if (err > maxerr) maxerr = err;
        }
      const bool ok = maxerr < 1e-3;
      bad += (!ok && shift % 8 == 0);
      printf("bf16 %s shift %3d : max abs err %g %s\n", mode == 0 ? "K-major " : "MN-major", shift, maxerr,
             ok ? "OK" : (shift % 8 ? "MISMATCH (unaligned start: matters only for tap aliasing)" : "MISMATCH"));
    }
  }
  printf(bad ? "RESULT: descriptor layout WRONG for an aligned start - fix tc_ptx / the kernels before anything else\n"
             : "RESULT: aligned starts OK\n");
  return bad ? 2 : 0;
}
