// Probe: cycles per tcgen05.mma.kind::tf32 instruction (SS mode, operands in shared memory) as a function of M and N.
// Issues `n_iss` back-to-back MMAs on the same operands from one elected lane and waits for the commit.
//   umma_rate_test          kind::tf32 (K = 8 per instruction)
//   umma_rate_test bf16     kind::f16 with bf16 operands (K = 16 per instruction, the same 32 bytes per operand row), K-major and
//                           MN-major (plain SWIZZLE_128B, 8-row k groups) — the numbers the bf16-operand mode is modelled on
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../mfvi_dip_mia_b200/csrc/tc_ptx.cuh"
using namespace mfvi::tc;

__global__ void __launch_bounds__(128) k_rate(int M, int N, int n_iss, int mn_major, int distinct_acc, int a_mode, long long* out,
                                              int bf16 = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(M, N, mn_major, mn_major, bf16 ? kFmtBF16 : kFmtTF32);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 48 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < n_iss; ++i) {
      const uint32_t k = (i & 3);
      // a_mode 0: same rows always; 1: new row window every 4 MMAs (k cycles inside); 2: new row window every MMA; 3: window shifts by 1 row every 4 MMAs
      uint32_t a_off = 0;
      if (a_mode == 1) a_off = ((i >> 2) * 37u % 200u) * 128u;
      if (a_mode == 2) a_off = (i * 37u % 200u) * 128u;
      if (a_mode == 3) a_off = ((i >> 2) % 200u) * 128u;
      // MN-major: one MMA reads 8 (tf32) / 16 (bf16) k rows of 128 bytes; blocks of 32 / 64 channels 4096 bytes apart
      const uint64_t ad = mn_major ? (bf16 ? make_desc(a + a_off + (k & 1) * 2048, 4096, 1024, kLayoutSw128)
                                           : make_desc(a + a_off + k * 1024, 4096, 512, kLayoutSw128Base32))
                                   : make_desc(a + a_off + k * 32, 16, 1024);
      const uint64_t bd = mn_major ? (bf16 ? make_desc(b + (k & 1) * 2048, 4096, 1024, kLayoutSw128)
                                           : make_desc(b + k * 1024, 4096, 512, kLayoutSw128Base32))
                                   : make_desc(b + k * 32, 16, 1024);
      const uint32_t d = tmem + (distinct_acc ? ((i & 1) * 256u) : 0u);
      if (bf16) tc_mma_f16_elect(d, ad, bd, idesc, 1u);
      else tc_mma_tf32_elect(d, ad, bd, idesc, 1u);
    }
    const long long t1 = clock64();
    tc_commit_elect(smem_u32(bar));
    mbar_wait(smem_u32(bar), 0);
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main(int argc, char** argv) {
  const int bf16 = (argc > 1 && std::string(argv[1]) == "bf16") ? 1 : 0;
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n_iss = 512;
  for (int mn = 0; mn < (bf16 ? 2 : 1); ++mn)
    for (int M : {128})
      for (int N : {16, 32, 64, 128, 192, 256}) {
        if (mn && N % 64) continue;                     // MN-major atoms: 64 channels per 128-byte row (bf16)
        for (int dist : {0, 1, 2, 3}) {
          k_rate<<<1, 128, 100 * 1024>>>(M, N, n_iss, mn, 0, dist, d, bf16);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s (M=%d N=%d)\n", cudaGetErrorString(e), M, N); return 1; }
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%s %s M=%3d N=%3d %s: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA\n", bf16 ? "bf16" : "tf32", mn ? "MN-major" : "K-major ", M, N,
                 dist == 0 ? "A fixed         " : dist == 1 ? "A new per 4 MMAs" : dist == 2 ? "A new per MMA   " : "A +1 row per 4  ", (double)h[0] / n_iss, (double)h[1] / n_iss);
        }
      }
  return 0;
}
