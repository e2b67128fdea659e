// Probe: cost of ISSUING tcgen05.mma from one elected lane when the descriptors change every instruction
// (as in an implicit-GEMM conv: new A rows per tap / M tile / k step).  Variants of the address arithmetic.
#include <cstdio>
#include "../mfvi_dip_mia_b200/csrc/tc_ptx.cuh"
using namespace mfvi::tc;

__global__ void __launch_bounds__(128) k_issue(int N, int n_iss, int variant, int stride_a, int stride_b, int n_inner, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, 0);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 128 * 1024);
    const uint32_t hi = desc_hi(1024, kLayoutSw128);
    const long long t0 = clock64();
    if (variant == 0) {            // constant descriptors
      const uint64_t ad = desc_pack(desc_lo(a, 16), hi), bd = desc_pack(desc_lo(b, 16), hi);
      for (int i = 0; i < n_iss; ++i) tc_mma_tf32_elect(tmem, ad, bd, idesc, 1u);
    } else if (variant == 1) {     // lo words advance by run-time strides, packed per instruction
      const uint32_t a_lo = desc_lo(a, 16), b_lo = desc_lo(b, 16);
      uint32_t oa = 0, ob = 0;
      for (int i = 0; i < n_iss; ++i, oa += stride_a, ob += stride_b) {
        if (oa > 5000) oa -= 5000;
        if (ob > 300) ob -= 300;
        tc_mma_tf32_elect(tmem, desc_pack(a_lo + oa, hi), desc_pack(b_lo + ob, hi), idesc, 1u);
      }
    } else if (variant == 2) {     // nested loops like the conv kernel: outer = tap, mid = M tile, inner = k step (x4 unrolled)
      int it = 0;
      uint32_t a_t = desc_lo(a, 16);
      while (it < n_iss) {
        uint32_t d_col = tmem;
        for (int j = 0; j < n_inner; ++j, d_col += N) {
          const uint32_t a_j = a_t + j * 1024u;
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_tf32_elect(d_col, desc_pack(a_j + 2u * k, hi), desc_pack(desc_lo(b, 16) + 2u * k, hi), idesc, 1u);
          it += 4;
        }
        a_t += stride_a;
        if (it % 512 == 0) a_t = desc_lo(a, 16);
      }
    } else if (variant >= 4) {     // narrow K-major rows: 4 = SWIZZLE_64B (16 fp32 rows), 5 = SWIZZLE_32B (8 fp32 rows), 6 = SW128 reference
      const int w = variant == 4 ? 16 : (variant == 5 ? 8 : 32);
      const uint32_t rb = w * 4, hi2 = desc_hi(8 * rb, kmajor_layout(w));
      const uint32_t a_lo = desc_lo(a, 16), b_lo = desc_lo(b, 16);
      uint32_t oa = 0;
      for (int i = 0; i < n_iss; ++i, oa += (rb >> 4)) {
        if (oa > 4000) oa -= 4000;
        tc_mma_tf32_elect(tmem, desc_pack(a_lo + oa, hi2), desc_pack(b_lo, hi2), idesc, 1u);
      }
    } else if (variant == 3) {     // 64-bit descriptors advanced with 64-bit adds
      uint64_t ad = desc_pack(desc_lo(a, 16), hi), bd = desc_pack(desc_lo(b, 16), hi);
      for (int i = 0; i < n_iss; ++i) {
        tc_mma_tf32_elect(tmem, ad, bd, idesc, 1u);
        ad += static_cast<uint64_t>(stride_a);
        bd += static_cast<uint64_t>(stride_b);
        if ((i & 63) == 63) { ad -= 64ull * stride_a; bd -= 64ull * stride_b; }
      }
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

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int n_iss = 1024;
  const char* names[] = {"constant descriptors", "lo += runtime stride", "conv-like nested loops", "64-bit desc += stride", "K-major SWIZZLE_64B rows", "K-major SWIZZLE_32B rows", "K-major SWIZZLE_128B rows"};
  for (int N : {16, 64, 128})
    for (int v = 0; v < 7; ++v) {
      k_issue<<<1, 128, 170 * 1024>>>(N, n_iss, v, 8, 2, 3, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("N=%3d %-24s: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA\n", N, names[v], (double)h[0] / n_iss, (double)h[1] / n_iss);
    }
  return 0;
}
