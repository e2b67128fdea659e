// Probe: may a UMMA shared-memory descriptor start at an arbitrary 128-byte row of a TMA-written swizzled region
// (row shift not a multiple of the 8-row swizzle atom)?  Decides whether a halo tile loaded ONCE can feed all
// conv taps by shifting descriptor start addresses.  Tests K-major SW128 (A operand of fwd/dgrad) and MN-major
// SW128/32B-atom (B operand of wgrad), with descriptor base_offset = 0 and = (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_test umma_shift_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mfvi_dip_mia_b200/csrc/tc_ptx.cuh"
#include <cuda_runtime.h>

using namespace mfvi::tc;

constexpr int R = 512;   // rows in the big operand region

static PFN_cuTensorMapEncodeTiled get_encode() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
}

static CUtensorMap map2d(const float* base, int rows, int box_rows, bool atom32) {
  CUtensorMap m;
  const uint64_t dims[2] = {32, static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {128};
  const uint32_t box[2] = {32, static_cast<uint32_t>(box_rows)};
  const uint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// mode 0: K-major.  D[128 x 16] = A[shift + m][0..31] * B[n][0..31]^T
// mode 1: MN-major. D[128 x 32] = sum_{k<128} A[k][m] * B[shift + k][n]   (A rows beyond 32 channels: blocks 1..3 reuse other rows)
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap tmBig, const __grid_constant__ CUtensorMap tmSmall,
                                               int mode, int shift, int use_bo, int pitch_rows, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* big = smem;                       // R rows x 128 B
  uint8_t* small = smem + R * 128;           // K-major B: 16 rows; MN-major A: 4 blocks x 128 rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(small + 4 * 128 * 128);
  uint64_t* mma_bar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_init(smem_u32(mma_bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const int N = mode == 0 ? 16 : (mode == 2 ? 96 : 32);
  if (threadIdx.x == 0) {
    const uint32_t small_bytes = mode == 0 ? 16 * 128 : 4 * 128 * 128;
    mbar_expect_tx(smem_u32(bar), R * 128 + small_bytes);
    for (int r0 = 0; r0 < R; r0 += 256) tma_load_2d(smem_u32(big + r0 * 128), &tmBig, smem_u32(bar), 0, r0);
    if (mode == 0) {
      tma_load_2d(smem_u32(small), &tmSmall, smem_u32(bar), 0, 0);
    } else {
      for (int j = 0; j < 4; ++j) tma_load_2d(smem_u32(small + j * 128 * 128), &tmSmall, smem_u32(bar), 0, j * 128);
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    if (mode == 0) {
      const uint32_t idesc = make_idesc(128, N, 0, 0);
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_addr = smem_u32(big) + shift * 128 + k * 32;
        uint64_t ad = make_desc(a_addr, 16, 1024);
        if (use_bo) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
        const uint64_t bd = make_desc(smem_u32(small) + k * 32, 16, 1024);
        tc_mma_tf32(tmem, ad, bd, idesc, k > 0);
      }
    } else if (mode == 2) {
      // aliasing: M blocks = A shifted by pitch_rows rows each (LBO = pitch_rows*128), N blocks = B shifted by 1 row each (LBO = 128)
      const uint32_t idesc = make_idesc(128, 96, 1, 1);
      for (int k = 0; k < 8; ++k) {
        const uint64_t ad = make_desc(smem_u32(small) + (shift + 8 * k) * 128, pitch_rows * 128, 512, kLayoutSw128Base32);
        const uint64_t bd = make_desc(smem_u32(big) + (8 * k) * 128, 128, 512, kLayoutSw128Base32);
        tc_mma_tf32(tmem, ad, bd, idesc, k > 0);
      }
    } else {
      const uint32_t idesc = make_idesc(128, N, 1, 1);
      for (int k = 0; k < 16; ++k) {
        const uint64_t ad = make_desc(smem_u32(small) + k * 1024, 128 * 128, 512, kLayoutSw128Base32);
        const uint32_t b_addr = smem_u32(big) + (shift + 8 * k) * 128;
        uint64_t bd = make_desc(b_addr, 128 * 128, 512, kLayoutSw128Base32);
        if (use_bo == 1) bd |= static_cast<uint64_t>((b_addr >> 7) & 7) << 49;
        if (use_bo == 2) bd |= static_cast<uint64_t>((b_addr >> 7) & 3) << 49;
        tc_mma_tf32(tmem, ad, bd, idesc, k > 0);
      }
    }
    tc_commit(smem_u32(mma_bar));
  }
  __syncthreads();
  mbar_wait(smem_u32(mma_bar), 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

int main() {
  std::vector<float> big(R * 32), smallK(16 * 32), smallMN(4 * 128 * 32);
  for (int r = 0; r < R; ++r) for (int k = 0; k < 32; ++k) big[r * 32 + k] = static_cast<float>((r * 7 + k * 3) % 31) - 15.f;
  for (int n = 0; n < 16; ++n) for (int k = 0; k < 32; ++k) smallK[n * 32 + k] = static_cast<float>((n * 5 + k * 11) % 13) - 6.f;
  for (int i = 0; i < 4 * 128 * 32; ++i) smallMN[i] = static_cast<float>((i * 13 + (i >> 5) * 3) % 11) - 5.f;
  float *dBig, *dK, *dMN, *dOut;
  cudaMalloc(&dBig, big.size() * 4); cudaMalloc(&dK, smallK.size() * 4); cudaMalloc(&dMN, smallMN.size() * 4);
  cudaMalloc(&dOut, 128 * 96 * 4);
  cudaMemcpy(dBig, big.data(), big.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dK, smallK.data(), smallK.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dMN, smallMN.data(), smallMN.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + R * 128 + 4 * 128 * 128 + 64;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 66, 67, 130, 131, 258, 259, 260};
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tmBig = map2d(dBig, R, 256, mode == 1);
    CUtensorMap tmSmall = mode == 0 ? map2d(dK, 16, 16, false) : map2d(dMN, 4 * 128, 128, true);
    const int N = mode == 0 ? 16 : 32;
    for (int bo = 0; bo < (mode == 0 ? 2 : 3); ++bo) {
      for (int shift : shifts) {
        cudaMemset(dOut, 0, 128 * 32 * 4);
        k_probe<<<1, 128, smem>>>(tmBig, tmSmall, mode, shift, bo, 0, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d bo %d shift %d: CUDA error %s\n", mode, bo, shift, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * N);
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            if (mode == 0) {
              for (int k = 0; k < 32; ++k) ref += (double)big[(shift + m) * 32 + k] * smallK[n * 32 + k];
            } else {
              // A[k][m]: block j = m / 32, row k, channel m % 32
              for (int k = 0; k < 128; ++k) ref += (double)smallMN[((m / 32) * 128 + k) * 32 + (m % 32)] * big[(shift + k) * 32 + n];
            }
            const double err = fabs(ref - out[m * N + n]);
            if (err > maxerr) maxerr = err;
          }
        printf("mode %s base_offset_mode %d shift %3d : max abs err %g %s\n", mode == 0 ? "K-major " : "MN-major", bo, shift, maxerr,
               maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
    }
  }
  {
    CUtensorMap tmBig = map2d(dBig, R, 256, true);
    CUtensorMap tmSmall = map2d(dMN, 4 * 128, 128, true);
    const int pitches[] = {10, 13, 66, 130};
    for (int pitch : pitches) {
      for (int shift : {0, 3}) {
        cudaMemset(dOut, 0, 128 * 96 * 4);
        k_probe<<<1, 128, smem>>>(tmBig, tmSmall, 2, shift, 0, pitch, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("alias pitch %d: CUDA error %s\n", pitch, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * 96);
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        // A rows: smallMN viewed as 512 consecutive rows of 32 (TMA loaded 4 boxes of 128 rows back to back)
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 96; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k)
              ref += (double)smallMN[(shift + k + (m / 32) * pitch) * 32 + (m % 32)] * big[(k + n / 32) * 32 + (n % 32)];
            const double err = fabs(ref - out[m * 96 + n]);
            if (err > maxerr) maxerr = err;
          }
        printf("alias MN-major: M blocks shifted by %3d rows, N blocks by 1 row, start shift %d : max abs err %g %s\n", pitch, shift, maxerr,
               maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
    }
  }
  return 0;
}
