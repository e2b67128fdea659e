// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (kind::tf32, fp32 accumulate in tensor memory).
//
// Every conv of the MFVI-DIP net is a "valid" convolution over a reflection-padded NHWC fp32 buffer, so tap (r,s) of
// an output tile is just the same TMA box shifted by (r,s): no im2col buffer exists anywhere.
//
//   forward : D[128 pixels x Cout] += A_tap[128 x 32ch] * W_tap[Cout x 32ch]^T        A K-major, B K-major
//   dgrad   : D[128 pixels x Cin ] += dY_tap[128 x 32co] * W_tap[32co x Cin]          A K-major, B MN-major
//             (tap offsets negative; TMA zero-fills out-of-bounds coordinates)
//   wgrad   : D[Cout x Cin] (per tap) += dY[pixels x Cout]^T * X_tap[pixels x Cin]    A MN-major, B MN-major, split-K
//             over pixel chunks, fp32 atomics into dw
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 TMEM allocator + single-thread MMA issuer, warps 2..5 epilogue
// (TMEM -> registers -> shared staging -> coalesced global stores, fused bias and BatchNorm (sum, sumsq) statistics).
// Operand tiles are 128-byte-swizzled rows of 32 fp32 (one swizzle atom), 3-4 stage mbarrier pipeline.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "conv_tc_wgrad.cuh"

namespace mfvi {
namespace tc {

constexpr int kBK = 32;         // fp32 per k-chunk = one 128-byte swizzle row   (kBM, kUmmaK, kThreads: conv_tc_wgrad.cuh)
constexpr int kABytes = kBM * kBK * 4;   // 16 KB

// ---------------------------------------------------------------------------------------------- fwd / dgrad
struct TcConvArgs {
  int Kc;             // contraction channels per tap (Cin fwd, Cout dgrad)
  int N;              // valid output channels (Cout fwd, Cin dgrad)
  int BN;             // UMMA N (multiple of 16; multiple of 32 when B is MN-major)
  int KH, KW;
  int Mh, Mw;         // output pixel space (Hout x Wout fwd, Hin x Win dgrad)
  int TH, TW;         // pixel tile, TH*TW <= 128
  int tiles_w;
  int a_bcast, b_bcast;
  int dgrad;          // tap offsets are negative, B is MN-major (original [tap][Cout][Cin] weights)
  int cstride;        // conv stride: fwd reads the input through a TMA map with elementStrides = cstride; dgrad with
                      // cstride 2 runs one output-parity class per blockIdx.y (taps of matching parity only)
  int Hfull, Wfull;   // dgrad: full dx size (Mh, Mw are per parity class)
  int a_rows2;        // stride-2 forward: Hin/2 (row offset of one sample in the parity-split 5-D input map)
  int stages;
  uint32_t b_bytes;   // bytes of one B stage
  uint32_t tmem_cols;
  MfviView o;
  const float* bias;  // [S][..] sampled bias (fwd only) or NULL
  long long bias_sstride;
  double* stats;      // [S][N][2] or NULL
  int accumulate;
  int vecO;
};

__global__ void __launch_bounds__(kThreads)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcConvArgs p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t stage_bytes = kABytes + p.b_bytes;
  uint8_t* ctrl = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = full_bar + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 17);
  int* row_off = reinterpret_cast<int*>(full_bar + 18);       // [128] element offset of each tile row, -1 = invalid
  double* col_acc = reinterpret_cast<double*>(row_off + kBM);  // [tpc][BN][2] partial BatchNorm statistics

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
  const int smp = blockIdx.z;
  const int tile = blockIdx.x;
  const int h0 = (tile / p.tiles_w) * p.TH, w0 = (tile % p.tiles_w) * p.TW;
  const int n0 = 0;
  // tap subset: all taps, or (dgrad, stride 2) the taps whose parity matches this CTA's output-parity class
  const bool par = p.dgrad && p.cstride == 2;
  const int ph = par ? (blockIdx.y >> 1) : 0, pw = par ? (blockIdx.y & 1) : 0;
  const int tstep = par ? 2 : 1;
  const int nr = par ? (p.KH - ph + 1) / 2 : p.KH, ns = par ? (p.KW - pw + 1) / 2 : p.KW;
  const int Mh = par ? (p.Hfull - ph + 1) / 2 : p.Mh, Mw = par ? (p.Wfull - pw + 1) / 2 : p.Mw;
  const int chunks = (p.Kc + kBK - 1) / kBK;
  const int n_iters = nr * ns * chunks;
  const uint32_t a_box_bytes = static_cast<uint32_t>(p.TH * p.TW) * 128u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      for (int it = 0; it < n_iters; ++it) {
        const int st = it % p.stages;
        const uint32_t phase = (it / p.stages) & 1;
        mbar_wait(smem_u32(&empty_bar[st]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[st]);
        mbar_expect_tx(fb, a_box_bytes + p.b_bytes);
        const int ti = it / chunks, kc = (it % chunks) * kBK;
        const int r = ph + tstep * (ti / ns), s = pw + tstep * (ti % ns);
        const int tap = r * p.KW + s;
        const uint32_t a_dst = smem_u32(smem + static_cast<size_t>(st) * stage_bytes);
        const uint32_t b_dst = a_dst + kABytes;
        int hh, ww;
        if (!p.dgrad) {
          hh = h0 * p.cstride + r; ww = w0 * p.cstride + s;
        } else if (!par) {
          hh = h0 - r; ww = w0 - s;
        } else {
          hh = h0 + (ph - r) / 2; ww = w0 + (pw - s) / 2;       // (ph - r), (pw - s) are even
        }
        if (!p.dgrad && p.cstride == 2) {
          // parity-split 5-D view (C, 2, W/2, 2, H/2 * S) of the padded input: pixel (2*w2 + pw, 2*h2 + ph)
          const int rows = p.a_bcast ? 0 : smp * p.a_rows2;
          tma_load_5d(a_dst, &tmA, fb, kc, s & 1, w0 + (s >> 1), r & 1, rows + h0 + (r >> 1));
        } else {
          tma_load_4d(a_dst, &tmA, fb, kc, ww, hh, p.a_bcast ? 0 : smp);
        }
        const int bs = p.b_bcast ? 0 : smp;
        if (!p.dgrad) {
          tma_load_4d(b_dst, &tmB, fb, kc, n0, tap, bs);            // box (32 k, BN n)
        } else {
          for (int j = 0; j < p.BN / 32; ++j)                        // boxes (32 n, 32 k), 4 KB each
            tma_load_4d(b_dst + j * 4096, &tmB, fb, n0 + 32 * j, kc, tap, bs);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread)
    const uint32_t idesc = make_idesc(kBM, p.BN, 0, p.dgrad ? 1 : 0);
    for (int it = 0; it < n_iters; ++it) {
      const int st = it % p.stages;
      const uint32_t phase = (it / p.stages) & 1;
      mbar_wait(smem_u32(&full_bar[st]), phase);
      tc_fence_after();
      {
        const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(st) * stage_bytes);
        const uint32_t b_addr = a_addr + kABytes;
        const uint32_t a_lo = desc_lo(a_addr, 16), a_hi = desc_hi(1024, kLayoutSw128);
        const uint32_t b_lo = p.dgrad ? desc_lo(b_addr, 4096) : desc_lo(b_addr, 16);
        const uint32_t b_hi = p.dgrad ? desc_hi(512, kLayoutSw128Base32) : a_hi;
        const uint32_t b_k = p.dgrad ? 64u : 2u;
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k)
          tc_mma_tf32_elect(tmem_base, desc_pack(a_lo + 2u * k, a_hi), desc_pack(b_lo + b_k * k, b_hi), idesc, (it > 0 || k > 0) ? 1u : 0u);
        tc_commit_elect(smem_u32(&empty_bar[st]));
        if (it == n_iters - 1) tc_commit_elect(smem_u32(tmem_full_bar));
      }
    }
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4
    const int q = warp & 3;
    const int m = q * 32 + lane;              // tile row == TMEM lane
    const int et = threadIdx.x - 64;          // 0..127
    {
      const int hl = m / p.TW, wl = m % p.TW;
      const bool ok = hl < p.TH && (h0 + hl) < Mh && (w0 + wl) < Mw;
      const long long oh = static_cast<long long>(h0 + hl) * tstep + ph, ow = static_cast<long long>(w0 + wl) * tstep + pw;
      row_off[m] = ok ? static_cast<int>(oh * p.o.hstride + ow * p.o.wstride) : -1;
    }
    if (n_iters > 0) {
      mbar_wait(smem_u32(tmem_full_bar), 0);
      tc_fence_after();
    }
    float* stg = reinterpret_cast<float*>(smem);      // [128][BN+1], reuses the (drained) pipeline stages
    const int ld = p.BN + 1;
    const bool row_ok = row_off[m] >= 0;
    const float* bias = (p.bias != nullptr) ? p.bias + static_cast<size_t>(smp) * p.bias_sstride + n0 : nullptr;
    for (int c = 0; c < p.BN; c += 16) {
      float v[16];
      if (n_iters > 0) {
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
      } else {                                   // parity class without taps (e.g. 1x1 stride 2): dx = 0 there
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = v[j];
        if (bias != nullptr && n0 + c + j < p.N) x += bias[c + j];
        stg[m * ld + c + j] = (row_ok && n0 + c + j < p.N) ? x : 0.f;
      }
    }
    tc_fence_before();
    epi_barrier();
    // ---- BatchNorm statistics of this tile (rows that are invalid hold zeros)
    if (p.stats != nullptr) {
      // column sums in double (var = E[x^2] - mean^2 cancels when |mean| >> std), fixed summation order
      const int tpc = p.BN <= 128 ? 128 / p.BN : 1;          // threads per column
      for (int col = et % p.BN, part = et / p.BN; col < p.BN && part < tpc; col += 128) {
        double s1 = 0.0, s2 = 0.0;
        for (int rr = part; rr < kBM; rr += tpc) {
          const double x = static_cast<double>(stg[rr * ld + col]);
          s1 += x;
          s2 += x * x;
        }
        col_acc[(part * p.BN + col) * 2] = s1;
        col_acc[(part * p.BN + col) * 2 + 1] = s2;
        if (p.BN <= 128) break;
      }
      epi_barrier();
      for (int col = et; col < p.BN; col += 128) {
        if (n0 + col < p.N) {
          double s1 = 0.0, s2 = 0.0;
          for (int part = 0; part < tpc; ++part) {
            s1 += col_acc[(part * p.BN + col) * 2];
            s2 += col_acc[(part * p.BN + col) * 2 + 1];
          }
          double* dst = p.stats + (static_cast<size_t>(smp) * p.N + n0 + col) * 2;
          atomicAdd(dst, s1);
          atomicAdd(dst + 1, s2);
        }
      }
    }
    // ---- coalesced store: BN/4 threads per row
    float* obase = p.o.ptr + static_cast<size_t>(smp) * p.o.sstride + n0;
    const int tpr = p.BN / 4;
    for (int idx = et; idx < kBM * tpr; idx += 128) {
      const int rr = idx / tpr, cq = (idx % tpr) * 4;
      const int off = row_off[rr];
      if (off < 0 || n0 + cq >= p.N) continue;
      float* dst = obase + off + cq;
      const float* src = stg + rr * ld + cq;
      if (p.vecO && n0 + cq + 3 < p.N) {
        float4 v = make_float4(src[0], src[1], src[2], src[3]);
        if (p.accumulate) {
          const float4 o = *reinterpret_cast<const float4*>(dst);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        for (int j = 0; j < 4; ++j)
          if (n0 + cq + j < p.N) dst[j] = p.accumulate ? dst[j] + src[j] : src[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------- wgrad: conv_tc_wgrad.cuh

// dbias[s][co] += sum over pixels of dy[s][.][.][co]   (column sum; V = 4 when C % 4 == 0, else scalar).  Few CTAs per sample
// (every CTA ends with C float atomics on the same cache line, which the L2 serialises: profiles/r02_stat_atomics_negative.txt),
// four independent loads per thread and trip, lanes of one channel group combined by shuffles before the shared-memory atomics.
template <int V>
__global__ void __launch_bounds__(256)
k_bias_grad(MfviView dy, int H, int W, int C, float* __restrict__ dbias, long long sstride) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm_part[];                 // [C]
  const int s = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sm_part[c] = 0.f;
  __syncthreads();
  const int G = C / V;                               // channel groups per pixel
  const int PPB = blockDim.x / G > 0 ? blockDim.x / G : 1;
  const int g = threadIdx.x % G, slot = threadIdx.x / G;
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  if (slot < PPB) {
    const int npix = H * W, step = gridDim.x * PPB;
    constexpr int U = 4;
    for (int px0 = blockIdx.x * PPB + slot; px0 < npix; px0 += U * step) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int px = px0 + u * step;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (px < npix) {
          const float* src = dy.ptr + view_off(dy, s, px / W, px % W) + V * g;
          if (V == 4) v[u] = *reinterpret_cast<const float4*>(src);
          else v[u].x = src[0];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[0] += v[u].x; acc[1 % V] += V > 1 ? v[u].y : 0.f; acc[2 % V] += V > 2 ? v[u].z : 0.f; acc[3 % V] += V > 3 ? v[u].w : 0.f;
      }
    }
  }
  // lanes l and l + k*G of a warp hold the same channel group when G divides 32: fold them before touching shared memory
  if (G <= 32 && (32 % G) == 0 && blockDim.x % 32 == 0) {
    for (int off = 16; off >= G; off >>= 1)
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], off);
    if ((threadIdx.x & 31) < G && slot < PPB)
#pragma unroll
      for (int j = 0; j < V; ++j) atomicAdd(&sm_part[V * g + j], acc[j]);
  } else if (slot < PPB) {
#pragma unroll
    for (int j = 0; j < V; ++j) atomicAdd(&sm_part[V * g + j], acc[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dbias + static_cast<size_t>(s) * sstride + c, sm_part[c]);
}

// ---------------------------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  });
  return fn;
}

// 4-D fp32 tensor map, 128B swizzle, zero fill. dims/box innermost first; strides in bytes for dims 1..3.
// bf16: 2-byte elements; MN-major bf16 operands use the plain 128-byte swizzle (64-element atoms).
static bool encode_map(CUtensorMap* m, const void* base, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                       const uint32_t* estr, bool mn_major = false, int rank = 4, bool bf16 = false) {
  if (dry_run() != nullptr) return true;
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  if (enc == nullptr) return false;
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (mn_major && !bf16) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Stride 1 and 2 are taken (stride 2: parity-split 5-D input maps for fwd/wgrad, one output-parity class per
// blockIdx.y for dgrad).  MFVI_TC_STRIDE2=0 sends stride-2 layers back to the fp32 CUDA-core kernels (debug switch).
static bool tc_stride_ok(int stride) {
  static const bool s2 = [] { const char* e = getenv("MFVI_TC_STRIDE2"); return e == nullptr || e[0] != '0'; }();
  return stride == 1 || (stride == 2 && s2);
}

static bool view_tma_ok(const MfviView& v, int C) {
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && (C % 4 == 0) && (v.wstride % 4 == 0) && (v.hstride % 4 == 0) &&
         (v.sstride % 4 == 0) && v.wstride >= C && v.hstride >= v.wstride;
}
// bf16 views: 16-byte aligned pixels = strides in multiples of 8 elements; the channel count itself is free (TMA zero-fills)
static bool view_tma_ok_bf16(const MfviView& v, int C) {
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && C >= 1 && (v.wstride % 8 == 0) && (v.hstride % 8 == 0) &&
         (v.sstride % 8 == 0) && v.wstride >= C && v.hstride >= v.wstride;
}

static uint32_t pow2_cols(int n) {
  uint32_t c = 32;
  while (c < static_cast<uint32_t>(n)) c <<= 1;
  return c;
}

static void pick_tile(int Mh, int Mw, int& TH, int& TW) {
  TW = Mw >= 128 ? 128 : Mw;
  TH = 128 / TW;
  if (TH < 1) TH = 1;
  if (TH > Mh) TH = Mh;
}

// a: activation view read through TMA (x for fwd, dy for dgrad), its channel count Ca and spatial size (Ha, Wa).
static bool map_activation(CUtensorMap* m, const MfviView& a, int Ca, int Ha, int Wa, int S, int TH, int TW, bool& bcast,
                           bool mn_major = false, int cstride = 1, bool bf16 = false) {
  bcast = (a.sstride == 0) || S == 1;
  if (bf16) {
    // bf16 view (weight-gradient operands only): one 128-byte row = 64 channels, strides in 2-byte elements
    if (TW > 256 || TH > 256) return false;
    if (cstride == 2) {
      if ((Ha & 1) || (Wa & 1) || (!bcast && a.sstride != static_cast<long long>(Ha) * a.hstride)) return false;
      const uint64_t dims[5] = {static_cast<uint64_t>(Ca), 2, static_cast<uint64_t>(Wa / 2), 2,
                                static_cast<uint64_t>(Ha / 2) * (bcast ? 1 : S)};
      const uint64_t strides[4] = {static_cast<uint64_t>(a.wstride) * 2, static_cast<uint64_t>(a.wstride) * 4,
                                   static_cast<uint64_t>(a.hstride) * 2, static_cast<uint64_t>(a.hstride) * 4};
      const uint32_t box[5] = {64, 1, static_cast<uint32_t>(TW), 1, static_cast<uint32_t>(TH)};
      const uint32_t estr[5] = {1, 1, 1, 1, 1};
      return encode_map(m, a.ptr, dims, strides, box, estr, mn_major, 5, true);
    }
    const uint64_t dims[4] = {static_cast<uint64_t>(Ca), static_cast<uint64_t>(Wa), static_cast<uint64_t>(Ha),
                              static_cast<uint64_t>(bcast ? 1 : S)};
    const uint64_t sbytes = bcast ? static_cast<uint64_t>(a.hstride) * Ha * 2 : static_cast<uint64_t>(a.sstride) * 2;
    const uint64_t strides[3] = {static_cast<uint64_t>(a.wstride) * 2, static_cast<uint64_t>(a.hstride) * 2, sbytes};
    const uint32_t box[4] = {64, static_cast<uint32_t>(TW), static_cast<uint32_t>(TH), 1};
    const uint32_t estr[4] = {1, 1, 1, 1};
    return encode_map(m, a.ptr, dims, strides, box, estr, mn_major, 4, true);
  }
  if (cstride == 2) {
    // parity-split view: dims (C, 2, W/2, 2, H/2 * S); the sample axis is folded into the row axis, which needs
    // densely stacked samples.  (TMA elementStrides are avoided on purpose.)
    if ((Ha & 1) || (Wa & 1) || (!bcast && a.sstride != static_cast<long long>(Ha) * a.hstride)) return false;
    const uint64_t dims[5] = {static_cast<uint64_t>(Ca), 2, static_cast<uint64_t>(Wa / 2), 2,
                              static_cast<uint64_t>(Ha / 2) * (bcast ? 1 : S)};
    const uint64_t strides[4] = {static_cast<uint64_t>(a.wstride) * 4, static_cast<uint64_t>(a.wstride) * 8,
                                 static_cast<uint64_t>(a.hstride) * 4, static_cast<uint64_t>(a.hstride) * 8};
    const uint32_t box[5] = {static_cast<uint32_t>(kBK), 1, static_cast<uint32_t>(TW), 1, static_cast<uint32_t>(TH)};
    const uint32_t estr[5] = {1, 1, 1, 1, 1};
    if (TW > 256 || TH > 256) return false;
    return encode_map(m, a.ptr, dims, strides, box, estr, mn_major, 5);
  }
  const uint64_t dims[4] = {static_cast<uint64_t>(Ca), static_cast<uint64_t>(Wa), static_cast<uint64_t>(Ha),
                            static_cast<uint64_t>(bcast ? 1 : S)};
  const uint64_t sbytes = bcast ? static_cast<uint64_t>(a.hstride) * Ha * 4 : static_cast<uint64_t>(a.sstride) * 4;
  const uint64_t strides[3] = {static_cast<uint64_t>(a.wstride) * 4, static_cast<uint64_t>(a.hstride) * 4, sbytes};
  const uint32_t box[4] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(TW), static_cast<uint32_t>(TH), 1};
  const uint32_t estr[4] = {1, 1, 1, 1};
  if (TW > 256 || TH > 256) return false;
  return encode_map(m, a.ptr, dims, strides, box, estr, mn_major);
}

static size_t conv_smem_bytes(int stages, uint32_t b_bytes, int BN) {
  size_t pipe = static_cast<size_t>(stages) * (kABytes + b_bytes);
  size_t stg = static_cast<size_t>(kBM) * (BN + 1) * 4;
  if (stg > pipe) pipe = (stg + 1023) / 1024 * 1024;
  return 1024 + pipe + 18 * 8 + kBM * 4 + static_cast<size_t>(BN > 128 ? BN : 128) * 16 + 64;
}

}  // namespace tc
}  // namespace mfvi

using namespace mfvi;

extern "C" {

// Returns 0 on success, -1 when this shape is not taken by the tensor-core kernel (caller falls back), >0 on error.
int mfvi_conv2d_fwd_tc(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                       double* stats, mfvi_stream_t st) {
  using namespace mfvi::tc;
  if (!tc_stride_ok(d->stride) || !view_tma_ok(x, d->Cin) || (reinterpret_cast<uintptr_t>(w) % 16) || (w_sstride % 4) || d->Cout > 256) return -1;
  TcConvArgs a{};
  a.Kc = d->Cin; a.N = d->Cout; a.BN = (d->Cout + 15) / 16 * 16;
  a.KH = d->KH; a.KW = d->KW; a.Mh = d->Hout; a.Mw = d->Wout;
  pick_tile(a.Mh, a.Mw, a.TH, a.TW);
  a.tiles_w = (a.Mw + a.TW - 1) / a.TW;
  a.dgrad = 0; a.cstride = d->stride; a.Hfull = a.Mh; a.Wfull = a.Mw; a.a_rows2 = d->Hin / 2;
  a.b_bytes = static_cast<uint32_t>(a.BN) * 128u;
  a.stages = a.BN >= 64 ? 3 : 4;
  if (const char* e = getenv("MFVI_TC_STAGES")) a.stages = atoi(e);
  a.tmem_cols = pow2_cols(a.BN);
  a.o = y; a.bias = bias; a.bias_sstride = w_sstride; a.stats = stats; a.accumulate = 0;
  a.vecO = ((reinterpret_cast<uintptr_t>(y.ptr) % 16 == 0) && y.sstride % 4 == 0 && y.hstride % 4 == 0 && y.wstride % 4 == 0 && d->Cout % 4 == 0) ? 1 : 0;
  CUtensorMap tmA, tmB;
  bool ab = false;
  if (!map_activation(&tmA, x, d->Cin, d->Hin, d->Win, d->S, a.TH, a.TW, ab, false, d->stride)) return -1;
  a.a_bcast = ab ? 1 : 0;
  a.b_bcast = (w_sstride == 0 || d->S == 1) ? 1 : 0;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(d->KH * d->KW),
                              static_cast<uint64_t>(a.b_bcast ? 1 : d->S)};
    const uint64_t tap_bytes = static_cast<uint64_t>(d->Cout) * d->Cin * 4;
    const uint64_t strides[3] = {static_cast<uint64_t>(d->Cin) * 4, tap_bytes,
                                 a.b_bcast ? tap_bytes * d->KH * d->KW : static_cast<uint64_t>(w_sstride) * 4};
    const uint32_t box[4] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(a.BN), 1, 1};
    const uint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmB, w, dims, strides, box, estr)) return -1;
  }
  const size_t smem = conv_smem_bytes(a.stages, a.b_bytes, a.BN);
  static unsigned long long attr_done = 0;
  if (dry_run() == nullptr) {
    const cudaError_t e = allow_dyn_smem(k_conv_tc, 200 * 1024, &attr_done);
    MFVI_REQUIRE(e == cudaSuccess, "conv2d_fwd_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  const int tiles_h = (a.Mh + a.TH - 1) / a.TH;
  dim3 grid(a.tiles_w * tiles_h, 1, d->S);
  dry_detail("TH=%d TW=%d BN=%d stages=%d tmem_cols=%u", a.TH, a.TW, a.BN, a.stages, a.tmem_cols);
  launch_k(k_conv_tc, grid, kThreads, smem, as_stream(st), tmA, tmB, a);
  return check_launch("conv2d_fwd_tc");
}

int mfvi_conv2d_dgrad_tc(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                         mfvi_stream_t st) {
  using namespace mfvi::tc;
  if (!tc_stride_ok(d->stride) || !view_tma_ok(dy, d->Cout) || (reinterpret_cast<uintptr_t>(w) % 16) || (w_sstride % 4) || d->Cin % 4 ||
      d->Cin > 256)
    return -1;
  TcConvArgs a{};
  a.Kc = d->Cout; a.N = d->Cin; a.BN = (d->Cin + 31) / 32 * 32;
  a.KH = d->KH; a.KW = d->KW;
  a.cstride = d->stride; a.Hfull = d->Hin; a.Wfull = d->Win;
  a.Mh = (d->Hin + d->stride - 1) / d->stride; a.Mw = (d->Win + d->stride - 1) / d->stride;   // largest parity class
  pick_tile(a.Mh, a.Mw, a.TH, a.TW);
  a.tiles_w = (a.Mw + a.TW - 1) / a.TW;
  a.dgrad = 1;
  a.b_bytes = static_cast<uint32_t>(a.BN) * 128u;
  a.stages = a.BN >= 64 ? 3 : 4;
  a.tmem_cols = pow2_cols(a.BN);
  a.o = dx; a.bias = nullptr; a.bias_sstride = 0; a.stats = nullptr; a.accumulate = accumulate;
  a.vecO = ((reinterpret_cast<uintptr_t>(dx.ptr) % 16 == 0) && dx.sstride % 4 == 0 && dx.hstride % 4 == 0 && dx.wstride % 4 == 0) ? 1 : 0;
  CUtensorMap tmA, tmB;
  bool ab = false;
  if (!map_activation(&tmA, dy, d->Cout, d->Hout, d->Wout, d->S, a.TH, a.TW, ab)) return -1;
  a.a_bcast = ab ? 1 : 0;
  a.b_bcast = (w_sstride == 0 || d->S == 1) ? 1 : 0;
  {
    // original weights [tap][Cout][Cin]: inner dim = Cin (the GEMM N), rows = Cout (the GEMM K): MN-major B
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(d->KH * d->KW),
                              static_cast<uint64_t>(a.b_bcast ? 1 : d->S)};
    const uint64_t tap_bytes = static_cast<uint64_t>(d->Cout) * d->Cin * 4;
    const uint64_t strides[3] = {static_cast<uint64_t>(d->Cin) * 4, tap_bytes,
                                 a.b_bcast ? tap_bytes * d->KH * d->KW : static_cast<uint64_t>(w_sstride) * 4};
    const uint32_t box[4] = {32, static_cast<uint32_t>(kBK), 1, 1};
    const uint32_t estr[4] = {1, 1, 1, 1};
    if (!encode_map(&tmB, w, dims, strides, box, estr, true)) return -1;
  }
  const size_t smem = conv_smem_bytes(a.stages, a.b_bytes, a.BN);
  static unsigned long long attr_done = 0;
  if (dry_run() == nullptr) {
    const cudaError_t e = allow_dyn_smem(k_conv_tc, 200 * 1024, &attr_done);
    MFVI_REQUIRE(e == cudaSuccess, "conv2d_dgrad_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  const int tiles_h = (a.Mh + a.TH - 1) / a.TH;
  dim3 grid(a.tiles_w * tiles_h, d->stride == 2 ? 4 : 1, d->S);
  dry_detail("TH=%d TW=%d BN=%d stages=%d tmem_cols=%u", a.TH, a.TW, a.BN, a.stages, a.tmem_cols);
  launch_k(k_conv_tc, grid, kThreads, smem, as_stream(st), tmA, tmB, a);
  return check_launch("conv2d_dgrad_tc");
}

// dbias[s][co] += sum over pixels of dy (requires Cout % 4 == 0 and 16-byte aligned rows: the same conditions as the TMA paths)
int mfvi_conv2d_bias_grad_tc(const MfviConvDesc* d, MfviView dy, float* dbias, long long w_sstride, mfvi_stream_t st) {
  using namespace mfvi::tc;
  const bool vec = d->Cout % 4 == 0 && (reinterpret_cast<uintptr_t>(dy.ptr) % 16 == 0) && dy.sstride % 4 == 0 && dy.hstride % 4 == 0 &&
                   dy.wstride % 4 == 0;
  const int groups = vec ? d->Cout / 4 : d->Cout;
  if (groups > 256) return -1;
  // one wave of about two CTAs per SM over all samples, at most 64 per sample (the float atomics at the end of every CTA)
  int blocks = std::max(1, std::min(std::min(64, (kNumSMs * 2 + d->S - 1) / d->S), (d->Hout * d->Wout * groups + 255) / 256));
  dim3 g2(blocks, d->S);
  if (vec)
    launch_k(k_bias_grad<4>, g2, 256, d->Cout * sizeof(float), as_stream(st), dy, d->Hout, d->Wout, d->Cout, dbias, w_sstride);
  else
    launch_k(k_bias_grad<1>, g2, 256, d->Cout * sizeof(float), as_stream(st), dy, d->Hout, d->Wout, d->Cout, dbias, w_sstride);
  return check_launch("conv2d_bias_grad_tc");
}

}  // extern "C"

static int wgrad_tc_launch(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                           mfvi_stream_t st) {
  using namespace mfvi::tc;
  constexpr bool bf16 = false;
  const bool views_ok = bf16 ? (view_tma_ok_bf16(x, d->Cin) && view_tma_ok_bf16(dy, d->Cout)) : (view_tma_ok(x, d->Cin) && view_tma_ok(dy, d->Cout));
  if (!tc_stride_ok(d->stride) || !views_ok || d->Cout > 128 || d->Cin > 256 ||
      (reinterpret_cast<uintptr_t>(dw) % 16) || (w_sstride % 4))
    return -1;
  const int cb = bf16 ? 64 : 32;                 // channels of one 128-byte operand row
  TcWgradArgs a{};
  a.Cout = d->Cout; a.Cin = d->Cin; a.KW = d->KW; a.Ho = d->Hout; a.Wo = d->Wout;
  a.MB = (d->Cout + cb - 1) / cb; a.NB = (d->Cin + cb - 1) / cb; a.cstride = d->stride; a.x_rows2 = d->Hin / 2;
  if (!bf16 && x.sstride == 0 && dy.sstride != 0 && d->Cout <= 32 && d->S >= 4 && d->S % 4 == 0) {      // see TcWgradArgs::sgrp
    a.sgrp = 4;
    a.MB = 4;
  }
  const int a_blocks = a.MB == 1 ? 1 : (bf16 ? 2 : 4);   // see k_wgrad_tc: one dy block is aliased across M when it is the only one
  int TP = a.MB == 1 ? 256 : 128;
  const size_t stage_cap = a.MB == 1 ? 64 * 1024 : 48 * 1024;
  while (TP > 8 && static_cast<size_t>(a_blocks + a.NB) * TP * 128 > stage_cap) TP >>= 1;
  // the pixel tile must cover exactly TP rows (every K row enters the sum): TW | TP
  int TW = d->Wout >= TP ? TP : d->Wout;
  while (TW > 1 && TP % TW) --TW;
  int TH = TP / TW;
  if (TH > d->Hout) {        // tiny images: shrink the tile
    TH = d->Hout;
    // bf16 has no CUDA-core fallback and contracts 16 pixel rows per MMA: keep the tile a multiple of 16 rows by letting its
    // box reach below the image (TMA zero-fills those rows of x and dy, so they add nothing).  TW is a power of two here.
    if (bf16 && TW < 16 && TH % (16 / TW)) TH = (TH + 16 / TW - 1) / (16 / TW) * (16 / TW);
    TP = TH * TW;
  }
  if (TP % (bf16 ? 16 : 8) || TW > 256 || TH > 256) return -1;
  a.TH = TH; a.TW = TW; a.TP = TP;
  a.tiles_w = (d->Wout + TW - 1) / TW;
  a.n_tiles = a.tiles_w * ((d->Hout + TH - 1) / TH);
  const int taps = d->KH * d->KW;
  const int Sz = a.sgrp > 0 ? d->S / a.sgrp : d->S;          // CTA columns along the sample axis
  int chunks = (2 * kNumSMs + taps * Sz - 1) / (taps * Sz);
  chunks = std::max(1, std::min(chunks, a.n_tiles));
  a.tiles_per_cta = (a.n_tiles + chunks - 1) / chunks;
  chunks = (a.n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
  a.stages = static_cast<size_t>(a_blocks + a.NB) * TP * 128 > 48 * 1024 ? 3 : 4;
  a.tmem_cols = pow2_cols(a.NB * cb);
  a.dw = dw; a.w_sstride = w_sstride;
  CUtensorMap tmDy, tmX;
  bool db = false, xb = false;
  if (!map_activation(&tmDy, dy, d->Cout, d->Hout, d->Wout, d->S, TH, TW, db, true, 1, bf16)) return -1;
  if (db && d->S > 1) return -1;
  if (!map_activation(&tmX, x, d->Cin, d->Hin, d->Win, d->S, TH, TW, xb, true, d->stride, bf16)) return -1;
  a.x_bcast = xb ? 1 : 0;
  const size_t smem = 1024 + static_cast<size_t>(a.stages) * (a_blocks + a.NB) * TP * 128 + 18 * 8 + 64;
  static unsigned long long attr_done = 0;
  if (dry_run() == nullptr) {
    const cudaError_t e = allow_dyn_smem(k_wgrad_tc<false>, 220 * 1024, &attr_done);
    MFVI_REQUIRE(e == cudaSuccess, "conv2d_wgrad_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  MFVI_REQUIRE(smem <= 220 * 1024, "conv2d_wgrad_tc: stage does not fit in shared memory");
  dim3 grid(chunks, taps, Sz);
  dry_detail("TH=%d TW=%d TP=%d MB=%d NB=%d sgrp=%d stages=%d tmem_cols=%u tiles_per_cta=%d", a.TH, a.TW, a.TP, a.MB, a.NB, a.sgrp,
             a.stages, a.tmem_cols, a.tiles_per_cta);
  launch_k(k_wgrad_tc<false>, grid, kThreads, smem, as_stream(st), tmDy, tmX, a);
  if (int rc = check_launch("conv2d_wgrad_tc")) return rc;
  if (dbias != nullptr) return mfvi_conv2d_bias_grad_tc(d, dy, dbias, w_sstride, st);
  return 0;
}


extern "C" {

int mfvi_conv2d_wgrad_tc(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                         mfvi_stream_t st) {
  return wgrad_tc_launch(d, x, dy, dw, dbias, w_sstride, st);
}

}  // extern "C"
