// Weight gradient of the small-channel (Cout <= 32) stride-1 convolutions on tcgen05 (kind::tf32), with the filter taps
// packed into ONE MMA by descriptor aliasing.
//
// wgrad is a GEMM with a tiny output (Cout x Cin per tap) and a huge contraction (all pixels), and kind::tf32 contracts
// only 8 pixels per instruction, so conv_tc.cu's k_wgrad_tc (one MMA stream per tap) is bound by the NUMBER of MMA
// instructions (~100 issue cycles each), not by flops or bytes.  Both operands are MN-major here (rows = pixels, 128-byte
// rows of 32 channels), and the UMMA descriptor places consecutive 32-wide M / N blocks LBO bytes apart -- with ANY LBO,
// because the swizzle is a function of the absolute shared-memory address (profiles/r01_umma_alias_probe.txt).  So:
//
//   3x3:  A = dy tile in "flat" pixel space (pitch Pw = TW+2, junk columns zero), M block m starts m rows later (LBO = 128 B);
//         B = x halo tile, N block n starts n*Pw rows later (LBO = Pw*128 B).  One instruction accumulates
//             D[(m,co)][(n,ci)] += sum_k dyS[k+m][co] * xS[k+n*Pw][ci]          = filter tap (r = n, s = 2-m)
//         i.e. all nine taps at once (M = 128: three tap columns + one ignored block, N = 96).
//   1x1:  M block m = dy rows 8m later, N block n = x rows 8n later, k advances 32 rows per instruction: the four diagonal
//         blocks hold four consecutive 8-pixel slices of the contraction and are summed in the epilogue (4x fewer MMAs).
//
// Cin is processed in blocks of 32 channels (one accumulator of 96 / 128 TMEM columns each).  Split-K: a CTA owns a
// contiguous run of pixel tiles of one sample, accumulates in TMEM and adds its result into dw[s] with fp32 atomics.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue.
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mfvi {
namespace tc3 {
using namespace mfvi::tc;

constexpr int kThreads = 192;
constexpr int kMaxCblk = 5;

struct WArgs {
  int Cout, Cin, K3;            // K3 = 1 for 3x3, 0 for 1x1;  Cout = the <= 32 output channels THIS launch covers
  int Cout_total, co0;          // the layer's output channels and the first one of this launch (dw row = co0 + co)
  int n_cblk;
  int Ho, Wo;
  int TH, TW, Pw;
  int tiles_h, tiles_w, tiles_per_sample, tiles_per_cta, ctas_per_sample;
  int n_k, k_rows;
  int dy_rows, x_rows;          // rows (128 B each) of the dy region and of one x channel-block region of a stage
  int x_bcast;
  int n_stages;
  uint32_t stage_bytes, tmem_cols;
  float* dw;                    // [S][taps][Cout][Cin]
  long long w_sstride;
};

__global__ void __launch_bounds__(kThreads, 1)
k_wgrad_alias(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const WArgs p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* ctrl = smem + static_cast<size_t>(p.n_stages) * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* acc_bar = full_bar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 9);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int smp = blockIdx.x / p.ctas_per_sample;
  const int part = blockIdx.x - smp * p.ctas_per_sample;
  const int t_begin = part * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.tiles_per_sample);
  const int n_tiles = t_end - t_begin;

  // every row the TMA never writes (junk columns of the flat dy tile, slack rows) must read as zero
  {
    float4* z = reinterpret_cast<float4*>(smem);
    const int n16 = static_cast<int>((static_cast<size_t>(p.n_stages) * p.stage_bytes) >> 4);
    for (int i = threadIdx.x; i < n16; i += kThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(acc_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  fence_proxy_async();            // the zero fill (generic proxy) must be ordered before the TMA writes (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();          // the zero fill / barrier / TMEM prologue overlapped the previous kernel
  const uint32_t dy_bytes = static_cast<uint32_t>(p.dy_rows) * 128u, x_bytes = static_cast<uint32_t>(p.x_rows) * 128u;
  const int BN = p.K3 ? 96 : 128;

  if (n_tiles > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < n_tiles; ++it) {
          const int st = it % p.n_stages;
          mbar_wait(smem_u32(&empty_bar[st]), ((it / p.n_stages) & 1) ^ 1);
          const uint32_t fb = smem_u32(&full_bar[st]);
          const int t = t_begin + it;
          const int h0 = (t / p.tiles_w) * p.TH, w0 = (t % p.tiles_w) * p.TW;
          const uint32_t dy_dst = smem_u32(smem + static_cast<size_t>(st) * p.stage_bytes);
          const int xs = p.x_bcast ? 0 : smp;
          if (p.K3) {
            mbar_expect_tx(fb, static_cast<uint32_t>(p.TH * p.TW + p.n_cblk * (p.TH + 2) * p.Pw) * 128u);
            for (int hl = 0; hl < p.TH; ++hl)          // box (32 co, TW px, 1 row) -> flat row 2 + hl*Pw
              tma_load_4d(dy_dst + static_cast<uint32_t>(2 + hl * p.Pw) * 128u, &tmDy, fb, 0, w0, h0 + hl, smp);
          } else {
            mbar_expect_tx(fb, static_cast<uint32_t>((1 + p.n_cblk) * p.TH * p.TW) * 128u);
            tma_load_4d(dy_dst, &tmDy, fb, 0, w0, h0, smp);                       // box (32 co, TW, TH)
          }
          for (int c = 0; c < p.n_cblk; ++c)           // box (32 ci, Pw, TH+2) / (32 ci, TW, TH)
            tma_load_4d(dy_dst + dy_bytes + c * x_bytes, &tmX, fb, 32 * c, w0, h0, xs);
        }
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(128, BN, 1, 1);
      const uint32_t hi = desc_hi(512, kLayoutSw128Base32);
      const uint32_t lbo_a = p.K3 ? 128u : 1024u, lbo_b = p.K3 ? static_cast<uint32_t>(p.Pw) * 128u : 1024u;
      const uint32_t k_adv = static_cast<uint32_t>(p.k_rows) * 8u;      // (k_rows * 128 B) >> 4
      for (int it = 0; it < n_tiles; ++it) {
        const int st = it % p.n_stages;
        mbar_wait(smem_u32(&full_bar[st]), (it / p.n_stages) & 1);
        tc_fence_after();
        const uint32_t dy_addr = smem_u32(smem + static_cast<size_t>(st) * p.stage_bytes);
        const uint32_t a_lo0 = desc_lo(dy_addr, lbo_a);
        for (int c = 0; c < p.n_cblk; ++c) {
          const uint32_t b_lo0 = desc_lo(dy_addr + dy_bytes + c * x_bytes, lbo_b);
          const uint32_t d = tmem_base + static_cast<uint32_t>(c * BN);
          uint32_t a_lo = a_lo0, b_lo = b_lo0;
          for (int kk = 0; kk < p.n_k; ++kk, a_lo += k_adv, b_lo += k_adv)
            tc_mma_tf32_elect(d, desc_pack(a_lo, hi), desc_pack(b_lo, hi), idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit_elect(smem_u32(&empty_bar[st]));
      }
      tc_commit_elect(smem_u32(acc_bar));
    } else {
      // ===== epilogue: TMEM lane = (m block, co); column = (channel block, n block, ci)
      const int q = warp & 3;
      mbar_wait(smem_u32(acc_bar), 0);
      tc_fence_after();
      const int co = lane;
      float* dws = p.dw + static_cast<size_t>(smp) * p.w_sstride;
      for (int c = 0; c < p.n_cblk; ++c) {
        for (int n = 0; n < BN / 32; ++n) {
          // 3x3: tap (r = n, s = 2 - q), lanes of block q = 3 are ignored;  1x1: only the diagonal block n == q counts
          const bool use = p.K3 ? (q < 3) : (n == q);
          const int tap = p.K3 ? n * 3 + (2 - q) : 0;
          float* dst = dws + (static_cast<size_t>(tap) * p.Cout_total + p.co0 + co) * p.Cin + c * 32;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float v[16];
            tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * BN + n * 32 + h * 16), v);
            if (use && co < p.Cout) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const int ci = c * 32 + h * 16 + i;
                if (ci + 3 < p.Cin) {
                  atomicAdd(reinterpret_cast<float4*>(dst + h * 16 + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                } else {
                  for (int e = 0; e < 4; ++e)
                    if (ci + e < p.Cin) atomicAdd(dst + h * 16 + i + e, v[i + e]);
                }
              }
            }
          }
        }
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

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int rup(int a, int b) { return cdiv(a, b) * b; }
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

static bool view_ok(const MfviView& v, int C) {
  // TMA needs a 16-byte aligned base and 16-byte multiples for every stride; the channel count itself is free
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && C >= 1 && (v.wstride % 4 == 0) && (v.hstride % 4 == 0) &&
         (v.sstride % 4 == 0) && v.wstride >= C && v.hstride >= v.wstride;
}

static bool encode_act(CUtensorMap* m, const MfviView& a, int C, int H, int W, int S, bool bcast, int box_w, int box_h) {
  const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(bcast ? 1 : S)};
  const uint64_t sbytes = bcast ? static_cast<uint64_t>(a.hstride) * H * 4 : static_cast<uint64_t>(a.sstride) * 4;
  const uint64_t strides[3] = {static_cast<uint64_t>(a.wstride) * 4, static_cast<uint64_t>(a.hstride) * 4, sbytes};
  const uint32_t box[4] = {32, static_cast<uint32_t>(box_w), static_cast<uint32_t>(box_h), 1};
  return tma_encode(m, a.ptr, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
}

}  // namespace tc3
}  // namespace mfvi

using namespace mfvi;

// One launch: output channels [co0, co0 + Cout) of the layer, dy already sliced to them.
static int wgrad_alias_launch(const MfviConvDesc* d_layer, int Cout, int co0, MfviView x, MfviView dy, float* dw, long long w_sstride,
                              mfvi_stream_t st) {
  using namespace mfvi::tc3;
  MfviConvDesc dd = *d_layer;
  dd.Cout = Cout;
  const MfviConvDesc* d = &dd;
  const bool k3 = d->KH == 3 && d->KW == 3;
  if (!view_ok(dy, d->Cout)) return -1;
  WArgs a{};
  a.Cout = d->Cout; a.Cin = d->Cin; a.K3 = k3 ? 1 : 0;
  a.Cout_total = d_layer->Cout; a.co0 = co0;
  a.n_cblk = cdiv(d->Cin, 32);
  a.Ho = d->Hout; a.Wo = d->Wout;
  const int BN = k3 ? 96 : 128;
  if (a.n_cblk * BN > 512) return -1;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(a.n_cblk * BN)) cols <<= 1;
  a.tmem_cols = cols;
  // ---- tile: full-width rows when they fit a TMA box, else equal strips; TH as large as the 2-stage shared-memory budget allows
  const int halo = k3 ? 2 : 0;
  int strips = 1;
  while (cdiv(d->Wout, strips) + halo > 256 || (k3 && a.n_cblk > 1 && cdiv(d->Wout, strips) > 64) || cdiv(d->Wout, strips) > 128) ++strips;
  if (const int f = env_int("MFVI_WGRAD2_STRIPS", 0)) strips = f;
  a.n_stages = 2;
  const size_t budget = 190 * 1024;
  int best_th = 0;
  for (;; ++strips) {              // narrower strips when even a one-row tile of 5 channel blocks exceeds the two-stage budget
  a.TW = cdiv(d->Wout, strips);
  a.Pw = a.TW + halo;
  a.tiles_w = cdiv(d->Wout, a.TW);
  best_th = 0;
  for (int TH = 1; TH <= std::min(d->Hout, k3 ? 8 : 64); ++TH) {
    if (a.TW > 256 || TH + halo > 256) break;
    int n_k, dy_rows, x_rows;
    if (k3) {
      n_k = cdiv((TH - 1) * a.Pw + a.TW + 2, 8);
      dy_rows = rup(8 * n_k + 8, 8);
      x_rows = rup(std::max((TH + 2) * a.Pw, 8 * n_k + 2 * a.Pw + 8), 8);
    } else {
      n_k = cdiv(TH * a.TW, 32);
      dy_rows = x_rows = 32 * n_k;
    }
    const size_t stage = static_cast<size_t>(dy_rows + a.n_cblk * x_rows) * 128;
    if (1024 + 2 * stage + 256 > budget) break;
    // keep enough tiles for all SMs
    const int tiles = d->S * cdiv(d->Hout, TH) * a.tiles_w;
    if (best_th > 0 && tiles < kNumSMs) break;
    best_th = TH;
  }
  if (best_th > 0 || strips >= 4 || a.TW <= 16) break;
  }
  if (const int f = env_int("MFVI_WGRAD2_TH", 0)) best_th = f;
  if (best_th == 0) return -1;
  a.TH = best_th;
  if (k3) {
    a.n_k = cdiv((a.TH - 1) * a.Pw + a.TW + 2, 8);
    a.k_rows = 8;
    a.dy_rows = rup(8 * a.n_k + 8, 8);
    a.x_rows = rup(std::max((a.TH + 2) * a.Pw, 8 * a.n_k + 2 * a.Pw + 8), 8);
  } else {
    a.n_k = cdiv(a.TH * a.TW, 32);
    a.k_rows = 32;
    a.dy_rows = a.x_rows = 32 * a.n_k;
  }
  a.stage_bytes = static_cast<uint32_t>(a.dy_rows + a.n_cblk * a.x_rows) * 128u;
  a.tiles_h = cdiv(d->Hout, a.TH);
  a.tiles_per_sample = a.tiles_h * a.tiles_w;
  a.ctas_per_sample = std::max(1, std::min(a.tiles_per_sample, kNumSMs / d->S));
  a.tiles_per_cta = cdiv(a.tiles_per_sample, a.ctas_per_sample);
  a.ctas_per_sample = cdiv(a.tiles_per_sample, a.tiles_per_cta);
  a.x_bcast = (x.sstride == 0 || d->S == 1) ? 1 : 0;
  a.dw = dw; a.w_sstride = w_sstride;
  if (dy.sstride == 0 && d->S > 1) return -1;
  CUtensorMap tmDy, tmX;
  if (!encode_act(&tmDy, dy, d->Cout, d->Hout, d->Wout, d->S, d->S == 1, a.TW, k3 ? 1 : a.TH)) return -1;
  if (!encode_act(&tmX, x, d->Cin, d->Hin, d->Win, d->S, a.x_bcast != 0, a.Pw, a.TH + halo)) return -1;
  const size_t smem = 1024 + static_cast<size_t>(a.n_stages) * a.stage_bytes + 256;
  static unsigned long long attr_done = 0;
  if (dry_run() == nullptr) {
    const cudaError_t e = allow_dyn_smem(k_wgrad_alias, 200 * 1024, &attr_done);
    MFVI_REQUIRE(e == cudaSuccess, "conv2d_wgrad_tc2: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  MFVI_REQUIRE(smem <= 200 * 1024, "conv2d_wgrad_tc2: stage does not fit in shared memory");
  if (env_int("MFVI_TC2_VERBOSE", 0))
    fprintf(stderr, "[wgrad2] %d->%d k%d %dx%d S=%d: TH=%d TW=%d Pw=%d n_k=%d cblk=%d stage=%u grid=%d tiles/cta=%d\n", d->Cin, d->Cout, d->KH,
            d->Hout, d->Wout, d->S, a.TH, a.TW, a.Pw, a.n_k, a.n_cblk, a.stage_bytes, d->S * a.ctas_per_sample, a.tiles_per_cta);
  dry_detail("TH=%d TW=%d Pw=%d n_k=%d cblk=%d stage_bytes=%u tmem_cols=%u tiles_per_cta=%d", a.TH, a.TW, a.Pw, a.n_k, a.n_cblk,
             a.stage_bytes, a.tmem_cols, a.tiles_per_cta);
  launch_k(k_wgrad_alias, d->S * a.ctas_per_sample, kThreads, smem, as_stream(st), tmDy, tmX, a);
  return check_launch("conv2d_wgrad_tc2");
}

extern "C" {

// dw only (the bias gradient is left to the caller).  Returns -1 when the shape is not taken.  Layers with 64 output channels
// (MFVI_WGRAD2_SPLIT = the largest number of 32-channel launches taken, default 2) run as one launch per 32 output channels
// over channel slices of dy.
int mfvi_conv2d_wgrad_tc2(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, long long w_sstride, mfvi_stream_t st) {
  using namespace mfvi::tc3;
  static const bool on = env_int("MFVI_WGRAD2", 1) != 0;
  static const int max_parts = env_int("MFVI_WGRAD2_SPLIT", 2);
  const bool k3 = d->KH == 3 && d->KW == 3, k1 = d->KH == 1 && d->KW == 1;
  // more than one launch only where the per-tap kernel is the slow one: long contractions (Cin x taps >= 512) over >= 4096
  // pixels of >= 4 samples (measured: 132->64 at 64^2, S = 8: 52.6 -> 41.2 us; with one sample per GPU the per-tap kernel wins)
  const bool big = d->Hout * d->Wout >= 4096 && d->Cin * d->KH * d->KW >= 512 && d->S >= 4;
  const int parts = d->Cout <= 32 ? 1 : ((d->Cout % 32 == 0 && big) ? d->Cout / 32 : 0);
  if (!on || d->stride != 1 || !(k3 || k1) || parts < 1 || parts > max_parts || d->Cin > 32 * kMaxCblk || !view_ok(x, d->Cin) ||
      (reinterpret_cast<uintptr_t>(dw) % 16) || (w_sstride % 4) || d->Cin % 4)
    return -1;
  for (int h = 0; h < parts; ++h) {
    MfviView dyh = dy;
    dyh.ptr = dy.ptr + 32 * h;
    const int rc = wgrad_alias_launch(d, parts == 1 ? d->Cout : 32, 32 * h, x, dyh, dw, w_sstride, st);
    if (rc != 0) return h == 0 ? rc : (rc < 0 ? 1 : rc);      // a later slice cannot be declined once the first one ran
  }
  return 0;
}

}  // extern "C"
