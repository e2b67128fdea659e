// Weight-gradient kernel of conv_tc.cu (tcgen05 kind::tf32, both operands MN-major).  The BF16 template parameter is the
// kind::f16 variant measured in round 2 (profiles/r02_bf16_mode_bench_negative.json: no faster on this net) — never instantiated.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mfvi {
namespace tc {

constexpr int kBM = 128;        // UMMA_M (cta_group::1)
constexpr int kUmmaK = 8;       // kind::tf32
constexpr int kThreads = 192;

// ---------------------------------------------------------------------------------------------- wgrad
struct TcWgradArgs {
  int Cout, Cin, KW;
  int Ho, Wo;           // dy spatial size
  int TH, TW, TP;       // pixel tile (K chunk) TP = TH*TW, multiple of 8
  int tiles_w, n_tiles, tiles_per_cta;
  int MB, NB;           // 32-channel blocks of dy (M) and x (N)
  int cstride;          // conv stride; 2 = x is read through the parity-split 5-D map
  int x_rows2;          // Hin/2
  int x_bcast;
  int sgrp;             // > 0: "sample-blocked" mode — x is ONE image shared by all samples and Cout <= 32, so each of the four
                        // 32-row blocks of the M = 128 operand holds the dy of a different sample (rows >= Cout zero-filled by
                        // TMA) and a CTA serves `sgrp` = 4 samples at once: 4x fewer MMAs and x loads (first-layer wgrads)
  int stages;
  uint32_t tmem_cols;
  float* dw;            // [S][taps][Cout][Cin]
  long long w_sstride;
};

// BF16: both operands are bf16 (kind::f16).  The kernel is written in bytes — a block is TP pixel rows of 128 bytes — so only
// the channels per row (64 instead of 32), the MN-major layout (plain SWIZZLE_128B, 8-row k groups, SBO 1024) and the rows per
// MMA (16 instead of 8) change; MB / NB then count 64-channel blocks and M = 128 is two of them.
template <bool BF16 = false>
__global__ void __launch_bounds__(kThreads)
k_wgrad_tc(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const TcWgradArgs p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t blk_bytes = static_cast<uint32_t>(p.TP) * 128u;          // one 32-channel block of TP pixel rows
  // M = 128 reads four 32-channel blocks of dy.  With Cout <= 32 only one exists: the descriptor's block stride (LBO) is then 0, so
  // the other three alias it (their output rows are duplicates that the epilogue drops) and the stage holds ONE dy block —
  // which lets the pixel tile (the K chunk per TMA round trip) be 4x longer for the latency-bound small-channel layers
  const uint32_t a_blocks = p.MB == 1 ? 1u : (BF16 ? 2u : 4u);
  const uint32_t a_bytes = a_blocks * blk_bytes;
  const uint32_t stage_bytes = a_bytes + static_cast<uint32_t>(p.NB) * blk_bytes;
  uint8_t* ctrl = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = full_bar + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 17);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for ptxas
  const int smp = p.sgrp > 0 ? blockIdx.z * p.sgrp : blockIdx.z, tap = blockIdx.y;     // (first) sample of this CTA
  const int r = tap / p.KW, s = tap % p.KW;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.n_tiles);
  const int n_iters = t_end - t_begin;
  const int BN = p.NB * (BF16 ? 64 : 32);
  constexpr int kCb = BF16 ? 64 : 32;          // channels of one 128-byte row

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
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
  if (n_iters <= 0) {           // nothing to do (uniform per CTA)
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < n_iters; ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(smem_u32(&empty_bar[st]), ph ^ 1);
        const uint32_t fb = smem_u32(&full_bar[st]);
        mbar_expect_tx(fb, static_cast<uint32_t>(p.MB + p.NB) * blk_bytes);
        const int t = t_begin + it;
        const int h0 = (t / p.tiles_w) * p.TH, w0 = (t % p.tiles_w) * p.TW;
        const uint32_t a_dst = smem_u32(smem + static_cast<size_t>(st) * stage_bytes);
        const uint32_t b_dst = a_dst + a_bytes;
        if (p.sgrp > 0) {
          for (int j = 0; j < p.sgrp; ++j) tma_load_4d(a_dst + j * blk_bytes, &tmDy, fb, 0, w0, h0, smp + j);   // block j = sample smp+j
        } else {
          for (int j = 0; j < p.MB; ++j) tma_load_4d(a_dst + j * blk_bytes, &tmDy, fb, kCb * j, w0, h0, smp);
        }
        if (p.cstride == 2) {
          const int rows = p.x_bcast ? 0 : smp * p.x_rows2;
          for (int j = 0; j < p.NB; ++j)
            tma_load_5d(b_dst + j * blk_bytes, &tmX, fb, kCb * j, s & 1, w0 + (s >> 1), r & 1, rows + h0 + (r >> 1));
        } else {
          for (int j = 0; j < p.NB; ++j) tma_load_4d(b_dst + j * blk_bytes, &tmX, fb, kCb * j, w0 + s, h0 + r, p.x_bcast ? 0 : smp);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(kBM, BN, 1, 1, BF16 ? kFmtBF16 : kFmtTF32);
    const int ksteps = p.TP / (BF16 ? 16 : kUmmaK);
    for (int it = 0; it < n_iters; ++it) {
      const int st = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(smem_u32(&full_bar[st]), ph);
      tc_fence_after();
      {
        const uint32_t a_addr = smem_u32(smem + static_cast<size_t>(st) * stage_bytes);
        const uint32_t b_addr = a_addr + a_bytes;
        const uint32_t hi = BF16 ? desc_hi(1024, kLayoutSw128) : desc_hi(512, kLayoutSw128Base32);
        constexpr uint32_t k_adv = BF16 ? 128u : 64u;          // one MMA's k rows (16 / 8) x 128 bytes >> 4
        uint32_t a_lo = desc_lo(a_addr, p.MB == 1 ? 0u : blk_bytes), b_lo = desc_lo(b_addr, blk_bytes);
        for (int k = 0; k < ksteps; ++k, a_lo += k_adv, b_lo += k_adv) {
          if (BF16) tc_mma_f16_elect(tmem_base, desc_pack(a_lo, hi), desc_pack(b_lo, hi), idesc, (it > 0 || k > 0) ? 1u : 0u);
          else tc_mma_tf32_elect(tmem_base, desc_pack(a_lo, hi), desc_pack(b_lo, hi), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit_elect(smem_u32(&empty_bar[st]));
        if (it == n_iters - 1) tc_commit_elect(smem_u32(tmem_full_bar));
      }
    }
  } else {
    const int q = warp & 3;
    // accumulator row -> (sample, output channel): one sample per 32-row block in sample-blocked mode
    const int osmp = p.sgrp > 0 ? smp + q : smp;
    const int co = p.sgrp > 0 ? lane : q * 32 + lane;
    mbar_wait(smem_u32(tmem_full_bar), 0);
    tc_fence_after();
    float* dst = p.dw + static_cast<size_t>(osmp) * p.w_sstride + (static_cast<size_t>(tap) * p.Cout + co) * p.Cin;
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
      if (co < p.Cout) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (c + j + 3 < p.Cin) {
            atomicAdd(reinterpret_cast<float4*>(dst + c + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
            for (int jj = j; jj < j + 4; ++jj)
              if (c + jj < p.Cin) atomicAdd(dst + c + jj, v[jj]);
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


}  // namespace tc
}  // namespace mfvi
