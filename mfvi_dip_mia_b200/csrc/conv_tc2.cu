// "Halo" implicit-GEMM convolution on tcgen05 / TMEM / TMA for sm_100a (kind::tf32, fp32 accumulate in tensor memory):
// forward and data-gradient of the stride-1 sampled-weight convolutions.
//
// conv_tc.cu loads one TMA box per filter tap, so a 3x3 layer streams its input nine times from L2 into shared memory and
// the 16..32-channel full-resolution layers (most of the step's bytes) are bound by that traffic and by per-CTA latency.
// Here a CTA loads the (TH+KH-1) x (TW+KW-1) input halo of a TH x TW output tile ONCE per 32-channel chunk, as one dense
// box whose rows are pixels (pitch Pw = TW+KW-1 pixels) and whose 128/64/32-byte rows are channels.  In that "flat" pixel
// space output position q = hl*Pw + wl reads, for tap (r,s), halo row q + r*Pw + s: every tap is the SAME operand shifted
// by a constant number of rows, and a UMMA shared-memory descriptor may start at any row because the hardware swizzle is a
// function of the absolute address (probe: scripts/umma_shift_test.cu, profiles/r01_umma_row_shift_probe.txt).  The
// 128-row M tiles therefore tile the flat space; the KW-1 junk positions at the end of each tile row are computed and dropped.
//
//   forward : D[q][co] += sum_tap  X[q + r*Pw + s][ci-chunk] * W[tap][co][ci-chunk]^T        A K-major, B K-major
//   dgrad   : D[q][ci] += sum_tap dY[q + (KH-1-r)*Pw + (KW-1-s)][co-chunk] * W[tap][co-chunk][ci]   A K-major, B MN-major
//             (the halo box starts at (h0-KH+1, w0-KW+1); TMA zero-fills outside dY)
//
// Channel chunks are 32 fp32 (SWIZZLE_128B) plus one narrower tail chunk of 8 or 16 (SWIZZLE_32B / 64B), so Cin = 36 costs
// 40 channels of traffic, not 64.  Persistent CTAs walk a static tile list; 7 warps: 0 = activation TMA producer,
// 1 = TMEM owner + single-thread MMA issuer, 2 = weight TMA producer, 3..6 = epilogue (TMEM -> registers -> global,
// fused bias and per-sample BatchNorm (sum, sumsq) in double).  Accumulators are double-buffered in TMEM when they fit, so
// the epilogue of tile i overlaps the loads and MMAs of tile i+1.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace mfvi {
namespace tc2 {
using namespace mfvi::tc;

constexpr int kThreads = 224;
constexpr int kMaxChunks = 9;
constexpr int kMaxStages = 8;
constexpr int kTrLd = 17;

struct Args {
  int n_chunks;
  int ck0[kMaxChunks], cw[kMaxChunks];   // first channel and width (8/16/32) of every contraction chunk
  int N, BN, n_nb;                       // valid output channels, UMMA N per CTA, number of N blocks
  int KH, KW, taps, g, n_groups;         // g = taps per weight stage
  int Mh, Mw, TH, TW, Pw, tiles_h, tiles_w;
  int n_mt, total_tiles, tiles_per_sample;
  int box_rows;                          // (TH+hh) * Pw : rows of one halo box
  int dgrad, a_bcast, b_bcast;
  int cstride;                           // conv stride (1 or 2)
  int hh, hw;                            // halo rows / columns of the box beyond the TH x TW tile
  int org_h, org_w;                      // the box starts at (tile origin - org): dgrad reaches up / left
  int planes, plane_rows;                // stride-2 forward: the input is read as 4 parity planes (row, col parity), each a halo
                                         // box of its own, plane_rows shared-memory rows apart
  int a_rows2;                           // stride-2 forward: Hin/2 (rows of one sample in the parity-split 5-D map)
  int n_cls;                             // stride-2 dgrad: 4 output-parity classes, each a stride-1 problem over a tap subset
  int tap_off[4][25];                    // [class][tap] -> operand row offset of the tap inside the A stage, -1 = tap not in class
  int n_a, n_b, acc_stages;
  uint32_t a_stage_bytes, b_stage_bytes, tmem_cols;
  MfviView o;
  const float* bias;
  long long bias_sstride;
  double* stats;
  int accumulate, vecO;
  long long* dbg;     // optional timeline of CTA 0 (clock64 stamps), 8 slots per tile
  int dbg_mode;       // timing experiments only (results are garbage): 1 = no epilogue work, 2 = no TMA after the first stages, 3 = both
};

#define TC2_STAMP(slot) do { if (p.dbg != nullptr && blockIdx.x == 0) p.dbg[(tile_i) * 8 + (slot)] = clock64(); } while (0)

struct TileCoord {
  int smp, nb, cls, th, tw;
};
__device__ __forceinline__ TileCoord decode_tile(const Args& p, int t) {
  TileCoord c;
  c.smp = t / p.tiles_per_sample;
  int r = t - c.smp * p.tiles_per_sample;
  const int per_cls = p.tiles_h * p.tiles_w, per_nb = per_cls * p.n_cls;
  c.nb = r / per_nb;
  r -= c.nb * per_nb;
  c.cls = r / per_cls;
  r -= c.cls * per_cls;
  c.th = r / p.tiles_w;
  c.tw = r - c.th * p.tiles_w;
  return c;
}

// BF16 = operands are bf16 (tcgen05 kind::f16, K = 16 channels per 32-byte k-step).  The kernel is written in BYTES — chunk widths
// `cw` count 4-byte words of an operand row (8 / 16 / 32 words = 32 / 64 / 128-byte rows), `ck0` counts elements (TMA coordinates)
// — so the K-major operands keep every descriptor constant; only the MMA kind, the instruction descriptor and the MN-major weight
// operand of the data gradient differ (64-element atoms of plain SWIZZLE_128B instead of the 32-element atoms tf32 needs).
template <bool BF16>
__device__ __forceinline__ void mma_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (BF16) tc_mma_f16_elect(d_tmem, adesc, bdesc, idesc, accumulate);
  else tc_mma_tf32_elect(d_tmem, adesc, bdesc, idesc, accumulate);
}

template <bool DGRAD, bool BF16 = false>
__global__ void __launch_bounds__(kThreads, 2)
k_conv_halo(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAt,
            const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBt, const Args p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_stages = smem;
  uint8_t* b_stages = a_stages + static_cast<size_t>(p.n_a) * p.a_stage_bytes;
  uint8_t* ctrl = b_stages + static_cast<size_t>(p.n_b) * p.b_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* a_empty = a_full + kMaxStages;
  uint64_t* b_full = a_empty + kMaxStages;
  uint64_t* b_empty = b_full + kMaxStages;
  uint64_t* acc_full = b_empty + kMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* tr_all = reinterpret_cast<float*>(ctrl + 512);      // [4 warps][32][kTrLd]

  // warp index through a shuffle: ptxas then treats it (and every role branch on it) as warp-uniform and keeps the UMMA /
  // TMA operands in uniform registers instead of R2UR-ing them before every instruction
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmAt);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmBt);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(smem_u32(&a_full[i]), 1);
      mbar_init(smem_u32(&a_empty[i]), 1);
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_cols = static_cast<uint32_t>(p.n_mt * p.BN);
  pdl_wait();          // prologue above overlapped the previous kernel; global memory is touched only from here on

  if (warp == 0) {
    // ===== activation producer: one halo box per (tile, chunk)
    if (lane == 0) {
      uint32_t ia = 0;
      int tile_i = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_i) {
        const TileCoord tc = decode_tile(p, t);
        const int h0 = tc.th * p.TH - p.org_h;
        const int w0 = tc.tw * p.TW - p.org_w;
        for (int c = 0; c < p.n_chunks; ++c, ++ia) {
          const uint32_t st = ia % p.n_a, ph = (ia / p.n_a) & 1;
          mbar_wait(smem_u32(&a_empty[st]), ph ^ 1);
          if (c == 0) TC2_STAMP(0);
          const uint32_t fb = smem_u32(&a_full[st]);
          if ((p.dbg_mode & 2) && ia >= static_cast<uint32_t>(p.n_a)) { mbar_arrive(fb); continue; }
          const uint32_t box_bytes = static_cast<uint32_t>(p.box_rows) * p.cw[c] * 4u;
          const uint32_t dst = smem_u32(a_stages + static_cast<size_t>(st) * p.a_stage_bytes);
          const CUtensorMap* map = p.cw[c] == 32 ? &tmA : &tmAt;
          if (p.planes == 1) {
            mbar_expect_tx(fb, box_bytes);
            tma_load_4d(dst, map, fb, p.ck0[c], w0, h0, p.a_bcast ? 0 : tc.smp);
          } else {
            // parity-split 5-D view (C, 2, W/2, 2, H/2 * S): plane (pr, ps) holds pixels (2i + pr, 2j + ps)
            mbar_expect_tx(fb, 4u * box_bytes);
            const int rows = (p.a_bcast ? 0 : tc.smp * p.a_rows2) + h0;
            const uint32_t pl_bytes = static_cast<uint32_t>(p.plane_rows) * p.cw[c] * 4u;
            for (int pl = 0; pl < 4; ++pl) tma_load_5d(dst + pl * pl_bytes, map, fb, p.ck0[c], pl & 1, w0, pl >> 1, rows);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===== weight producer: g taps of one chunk per stage
    if (lane == 0) {
      uint32_t ib = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int bs = p.b_bcast ? 0 : tc.smp;
        const int n0 = tc.nb * p.BN;
        for (int c = 0; c < p.n_chunks; ++c) {
          const CUtensorMap* map = p.cw[c] == 32 ? &tmB : &tmBt;
          for (int grp = 0; grp < p.n_groups; ++grp, ++ib) {
            const uint32_t st = ib % p.n_b, ph = (ib / p.n_b) & 1;
            mbar_wait(smem_u32(&b_empty[st]), ph ^ 1);
            const uint32_t fb = smem_u32(&b_full[st]);
            const uint32_t dst = smem_u32(b_stages + static_cast<size_t>(st) * p.b_stage_bytes);
            if ((p.dbg_mode & 2) && ib >= static_cast<uint32_t>(p.n_b)) { mbar_arrive(fb); continue; }
            if (!DGRAD) {
              mbar_expect_tx(fb, static_cast<uint32_t>(p.g * p.BN * p.cw[c]) * 4u);
              tma_load_4d(dst, map, fb, p.ck0[c], n0, grp * p.g, bs);                      // box (w k, BN n, g taps)
            } else {
              // one 128-byte row of n per k row: 32 (tf32) / 64 (bf16) input channels; a chunk has w (tf32) / 2w (bf16) k rows
              if (!BF16) {
                const uint32_t blk = static_cast<uint32_t>(p.g * p.cw[c]) * 128u;            // one 32-wide n block
                mbar_expect_tx(fb, blk * static_cast<uint32_t>(p.BN / 32));
                for (int j = 0; j < p.BN / 32; ++j)
                  tma_load_4d(dst + j * blk, map, fb, n0 + 32 * j, p.ck0[c], grp * p.g, bs);  // box (32 n, w k, g taps)
              } else {
                const uint32_t blk = static_cast<uint32_t>(p.g * p.cw[c]) * 256u;            // one 64-wide n block
                mbar_expect_tx(fb, blk * static_cast<uint32_t>(p.BN / 64));
                for (int j = 0; j < p.BN / 64; ++j)
                  tma_load_4d(dst + j * blk, map, fb, n0 + 64 * j, p.ck0[c], grp * p.g, bs);  // box (64 n, 2w k, g taps)
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t idesc = make_idesc(128, p.BN, 0, DGRAD ? 1 : 0, BF16 ? kFmtBF16 : kFmtTF32);
    uint32_t ia = 0, ib = 0, it = 0;
    int tile_i = 0;
    long long cyc_wait_a = 0, cyc_wait_b = 0, cyc_issue = 0, n_mma = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it, ++tile_i) {
      const uint32_t as = it % p.acc_stages, aph = (it / p.acc_stages) & 1;
      mbar_wait(smem_u32(&acc_empty[as]), aph ^ 1);
      tc_fence_after();
      if (lane == 0) TC2_STAMP(1);
      const uint32_t d_base = tmem_base + as * acc_cols;
      const int tcls = decode_tile(p, t).cls;
      bool started = false;                            // the first MMA of every accumulator overwrites, the rest accumulate
      for (int c = 0; c < p.n_chunks; ++c, ++ia) {
        const uint32_t sa = ia % p.n_a;
        long long tw0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&a_full[sa]), (ia / p.n_a) & 1);
        tc_fence_after();
        if (p.dbg) cyc_wait_a += clock64() - tw0;
        if (lane == 0 && c == 0) TC2_STAMP(2);
        const int w = p.cw[c];
        const uint32_t rb = static_cast<uint32_t>(w) * 4u, sbo = 8u * rb, layout = kmajor_layout(w);
        const int ksteps = w / 8;
        const uint32_t a_base = smem_u32(a_stages + static_cast<size_t>(sa) * p.a_stage_bytes);
        for (int grp = 0; grp < p.n_groups; ++grp, ++ib) {
          const uint32_t sb = ib % p.n_b;
          long long tb0 = p.dbg ? clock64() : 0;
          mbar_wait(smem_u32(&b_full[sb]), (ib / p.n_b) & 1);
          tc_fence_after();
          long long ti0 = p.dbg ? clock64() : 0;
          if (p.dbg) cyc_wait_b += ti0 - tb0;
          {
            const uint32_t b_base = smem_u32(b_stages + static_cast<size_t>(sb) * p.b_stage_bytes);
            // descriptors: hi words fixed per chunk, lo words advance by (bytes >> 4)
            const uint32_t a_hi = desc_hi(sbo, layout);
            // MN-major weight operand of the data gradient: rows = k (output channels), 128 bytes of n per row, n blocks LBO
            // apart; tf32: 32-element atoms, 4-row k groups (SBO 512), 8 rows per k-step; bf16: 64-element atoms of plain
            // SWIZZLE_128B, 8-row k groups (SBO 1024), 16 rows per k-step
            const uint32_t b_hi = DGRAD ? (BF16 ? desc_hi(1024, kLayoutSw128) : desc_hi(512, kLayoutSw128Base32)) : a_hi;
            const uint32_t b_lo0 = DGRAD ? desc_lo(b_base, static_cast<uint32_t>(p.g * w) * (BF16 ? 256u : 128u)) : desc_lo(b_base, 16);
            const uint32_t b_tap = DGRAD ? static_cast<uint32_t>(w) * (BF16 ? 16u : 8u) : (static_cast<uint32_t>(p.BN) * rb) >> 4;   // per tap
            constexpr uint32_t b_k = DGRAD ? (BF16 ? 128u : 64u) : 2u;                                                // per k-step
            const uint32_t a_lo0 = desc_lo(a_base, 16);
            const uint32_t a_mt = (128u * rb) >> 4, rb16 = rb >> 4;
            const int tap0 = grp * p.g;
            const int ntap = min(p.g, p.taps - tap0);
            // One specialised tap loop per chunk width (ksteps = 4 / 2 / 1): the single issuing thread spends most of its time
            // in uniform-datapath bookkeeping, so everything loop-invariant (n_mt, BN, descriptor strides) sits in registers,
            // the weight descriptor advances incrementally, and the tap's operand offset — a kernel-parameter table read with
            // a run-time index in front of a branch — is fetched one tap ahead.
            const int n_mt = p.n_mt;
            const uint32_t BNu = static_cast<uint32_t>(p.BN);
            int off_next = p.tap_off[tcls][tap0];
            uint32_t b_t = b_lo0;
            auto taps = [&](auto ks_tag) {
              constexpr int KS = decltype(ks_tag)::value;
              for (int tt = 0; tt < ntap; ++tt, b_t += b_tap) {
                const int off_rows = off_next;
                off_next = p.tap_off[tcls][tap0 + (tt + 1 < ntap ? tt + 1 : tt)];
                if (off_rows < 0) continue;          // tap belongs to another output-parity class (stride-2 dgrad)
                uint32_t al = a_lo0 + static_cast<uint32_t>(off_rows) * rb16;
                const uint32_t acc0 = started ? 1u : 0u;
                started = true;
                uint32_t d_col = d_base;
                for (int jm = 0; jm < n_mt; ++jm, al += a_mt, d_col += BNu) {
                  mma_elect<BF16>(d_col, desc_pack(al, a_hi), desc_pack(b_t, b_hi), idesc, acc0);
                  if (KS >= 2) mma_elect<BF16>(d_col, desc_pack(al + 2u, a_hi), desc_pack(b_t + b_k, b_hi), idesc, 1u);
                  if (KS >= 4) {
                    mma_elect<BF16>(d_col, desc_pack(al + 4u, a_hi), desc_pack(b_t + 2u * b_k, b_hi), idesc, 1u);
                    mma_elect<BF16>(d_col, desc_pack(al + 6u, a_hi), desc_pack(b_t + 3u * b_k, b_hi), idesc, 1u);
                  }
                }
              }
            };
            if (ksteps == 4) taps(std::integral_constant<int, 4>{});
            else if (ksteps == 2) taps(std::integral_constant<int, 2>{});
            else taps(std::integral_constant<int, 1>{});
            tc_commit_elect(smem_u32(&b_empty[sb]));
            if (p.dbg) { cyc_issue += clock64() - ti0; n_mma += static_cast<long long>(p.g) * p.n_mt * ksteps; }
          }
        }
        tc_commit_elect(smem_u32(&a_empty[sa]));
      }
      tc_commit_elect(smem_u32(&acc_full[as]));
      if (lane == 0) TC2_STAMP(3);
    }
    if (p.dbg != nullptr && blockIdx.x == 0 && lane == 0) {
      p.dbg[63 * 8 + 0] = cyc_wait_a; p.dbg[63 * 8 + 1] = cyc_wait_b; p.dbg[63 * 8 + 2] = cyc_issue; p.dbg[63 * 8 + 3] = n_mma;
    }
  } else {
    // ===== epilogue warps 3..6: TMEM lane quarter = warp % 4.  Lane l owns flat position (j*128 + q*32 + l) of M tile j.
    const int q = warp & 3;
    uint32_t it = 0;
    int tile_i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it, ++tile_i) {
      const TileCoord tc = decode_tile(p, t);
      const uint32_t as = it % p.acc_stages;
      if (warp == 3 && lane == 0) TC2_STAMP(4);
      mbar_wait(smem_u32(&acc_full[as]), (it / p.acc_stages) & 1);
      tc_fence_after();
      if (warp == 3 && lane == 0) TC2_STAMP(5);
      const int n0 = tc.nb * p.BN;
      const float* bias = p.bias != nullptr ? p.bias + static_cast<size_t>(tc.smp) * p.bias_sstride + n0 : nullptr;
      float* obase = p.o.ptr + static_cast<size_t>(tc.smp) * p.o.sstride + n0;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * acc_cols;
      const int ostr = p.n_cls == 4 ? 2 : 1;
#pragma unroll 1
      for (int c = 0; c < ((p.dbg_mode & 1) ? 0 : p.BN); c += 16) {
        float b[16], s1[16], s2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          b[i] = (bias != nullptr && n0 + c + i < p.N) ? __ldg(bias + c + i) : 0.f;
          s1[i] = 0.f;
          s2[i] = 0.f;
        }
        if (n0 + c >= p.N) break;          // column groups beyond the valid channels (N padded up to the UMMA granule)
        const bool full16 = n0 + c + 15 < p.N;
        // one M tile: bias, BatchNorm partials, store.  The TMEM load of the NEXT M tile is in flight meanwhile.
        auto consume = [&](const uint32_t (&r)[16], int j) {
          const int pos = j * 128 + q * 32 + lane;
          const int hl = pos / p.Pw, wl = pos - hl * p.Pw;
          // stride-2 dgrad: this tile is output-parity class (cls >> 1, cls & 1) of dx, written with pixel stride 2
          const int gh = (tc.th * p.TH + hl) * ostr + (tc.cls >> 1), gw = (tc.tw * p.TW + wl) * ostr + (tc.cls & 1);
          const bool valid = hl < p.TH && wl < p.TW && gh < p.Mh && gw < p.Mw;
          if (valid) {
            float v[16];
            float* optr = obase + static_cast<size_t>(gh) * p.o.hstride + static_cast<size_t>(gw) * p.o.wstride + c;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              v[i] = __uint_as_float(r[i]);
              if (!DGRAD) {                      // bias and BatchNorm partials exist in the forward only
                v[i] += b[i];
                s1[i] += v[i];
                s2[i] = fmaf(v[i], v[i], s2[i]);
              }
            }
            if (p.vecO && full16) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                float4 o4 = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                float4* dst = reinterpret_cast<float4*>(optr + i);
                if (p.accumulate) {
                  const float4 old = *dst;
                  o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w;
                }
                *dst = o4;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n0 + c + i < p.N) optr[i] = p.accumulate ? optr[i] + v[i] : v[i];
            }
          }
        };
        uint32_t ra[16], rb[16];
        tmem_ld16_issue(tbase + static_cast<uint32_t>(c), ra);
#pragma unroll 1
        for (int j = 0; j < p.n_mt; j += 2) {
          tmem_ld16_wait(ra);
          if (j + 1 < p.n_mt) tmem_ld16_issue(tbase + static_cast<uint32_t>((j + 1) * p.BN + c), rb);
          consume(ra, j);
          if (j + 1 < p.n_mt) {
            tmem_ld16_wait(rb);
            if (j + 2 < p.n_mt) tmem_ld16_issue(tbase + static_cast<uint32_t>((j + 2) * p.BN + c), ra);
            consume(rb, j + 1);
          }
        }
        if (!DGRAD && p.stats != nullptr) {
          // (sum, sumsq) of 16 channels: per-lane fp32 partials over <= n_mt rows, then a transposing butterfly across the
          // 32 lanes in double (fixed order).  Lane l ends up with the total of cell l: l < 16 -> sum of channel c+l,
          // l >= 16 -> sum of squares of channel c+l-16.
          double d[16];
          {
            const bool hi = lane & 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const double keep = hi ? static_cast<double>(s2[i]) : static_cast<double>(s1[i]);
              const double send = hi ? static_cast<double>(s1[i]) : static_cast<double>(s2[i]);
              d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
          }
#pragma unroll
          for (int off = 8; off >= 1; off >>= 1) {
            const bool hi = lane & off;
#pragma unroll
            for (int i = 0; i < off; ++i) {
              const double keep = hi ? d[i + off] : d[i];
              const double send = hi ? d[i] : d[i + off];
              d[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
          const int col = n0 + c + (lane & 15);
          if (col < p.N) atomicAdd(p.stats + (static_cast<size_t>(tc.smp) * p.N + col) * 2 + (lane >> 4), d[0]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 3 && lane == 0) TC2_STAMP(6);
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[as]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------- host side
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int rup(int a, int b) { return cdiv(a, b) * b; }

// q = elements per 16 bytes (4 for fp32 views, 8 for bf16 views)
static bool view_ok(const MfviView& v, int C, int q = 4) {
  // TMA needs a 16-byte aligned base and 16-byte multiples for every stride; the channel count itself is free
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && C >= 1 && (v.wstride % q == 0) && (v.hstride % q == 0) &&
         (v.sstride % q == 0) && v.wstride >= C && v.hstride >= v.wstride;
}

static CUtensorMapSwizzle swz_of(int width) {
  return width == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (width == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

struct Plan {
  bool ok = false;
  int TH, TW, Pw, tiles_h, tiles_w, n_mt, BN, n_nb, g, n_groups, n_a, n_b, acc_stages, box_rows, plane_rows, grid;
  uint32_t a_stage, b_stage, tmem_cols;
  size_t smem;
  double cost;
};

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

// (Mh, Mw): the pixel space the M tiles cover (output pixels; for stride-2 dgrad one output-parity class);  (hh, hw): halo
// of the box;  planes: 4 parity planes per stage for the stride-2 forward;  n_cls: 4 parity classes for the stride-2 dgrad
static Plan make_plan(int S, int Mh, int Mw, int hh, int hw, int taps, int planes, int n_cls, int n_chunks, const int* cw,
                      int Nvalid, bool dgrad, bool bf16 = false) {
  Plan best;
  best.cost = 1e30;
  int kw_total = 0, rb_max = 0;
  for (int i = 0; i < n_chunks; ++i) {
    kw_total += cw[i];
    rb_max = std::max(rb_max, cw[i] * 4);
  }
  const int nq = dgrad ? (bf16 ? 64 : 32) : 16;      // N granule: one MN-major atom (dgrad), the UMMA N step (forward)
  const int BN_full = rup(Nvalid, nq);
  const int force_th = env_int("MFVI_TC2_TH", 0), force_strips = env_int("MFVI_TC2_STRIPS", 0), force_bn = env_int("MFVI_TC2_BN", 0);
  for (int split = 1; split <= 8; split *= 2) {
    if (BN_full % split) break;
    const int BN = BN_full / split;
    if (BN % nq || BN > 256) continue;
    if (force_bn && BN != force_bn) continue;
    for (int strips = 1; strips <= 4; ++strips) {
      if (force_strips && strips != force_strips) continue;
      const int TW = cdiv(Mw, strips);
      const int Pw = TW + hw;
      if (Pw > 256) continue;
      if (strips > 1 && cdiv(Mw, TW) != strips) continue;
      for (int TH = 1; TH <= std::min(Mh, 64); ++TH) {
        if (force_th && TH != force_th) continue;
        if (TH + hh > 256) break;
        Plan pl;
        pl.TH = TH; pl.TW = TW; pl.Pw = Pw; pl.BN = BN; pl.n_nb = split;
        pl.tiles_h = cdiv(Mh, TH); pl.tiles_w = strips;
        pl.box_rows = (TH + hh) * Pw;
        pl.n_mt = cdiv((TH - 1) * Pw + TW, 128);
        const int acc_cols = pl.n_mt * BN;
        if (acc_cols > 512) break;
        pl.acc_stages = 2 * acc_cols <= 512 ? 2 : 1;
        uint32_t cols = 32;
        while (cols < static_cast<uint32_t>(pl.acc_stages * acc_cols)) cols <<= 1;
        pl.tmem_cols = cols;
        pl.plane_rows = rup(std::max(pl.box_rows, pl.n_mt * 128 + hh * Pw + hw), 8);
        pl.a_stage = static_cast<uint32_t>(rup(planes * pl.plane_rows * rb_max, 1024));
        int g = taps;
        while (g > 1 && g * BN * rb_max > 32 * 1024) --g;
        while (taps % g) --g;           // equal groups
        pl.g = g; pl.n_groups = taps / g;
        pl.b_stage = static_cast<uint32_t>(rup(g * BN * rb_max, 1024));
        pl.n_a = 2;
        pl.n_b = (pl.n_groups * n_chunks > 1) ? 3 : 2;
        auto smem_of = [&](int na, int nb) {
          return 1024 + static_cast<size_t>(na) * pl.a_stage + static_cast<size_t>(nb) * pl.b_stage + 512 + 4 * 32 * kTrLd * 4 +
                 4 * BN * 16 + 64;
        };
        pl.smem = smem_of(pl.n_a, pl.n_b);
        if (pl.smem > 200 * 1024) break;
        {
          // Latency-bound tiles (few tiles per CTA, many operand stages per tile): every extra stage in flight saves a TMA round
          // trip (~1.5 us).  Deepen the rings up to what a tile consumes while the CTA stays within `deep_cap` bytes
          // (MFVI_TC2_DEEP, default 0 = the two/three-stage rings of round 1).
          static const int deep_cap = env_int("MFVI_TC2_DEEP", 0) * 1024;
          const int want_b = std::min(kMaxStages, pl.n_groups * n_chunks), want_a = std::min(4, n_chunks);
          while (deep_cap > 0) {
            if (pl.n_b < want_b && smem_of(pl.n_a, pl.n_b + 1) <= static_cast<size_t>(deep_cap)) { ++pl.n_b; continue; }
            if (pl.n_a < want_a && smem_of(pl.n_a + 1, pl.n_b) <= static_cast<size_t>(deep_cap)) { ++pl.n_a; continue; }
            break;
          }
          pl.smem = smem_of(pl.n_a, pl.n_b);
        }
        const int tiles = S * split * n_cls * pl.tiles_h * pl.tiles_w;
        const int cpsm = (pl.smem <= 100 * 1024 && cols <= 256) ? 2 : 1;
        const int slots = kNumSMs * cpsm;
        const int waves = cdiv(tiles, slots);
        const double t_load = (static_cast<double>(planes) * pl.box_rows * kw_total * 4 + static_cast<double>(taps) * BN * kw_total * 4) / 40.0 * cpsm;
        // measured (profiles/r01_umma_rate_probe.txt, conv timelines): a kind::tf32 M=128 MMA issues every ~39 cycles up to
        // N=48 and ~N/2 beyond, plus ~6 cycles of loop overhead; an epilogue unit (128 rows x 16 columns) costs ~300 cycles
        // The forward keeps the older model: its constants under-state the MMA and epilogue costs, but it was tuned on
        // measured layer times and its preference for large tiles hides the TMA latency that the two activation stages
        // cannot prefetch across tiles (the "realistic" model picks 1-row tiles for 36->16 at 256^2: 124 us instead of 84).
        // For the data gradient the realistic model wins (36->16: 76 -> 61 us): it knows that a single accumulator stage
        // serialises the wide dgrad epilogue behind the MMAs.
        static const int forced = env_int("MFVI_TC2_COST", -1);
        const int model = forced >= 0 ? forced : (dgrad ? 1 : 0);
        double t_tile;
        if (model == 0) {
          const double t_mma = static_cast<double>(pl.n_mt) * taps * (kw_total / 8) * std::max(BN / 2, 16) * cpsm;
          const double t_epi = static_cast<double>(pl.n_mt) * (BN / 16) * 120.0;
          t_tile = std::max(t_load, std::max(t_mma, t_epi)) + 500.0;
        } else {
          const double mma_cyc = std::max(39.0 + 0.2 * (BN - 16), 0.5 * BN) + 6.0;
          const double t_mma = static_cast<double>(pl.n_mt) * taps * (kw_total / 8) * mma_cyc * cpsm;
          const int n_units = cdiv(std::min(Nvalid, BN), 16);
          const double t_epi = static_cast<double>(pl.n_mt) * n_units * 300.0;
          // with a single accumulator stage the epilogue of tile i cannot overlap the MMAs of tile i+1
          t_tile = (pl.acc_stages == 2 ? std::max(t_load, std::max(t_mma, t_epi)) : std::max(t_load, t_mma + t_epi)) + 500.0;
        }
        pl.cost = 5000.0 + waves * t_tile + std::min(t_load, 4000.0);
        pl.grid = std::min(tiles, slots);
        pl.ok = true;
        if (pl.cost < best.cost) best = pl;
      }
    }
  }
  return best;
}

static int split_chunks(int Kc, int* ck0, int* cw) {
  int n = 0, k = 0;
  while (Kc - k >= 32 && n < kMaxChunks) {
    ck0[n] = k; cw[n] = 32; ++n; k += 32;
  }
  if (Kc - k >= 32) return -1;
  const int rem = Kc - k;
  if (rem > 0) {
    if (n >= kMaxChunks) return -1;
    ck0[n] = k;
    cw[n] = rem <= 8 ? 8 : (rem <= 16 ? 16 : 32);
    ++n;
  }
  return n;
}

// a: the activation read through TMA (x for forward, dy for dgrad); Ca channels, (Ha, Wa) pixels.
// bf16: `a` and `w` hold bf16 elements (view strides, w_sstride and the weight row pitch `w_cpitch` count bf16 elements; the
// weight block is [tap][Cout][w_cpitch] with w_cpitch a multiple of 8 so that every TMA stride is a multiple of 16 bytes); the
// output, bias and statistics stay fp32.
static int launch(const MfviConvDesc* d, bool dgrad, MfviView a, int Ca, int Ha, int Wa, const void* w, long long w_sstride,
                  MfviView o, int Mh, int Mw, int Nvalid, const float* bias, double* stats, int accumulate, mfvi_stream_t st,
                  const char* what, bool bf16 = false, int w_cpitch = 0, long long bias_sstride = -1) {
  Args p{};
  const int esz = bf16 ? 2 : 4, epw = bf16 ? 2 : 1;          // bytes per element, elements per 4-byte word
  if (w_cpitch == 0) w_cpitch = d->Cin;
  p.n_chunks = split_chunks(cdiv(Ca, epw), p.ck0, p.cw);   // chunk widths in words ...
  if (p.n_chunks <= 0) return -1;
  for (int i = 0; i < p.n_chunks; ++i) p.ck0[i] *= epw;      // ... first channel of a chunk in elements
  const int taps = d->KH * d->KW;
  if (taps > 25) return -1;
  const bool s2 = d->stride == 2;
  const bool a_bcast = (a.sstride == 0 || d->S == 1);
  int hh, hw, planes = 1, n_cls = 1, Th_space = Mh, Tw_space = Mw;
  if (!s2) {
    hh = d->KH - 1; hw = d->KW - 1;
  } else {
    if (d->KH < 2 || d->KW < 2) return -1;                 // 1x1 stride 2: parity classes without taps; left to conv_tc.cu
    hh = (d->KH - 1) >> 1; hw = (d->KW - 1) >> 1;
    if (!dgrad) {
      // forward reads 4 parity planes through a 5-D view that folds the sample axis into the row axis
      if ((Ha & 1) || (Wa & 1) || (!a_bcast && a.sstride != static_cast<long long>(Ha) * a.hstride)) return -1;
      planes = 4;
    } else {
      n_cls = 4;
      Th_space = (Mh + 1) / 2; Tw_space = (Mw + 1) / 2;    // largest parity class of dx
    }
  }
  const Plan pl = make_plan(d->S, Th_space, Tw_space, hh, hw, taps, planes, n_cls, p.n_chunks, p.cw, Nvalid, dgrad, bf16);
  if (!pl.ok) return -1;
  p.cstride = d->stride; p.hh = hh; p.hw = hw; p.planes = planes; p.plane_rows = pl.plane_rows; p.n_cls = n_cls;
  p.org_h = dgrad ? hh : 0; p.org_w = dgrad ? hw : 0;
  p.a_rows2 = Ha / 2;
  for (int cls = 0; cls < 4; ++cls)
    for (int tap = 0; tap < 25; ++tap) {
      int off = -1;
      if (tap < taps && cls < n_cls) {
        const int r = tap / d->KW, sx = tap % d->KW;
        if (!s2) {
          off = dgrad ? (d->KH - 1 - r) * pl.Pw + (d->KW - 1 - sx) : r * pl.Pw + sx;
        } else if (!dgrad) {
          off = ((r & 1) * 2 + (sx & 1)) * pl.plane_rows + (r >> 1) * pl.Pw + (sx >> 1);
        } else {
          const int ph = cls >> 1, pw = cls & 1;
          if ((r & 1) == ph && (sx & 1) == pw) off = (hh + (ph - r) / 2) * pl.Pw + (hw + (pw - sx) / 2);
        }
      }
      p.tap_off[cls][tap] = off;
    }
  p.N = Nvalid; p.BN = pl.BN; p.n_nb = pl.n_nb;
  p.KH = d->KH; p.KW = d->KW; p.taps = d->KH * d->KW; p.g = pl.g; p.n_groups = pl.n_groups;
  p.Mh = Mh; p.Mw = Mw; p.TH = pl.TH; p.TW = pl.TW; p.Pw = pl.Pw; p.tiles_h = pl.tiles_h; p.tiles_w = pl.tiles_w;
  p.n_mt = pl.n_mt;
  p.tiles_per_sample = pl.n_nb * n_cls * pl.tiles_h * pl.tiles_w;
  p.total_tiles = d->S * p.tiles_per_sample;
  p.box_rows = pl.box_rows;
  p.dgrad = dgrad ? 1 : 0;
  p.a_bcast = a_bcast ? 1 : 0;
  p.b_bcast = (w_sstride == 0 || d->S == 1) ? 1 : 0;
  p.n_a = pl.n_a; p.n_b = pl.n_b; p.acc_stages = pl.acc_stages;
  p.a_stage_bytes = pl.a_stage; p.b_stage_bytes = pl.b_stage; p.tmem_cols = pl.tmem_cols;
  p.o = o; p.bias = bias; p.bias_sstride = bias_sstride >= 0 ? bias_sstride : w_sstride; p.stats = stats; p.accumulate = accumulate;
  p.vecO = ((reinterpret_cast<uintptr_t>(o.ptr) % 16 == 0) && o.sstride % 4 == 0 && o.hstride % 4 == 0 && o.wstride % 4 == 0) ? 1 : 0;
  if (const char* e = getenv("MFVI_TC2_DBG")) p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));   // device pointer (debug)
  p.dbg_mode = env_int("MFVI_TC2_DBGMODE", 0);

  // ---- tensor maps: wide (32-channel) and tail chunk variants
  int tail_w = 0;
  bool has32 = false;
  for (int i = 0; i < p.n_chunks; ++i) {
    if (p.cw[i] == 32) has32 = true; else tail_w = p.cw[i];
  }
  // `width` = chunk width in words (it selects the swizzle); boxes, dims and coordinates count elements, strides count bytes
  CUtensorMap tmA, tmAt, tmB, tmBt;
  const CUtensorMapDataType dt = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const uint64_t ez = static_cast<uint64_t>(esz);
  auto enc_a = [&](CUtensorMap* m, int width) {
    if (planes == 4) {
      const uint64_t dims[5] = {static_cast<uint64_t>(Ca), 2, static_cast<uint64_t>(Wa / 2), 2,
                                static_cast<uint64_t>(Ha / 2) * (a_bcast ? 1 : d->S)};
      const uint64_t strides[4] = {static_cast<uint64_t>(a.wstride) * ez, static_cast<uint64_t>(a.wstride) * 2 * ez,
                                   static_cast<uint64_t>(a.hstride) * ez, static_cast<uint64_t>(a.hstride) * 2 * ez};
      const uint32_t box[5] = {static_cast<uint32_t>(width * epw), 1, static_cast<uint32_t>(pl.Pw), 1, static_cast<uint32_t>(pl.TH + hh)};
      return tma_encode(m, a.ptr, 5, dims, strides, box, swz_of(width), dt);
    }
    const uint64_t dims[4] = {static_cast<uint64_t>(Ca), static_cast<uint64_t>(Wa), static_cast<uint64_t>(Ha),
                              static_cast<uint64_t>(p.a_bcast ? 1 : d->S)};
    const uint64_t sbytes = p.a_bcast ? static_cast<uint64_t>(a.hstride) * Ha * ez : static_cast<uint64_t>(a.sstride) * ez;
    const uint64_t strides[3] = {static_cast<uint64_t>(a.wstride) * ez, static_cast<uint64_t>(a.hstride) * ez, sbytes};
    const uint32_t box[4] = {static_cast<uint32_t>(width * epw), static_cast<uint32_t>(pl.Pw), static_cast<uint32_t>(pl.TH + hh), 1};
    return tma_encode(m, a.ptr, 4, dims, strides, box, swz_of(width), dt);
  };
  auto enc_b = [&](CUtensorMap* m, int width) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->Cin), static_cast<uint64_t>(d->Cout), static_cast<uint64_t>(p.taps),
                              static_cast<uint64_t>(p.b_bcast ? 1 : d->S)};
    const uint64_t tap_bytes = static_cast<uint64_t>(d->Cout) * w_cpitch * ez;
    const uint64_t strides[3] = {static_cast<uint64_t>(w_cpitch) * ez, tap_bytes,
                                 p.b_bcast ? tap_bytes * p.taps : static_cast<uint64_t>(w_sstride) * ez};
    if (!dgrad) {
      const uint32_t box[4] = {static_cast<uint32_t>(width * epw), static_cast<uint32_t>(pl.BN), static_cast<uint32_t>(pl.g), 1};
      return tma_encode(m, w, 4, dims, strides, box, swz_of(width), dt);
    }
    // MN-major: one 128-byte row of n (input channels) per k (output channel)
    const uint32_t box[4] = {bf16 ? 64u : 32u, static_cast<uint32_t>(width * epw), static_cast<uint32_t>(pl.g), 1};
    return tma_encode(m, w, 4, dims, strides, box, bf16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, dt);
  };
  const int wide = has32 ? 32 : tail_w, narrow = tail_w ? tail_w : 32;
  if (!enc_a(&tmA, wide) || !enc_a(&tmAt, narrow) || !enc_b(&tmB, wide) || !enc_b(&tmBt, narrow)) return -1;
  if (!has32) {          // only a narrow chunk exists: the kernel picks the *t maps for width != 32
    tmA = tmAt;
    tmB = tmBt;
  }
  static unsigned long long attr_done[2] = {0, 0};
  if (dry_run() == nullptr) {
    cudaError_t e = allow_dyn_smem(k_conv_halo<false>, 200 * 1024, &attr_done[0]);
    if (e == cudaSuccess) e = allow_dyn_smem(k_conv_halo<true>, 200 * 1024, &attr_done[1]);
    MFVI_REQUIRE(e == cudaSuccess, "%s: cannot raise dynamic shared memory: %s", what, cudaGetErrorString(e));
  }
  if (env_int("MFVI_TC2_VERBOSE", 0))
    fprintf(stderr, "[tc2] %s Kc=%d N=%d k%d M=%dx%d S=%d: TH=%d TW=%d Pw=%d n_mt=%d BN=%d nb=%d g=%d acc_stages=%d smem=%zu grid=%d tiles=%d\n",
            what, Ca, Nvalid, d->KH, Mh, Mw, d->S, pl.TH, pl.TW, pl.Pw, pl.n_mt, pl.BN, pl.n_nb, pl.g, pl.acc_stages, pl.smem, pl.grid,
            p.total_tiles);
  dry_detail("TH=%d TW=%d Pw=%d n_mt=%d BN=%d nb=%d chunks=%d g=%d acc_stages=%d tmem_cols=%u tiles=%d cls=%d planes=%d", pl.TH, pl.TW,
             pl.Pw, pl.n_mt, pl.BN, pl.n_nb, p.n_chunks, pl.g, pl.acc_stages, pl.tmem_cols, p.total_tiles, n_cls, planes);
  if (bf16) return -1;          // the kind::f16 instantiation was measured and dropped
  if (dgrad)
    launch_k(k_conv_halo<true>, pl.grid, kThreads, pl.smem, as_stream(st), tmA, tmAt, tmB, tmBt, p);
  else
    launch_k(k_conv_halo<false>, pl.grid, kThreads, pl.smem, as_stream(st), tmA, tmAt, tmB, tmBt, p);
  return check_launch(what);
}

}  // namespace tc2
}  // namespace mfvi

using namespace mfvi;

extern "C" {

// Return 0 on success, -1 when the shape is not taken (caller falls back to conv_tc.cu / conv_simt.cu), >0 on error.
int mfvi_conv2d_fwd_tc2(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                        double* stats, mfvi_stream_t st) {
  static const bool on = tc2::env_int("MFVI_TC2", 1) != 0;
  static const bool on2 = tc2::env_int("MFVI_TC2_S2", 1) != 0;
  if (!on || !(d->stride == 1 || (d->stride == 2 && on2)) || !tc2::view_ok(x, d->Cin) || (reinterpret_cast<uintptr_t>(w) % 16) || (w_sstride % 4) || d->Cin % 4 ||
      d->Cout > 256)
    return -1;
  return tc2::launch(d, false, x, d->Cin, d->Hin, d->Win, w, w_sstride, y, d->Hout, d->Wout, d->Cout, bias, stats, 0, st,
                     "conv2d_fwd_tc2");
}

int mfvi_conv2d_dgrad_tc2(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                          mfvi_stream_t st) {
  static const bool on = tc2::env_int("MFVI_TC2", 1) != 0;
  static const bool on2 = tc2::env_int("MFVI_TC2_S2", 1) != 0;
  if (!on || !(d->stride == 1 || (d->stride == 2 && on2)) || !tc2::view_ok(dy, d->Cout) || (reinterpret_cast<uintptr_t>(w) % 16) || (w_sstride % 4) || d->Cin % 4 ||
      d->Cin > 256)
    return -1;
  return tc2::launch(d, true, dy, d->Cout, d->Hout, d->Wout, w, w_sstride, dx, d->Hin, d->Win, d->Cin, nullptr, nullptr, accumulate, st,
                     "conv2d_dgrad_tc2");
}

}  // extern "C"
