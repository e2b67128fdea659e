// Persistent multi-stage kernel for the coarse scales of the hour-glass net.
//
// Why: at the <= 16 x 16 scales of the MFVI-DIP net every kernel of the plan has microseconds of work, and the step is the serial
// sum of ~100 such kernels, each paying its launch, prologue, first-load latency and drain (~8 us per link of the chain at one MC
// sample per GPU).  A grid-wide barrier costs as much as a kernel boundary (measured 3.5 us: profiles/r02_mega_grid_barrier.txt),
// so the fusion is done per MC SAMPLE instead: samples never interact inside the network (per-sample weights, per-sample
// BatchNorm statistics), so one thread-block cluster of 8 CTAs owns a sample, walks the stage list of the plan for it, and a
// cluster barrier (~0.2 us) stands where a kernel boundary used to be.  Activations of these scales stay in L2.
//
//   * elementwise stages run the bodies of the stand-alone kernels (elementwise_body.cuh) over the virtual blocks of the
//     sample: identical arithmetic, summation order included;
//   * convolution stages are implicit GEMMs on the warp-level tensor-core path (mma.sync m16n8k8 tf32, fp32 accumulate) with
//     64 x {16,32,64} CTA tiles (one sample at 8 x 8 is half a 128-row tcgen05 tile), operands streamed global -> shared
//     through a 6-deep cp.async ring.  In the exact-fp32 mode every product is formed as hi*hi + hi*lo + lo*hi of the tf32
//     splits (3xTF32), which restores fp32 accuracy on the tensor cores.
//
// The gradients of the BatchNorm affine parameters sum over all samples: they are formed by mfvi_bn_param_grads after the launch.
#include <algorithm>
#include <vector>

#include "elementwise_body.cuh"
#include "mega.cuh"

namespace mfvi {
namespace mega {

// ---------------------------------------------------------------------------------------------- recorder
static thread_local std::vector<Stage>* g_rec = nullptr;
static thread_local int g_nosync_next = 0;

bool recording() { return g_rec != nullptr; }

Stage* append(int op) {
  if (g_rec == nullptr) return nullptr;
  g_rec->emplace_back();
  Stage* s = &g_rec->back();
  memset(s, 0, sizeof(*s));
  s->op = op;
  s->nosync = g_nosync_next;
  g_nosync_next = 0;
  return s;
}

// ---------------------------------------------------------------------------------------------- device side
constexpr int kThreads = 256;
constexpr int kCluster = 8;        // CTAs per cluster = per MC sample
constexpr int kBM = 64;            // rows of a CTA tile
constexpr int kBK = 32;            // contraction chunk
constexpr int kOpFloats = 2304;    // one operand buffer: 64 x (32 + 4) = 32 x (64 + 8) floats
constexpr int kStages = 6;         // cp.async ring depth: ~55 KB in flight per CTA hides the L2 round trip of the tiny tiles

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// all CTAs of the cluster (= all CTAs working on this MC sample): release / acquire at cluster scope orders their global
// memory traffic, and the acquire drops stale L1 lines
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// 16-byte (or 4-byte) asynchronous copy global -> shared; bytes beyond src_bytes are zero-filled
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct Line {
  const float* p;    // first element of the line's contiguous run (NULL = all zeros)
  int n;             // valid elements
};

struct Tile {
  int s, m0, n0, tap, k_begin, k_end;
};

// Position of the contraction loop: chunk `it` = (tap, channel chunk) for forward / dgrad, 32 consecutive pixels for wgrad.
// Advanced incrementally (no divisions in the loop).
struct KPos {
  int kr, kt, tap, c0;
  __device__ __forceinline__ void reset() { kr = kt = tap = c0 = 0; }
  __device__ __forceinline__ void next(int Kc, int KW) {          // Kc = channels of the contraction
    c0 += kBK;
    if (c0 >= Kc) {
      c0 = 0;
      ++tap;
      if (++kt == KW) { kt = 0; ++kr; }
    }
  }
};

// One operand slot of a thread: line `line` of the tile, float4 `q` of it.  For forward / dgrad A-lines are tile rows (output /
// input pixels, fixed for the tile): (ph, pw) is the pixel, `ok` whether the row exists.  For wgrad lines are contraction
// indices (pixels that advance by 32 per chunk).
struct Slot {
  int line, q, ph, pw;
  bool ok;
};

// ---- operand lines.  Forward:  A[m = output pixel][k = (tap, ci)] = x,  B[n = co][k] = w            (both k-contiguous)
//                      dgrad:    A[m = input pixel][k = (tap, co)] = dy, B[k][n = ci] = w            (B n-contiguous)
//                      wgrad:    A[k = pixel][m = co] = dy,             B[k = pixel][n = ci] = x      (both row-contiguous), per tap
__device__ __forceinline__ Line a_line(const Stage& st, const Tile& t, const KPos& k, const Slot& sl) {
  const MfviConvDesc& d = st.d;
  Line l{nullptr, 0};
  if (!sl.ok) return l;
  if (st.op == OP_CONV_FWD) {
    l.p = st.a.ptr + (size_t)t.s * st.a.sstride + (size_t)(sl.ph * d.stride + k.kr) * st.a.hstride +
          (size_t)(sl.pw * d.stride + k.kt) * st.a.wstride + k.c0;
    l.n = min(kBK, d.Cin - k.c0);
  } else if (st.op == OP_CONV_DGRAD) {
    const int th = sl.ph - k.kr, tw = sl.pw - k.kt;
    if (th < 0 || tw < 0) return l;
    int ho = th, wo = tw;
    if (d.stride == 2) {
      if ((th | tw) & 1) return l;
      ho >>= 1;
      wo >>= 1;
    }
    if (ho >= d.Hout || wo >= d.Wout) return l;
    l.p = st.a.ptr + (size_t)t.s * st.a.sstride + (size_t)ho * st.a.hstride + (size_t)wo * st.a.wstride + k.c0;
    l.n = min(kBK, d.Cout - k.c0);
  } else {      // wgrad: line = contraction index (pixel (ph, pw) of dy), run = 64 output channels
    l.p = st.b.ptr + (size_t)t.s * st.b.sstride + (size_t)sl.ph * st.b.hstride + (size_t)sl.pw * st.b.wstride + t.m0;
    l.n = min(kBM, d.Cout - t.m0);
  }
  return l;
}

__device__ __forceinline__ Line b_line(const Stage& st, const Tile& t, const KPos& k, const Slot& sl, int BN) {
  const MfviConvDesc& d = st.d;
  Line l{nullptr, 0};
  if (st.op == OP_CONV_FWD) {
    const int co = t.n0 + sl.line;
    if (co >= d.Cout) return l;
    l.p = st.w + (size_t)t.s * st.w_sstride + ((size_t)k.tap * d.Cout + co) * d.Cin + k.c0;
    l.n = min(kBK, d.Cin - k.c0);
  } else if (st.op == OP_CONV_DGRAD) {
    const int co = k.c0 + sl.line;
    if (co >= d.Cout) return l;
    l.p = st.w + (size_t)t.s * st.w_sstride + ((size_t)k.tap * d.Cout + co) * d.Cin + t.n0;
    l.n = min(BN, d.Cin - t.n0);
  } else {
    if (!sl.ok) return l;
    const int kr = t.tap / d.KW, kt = t.tap - kr * d.KW;
    l.p = st.a.ptr + (size_t)t.s * st.a.sstride + (size_t)(sl.ph * d.stride + kr) * st.a.hstride +
          (size_t)(sl.pw * d.stride + kt) * st.a.wstride + t.n0;
    l.n = min(BN, d.Cin - t.n0);
  }
  return l;
}

// 4 consecutive floats q*4 .. q*4+3 of a line -> shared memory, asynchronously (zeros beyond the line's valid run)
__device__ __forceinline__ void copy4(uint32_t dst, const Line& l, int q, bool vec, const float* dummy) {
  const int e = q * 4;
  const int left = l.p != nullptr ? l.n - e : 0;
  if (vec) {
    cp_async16(dst, left > 0 ? l.p + e : dummy, left >= 4 ? 16 : (left > 0 ? left * 4 : 0));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) cp_async4(dst + 4 * j, left > j ? l.p + e + j : dummy, left > j ? 4 : 0);
  }
}

// One convolution stage for MC sample `smp`, shared by the `csize` CTAs of its cluster (this one is `rank`).
// NT = 8-column MMA tiles per warp (CTA tile 64 x 16*NT; warps 4 (rows) x 2 (columns)).  Operands travel global -> shared with
// cp.async through a kStages-deep ring as raw fp32; the tensor core reads the tf32 part of an fp32 word, and the 3xTF32 mode
// splits hi / lo when the fragments are loaded.
template <int NT, bool SPLIT3>
__device__ void conv_stage(const Stage& st, float* smem_f, int smp, int rank, int csize) {
  constexpr int BN = 16 * NT;
  constexpr int kBSlots = (BN * kBK / 4 + kThreads - 1) / kThreads;       // float4 slots of the B tile per thread (2, 1, 1)
  const MfviConvDesc& d = st.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  const bool fwd = st.op == OP_CONV_FWD, wgrad = st.op == OP_CONV_WGRAD;
  const int taps = d.KH * d.KW;
  // operand layouts in shared memory: line-major; "kc" = a line is a row with 32 contraction elements (pitch 36), "rc" = a line
  // is a contraction index with the tile's rows (pitch rows + 8).  Both make the fragment reads bank-conflict free.
  const bool a_kc = !wgrad, b_kc = fwd;
  const int a_len = a_kc ? kBK : kBM, b_len = b_kc ? kBK : BN;
  const int a_ld = a_kc ? kBK + 4 : kBM + 8, b_ld = b_kc ? kBK + 4 : BN + 8;
  const int a_rs = a_kc ? a_ld : 1, a_ks = a_kc ? 1 : a_ld;          // element (row r, k) of A at r * a_rs + k * a_ks
  const int b_rs = b_kc ? b_ld : 1, b_ks = b_kc ? 1 : b_ld;
  const int a_lpl = a_len / 4, b_lpl = b_len / 4;                    // float4 slots per line
  const int kch = fwd ? (d.Cin + kBK - 1) / kBK : (d.Cout + kBK - 1) / kBK;
  const bool vec_a = st.vec_a != 0, vec_b = st.vec_b != 0;
  const int per_s = wgrad ? taps * st.m_tiles * st.n_tiles * st.k_splits : st.m_tiles * st.n_tiles;
  const int K_total = d.Hout * d.Wout;
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem_f));
  const float* dummy = st.a.ptr;                                   // any mapped address: a zero-byte copy reads nothing

  for (int r0 = rank; r0 < per_s; r0 += csize) {
    Tile t;
    t.s = smp;
    int r = r0;
    int n_it;
    if (wgrad) {
      const int ks = r % st.k_splits;
      r /= st.k_splits;
      const int nb = r % st.n_tiles;
      r /= st.n_tiles;
      const int mb = r % st.m_tiles;
      t.tap = r / st.m_tiles;
      t.m0 = mb * kBM;
      t.n0 = nb * BN;
      t.k_begin = ks * st.k_len;
      t.k_end = min(t.k_begin + st.k_len, K_total);
      n_it = (t.k_end - t.k_begin + kBK - 1) / kBK;
    } else {
      const int mb = r / st.n_tiles;
      t.m0 = mb * kBM;
      t.n0 = (r - mb * st.n_tiles) * BN;
      t.tap = 0;
      t.k_begin = t.k_end = 0;
      n_it = taps * kch;
    }
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    // this thread's operand slots and the loop position of the NEXT chunk to issue
    Slot sa[2], sb[kBSlots];
    const int Mw_rows = fwd ? d.Wout : d.Win, M_rows = fwd ? d.Hout * d.Wout : d.Hin * d.Win;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = tid + j * kThreads;
      sa[j].line = idx / a_lpl;
      sa[j].q = idx - sa[j].line * a_lpl;
      const int m = wgrad ? t.k_begin + sa[j].line : t.m0 + sa[j].line;
      sa[j].ok = wgrad ? m < t.k_end : m < M_rows;
      const int wdiv = wgrad ? d.Wout : Mw_rows;
      sa[j].ph = m / wdiv;
      sa[j].pw = m - sa[j].ph * wdiv;
    }
#pragma unroll
    for (int j = 0; j < kBSlots; ++j) {
      const int idx = tid + j * kThreads;
      sb[j].line = idx / b_lpl;
      sb[j].q = idx - sb[j].line * b_lpl;
      const int m = t.k_begin + sb[j].line;            // wgrad only: the pixel of the line
      sb[j].ok = m < t.k_end;
      sb[j].ph = m / d.Wout;
      sb[j].pw = m - sb[j].ph * d.Wout;
    }
    const int adv_h = kBK / d.Wout, adv_w = kBK - adv_h * d.Wout;       // wgrad: a chunk advances every line by 32 pixels
    KPos kp;
    kp.reset();
    int issued = 0;
    auto issue = [&]() {            // operands of the next contraction chunk -> ring slot issued % kStages
      const uint32_t a_dst = sbase + static_cast<uint32_t>((issued % kStages) * 2 * kOpFloats) * 4u;
      const uint32_t b_dst = a_dst + kOpFloats * 4u;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        copy4(a_dst + static_cast<uint32_t>(sa[j].line * a_ld + sa[j].q * 4) * 4u, a_line(st, t, kp, sa[j]), sa[j].q, vec_a, dummy);
#pragma unroll
      for (int j = 0; j < kBSlots; ++j)
        if (tid + j * kThreads < BN * kBK / 4)
          copy4(b_dst + static_cast<uint32_t>(sb[j].line * b_ld + sb[j].q * 4) * 4u, b_line(st, t, kp, sb[j], BN), sb[j].q, vec_b,
                dummy);
      ++issued;
      if (wgrad) {
        auto adv = [&](Slot& x) {
          x.pw += adv_w;
          x.ph += adv_h;
          if (x.pw >= d.Wout) { x.pw -= d.Wout; ++x.ph; }
          x.ok = x.ok && (x.ph * d.Wout + x.pw) < t.k_end;
        };
#pragma unroll
        for (int j = 0; j < 2; ++j) adv(sa[j]);
#pragma unroll
        for (int j = 0; j < kBSlots; ++j) adv(sb[j]);
      } else {
        kp.next(fwd ? d.Cin : d.Cout, d.KW);
      }
    };

    __syncthreads();                 // the previous item's epilogue / the previous stage is done with shared memory
#pragma unroll
    for (int pre = 0; pre < kStages - 1; ++pre) {
      if (pre < n_it) issue();
      cp_async_commit();
    }
    for (int it = 0; it < n_it; ++it) {
      cp_async_wait<kStages - 2>();     // chunk `it` has landed (for this thread's copies) ...
      __syncthreads();                  // ... and for everybody's; slot (it - 1) % kStages is free again
      if (it + kStages - 1 < n_it) issue();
      cp_async_commit();
      const uint32_t* A = reinterpret_cast<const uint32_t*>(smem_f) + (it % kStages) * 2 * kOpFloats + (wm * 16 + g) * a_rs;
      const uint32_t* B = reinterpret_cast<const uint32_t*>(smem_f) + (it % kStages) * 2 * kOpFloats + kOpFloats +
                          (wn * 8 * NT + g) * b_rs;
#pragma unroll
      for (int k0 = 0; k0 < kBK; k0 += 8) {
        const int ka = (k0 + tq) * a_ks, ka4 = (k0 + tq + 4) * a_ks;
        uint32_t a0 = A[ka], a1 = A[8 * a_rs + ka], a2 = A[ka4], a3 = A[8 * a_rs + ka4];
        uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0;
        if (SPLIT3) {                    // hi = tf32 rounding of x, lo = tf32 rounding of the remainder
          const float f0 = __uint_as_float(a0), f1 = __uint_as_float(a1), f2 = __uint_as_float(a2), f3 = __uint_as_float(a3);
          a0 = to_tf32(f0); a1 = to_tf32(f1); a2 = to_tf32(f2); a3 = to_tf32(f3);
          l0 = to_tf32(f0 - __uint_as_float(a0)); l1 = to_tf32(f1 - __uint_as_float(a1));
          l2 = to_tf32(f2 - __uint_as_float(a2)); l3 = to_tf32(f3 - __uint_as_float(a3));
        }
        const int kb = (k0 + tq) * b_ks, kb4 = (k0 + tq + 4) * b_ks;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          uint32_t b0 = B[i * 8 * b_rs + kb], b1 = B[i * 8 * b_rs + kb4];
          if (SPLIT3) {
            const float f0 = __uint_as_float(b0), f1 = __uint_as_float(b1);
            b0 = to_tf32(f0); b1 = to_tf32(f1);
            const uint32_t m0 = to_tf32(f0 - __uint_as_float(b0)), m1 = to_tf32(f1 - __uint_as_float(b1));
            mma_tf32(acc[i], l0, l1, l2, l3, b0, b1);       // small terms first
            mma_tf32(acc[i], a0, a1, a2, a3, m0, m1);
          }
          mma_tf32(acc[i], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();                   // all fragments read: shared memory may be reused by the epilogue

    // ---- epilogue
    if (wgrad) {
      float* dwp = st.dw + (size_t)t.s * st.w_sstride + (size_t)t.tap * d.Cout * d.Cin;
#pragma unroll
      for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = t.m0 + wm * 16 + g + (e >> 1) * 8;
          const int ci = t.n0 + wn * 8 * NT + i * 8 + 2 * tq + (e & 1);
          if (co < d.Cout && ci < d.Cin && n_it > 0) atomicAdd(dwp + (size_t)co * d.Cin + ci, acc[i][e]);
        }
      continue;
    }
    const int Mw = fwd ? d.Wout : d.Win, Mtot = fwd ? d.Hout * d.Wout : d.Hin * d.Win;
    const int Nvalid = fwd ? d.Cout : d.Cin;
    const MfviView& o = st.b;                        // forward: y, dgrad: dx (both recorded in `b`)
    float s1[NT][2], s2[NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i) s1[i][0] = s1[i][1] = s2[i][0] = s2[i][1] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = t.m0 + wm * 16 + g + h * 8;
      if (m >= Mtot) continue;
      const int ph = m / Mw, pw = m - ph * Mw;
      float* orow = o.ptr + (size_t)t.s * o.sstride + (size_t)ph * o.hstride + (size_t)pw * o.wstride;
#pragma unroll
      for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n = t.n0 + wn * 8 * NT + i * 8 + 2 * tq + c;
          if (n >= Nvalid) continue;
          float v = acc[i][h * 2 + c];
          if (fwd) {
            if (st.bias != nullptr) v += __ldg(st.bias + (size_t)t.s * st.w_sstride + n);
            s1[i][c] += v;
            s2[i][c] = fmaf(v, v, s2[i][c]);
            orow[n] = v;
          } else {
            orow[n] = st.accumulate ? __ldcg(orow + n) + v : v;
          }
        }
    }
    if (fwd && st.stats != nullptr) {
      // per-sample BatchNorm partials of the tile: rows of the warp (xor over g), then the four row warps through shared memory
      // (fixed order), one double atomic per column
#pragma unroll
      for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int off = 4; off < 32; off <<= 1) {
            s1[i][c] += __shfl_xor_sync(0xffffffffu, s1[i][c], off);
            s2[i][c] += __shfl_xor_sync(0xffffffffu, s2[i][c], off);
          }
        }
      float* red = smem_f;                         // [4 row warps][BN][2]
      if (g == 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col = wn * 8 * NT + i * 8 + 2 * tq + c;
            red[(wm * BN + col) * 2 + 0] = s1[i][c];
            red[(wm * BN + col) * 2 + 1] = s2[i][c];
          }
      }
      __syncthreads();
      if (tid < BN * 2) {
        const int col = tid >> 1, which = tid & 1;
        const int n = t.n0 + col;
        if (n < d.Cout) {
          const double v = (double)red[(0 * BN + col) * 2 + which] + (double)red[(1 * BN + col) * 2 + which] +
                           (double)red[(2 * BN + col) * 2 + which] + (double)red[(3 * BN + col) * 2 + which];
          atomicAdd(st.stats + ((size_t)t.s * d.Cout + n) * 2 + which, v);
        }
      }
    }
  }
}

// elementwise stage of MC sample `smp`: the virtual blocks (bx, smp) of the stand-alone kernel's grid, shared by the cluster
template <int V>
__device__ void ew_stage(const Stage& st, EwSmem sm, int smp, int rank, int csize) {
  for (int bx = rank; bx < st.gx; bx += csize) {
    const VGrid vg{bx, smp, st.gx};
    switch (st.op) {
      case OP_BN_ACT_PAD_FWD:
        body_bn_act_pad_fwd<V, false>(vg, sm, st.a, st.H, st.W, st.C, st.sums, st.gamma, st.beta, st.act, st.pad, st.b, st.G, st.PPB);
        break;
      case OP_CAT_UP_FWD:
        body_cat_up_fwd<V>(vg, sm, st.a, st.C, st.sums, st.gamma, st.beta, st.b, st.C2, st.sums2, st.gamma2, st.beta2, st.H, st.W,
                           st.mode, st.c, st.red, st.G, st.PPB);
        break;
      case OP_PAD_ACT_BWD:
        body_pad_act_bwd<V>(vg, sm, st.a, st.H, st.W, st.C, st.pad, st.b, st.sums, st.gamma, st.beta, st.act, st.c, st.red, st.G,
                            st.PPB);
        break;
      case OP_BN_BWD_APPLY:
        // the affine-parameter gradients sum over ALL samples: they are formed after the launch (mfvi_bn_param_grads), when
        // every cluster has finished
        body_bn_bwd_apply<V, false>(vg, sm, st.a, st.b, st.S, st.H, st.W, st.C, st.sums, st.red, st.gamma, st.c, nullptr, nullptr,
                                    st.G, st.PPB);
        break;
      case OP_CAT_BWD_SKIP:
        body_cat_bwd_skip<V>(vg, sm, st.a, st.H, st.W, st.b, st.C, st.sums, st.gamma, st.beta, st.c, st.red, st.G, st.PPB);
        break;
      case OP_CAT_BWD_UP:
        body_cat_bwd_up<V>(vg, sm, st.a, st.H, st.W, st.mode, st.C, st.b, st.C2, st.sums, st.gamma, st.beta, st.c, st.red, st.G,
                           st.PPB);
        break;
      default:
        break;
    }
    __syncthreads();                 // the next virtual block reuses the shared-memory tables
  }
}

// One cluster per MC sample (looping when there are more samples than clusters): the samples of an MFVI-DIP step never
// interact inside the network (per-sample weights, per-sample BatchNorm statistics), so every dependency between stages is
// between the CTAs of ONE cluster and a cluster barrier (~0.2 us) stands where a kernel boundary (~5 us) used to be.
__global__ void __launch_bounds__(kThreads, 1)
k_mega(const Stage* __restrict__ prog, int n_stages, int S, long long* prof) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ Stage st;
  pdl_trigger();
  pdl_wait();
  float* smem_f = reinterpret_cast<float*>(smem_raw);
  // elementwise pieces overlay the convolution ring (stages are sequential)
  EwSmem sm;
  sm.red = reinterpret_cast<double*>(smem_raw);
  sm.tab = reinterpret_cast<BnTable*>(smem_raw + 2 * kEwThreads * 4 * sizeof(double));
  sm.misc = reinterpret_cast<float*>(smem_raw + 2 * kEwThreads * 4 * sizeof(double) + sizeof(BnTable));
  const int rank = static_cast<int>(cluster_rank());
  const int n_clusters = gridDim.x / kCluster, cluster_id = blockIdx.x / kCluster;
  for (int smp = cluster_id; smp < S; smp += n_clusters) {
    for (int i = 0; i < n_stages; ++i) {
      __syncthreads();
      {
        const int* src = reinterpret_cast<const int*>(prog + i);
        int* dst = reinterpret_cast<int*>(&st);
        for (int k = threadIdx.x; k < (int)(sizeof(Stage) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
      }
      __syncthreads();
      if (i > 0 && !st.nosync) cluster_barrier();
      if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {        // stage start time of CTA 0 (ns): a profiling aid
        long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        prof[i] = tns;
      }
      switch (st.op) {
        case OP_CONV_FWD:
        case OP_CONV_DGRAD:
        case OP_CONV_WGRAD:
          if (st.split3) {
            if (st.nt == 4) conv_stage<4, true>(st, smem_f, smp, rank, kCluster);
            else if (st.nt == 2) conv_stage<2, true>(st, smem_f, smp, rank, kCluster);
            else conv_stage<1, true>(st, smem_f, smp, rank, kCluster);
          } else {
            if (st.nt == 4) conv_stage<4, false>(st, smem_f, smp, rank, kCluster);
            else if (st.nt == 2) conv_stage<2, false>(st, smem_f, smp, rank, kCluster);
            else conv_stage<1, false>(st, smem_f, smp, rank, kCluster);
          }
          break;
        case OP_FILL: {      // the buffer holds S equal per-sample slices
          const size_t per = st.fill_n / S;
          float* p = st.fill_ptr + (size_t)smp * per;
          for (size_t k = (size_t)rank * blockDim.x + threadIdx.x; k < per; k += (size_t)kCluster * blockDim.x) p[k] = st.fill_v;
          break;
        }
        default:
          if (st.V == 4) ew_stage<4>(st, sm, smp, rank, kCluster);
          else ew_stage<1>(st, sm, smp, rank, kCluster);
          break;
      }
    }
    cluster_barrier();      // a further sample of this cluster starts from a clean slate (and stage 0 has no barrier of its own)
  }
  if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    long long tns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
    prof[n_stages] = tns;
  }
}

// gradients of the BatchNorm affine parameters of the fused layers: dgamma[c] = sum_s red[s][c][1], dbeta[c] = sum_s red[s][c][0]
__global__ void k_bn_param_grads(const long long* __restrict__ red_ptr, const long long* __restrict__ dgamma_ptr,
                                 const long long* __restrict__ dbeta_ptr, const int* __restrict__ Cs, int S) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const double* red = reinterpret_cast<const double*>(red_ptr[b]);
  float* dgamma = reinterpret_cast<float*>(dgamma_ptr[b]);
  float* dbeta = reinterpret_cast<float*>(dbeta_ptr[b]);
  const int C = Cs[b];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double dg = 0.0, db = 0.0;
    for (int s = 0; s < S; ++s) {
      db += red[((size_t)s * C + c) * 2 + 0];
      dg += red[((size_t)s * C + c) * 2 + 1];
    }
    dgamma[c] = (float)dg;
    dbeta[c] = (float)db;
  }
}

constexpr size_t kEwSmemBytes = 2 * kEwThreads * 4 * sizeof(double) + sizeof(BnTable) + 5 * kMaxC * sizeof(float);
constexpr size_t kConvSmemBytes = static_cast<size_t>(kStages) * 2 * kOpFloats * sizeof(float);      // the A / B ring

static bool view_vec(const MfviView& v) {
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && v.sstride % 4 == 0 && v.hstride % 4 == 0 && v.wstride % 4 == 0;
}

static int cdiv(int a, int b) { return (a + b - 1) / b; }

// tile plan of a convolution stage: the widest column tile that still gives every SM two work items, else the narrowest
static void plan_conv(Stage* s, int target) {
  const MfviConvDesc& d = s->d;
  if (s->op == OP_CONV_WGRAD) {
    const int taps = d.KH * d.KW, K = d.Hout * d.Wout;
    s->m_tiles = cdiv(d.Cout, kBM);
    int nt = 4;
    while (nt > 1 && 16 * nt > ((d.Cin + 15) / 16) * 16) nt >>= 1;
    s->nt = nt;
    s->n_tiles = cdiv(d.Cin, 16 * nt);
    const int base = d.S * taps * s->m_tiles * s->n_tiles;
    int splits = std::max(1, std::min(cdiv(K, kBK), cdiv(target, base)));
    s->k_len = cdiv(cdiv(K, splits), kBK) * kBK;
    s->k_splits = cdiv(K, s->k_len);
    s->items = base * s->k_splits;
    return;
  }
  const int M = s->op == OP_CONV_FWD ? d.Hout * d.Wout : d.Hin * d.Win;
  const int N = s->op == OP_CONV_FWD ? d.Cout : d.Cin;
  s->m_tiles = cdiv(M, kBM);
  const int Nfull = ((N + 15) / 16) * 16;
  int nt = 4;
  for (; nt >= 1; nt >>= 1) {
    if (16 * nt > Nfull && nt > 1) continue;
    if (d.S * s->m_tiles * cdiv(N, 16 * nt) >= target || nt == 1) break;
  }
  s->nt = nt;
  s->n_tiles = cdiv(N, 16 * nt);
  s->k_splits = 1;
  s->k_len = 0;
  s->items = d.S * s->m_tiles * s->n_tiles;
}

// fills the convolution part of a stage; `target` = work items to aim for over all samples
static void fill_conv(Stage* s, int op, const MfviConvDesc* d, MfviView act_in, MfviView act_out, const float* w, const float* bias,
                      long long w_sstride, float* dw, double* stats, int accumulate, int target) {
  s->op = op;
  s->d = *d;
  s->a = act_in;       // fwd: x (padded input); dgrad: dy; wgrad: x
  s->b = act_out;      // fwd: y; dgrad: dx; wgrad: dy
  s->w = w;
  s->bias = bias;
  s->w_sstride = w_sstride;
  s->dw = dw;
  s->stats = stats;
  s->accumulate = accumulate;
  s->split3 = d->math == MFVI_MATH_FP32 ? 1 : 0;
  const bool w_vec = w != nullptr && reinterpret_cast<uintptr_t>(w) % 16 == 0 && w_sstride % 4 == 0 && d->Cin % 4 == 0;
  if (op == OP_CONV_FWD) {
    s->vec_a = view_vec(act_in) && d->Cin % 4 == 0;
    s->vec_b = w_vec;
  } else if (op == OP_CONV_DGRAD) {
    s->vec_a = view_vec(act_in) && d->Cout % 4 == 0;
    s->vec_b = w_vec;
  } else {
    s->vec_a = view_vec(act_out) && d->Cout % 4 == 0;      // dy rows
    s->vec_b = view_vec(act_in) && d->Cin % 4 == 0;        // x rows
  }
  plan_conv(s, target);
}

int record_conv(int op, const MfviConvDesc* d, MfviView act_in, MfviView act_out, const float* w, const float* bias,
                long long w_sstride, float* dw, double* stats, int accumulate) {
  MFVI_REQUIRE(d != nullptr && (d->stride == 1 || d->stride == 2), "mega: convolution stride must be 1 or 2");
  MFVI_REQUIRE(act_in.ptr != nullptr && act_out.ptr != nullptr, "mega: null activation view");
  Stage* s = append(op);
  // one work item per CTA of every sample's cluster
  fill_conv(s, op, d, act_in, act_out, w, bias, w_sstride, dw, stats, accumulate, kCluster * d->S);
  return 0;
}

// ---- the same convolution tiles as a stand-alone kernel: grid (work items of a sample, S).  Used by the exact-fp32 mode, where
// the 3xTF32 products put the convolutions on the tensor cores at fp32 accuracy (the CUDA-core kernels of conv_simt.cu remain
// the fallback for shapes outside stride 1 / 2).
template <bool SPLIT3>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_mma(const __grid_constant__ Stage st) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  pdl_trigger();
  pdl_wait();
  float* smem_f = reinterpret_cast<float*>(smem_raw);
  const int smp = blockIdx.y, rank = blockIdx.x, csize = gridDim.x;
  if (st.nt == 4) conv_stage<4, SPLIT3>(st, smem_f, smp, rank, csize);
  else if (st.nt == 2) conv_stage<2, SPLIT3>(st, smem_f, smp, rank, csize);
  else conv_stage<1, SPLIT3>(st, smem_f, smp, rank, csize);
}

int launch_conv_mma(int op, const MfviConvDesc* d, MfviView act_in, MfviView act_out, const float* w, const float* bias,
                    long long w_sstride, float* dw, double* stats, int accumulate, mfvi_stream_t stream, const char* what) {
  if (d == nullptr || !(d->stride == 1 || d->stride == 2) || act_in.ptr == nullptr || act_out.ptr == nullptr) return -1;
  Stage s;
  memset(&s, 0, sizeof(s));
  fill_conv(&s, op, d, act_in, act_out, w, bias, w_sstride, dw, stats, accumulate, 2 * kNumSMs);
  const int per_s = s.items / d->S;
  if (per_s < 1 || per_s > 65535 * 32) return -1;
  static unsigned long long attr_done[2] = {0, 0};
  if (dry_run() == nullptr) {
    cudaError_t e = allow_dyn_smem(k_conv_mma<true>, (int)(kStages * 2 * kOpFloats * 4), &attr_done[0]);
    if (e == cudaSuccess) e = allow_dyn_smem(k_conv_mma<false>, (int)(kStages * 2 * kOpFloats * 4), &attr_done[1]);
    MFVI_REQUIRE(e == cudaSuccess, "%s: cannot raise dynamic shared memory: %s", what, cudaGetErrorString(e));
  }
  const size_t smem = static_cast<size_t>(kStages) * 2 * kOpFloats * sizeof(float);
  const dim3 grid(std::min(per_s, 4 * kNumSMs), d->S);
  dry_detail("mma tile 64x%d m_tiles=%d n_tiles=%d k_splits=%d items=%d split3=%d", 16 * s.nt, s.m_tiles, s.n_tiles, s.k_splits, s.items,
             s.split3);
  if (s.split3) launch_k(k_conv_mma<true>, grid, kThreads, smem, as_stream(stream), s);
  else launch_k(k_conv_mma<false>, grid, kThreads, smem, as_stream(stream), s);
  return check_launch(what);
}

}  // namespace mega
}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_mega_begin(void) {
  MFVI_REQUIRE(mega::g_rec == nullptr, "mega_begin: a program is already being recorded on this thread");
  mega::g_rec = new std::vector<mega::Stage>();
  mega::g_nosync_next = 0;
  return 0;
}

int mfvi_mega_mark_nosync(void) {
  MFVI_REQUIRE(mega::g_rec != nullptr, "mega_mark_nosync: not recording");
  mega::g_nosync_next = 1;
  return 0;
}

size_t mfvi_mega_stage_bytes(void) { return sizeof(mega::Stage); }

// Ends the recording and copies the program into `program_dev` (device memory of at least capacity_bytes, owned by the caller;
// a synchronous copy — plan-build time).  Reports the number of stages.
int mfvi_mega_end(void* program_dev, size_t capacity_bytes, int* n_stages) {
  MFVI_REQUIRE(mega::g_rec != nullptr, "mega_end: not recording");
  std::vector<mega::Stage>* rec = mega::g_rec;
  mega::g_rec = nullptr;
  const size_t bytes = rec->size() * sizeof(mega::Stage);
  if (n_stages) *n_stages = static_cast<int>(rec->size());
  int rc = 0;
  if (bytes > capacity_bytes || (bytes > 0 && program_dev == nullptr)) {
    set_error("mega_end: program of %zu bytes does not fit the %zu-byte buffer", bytes, capacity_bytes);
    rc = 1;
  } else if (bytes > 0) {
    const cudaError_t e = cudaMemcpy(program_dev, rec->data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("mega_end: copy failed: %s", cudaGetErrorString(e));
      rc = 2;
    }
  }
  delete rec;
  return rc;
}

// Runs a recorded program for S MC samples: ONE launch, one 8-CTA cluster per sample (at most 16 clusters; they loop beyond).
int mfvi_mega_run(const void* program_dev, int n_stages, int S, long long* stage_times, mfvi_stream_t st) {
  MFVI_REQUIRE(program_dev != nullptr && n_stages >= 1 && S >= 1, "mega_run: null program / no samples");
  const size_t smem = std::max(mega::kEwSmemBytes, mega::kConvSmemBytes);
  static unsigned long long attr_done = 0;
  {
    const cudaError_t e = allow_dyn_smem(mega::k_mega, (int)smem, &attr_done);
    MFVI_REQUIRE(e == cudaSuccess, "mega_run: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  const int clusters = std::min(S, 16);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * mega::kCluster);
  cfg.blockDim = dim3(mega::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(st);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = mega::kCluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, mega::k_mega, static_cast<const mega::Stage*>(program_dev), n_stages, S, stage_times);
  return check_launch("mega_run");
}

// dgamma / dbeta of n BatchNorm layers whose backward ran inside a program: HOST arrays of n device pointers / channel counts
// are not used here — the tables are DEVICE arrays (int64 addresses) prepared once by the caller.
int mfvi_bn_param_grads(const long long* red_ptrs, const long long* dgamma_ptrs, const long long* dbeta_ptrs, const int* Cs, int n,
                        int S, mfvi_stream_t st) {
  MFVI_REQUIRE(red_ptrs && dgamma_ptrs && dbeta_ptrs && Cs, "bn_param_grads: null table");
  if (n == 0) return 0;
  launch_k(mega::k_bn_param_grads, n, 128, 0, as_stream(st), red_ptrs, dgamma_ptrs, dbeta_ptrs, Cs, S);
  return check_launch("bn_param_grads");
}

}  // extern "C"
