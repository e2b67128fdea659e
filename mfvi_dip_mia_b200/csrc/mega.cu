// Persistent multi-stage kernel for the coarse scales of the hour-glass net.
//
// Why: at the <= 32 x 32 scales of the MFVI-DIP net every kernel of the plan has microseconds of work, and the step is the serial
// sum of ~100 such kernels, each paying its launch, prologue, first-load latency and drain (~0.85 ms of the
// 2 ms step).  Here the whole sub-network below a chosen scale runs as ONE launch per direction: the CTAs stay resident, walk the
// stage list of the plan and meet at a grid barrier (one L2 atomic + a spin on an L2 line, ~1 us) where a kernel boundary used to
// be.  All activations of these scales stay in L2 (a few MB).
//
//   * elementwise stages run the bodies of the stand-alone kernels (elementwise_body.cuh) over virtual block indices: identical
//     arithmetic, summation order included;
//   * convolution stages are implicit GEMMs on the warp-level tensor-core path (mma.sync m16n8k8 tf32, fp32 accumulate): with
//     64 x {16,32,64} CTA tiles the 8 x 8 .. 32 x 32 maps fill the machine, which the 128-row tcgen05 tiles cannot at these sizes
//     (one MC sample at 8 x 8 is half a tile).  Operands are rounded to tf32 with cvt.rna (round to nearest, unlike the
//     truncation of the tcgen05 path); in the exact-fp32 mode every product is formed as hi*hi + hi*lo + lo*hi of the tf32
//     splits (3xTF32), which restores fp32 accuracy on the tensor cores.
//
// Synchronisation: bar[0] is a monotonic count of arrivals (epoch k completes at k * gridDim.x), bar[1] counts CTAs that have
// finished; the last one resets both, so the pair is zero again when the launch ends (the caller allocates it zeroed, once).  The grid never exceeds the number of co-resident CTAs (occupancy x SMs), so the spin cannot deadlock; a
// kernel launched behind it with programmatic dependent launch starts only once every CTA of this one is resident.
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
constexpr int kBM = 64;            // rows of a CTA tile
constexpr int kBK = 32;            // contraction chunk
constexpr int kOpFloats = 2304;    // one operand buffer: 64 x (32 + 4) = 32 x (64 + 8) floats

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

__device__ __forceinline__ uint32_t ld_acquire(const unsigned* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// every CTA arrives once per epoch; the counter only grows, so epoch k is complete at k * gridDim.x arrivals
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                       // this CTA's writes are visible device-wide before it arrives
    atomicAdd(bar, 1u);
    while (ld_acquire(bar) < target) __nanosleep(32);
    __threadfence();                       // gpu-scope fence: also drops stale L1 lines of data other CTAs rewrote
  }
  __syncthreads();
}

struct Line {
  const float* p;    // first element of the line's contiguous run (NULL = all zeros)
  int n;             // valid elements
};

struct Tile {
  int s, m0, n0, tap, k_begin, k_end;
};

// ---- operand lines.  Forward:  A[m = output pixel][k = (tap, ci)] = x,  B[n = co][k] = w            (both k-contiguous)
//                      dgrad:    A[m = input pixel][k = (tap, co)] = dy, B[k][n = ci] = w            (B n-contiguous)
//                      wgrad:    A[k = pixel][m = co] = dy,             B[k = pixel][n = ci] = x      (both row-contiguous), per tap
__device__ __forceinline__ Line a_line(const Stage& st, const Tile& t, int it, int line, int kch) {
  const MfviConvDesc& d = st.d;
  Line l{nullptr, 0};
  if (st.op == OP_CONV_FWD) {
    const int tap = it / kch, c0 = (it - tap * kch) * kBK;
    const int m = t.m0 + line;
    if (m >= d.Hout * d.Wout) return l;
    const int ho = m / d.Wout, wo = m - ho * d.Wout;
    const int kr = tap / d.KW, kt = tap - kr * d.KW;
    l.p = st.a.ptr + (size_t)t.s * st.a.sstride + (size_t)(ho * d.stride + kr) * st.a.hstride +
          (size_t)(wo * d.stride + kt) * st.a.wstride + c0;
    l.n = min(kBK, d.Cin - c0);
  } else if (st.op == OP_CONV_DGRAD) {
    const int tap = it / kch, c0 = (it - tap * kch) * kBK;
    const int m = t.m0 + line;
    if (m >= d.Hin * d.Win) return l;
    const int hi = m / d.Win, wi = m - hi * d.Win;
    const int kr = tap / d.KW, kt = tap - kr * d.KW;
    const int th = hi - kr, tw = wi - kt;
    if (th < 0 || tw < 0) return l;
    int ho = th, wo = tw;
    if (d.stride == 2) {
      if ((th | tw) & 1) return l;
      ho >>= 1;
      wo >>= 1;
    }
    if (ho >= d.Hout || wo >= d.Wout) return l;
    l.p = st.a.ptr + (size_t)t.s * st.a.sstride + (size_t)ho * st.a.hstride + (size_t)wo * st.a.wstride + c0;
    l.n = min(kBK, d.Cout - c0);
  } else {      // wgrad: line = contraction index (pixel), run = 64 output channels of dy
    const int pix = t.k_begin + it * kBK + line;
    if (pix >= t.k_end) return l;
    const int ho = pix / d.Wout, wo = pix - ho * d.Wout;
    l.p = st.b.ptr + (size_t)t.s * st.b.sstride + (size_t)ho * st.b.hstride + (size_t)wo * st.b.wstride + t.m0;
    l.n = min(kBM, d.Cout - t.m0);
  }
  return l;
}

__device__ __forceinline__ Line b_line(const Stage& st, const Tile& t, int it, int line, int kch, int BN) {
  const MfviConvDesc& d = st.d;
  Line l{nullptr, 0};
  if (st.op == OP_CONV_FWD) {
    const int tap = it / kch, c0 = (it - tap * kch) * kBK;
    const int co = t.n0 + line;
    if (co >= d.Cout) return l;
    l.p = st.w + (size_t)t.s * st.w_sstride + ((size_t)tap * d.Cout + co) * d.Cin + c0;
    l.n = min(kBK, d.Cin - c0);
  } else if (st.op == OP_CONV_DGRAD) {
    const int tap = it / kch, c0 = (it - tap * kch) * kBK;
    const int co = c0 + line;
    if (co >= d.Cout) return l;
    l.p = st.w + (size_t)t.s * st.w_sstride + ((size_t)tap * d.Cout + co) * d.Cin + t.n0;
    l.n = min(BN, d.Cin - t.n0);
  } else {
    const int pix = t.k_begin + it * kBK + line;
    if (pix >= t.k_end) return l;
    const int ho = pix / d.Wout, wo = pix - ho * d.Wout;
    const int kr = t.tap / d.KW, kt = t.tap - kr * d.KW;
    l.p = st.a.ptr + (size_t)t.s * st.a.sstride + (size_t)(ho * d.stride + kr) * st.a.hstride +
          (size_t)(wo * d.stride + kt) * st.a.wstride + t.n0;
    l.n = min(BN, d.Cin - t.n0);
  }
  return l;
}

// 4 consecutive floats q*4 .. q*4+3 of a line (zeros beyond its valid run).  L2 loads (ld.global.cg): the data may have been
// written by another CTA in an earlier stage of this launch.
__device__ __forceinline__ float4 load4(const Line& l, int q, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const int e = q * 4;
  if (l.p == nullptr || e >= l.n) return v;
  if (vec && e + 4 <= l.n) return __ldcg(reinterpret_cast<const float4*>(l.p + e));
  v.x = __ldcg(l.p + e);
  if (e + 1 < l.n) v.y = __ldcg(l.p + e + 1);
  if (e + 2 < l.n) v.z = __ldcg(l.p + e + 2);
  if (e + 3 < l.n) v.w = __ldcg(l.p + e + 3);
  return v;
}

template <bool SPLIT3>
__device__ __forceinline__ void store4(uint32_t* hi, uint32_t* lo, int off, const float4& v) {
  uint4 h;
  h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
  *reinterpret_cast<uint4*>(hi + off) = h;
  if (SPLIT3) {
    uint4 l;
    l.x = to_tf32(v.x - __uint_as_float(h.x)); l.y = to_tf32(v.y - __uint_as_float(h.y));
    l.z = to_tf32(v.z - __uint_as_float(h.z)); l.w = to_tf32(v.w - __uint_as_float(h.w));
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

// One convolution stage.  NT = 8-column MMA tiles per warp (CTA tile 64 x 16*NT; warps 4 (rows) x 2 (columns)).
template <int NT, bool SPLIT3>
__device__ void conv_stage(const Stage& st, float* smem_f) {
  constexpr int BN = 16 * NT;
  constexpr int kBSlots = (BN * kBK / 4 + kThreads - 1) / kThreads;       // float4 slots of the B tile per thread (2, 1, 1)
  const MfviConvDesc& d = st.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int g = lane >> 2, tq = lane & 3;
  const bool fwd = st.op == OP_CONV_FWD, dgrad = st.op == OP_CONV_DGRAD, wgrad = st.op == OP_CONV_WGRAD;
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
  uint32_t* sA[2] = {reinterpret_cast<uint32_t*>(smem_f), reinterpret_cast<uint32_t*>(smem_f) + kOpFloats};
  uint32_t* sB[2] = {sA[1] + kOpFloats, sA[1] + 2 * kOpFloats};
  uint32_t* sAl[2] = {sB[1] + kOpFloats, sB[1] + 2 * kOpFloats};      // lo parts (SPLIT3 only)
  uint32_t* sBl[2] = {sAl[1] + kOpFloats, sAl[1] + 2 * kOpFloats};
  const bool vec_a = st.vec_a != 0, vec_b = st.vec_b != 0;
  const int per_s = wgrad ? taps * st.m_tiles * st.n_tiles * st.k_splits : st.m_tiles * st.n_tiles;
  const int K_total = d.Hout * d.Wout;

  for (int item = blockIdx.x; item < st.items; item += gridDim.x) {
    Tile t;
    t.s = item / per_s;
    int r = item - t.s * per_s;
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

    float4 ra[2], rb[kBSlots];
    auto fetch = [&](int it) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * kThreads;
        const int line = idx / a_lpl, q = idx - line * a_lpl;
        ra[j] = load4(a_line(st, t, it, line, kch), q, vec_a);
      }
#pragma unroll
      for (int j = 0; j < kBSlots; ++j) {
        const int idx = tid + j * kThreads;
        if (idx < BN * kBK / 4) {
          const int line = idx / b_lpl, q = idx - line * b_lpl;
          rb[j] = load4(b_line(st, t, it, line, kch, BN), q, vec_b);
        }
      }
    };
    auto stash = [&](int buf) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = tid + j * kThreads;
        const int line = idx / a_lpl, q = idx - line * a_lpl;
        store4<SPLIT3>(sA[buf], sAl[buf], line * a_ld + q * 4, ra[j]);
      }
#pragma unroll
      for (int j = 0; j < kBSlots; ++j) {
        const int idx = tid + j * kThreads;
        if (idx < BN * kBK / 4) {
          const int line = idx / b_lpl, q = idx - line * b_lpl;
          store4<SPLIT3>(sB[buf], sBl[buf], line * b_ld + q * 4, rb[j]);
        }
      }
    };

    __syncthreads();                 // the previous item's epilogue is done with shared memory
    if (n_it > 0) {
      fetch(0);
      stash(0);
    }
    __syncthreads();
    for (int it = 0; it < n_it; ++it) {
      const int buf = it & 1;
      if (it + 1 < n_it) fetch(it + 1);
      const uint32_t* A = sA[buf] + (wm * 16 + g) * a_rs;
      const uint32_t* B = sB[buf] + (wn * 8 * NT + g) * b_rs;
      const uint32_t* Al = sAl[buf] + (wm * 16 + g) * a_rs;
      const uint32_t* Bl = sBl[buf] + (wn * 8 * NT + g) * b_rs;
#pragma unroll
      for (int k0 = 0; k0 < kBK; k0 += 8) {
        const int ka = (k0 + tq) * a_ks, ka4 = (k0 + tq + 4) * a_ks;
        const uint32_t a0 = A[ka], a1 = A[8 * a_rs + ka], a2 = A[ka4], a3 = A[8 * a_rs + ka4];
        uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0;
        if (SPLIT3) {
          l0 = Al[ka]; l1 = Al[8 * a_rs + ka]; l2 = Al[ka4]; l3 = Al[8 * a_rs + ka4];
        }
        const int kb = (k0 + tq) * b_ks, kb4 = (k0 + tq + 4) * b_ks;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const uint32_t b0 = B[i * 8 * b_rs + kb], b1 = B[i * 8 * b_rs + kb4];
          if (SPLIT3) {
            const uint32_t m0 = Bl[i * 8 * b_rs + kb], m1 = Bl[i * 8 * b_rs + kb4];
            mma_tf32(acc[i], l0, l1, l2, l3, b0, b1);       // small terms first
            mma_tf32(acc[i], a0, a1, a2, a3, m0, m1);
          }
          mma_tf32(acc[i], a0, a1, a2, a3, b0, b1);
        }
      }
      if (it + 1 < n_it) stash(buf ^ 1);
      __syncthreads();
    }

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
    const MfviView& o = fwd ? st.b : st.b;           // forward: y, dgrad: dx (both recorded in `b`)
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
            orow[n] = st.accumulate ? orow[n] + v : v;
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
      float* red = smem_f;                         // [4 row warps][BN][2]; the operand buffers are idle (loop ended with a barrier)
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

template <int V>
__device__ void ew_stage(const Stage& st, EwSmem sm) {
  const int nvb = st.gx * st.S;
  for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
    const VGrid vg{vb % st.gx, vb / st.gx, st.gx};
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
        body_bn_bwd_apply<V, false>(vg, sm, st.a, st.b, st.S, st.H, st.W, st.C, st.sums, st.red, st.gamma, st.c, st.dgamma, st.dbeta,
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

__global__ void __launch_bounds__(kThreads, 2)
k_mega(const Stage* __restrict__ prog, int n_stages, unsigned* bar) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ Stage st;
  pdl_trigger();
  pdl_wait();
  float* smem_f = reinterpret_cast<float*>(smem_raw);
  // elementwise pieces overlay the convolution buffers (stages are sequential)
  EwSmem sm;
  sm.red = reinterpret_cast<double*>(smem_raw);
  sm.tab = reinterpret_cast<BnTable*>(smem_raw + 2 * kEwThreads * 4 * sizeof(double));
  sm.misc = reinterpret_cast<float*>(smem_raw + 2 * kEwThreads * 4 * sizeof(double) + sizeof(BnTable));
  unsigned epoch = 0;
  for (int i = 0; i < n_stages; ++i) {
    __syncthreads();
    {
      const int* src = reinterpret_cast<const int*>(prog + i);
      int* dst = reinterpret_cast<int*>(&st);
      for (int k = threadIdx.x; k < (int)(sizeof(Stage) / sizeof(int)); k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    if (i > 0 && !st.nosync) {
      ++epoch;
      grid_barrier(bar, epoch * gridDim.x);
    }
    switch (st.op) {
      case OP_CONV_FWD:
      case OP_CONV_DGRAD:
      case OP_CONV_WGRAD:
        if (st.split3) {
          if (st.nt == 4) conv_stage<4, true>(st, smem_f);
          else if (st.nt == 2) conv_stage<2, true>(st, smem_f);
          else conv_stage<1, true>(st, smem_f);
        } else {
          if (st.nt == 4) conv_stage<4, false>(st, smem_f);
          else if (st.nt == 2) conv_stage<2, false>(st, smem_f);
          else conv_stage<1, false>(st, smem_f);
        }
        break;
      case OP_FILL:
        for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < st.fill_n; k += (size_t)gridDim.x * blockDim.x)
          st.fill_ptr[k] = st.fill_v;
        break;
      default:
        if (st.V == 4) ew_stage<4>(st, sm);
        else ew_stage<1>(st, sm);
        break;
    }
  }
  // self-reset: a CTA gets here after its last barrier, so the last one to arrive knows nobody is waiting on bar[0] any more
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(bar + 1, 1u);
    if (done == gridDim.x - 1) {
      bar[0] = 0u;
      bar[1] = 0u;
    }
  }
}

constexpr size_t kEwSmemBytes = 2 * kEwThreads * 4 * sizeof(double) + sizeof(BnTable) + 5 * kMaxC * sizeof(float);
constexpr size_t kConvSmemBytes = 4 * kOpFloats * sizeof(float);          // A, B double-buffered
constexpr size_t kConvSmemBytes3 = 8 * kOpFloats * sizeof(float);         // + the lo parts

static bool view_vec(const MfviView& v) {
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && v.sstride % 4 == 0 && v.hstride % 4 == 0 && v.wstride % 4 == 0;
}

static int cdiv(int a, int b) { return (a + b - 1) / b; }

// tile plan of a convolution stage: the widest column tile that still gives every SM two work items, else the narrowest
static void plan_conv(Stage* s) {
  const MfviConvDesc& d = s->d;
  const int target = 2 * kNumSMs;
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

int record_conv(int op, const MfviConvDesc* d, MfviView act_in, MfviView act_out, const float* w, const float* bias,
                long long w_sstride, float* dw, double* stats, int accumulate) {
  MFVI_REQUIRE(d != nullptr && (d->stride == 1 || d->stride == 2), "mega: convolution stride must be 1 or 2");
  MFVI_REQUIRE(act_in.ptr != nullptr && act_out.ptr != nullptr, "mega: null activation view");
  Stage* s = append(op);
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
  plan_conv(s);
  return 0;
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
// a synchronous copy — plan-build time).  Reports the number of stages, the largest number of work items of a stage and whether
// any stage needs the 3xTF32 buffers.
int mfvi_mega_end(void* program_dev, size_t capacity_bytes, int* n_stages, int* max_items, int* any_split3) {
  MFVI_REQUIRE(mega::g_rec != nullptr, "mega_end: not recording");
  std::vector<mega::Stage>* rec = mega::g_rec;
  mega::g_rec = nullptr;
  const size_t bytes = rec->size() * sizeof(mega::Stage);
  int items = 1, s3 = 0;
  for (const mega::Stage& s : *rec) {
    const int it = (s.op <= mega::OP_CONV_WGRAD) ? s.items : (s.op == mega::OP_FILL ? kNumSMs : s.gx * s.S);
    items = std::max(items, it);
    s3 |= s.split3;
  }
  if (n_stages) *n_stages = static_cast<int>(rec->size());
  if (max_items) *max_items = items;
  if (any_split3) *any_split3 = s3;
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

// Runs a recorded program: ONE launch.  `barrier`: two device counters, zero on entry; the kernel leaves them zero.
int mfvi_mega_run(const void* program_dev, int n_stages, int max_items, int any_split3, unsigned* barrier, mfvi_stream_t st) {
  MFVI_REQUIRE(program_dev != nullptr && barrier != nullptr && n_stages >= 1, "mega_run: null program / barrier");
  const size_t smem = std::max(mega::kEwSmemBytes, any_split3 ? mega::kConvSmemBytes3 : mega::kConvSmemBytes);
  static size_t attr = 0;
  static int occ[2] = {0, 0};
  if (smem > attr) {
    const cudaError_t e = cudaFuncSetAttribute(mega::k_mega, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega::kConvSmemBytes3);
    MFVI_REQUIRE(e == cudaSuccess, "mega_run: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    attr = mega::kConvSmemBytes3;
  }
  int& oc = occ[any_split3 ? 1 : 0];
  if (oc == 0) {
    const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc, mega::k_mega, mega::kThreads, smem);
    MFVI_REQUIRE(e == cudaSuccess && oc >= 1, "mega_run: occupancy query failed: %s", cudaGetErrorString(e));
    if (oc > 2) oc = 2;
  }
  int dev = 0, sms = kNumSMs;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // never more CTAs than can be co-resident (the grid barrier spins), never more than the widest stage can use
  const int grid = std::max(1, std::min(oc * sms, max_items));
  launch_k(mega::k_mega, grid, mega::kThreads, smem, as_stream(st), static_cast<const mega::Stage*>(program_dev), n_stages, barrier);
  return check_launch("mega_run");
}

}  // extern "C"
