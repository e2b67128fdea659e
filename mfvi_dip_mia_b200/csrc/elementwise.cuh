// Shared building blocks of the skip-net elementwise kernels (elementwise.cu, elementwise_fused.cu): vector access, BatchNorm
// tables, the pixel iteration scheme, per-channel reductions, launch geometry.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mfvi {

constexpr int kEwThreads = 256;
constexpr int kMaxC = 512;  // per-kernel channel limit of the smem scale/shift tables

struct EwGeom {
  int V;      // vector width (4 or 1)
  int G;      // channel groups = ceil(C / V)
  int PPB;    // pixel slots per CTA iteration
};

static inline EwGeom ew_geom(int C, bool aligned) {
  EwGeom g;
  g.V = (C % 4 == 0 && aligned) ? 4 : 1;
  g.G = (C + g.V - 1) / g.V;
  g.PPB = kEwThreads / g.G;
  if (g.PPB < 1) g.PPB = 1;
  return g;
}

static inline bool view_vec_ok(const MfviView& v) {
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && (v.sstride % 4 == 0) && (v.hstride % 4 == 0) &&
         (v.wstride % 4 == 0);
}

template <int V>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

// bf16 outputs (bf16-operand mode, DESIGN.md section 8, stage C): the value is computed in fp32 exactly as for an fp32 output
// and rounded to nearest-even at the store; V = 4 is one 8-byte store.
template <int V>
__device__ __forceinline__ void store_bf16(__nv_bfloat16* p, const float (&v)[V]) {
  if (V == 4) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[V > 1 ? 2 : 0], v[V > 1 ? 3 : 0]);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&lo);
    u.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) p[j] = __float2bfloat16_rn(v[j]);
  }
}

// Per-CTA tables: scale[c] = gamma*invstd, shift[c] = beta - mean*scale  (z = y*scale + shift),
// mean[c], invstd[c] for sample s.
__device__ __forceinline__ void load_bn_tables(const double* __restrict__ sums, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, int s, int C, double inv_count,
                                               float* sm_scale, float* sm_shift, float* sm_mean, float* sm_invstd) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean = 0.f, invstd = 1.f;
    if (sums != nullptr) bn_mean_invstd(sums + ((size_t)s * C + c) * 2, inv_count, mean, invstd);
    const float g = gamma != nullptr ? gamma[c] : 1.f;
    const float b = beta != nullptr ? beta[c] : 0.f;
    const float sc = g * invstd;
    sm_scale[c] = sc;
    sm_shift[c] = b - mean * sc;
    if (sm_mean) sm_mean[c] = mean;
    if (sm_invstd) sm_invstd[c] = invstd;
  }
}

// CTA-level reduction of per-thread partial sums for a fixed channel (threads [slot][group] layout) without
// shared-memory atomics: every thread parks its V partial sums in a [slot][C] table, thread c then adds the column
// of channel c over the slots, and issues one double atomicAdd per channel per CTA into dst[c*2 + {0,1}].
// `scratch` must hold 2 * kEwThreads * 4 doubles.  Partial sums are carried in double: a CTA now covers thousands of
// pixels and the BatchNorm backward subtracts these means from values of the same size (cancellation).
template <int V>
__device__ __forceinline__ void cta_reduce_2(double (&a)[V], double (&b)[V], int group, int slot, int G, int PPB, int C,
                                             double* scratch, double* __restrict__ dst, bool active) {
  double* ta = scratch;
  double* tb = scratch + kEwThreads * 4;
  const int ld = G * V;                       // >= C
  __syncthreads();
  if (active) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      ta[slot * ld + group * V + j] = a[j];
      tb[slot * ld + group * V + j] = b[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double sa = 0.0, sb = 0.0;
    for (int sl = 0; sl < PPB; ++sl) {
      sa += ta[sl * ld + c];
      sb += tb[sl * ld + c];
    }
    atomicAdd(&dst[(size_t)c * 2 + 0], sa);
    atomicAdd(&dst[(size_t)c * 2 + 1], sb);
  }
}

// ---------------------------------------------------------------------------------------------
// Common iteration scheme of the kernels below.  grid = (chunks, S).  A CTA owns a CONTIGUOUS range of the pixel space
// [p0, p1); thread (slot, group) owns channel quad `group` and walks pixels p0+slot, p0+slot+PPB, ... keeping (h, w)
// incrementally (no division in the loop).  The BatchNorm constants of the thread's channels live in registers, and the
// per-channel reductions are carried in fp32 for kFlush pixels at a time before they are added to double accumulators.
constexpr int kFlush = 8;

// Virtual block coordinates: (bx, by) of a (gx, S) grid.  A stand-alone kernel passes its real blockIdx / gridDim; the persistent
// multi-stage kernel (mega.cu) loops over virtual blocks.
struct VGrid {
  int bx, by, gx;
};

// shared-memory pieces a kernel body may use (each wrapper allocates only what its body touches)
struct BnTable;
struct EwSmem {
  double* red;      // [2 * kEwThreads * 4] doubles: thread-private reduction cells
  BnTable* tab;     // BatchNorm constants of one sample
  float* misc;      // [5 * kMaxC] floats (bn_bwd_apply)
  float* pipe = nullptr;   // [kPipeBytes] thread-private cp.async slots of the PIPE variants (16-byte aligned)
};

// cp.async load pipeline of the stand-alone kernels.  Every thread owns D x NL slots of V floats in shared memory
// ([stage][load][thread]: conflict-free) and keeps D pixels' loads in flight without holding registers for them; the slots are
// thread-private, so the only synchronisation is cp.async.wait_group.  These kernels are latency-bound (one or two pixels of
// register loads in flight per thread, profiles/r02_ew_load_batching.txt), which is what the deeper pipeline addresses.
constexpr int kPipeBytes = 32 * 1024;
template <int V>
__device__ __forceinline__ void ew_cp_async(float* smem_dst, const float* g) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  if constexpr (V == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(g) : "memory");
}
__device__ __forceinline__ void ew_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void ew_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int V, int NL>
struct LdPipe {
  static constexpr int D = kPipeBytes / (NL * kEwThreads * V * 4) > 8 ? 8 : kPipeBytes / (NL * kEwThreads * V * 4);   // stages
  float* base;
  __device__ __forceinline__ explicit LdPipe(float* pipe) : base(pipe + threadIdx.x * V) {}
  __device__ __forceinline__ float* slot(int d, int l) const { return base + (d * NL + l) * (kEwThreads * V); }
};

struct PixIter {
  int p, npix, h, w, W, step, dh, dw;
  // grid-stride over pixels (all CTAs sweep the image together, which keeps DRAM pages hot); (h, w) advance incrementally
  __device__ __forceinline__ PixIter(const VGrid& vg, int npix_, int Wd, int PPB, int slot) {
    npix = npix_;
    p = vg.bx * PPB + slot;
    W = Wd;
    step = vg.gx * PPB;
    dh = step / Wd;
    dw = step - dh * Wd;
    h = p / Wd;
    w = p - h * Wd;
  }
  __device__ __forceinline__ bool valid() const { return p < npix; }
  __device__ __forceinline__ void next() {
    p += step;
    w += dw;
    h += dh;
    if (w >= W) {
      w -= W;
      ++h;
    }
  }
};

// U consecutive positions of a PixIter: the kernels issue the loads of a whole batch before they use the first one (a thread
// otherwise has one pixel's loads in flight -- the store to a possibly aliasing view keeps the compiler from hoisting the
// next trip's loads -- and these kernels are latency-bound: profiles/r02_ew_load_batching.txt).  Positions past the end
// repeat position 0 (their loads are harmless duplicates), n = the valid ones.
template <int U>
struct PixBatch {
  int h[U], w[U], n;
  __device__ __forceinline__ bool fill(PixIter& it) {
    if (!it.valid()) return false;
    h[0] = it.h; w[0] = it.w; n = 1;
    it.next();
#pragma unroll
    for (int u = 1; u < U; ++u) {
      if (it.valid()) {
        h[u] = it.h; w[u] = it.w; n = u + 1;
        it.next();
      } else {
        h[u] = h[0]; w[u] = w[0];
      }
    }
    return true;
  }
};

// BatchNorm constants of one sample: computed once per CTA (one thread per channel, double rsqrt) into shared memory,
// then every thread keeps the V channels it owns in registers.  z = y*sc + sh ; xhat = (y - mean)*invstd
struct BnTable {
  float sc[kMaxC], sh[kMaxC], mean[kMaxC], invstd[kMaxC];
  __device__ __forceinline__ void fill(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       int s, int C, double inv_count, int dst0 = 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float m = 0.f, is = 1.f;
      if (sums != nullptr) bn_mean_invstd(sums + ((size_t)s * C + c) * 2, inv_count, m, is);
      const float g = gamma != nullptr ? gamma[c] : 1.f;
      const float b = beta != nullptr ? beta[c] : 0.f;
      sc[dst0 + c] = g * is;
      sh[dst0 + c] = b - m * g * is;
      mean[dst0 + c] = m;
      invstd[dst0 + c] = is;
    }
  }
};

template <int V>
struct BnRegs {
  float sc[V], sh[V], mean[V], invstd[V];
  __device__ __forceinline__ void load(const BnTable& t, int c0, int C) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const bool ok = c0 + j < C;
      sc[j] = ok ? t.sc[c0 + j] : 0.f;
      sh[j] = ok ? t.sh[c0 + j] : 0.f;
      mean[j] = ok ? t.mean[c0 + j] : 0.f;
      invstd[j] = ok ? t.invstd[c0 + j] : 1.f;
    }
  }
};

template <int V>
struct Acc2 {
  float fa[V], fb[V];
  double* cell;        // thread-private doubles in shared memory, laid out [2V][kEwThreads] (bank-conflict free); keeps 4V
                       // registers free -> higher occupancy
  int n;
  __device__ __forceinline__ explicit Acc2(double* scratch) : n(0) {
    cell = scratch + threadIdx.x;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      fa[j] = fb[j] = 0.f;
      cell[j * kEwThreads] = 0.0;
      cell[(V + j) * kEwThreads] = 0.0;
    }
  }
  __device__ __forceinline__ void flush() {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      cell[j * kEwThreads] += (double)fa[j];
      cell[(V + j) * kEwThreads] += (double)fb[j];
      fa[j] = fb[j] = 0.f;
    }
    n = 0;
  }
  __device__ __forceinline__ void tick() {
    if (++n == kFlush) flush();
  }
};

// CTA-level reduction of the thread-private (sum a, sum b) cells ([2V][thread] doubles, threads laid out [slot][group]):
// thread c adds the cells of channel c over the pixel slots and issues one double atomicAdd per channel into dst[c*2+{0,1}].
template <int V>
__device__ __forceinline__ void cta_reduce_cells(const double* cells, int G, int PPB, int C, double* __restrict__ dst) {
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int group = c / V, j = c - group * V;
    double sa = 0.0, sb = 0.0;
    for (int sl = 0; sl < PPB; ++sl) {
      const double* cell = cells + (sl * G + group);
      sa += cell[j * kEwThreads];
      sb += cell[(V + j) * kEwThreads];
    }
    atomicAdd(&dst[(size_t)c * 2 + 0], sa);
    atomicAdd(&dst[(size_t)c * 2 + 1], sb);
  }
}

// number of padded positions (per dimension) that reflect onto source index h: fills q[0..n)
__device__ __forceinline__ int fold_sources(int h, int n, int pad, int (&q)[3]) {
  int cnt = 0;
  q[cnt++] = h + pad;
  if (h >= 1 && h <= pad) q[cnt++] = pad - h;
  if (h <= n - 2 && h >= n - 1 - pad) q[cnt++] = 2 * (n - 1) - h + pad;
  return cnt;
}

// chunks per sample.  The whole launch (all S samples) is ONE resident wave: at most `occ` CTAs per SM -- the kernel's own
// occupancy (MFVI_EW_OCC), capped by MFVI_EW_CTAS -- so that no partial second wave runs at a fraction of the machine (the
// grid-stride sweep gives every CTA the same work).  Small launches take fewer CTAs still: every thread should sweep at least
// kMinIter pixels, which amortises the per-CTA prologue (BN tables) and epilogue (reductions, double atomics), down to a floor
// of kFloorCtas CTAs per SM.  MFVI_EW_CTAS / MFVI_EW_MINITER / MFVI_EW_FLOOR override the three constants (measurement knobs).
constexpr int kCtasPerSM = 8, kMinIter = 4, kFloorCtas = 2, kRedCtas = 128;
static inline int ew_knob(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : 0;
  return v > 0 ? v : dflt;
}
static inline int ew_grid(int npix, int PPB, int S, int occ = kCtasPerSM, bool reducing = false) {
  static const int cap = ew_knob("MFVI_EW_CTAS", kCtasPerSM), miniter = ew_knob("MFVI_EW_MINITER", kMinIter),
                   floor_ctas = ew_knob("MFVI_EW_FLOOR", kFloorCtas);
  const int blocks = (npix + PPB - 1) / PPB;                      // per sample, one pixel per thread
  if (occ > cap) occ = cap;
  if (occ < 1) occ = 1;
  const long total = (long)blocks * S;
  long want = (total + miniter - 1) / miniter;
  const long lo = (long)kNumSMs * (floor_ctas < occ ? floor_ctas : occ), hi = (long)kNumSMs * occ;
  if (want < lo) want = lo;
  if (want > hi) want = hi;
  long gx = want / S;
  if (gx > blocks) gx = blocks;
  // a reducing kernel ends with one double atomic per channel and CTA on the sample's accumulators, which the L2 serialises
  // per address (~12 ns each, profiles/r02_stat_atomics_negative.txt): more than kRedCtas CTAs per sample cost more in that
  // burst than they save in the sweep (only reached with few MC samples per GPU)
  static const int red_ctas = ew_knob("MFVI_EW_RED_CTAS", kRedCtas);
  if (reducing && gx > red_ctas) gx = red_ctas;
  return gx < 1 ? 1 : (int)gx;
}

}  // namespace mfvi
