// fp32 CUDA-core implementation of the sampled-weight convolution (forward, dgrad, wgrad).
// This is the exact-fp32 parity path ("within rtol 1e-3 in fp32" of BASELINE.json) and the fallback for
// shapes the tcgen05 kernel (conv_tc.cu) does not take (Cout < 16, tiny images).
//
//   forward / dgrad : implicit GEMM, CTA tile 128 pixels x BN channels, K = taps x channels in chunks of 16,
//                     register tile TM x 4 per thread, operands staged through shared memory.
//   wgrad           : per-CTA strip of output pixels staged in shared memory (dy strip + input patch with halo),
//                     every thread owns a 4x4 (co x ci) block of ONE filter tap and accumulates over the strip;
//                     split-K over rows, fp32 atomics into dw.
#include <algorithm>

#include "common.cuh"

namespace mfvi {

constexpr int BM = 128;
constexpr int BK = 16;

struct ConvArgs {
  MfviConvDesc d;
  MfviView a;       // fwd: x (padded input)      dgrad: dy
  MfviView o;       // fwd: y                     dgrad: dx (padded-input sized)
  const float* w;   // [S][T][Cout][Cin]
  const float* bias;
  long long w_sstride;
  double* stats;
  int accumulate;
  int vecA;         // float4 loads of A allowed
  int vecO;         // float4 stores of the output allowed
};

template <int BN, bool DGRAD>
__global__ void __launch_bounds__(256)
k_conv_igemm(const ConvArgs p) {
  constexpr int TX = BN / 4;
  constexpr int TY = 256 / TX;
  constexpr int TM = BM / TY;
  constexpr int KQ = 256 / BN;       // k-slices among the B loader threads
  constexpr int KPER = BK / KQ;      // k values per B loader thread
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ double sm_stat[TY][BN][2];   // per-row-thread partial (sum, sumsq) of the BatchNorm statistics

  const MfviConvDesc& d = p.d;
  const int s = blockIdx.z;
  const int Mw = DGRAD ? d.Win : d.Wout;                 // width of the M pixel space
  const int M = DGRAD ? d.Hin * d.Win : d.Hout * d.Wout;
  const int N = DGRAD ? d.Cin : d.Cout;
  const int KC = DGRAD ? d.Cout : d.Cin;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;

  // A-loader role: one pixel, 8 consecutive k
  const int lm = tid % BM, lhalf = tid / BM;
  const int gm = m0 + lm;
  const bool m_ok = gm < M;
  const int ph = m_ok ? gm / Mw : 0, pw = m_ok ? gm % Mw : 0;
  const float* a_base = p.a.ptr + (size_t)s * p.a.sstride;
  // B-loader role
  const int bn = tid % BN, bkq = tid / BN;
  const float* w_base = p.w + (size_t)s * p.w_sstride;

  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int T = d.KH * d.KW;
  for (int tap = 0; tap < T; ++tap) {
    const int r = tap / d.KW, sx = tap % d.KW;
    // source pixel for this tap
    bool src_ok = m_ok;
    int ih, iw;
    if (!DGRAD) {
      ih = ph * d.stride + r;
      iw = pw * d.stride + sx;
    } else {
      const int hh = ph - r, ww = pw - sx;
      src_ok = src_ok && hh >= 0 && ww >= 0 && (hh % d.stride == 0) && (ww % d.stride == 0);
      ih = hh / d.stride;
      iw = ww / d.stride;
      src_ok = src_ok && ih < d.Hout && iw < d.Wout;
    }
    const float* a_pix = a_base + (size_t)ih * p.a.hstride + (size_t)iw * p.a.wstride;
    const float* w_tap = w_base + (size_t)tap * d.Cout * d.Cin;
    for (int c0 = 0; c0 < KC; c0 += BK) {
      // ---- stage A: As[k][m] = a[pixel m][c0 + k]
      {
        const int kb = lhalf * 8;
        float v[8];
        if (src_ok && p.vecA && c0 + kb + 8 <= KC) {
          const float4 t0 = *reinterpret_cast<const float4*>(a_pix + c0 + kb);
          const float4 t1 = *reinterpret_cast<const float4*>(a_pix + c0 + kb + 4);
          v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w;
          v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (src_ok && c0 + kb + j < KC) ? a_pix[c0 + kb + j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) As[kb + j][lm] = v[j];
      }
      // ---- stage B: Bs[k][n]
      {
#pragma unroll
        for (int j = 0; j < KPER; ++j) {
          const int k = bkq * KPER + j;
          const int kc = c0 + k, nn = n0 + bn;
          float v = 0.f;
          if (kc < KC && nn < N) {
            // storage [tap][co][ci]; fwd: k = ci, n = co; dgrad: k = co, n = ci
            v = DGRAD ? w_tap[(size_t)kc * d.Cin + nn] : w_tap[(size_t)nn * d.Cin + kc];
          }
          Bs[k][bn] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
          acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
          acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
          acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
        }
      }
      __syncthreads();
    }
  }

  // ---- epilogue
  const int nb = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (!DGRAD && p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (nb + j < N) bv[j] = p.bias[(size_t)s * p.w_sstride + nb + j];
  }
  // statistics in double: var = E[x^2] - mean^2 cancels badly when |mean| >> std, so no fp32 partial sums
  double st1[4] = {0.0, 0.0, 0.0, 0.0}, st2[4] = {0.0, 0.0, 0.0, 0.0};
  float* o_base = p.o.ptr + (size_t)s * p.o.sstride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
    const int oh = m / Mw, ow = m % Mw;
    float* dst = o_base + (size_t)oh * p.o.hstride + (size_t)ow * p.o.wstride + nb;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bv[j];
    if (p.vecO && nb + 3 < N) {
      if (DGRAD && p.accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(dst);
        v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
      }
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (nb + j < N) dst[j] = (DGRAD && p.accumulate) ? dst[j] + v[j] : v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      st1[j] += (double)v[j];
      st2[j] += (double)v[j] * (double)v[j];
    }
  }
  if (!DGRAD && p.stats != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sm_stat[ty][tx * 4 + j][0] = st1[j];
      sm_stat[ty][tx * 4 + j][1] = st2[j];
    }
    __syncthreads();
    for (int c = tid; c < BN; c += 256) {
      if (n0 + c < N) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < TY; ++r) {       // fixed order: deterministic per CTA
          a += sm_stat[r][c][0];
          b += sm_stat[r][c][1];
        }
        double* dst = p.stats + ((size_t)s * N + n0 + c) * 2;
        atomicAdd(dst + 0, a);
        atomicAdd(dst + 1, b);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct WgradArgs {
  MfviConvDesc d;
  MfviView x, dy;
  float* dw;
  float* dbias;
  long long w_sstride;
  int BCO, BCI, BW;       // tile sizes
  int rows_per_cta;
  int n_co_tiles, n_ci_tiles;
  int vec;
};

__global__ void k_conv_wgrad(const WgradArgs p) {
  extern __shared__ __align__(16) float smem[];
  const MfviConvDesc& d = p.d;
  const int T = d.KH * d.KW;
  const int NCOG = p.BCO / 4, NCIG = p.BCI / 4;
  const int XW = (p.BW - 1) * d.stride + d.KW;
  float* dy_s = smem;                              // [BW][BCO]
  float* x_s = smem + (size_t)p.BW * p.BCO;        // [KH][XW][BCI]

  const int s = blockIdx.z;
  const int tile = blockIdx.y;
  const int co0 = (tile / p.n_ci_tiles) * p.BCO;
  const int ci0 = (tile % p.n_ci_tiles) * p.BCI;
  const int row0 = blockIdx.x * p.rows_per_cta;
  const int row1 = min(row0 + p.rows_per_cta, d.Hout);

  const int tid = threadIdx.x;
  const int nblk = T * NCOG * NCIG;
  const bool worker = tid < nblk;
  const int cig = tid % NCIG;
  const int cog = (tid / NCIG) % NCOG;
  const int tap = worker ? tid / (NCIG * NCOG) : 0;
  const int r = tap / d.KW, sx = tap % d.KW;
  const bool do_bias = worker && p.dbias != nullptr && tap == 0 && cig == 0 && ci0 == 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};

  const float* x_base = p.x.ptr + (size_t)s * p.x.sstride;
  const float* dy_base = p.dy.ptr + (size_t)s * p.dy.sstride;

  for (int oh = row0; oh < row1; ++oh) {
    for (int w0 = 0; w0 < d.Wout; w0 += p.BW) {
      // ---- stage dy strip
      const int n_dy = p.BW * NCOG;
      for (int idx = tid; idx < n_dy; idx += blockDim.x) {
        const int pp = idx / NCOG, cg = idx % NCOG;
        const int ow = w0 + pp, co = co0 + cg * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ow < d.Wout) {
          const float* src = dy_base + (size_t)oh * p.dy.hstride + (size_t)ow * p.dy.wstride + co;
          if (p.vec && co + 3 < d.Cout) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            if (co + 0 < d.Cout) v.x = src[0];
            if (co + 1 < d.Cout) v.y = src[1];
            if (co + 2 < d.Cout) v.z = src[2];
            if (co + 3 < d.Cout) v.w = src[3];
          }
        }
        *reinterpret_cast<float4*>(dy_s + (size_t)pp * p.BCO + cg * 4) = v;
      }
      // ---- stage input patch (KH rows, XW cols)
      const int n_x = d.KH * XW * NCIG;
      for (int idx = tid; idx < n_x; idx += blockDim.x) {
        const int cg = idx % NCIG;
        const int col = (idx / NCIG) % XW;
        const int rr = idx / (NCIG * XW);
        const int ih = oh * d.stride + rr, iw = w0 * d.stride + col, ci = ci0 + cg * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iw < d.Win && ih < d.Hin) {
          const float* src = x_base + (size_t)ih * p.x.hstride + (size_t)iw * p.x.wstride + ci;
          if (p.vec && ci + 3 < d.Cin) {
            v = *reinterpret_cast<const float4*>(src);
          } else {
            if (ci + 0 < d.Cin) v.x = src[0];
            if (ci + 1 < d.Cin) v.y = src[1];
            if (ci + 2 < d.Cin) v.z = src[2];
            if (ci + 3 < d.Cin) v.w = src[3];
          }
        }
        *reinterpret_cast<float4*>(x_s + ((size_t)rr * XW + col) * p.BCI + cg * 4) = v;
      }
      __syncthreads();
      if (worker) {
        const float* xa = x_s + ((size_t)r * XW + sx) * p.BCI + cig * 4;
        const float* da = dy_s + cog * 4;
        const int xstep = d.stride * p.BCI;
#pragma unroll 4
        for (int pp = 0; pp < p.BW; ++pp) {
          const float4 a = *reinterpret_cast<const float4*>(da + (size_t)pp * p.BCO);
          const float4 b = *reinterpret_cast<const float4*>(xa + (size_t)pp * xstep);
          acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
          acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
          acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
          acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
          acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
          acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
          acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
          acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
          if (do_bias) {
            bacc[0] += a.x; bacc[1] += a.y; bacc[2] += a.z; bacc[3] += a.w;
          }
        }
      }
      __syncthreads();
    }
  }
  if (worker) {
    float* dst = p.dw + (size_t)s * p.w_sstride + (size_t)tap * d.Cout * d.Cin;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + cog * 4 + i;
      if (co >= d.Cout) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = ci0 + cig * 4 + j;
        if (ci < d.Cin) atomicAdd(dst + (size_t)co * d.Cin + ci, acc[i][j]);
      }
      if (do_bias) atomicAdd(p.dbias + (size_t)s * p.w_sstride + co, bacc[i]);
    }
  }
}

static inline bool view_vec4(const MfviView& v) {
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && (v.sstride % 4 == 0) && (v.hstride % 4 == 0) &&
         (v.wstride % 4 == 0);
}

static int validate_desc(const MfviConvDesc* d, const char* who) {
  MFVI_REQUIRE(d != nullptr, "%s: null descriptor", who);
  MFVI_REQUIRE(d->S >= 1 && d->S <= 65535, "%s: S=%d out of range", who, d->S);
  MFVI_REQUIRE(d->Cin >= 1 && d->Cout >= 1, "%s: bad channel counts", who);
  MFVI_REQUIRE(d->KH >= 1 && d->KW >= 1 && d->KH * d->KW <= 49, "%s: bad kernel size", who);
  MFVI_REQUIRE(d->stride >= 1, "%s: bad stride", who);
  MFVI_REQUIRE(d->Hin >= d->KH && d->Win >= d->KW, "%s: input smaller than the filter", who);
  MFVI_REQUIRE(d->Hout == (d->Hin - d->KH) / d->stride + 1 && d->Wout == (d->Win - d->KW) / d->stride + 1,
               "%s: Hout/Wout inconsistent with Hin/Win/K/stride", who);
  return 0;
}

template <bool DGRAD>
static int launch_igemm(const ConvArgs& a, cudaStream_t st) {
  const MfviConvDesc& d = a.d;
  const int M = DGRAD ? d.Hin * d.Win : d.Hout * d.Wout;
  const int N = DGRAD ? d.Cin : d.Cout;
  const int gm = (M + BM - 1) / BM;
  const dim3 grid(gm, N > 32 ? (N + 63) / 64 : 1, d.S);
  if (dry_run() != nullptr) {
    dry_detail("igemm BM=%d BN=%d", BM, N > 32 ? 64 : (N > 16 ? 32 : 16));
    dry_note(grid, 256, 0);
  } else if (N > 32) {
    k_conv_igemm<64, DGRAD><<<grid, 256, 0, st>>>(a);
  } else if (N > 16) {
    k_conv_igemm<32, DGRAD><<<grid, 256, 0, st>>>(a);
  } else {
    k_conv_igemm<16, DGRAD><<<grid, 256, 0, st>>>(a);
  }
  return check_launch(DGRAD ? "conv2d_dgrad" : "conv2d_fwd");
}

}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_conv2d_fwd_simt(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride,
                         MfviView y, double* stats, mfvi_stream_t st) {
  if (int rc = validate_desc(d, "conv2d_fwd")) return rc;
  MFVI_REQUIRE(x.ptr && w && y.ptr, "conv2d_fwd: null pointer");
  ConvArgs a;
  a.d = *d; a.a = x; a.o = y; a.w = w; a.bias = bias; a.w_sstride = w_sstride; a.stats = stats; a.accumulate = 0;
  a.vecA = view_vec4(x) ? 1 : 0;
  a.vecO = view_vec4(y) ? 1 : 0;
  return launch_igemm<false>(a, as_stream(st));
}

int mfvi_conv2d_dgrad_simt(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx,
                           int accumulate, mfvi_stream_t st) {
  if (int rc = validate_desc(d, "conv2d_dgrad")) return rc;
  MFVI_REQUIRE(dy.ptr && w && dx.ptr, "conv2d_dgrad: null pointer");
  MFVI_REQUIRE(dx.sstride != 0 || d->S == 1, "conv2d_dgrad: dx cannot be a broadcast view");
  ConvArgs a;
  a.d = *d; a.a = dy; a.o = dx; a.w = w; a.bias = nullptr; a.w_sstride = w_sstride; a.stats = nullptr;
  a.accumulate = accumulate;
  a.vecA = view_vec4(dy) ? 1 : 0;
  a.vecO = view_vec4(dx) ? 1 : 0;
  return launch_igemm<true>(a, as_stream(st));
}

int mfvi_conv2d_wgrad_simt(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias,
                           long long w_sstride, mfvi_stream_t st) {
  if (int rc = validate_desc(d, "conv2d_wgrad")) return rc;
  MFVI_REQUIRE(x.ptr && dy.ptr && dw, "conv2d_wgrad: null pointer");
  WgradArgs a;
  a.d = *d; a.x = x; a.dy = dy; a.dw = dw; a.dbias = dbias; a.w_sstride = w_sstride;
  const int T = d->KH * d->KW;
  auto up4 = [](int v) { return (v + 3) / 4 * 4; };
  int cap = T == 1 ? 64 : 32;
  a.BCO = std::min(cap, up4(d->Cout));
  a.BCI = std::min(cap, up4(d->Cin));
  while (T * (a.BCO / 4) * (a.BCI / 4) > 1024) {
    if (a.BCI >= a.BCO && a.BCI > 4) a.BCI = up4(a.BCI / 2); else a.BCO = up4(a.BCO / 2);
  }
  a.BW = d->stride == 1 ? 64 : 32;
  if (a.BW > d->Wout) a.BW = std::max(4, up4(d->Wout));
  const int XW = (a.BW - 1) * d->stride + d->KW;
  size_t smem = ((size_t)a.BW * a.BCO + (size_t)d->KH * XW * a.BCI) * sizeof(float);
  while (smem > 96 * 1024 && a.BW > 8) {
    a.BW /= 2;
    const int xw = (a.BW - 1) * d->stride + d->KW;
    smem = ((size_t)a.BW * a.BCO + (size_t)d->KH * xw * a.BCI) * sizeof(float);
  }
  MFVI_REQUIRE(smem <= 200 * 1024, "conv2d_wgrad: tile does not fit in shared memory");
  a.n_co_tiles = (d->Cout + a.BCO - 1) / a.BCO;
  a.n_ci_tiles = (d->Cin + a.BCI - 1) / a.BCI;
  const int tiles = a.n_co_tiles * a.n_ci_tiles;
  // aim for ~4 waves of CTAs
  int want_chunks = std::max(1, (kNumSMs * 4) / std::max(1, tiles * d->S));
  a.rows_per_cta = std::max(1, (d->Hout + want_chunks - 1) / want_chunks);
  const int chunks = (d->Hout + a.rows_per_cta - 1) / a.rows_per_cta;
  a.vec = (view_vec4(x) && view_vec4(dy)) ? 1 : 0;
  int threads = T * (a.BCO / 4) * (a.BCI / 4);
  threads = std::max(64, (threads + 31) / 32 * 32);
  static unsigned long long attr_done = 0;
  if (smem > 48 * 1024 && dry_run() == nullptr) {
    const cudaError_t e = allow_dyn_smem(k_conv_wgrad, 200 * 1024, &attr_done);
    MFVI_REQUIRE(e == cudaSuccess, "conv2d_wgrad: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
  }
  dim3 grid(chunks, tiles, d->S);
  if (dry_run() != nullptr) {
    dry_detail("BCO=%d BCI=%d BW=%d rows_per_cta=%d", a.BCO, a.BCI, a.BW, a.rows_per_cta);
    dry_note(grid, threads, smem);
  } else {
    k_conv_wgrad<<<grid, threads, smem, as_stream(st)>>>(a);
  }
  return check_launch("conv2d_wgrad");
}

}  // extern "C"
