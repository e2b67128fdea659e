// 1x1 (pointwise), stride-1 sampled-weight convolution with FEW output channels (Cout <= 4, Cin <= 32): forward, data
// gradient and weight (+ bias) gradient on CUDA cores, exact fp32.
//
// These layers (the 16->4 / 32->4 skip convs and the final 16->2 conv of the metric net) are pure streaming: 70..150 bytes
// per pixel against <= 256 FLOP, and an M=128 x N=16 tensor-core tile would compute 4..8x padding.  Measured against the
// tcgen05 path (eager, 256^2, S=8): 16->2 forward 24 -> 15 us, its dgrad 22 -> 17 us, 16->4 forward 32 -> 24 us.  With 16 or 32
// output channels the same kernel loses (16->16: 23 -> 35 us; register- and LDS-bound), so those stay on the tensor cores.
// One thread owns one pixel: it loads the pixel's Cin values as float4s, multiplies by the
// sample's weight matrix held in shared memory (broadcast reads), and stores Cout values as float4s — coalesced both ways.
//   forward : y[s,p,co] = b[s,co] + sum_ci x[s,p,ci] * w[s][co][ci]        (+ per-sample BatchNorm (sum, sumsq) in double)
//   dgrad   : dx[s,p,ci] (+)= sum_co dy[s,p,co] * w[s][co][ci]
// Both are out[p][n] = sum_k in[p][k] * M[k][n] with M = w^T (forward) or w (dgrad), zero-padded to [4*KP][4*NP].
//   wgrad   : dw[s][co][ci] += sum_p dy[s,p,co] * x[s,p,ci],  dbias[s][co] += sum_p dy[s,p,co]   (k_wgrad_pointwise below)
#include "common.cuh"

namespace mfvi {
namespace pw {

constexpr int kThreads = 256;

struct Args {
  MfviView in, out;
  const float* w;          // [S][Cout][Cin] (one tap)
  const float* bias;       // [S][Cout] or null
  long long w_sstride;
  double* stats;           // [S][N][2] or null
  int K, N, HW, W, dgrad, accumulate, Cout, Cin;
};

template <int KP, int NP, bool STATS>
__global__ void __launch_bounds__(kThreads)
k_conv_pointwise(const Args p) {
  __shared__ __align__(16) float M[4 * KP][4 * NP];
  __shared__ float bias_s[4 * NP];
  __shared__ float red[kThreads / 32][8 * NP];
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.y;
  const float* w = p.w + static_cast<size_t>(s) * p.w_sstride;
  for (int i = threadIdx.x; i < 16 * KP * NP; i += kThreads) {
    const int k = i / (4 * NP), n = i - k * (4 * NP);
    float v = 0.f;
    if (k < p.K && n < p.N) v = p.dgrad ? w[static_cast<size_t>(k) * p.Cin + n] : w[static_cast<size_t>(n) * p.Cin + k];
    M[k][n] = v;
  }
  if (threadIdx.x < 4 * NP)
    bias_s[threadIdx.x] = (p.bias != nullptr && threadIdx.x < p.N) ? p.bias[static_cast<size_t>(s) * p.w_sstride + threadIdx.x] : 0.f;
  __syncthreads();
  const bool vec_in = (p.K % 4 == 0), vec_out = (p.N % 4 == 0);
  float s1[4 * NP], s2[4 * NP];
#pragma unroll
  for (int n = 0; n < 4 * NP; ++n) s1[n] = s2[n] = 0.f;
  const float* ibase = p.in.ptr + static_cast<size_t>(s) * p.in.sstride;
  float* obase = p.out.ptr + static_cast<size_t>(s) * p.out.sstride;
  for (int pix = blockIdx.x * kThreads + threadIdx.x; pix < p.HW; pix += gridDim.x * kThreads) {
    const int h = pix / p.W, wq = pix - h * p.W;
    const float* ip = ibase + static_cast<size_t>(h) * p.in.hstride + static_cast<size_t>(wq) * p.in.wstride;
    float* op = obase + static_cast<size_t>(h) * p.out.hstride + static_cast<size_t>(wq) * p.out.wstride;
    float in[4 * KP];
    if (vec_in) {
#pragma unroll
      for (int k4 = 0; k4 < KP; ++k4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * k4 < p.K) t = __ldg(reinterpret_cast<const float4*>(ip) + k4);
        in[4 * k4] = t.x; in[4 * k4 + 1] = t.y; in[4 * k4 + 2] = t.z; in[4 * k4 + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4 * KP; ++k) in[k] = k < p.K ? __ldg(ip + k) : 0.f;
    }
    float acc[4 * NP];
#pragma unroll
    for (int n = 0; n < 4 * NP; ++n) acc[n] = bias_s[n];
#pragma unroll
    for (int k = 0; k < 4 * KP; ++k) {
#pragma unroll
      for (int n4 = 0; n4 < NP; ++n4) {
        const float4 m = *reinterpret_cast<const float4*>(&M[k][4 * n4]);
        acc[4 * n4] = fmaf(in[k], m.x, acc[4 * n4]);
        acc[4 * n4 + 1] = fmaf(in[k], m.y, acc[4 * n4 + 1]);
        acc[4 * n4 + 2] = fmaf(in[k], m.z, acc[4 * n4 + 2]);
        acc[4 * n4 + 3] = fmaf(in[k], m.w, acc[4 * n4 + 3]);
      }
    }
    if (STATS) {
#pragma unroll
      for (int n = 0; n < 4 * NP; ++n) {
        s1[n] += acc[n];
        s2[n] = fmaf(acc[n], acc[n], s2[n]);
      }
    }
    if (vec_out) {
#pragma unroll
      for (int n4 = 0; n4 < NP; ++n4) {
        if (4 * n4 >= p.N) break;
        float4 o = make_float4(acc[4 * n4], acc[4 * n4 + 1], acc[4 * n4 + 2], acc[4 * n4 + 3]);
        float4* dst = reinterpret_cast<float4*>(op) + n4;
        if (p.accumulate) {
          const float4 old = *dst;
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *dst = o;
      }
    } else {
#pragma unroll
      for (int n = 0; n < 4 * NP; ++n)
        if (n < p.N) op[n] = p.accumulate ? op[n] + acc[n] : acc[n];
    }
  }
  if (STATS) {
    // (sum, sumsq) per channel: warp shuffle, one row per warp in shared memory, double across warps, one atomic per cell
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < 4 * NP; ++n) {
      const float a = warp_sum(s1[n]), b = warp_sum(s2[n]);
      if (lane == 0) {
        red[warp][2 * n] = a;
        red[warp][2 * n + 1] = b;
      }
    }
    __syncthreads();
    if (threadIdx.x < 8 * NP) {
      const int n = threadIdx.x >> 1;
      if (n < p.N) {
        double t = 0.0;
        for (int wi = 0; wi < kThreads / 32; ++wi) t += static_cast<double>(red[wi][threadIdx.x]);
        atomicAdd(p.stats + (static_cast<size_t>(s) * p.N + n) * 2 + (threadIdx.x & 1), t);
      }
    }
  }
}

// Weight gradient of the same layers: a reduction over all pixels into Cout x Cin <= 128 numbers — 33 MB of x and 4..8 MB of
// dy per launch at 256^2 / S = 8, i.e. streaming work (in the step's graph the tcgen05 alias kernel + bias-gradient kernel
// took 29 us for 16->2 and 18 us for 16->4: profiles/r02_plan_link_cost_mc8.txt).  A thread owns one pixel slot and one
// quad of input channels: per pixel one float4 of x, the pixel's <= 4 dy values (the lanes of a pixel read the same
// addresses), 16 FMAs into acc[co][ci]; four pixels' loads are issued per trip.  Lanes of one channel quad are folded with
// shuffles, the warps through shared memory, and a CTA ends with Cout x Cin (+ Cout) float atomics — at most 64 CTAs per sample.
struct WArgs {
  MfviView x, dy;
  float* dw;               // [S][Cout][Cin], sample stride w_sstride
  float* dbias;            // [S][Cout] (sample stride w_sstride) or null
  long long w_sstride;
  int Cin, Cout, HW, W, G;
};

__global__ void __launch_bounds__(kThreads)
k_wgrad_pointwise(const WArgs p) {
  __shared__ float part[kThreads / 32][8][20];          // [warp][channel quad][16 dw + 4 dbias]
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.y;
  const int G = p.G, PPB = kThreads / G;                // G = Cin / 4 in {1, 2, 4, 8}: divides the warp
  const int g = threadIdx.x % G, slot = threadIdx.x / G;
  float acc[4][4], bsum[4];
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    bsum[co] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[co][j] = 0.f;
  }
  const float* xb = p.x.ptr + static_cast<size_t>(s) * p.x.sstride + 4 * g;
  const float* db = p.dy.ptr + static_cast<size_t>(s) * p.dy.sstride;
  const int step = gridDim.x * PPB;
  constexpr int U = 4;
  for (int px0 = blockIdx.x * PPB + slot; px0 < p.HW; px0 += U * step) {
    float4 xv[U];
    float dv[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int px = px0 + u * step;
      const bool ok = px < p.HW;
      const int h = ok ? px / p.W : 0, w = ok ? px - h * p.W : 0;
      xv[u] = ok ? __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(h) * p.x.hstride + static_cast<size_t>(w) * p.x.wstride))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
      const float* dp = db + static_cast<size_t>(h) * p.dy.hstride + static_cast<size_t>(w) * p.dy.wstride;
#pragma unroll
      for (int co = 0; co < 4; ++co) dv[u][co] = (ok && co < p.Cout) ? __ldg(dp + co) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        acc[co][0] = fmaf(dv[u][co], xv[u].x, acc[co][0]);
        acc[co][1] = fmaf(dv[u][co], xv[u].y, acc[co][1]);
        acc[co][2] = fmaf(dv[u][co], xv[u].z, acc[co][2]);
        acc[co][3] = fmaf(dv[u][co], xv[u].w, acc[co][3]);
        bsum[co] += dv[u][co];
      }
  }
  // lanes l, l + G, l + 2G, ... of a warp hold the same channel quad
  for (int off = 16; off >= G; off >>= 1) {
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      bsum[co] += __shfl_down_sync(0xffffffffu, bsum[co], off);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[co][j] += __shfl_down_sync(0xffffffffu, acc[co][j], off);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < G) {
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      part[warp][lane][16 + co] = bsum[co];
#pragma unroll
      for (int j = 0; j < 4; ++j) part[warp][lane][4 * co + j] = acc[co][j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * 20; i += kThreads) {
    const int gg = i / 20, v = i - gg * 20;
    float t = 0.f;
#pragma unroll
    for (int wi = 0; wi < kThreads / 32; ++wi) t += part[wi][gg][v];
    if (v < 16) {
      const int co = v >> 2, j = v & 3;
      if (co < p.Cout) atomicAdd(p.dw + static_cast<size_t>(s) * p.w_sstride + static_cast<size_t>(co) * p.Cin + 4 * gg + j, t);
    } else if (gg == 0 && p.dbias != nullptr && v - 16 < p.Cout) {
      atomicAdd(p.dbias + static_cast<size_t>(s) * p.w_sstride + (v - 16), t);
    }
  }
}

static inline int pad_quads(int c) { return c <= 4 ? 1 : (c <= 16 ? 4 : 8); }

static bool ok_view(const MfviView& v, int C) {
  // float4 access needs 16-byte aligned pixels; otherwise the kernel's scalar path is taken per channel count
  const bool al = (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && v.wstride % 4 == 0 && v.hstride % 4 == 0 && v.sstride % 4 == 0;
  return C % 4 != 0 || al;
}

static int launch(const MfviConvDesc* d, bool dgrad, MfviView in, MfviView out, const float* w, const float* bias,
                  long long w_sstride, double* stats, int accumulate, mfvi_stream_t st, const char* what) {
  Args p{};
  p.in = in; p.out = out; p.w = w; p.bias = bias; p.w_sstride = (d->S == 1) ? 0 : w_sstride; p.stats = stats;
  p.K = dgrad ? d->Cout : d->Cin;
  p.N = dgrad ? d->Cin : d->Cout;
  p.HW = d->Hout * d->Wout; p.W = d->Wout; p.dgrad = dgrad ? 1 : 0; p.accumulate = accumulate;
  p.Cout = d->Cout; p.Cin = d->Cin;
  const int KP = pad_quads(p.K), NP = pad_quads(p.N);
  int chunks = (p.HW + kThreads - 1) / kThreads;
  const int cap = (kNumSMs * 8 + d->S - 1) / d->S;
  if (chunks > cap) chunks = cap;
  dim3 grid(chunks, d->S);
  dry_detail("KP=%d NP=%d", KP, NP);
#define PW_CASE(KQ, NQ)                                                                          \
  if (KP == KQ && NP == NQ) {                                                                    \
    if (stats != nullptr) launch_k(k_conv_pointwise<KQ, NQ, true>, grid, kThreads, 0, as_stream(st), p);  \
    else launch_k(k_conv_pointwise<KQ, NQ, false>, grid, kThreads, 0, as_stream(st), p);          \
  }
  PW_CASE(1, 1); PW_CASE(1, 4); PW_CASE(1, 8);       // dgrad: K = Cout <= 4
  PW_CASE(4, 1); PW_CASE(8, 1);                      // forward: N = Cout <= 4
#undef PW_CASE
  return check_launch(what);
}

static bool enabled() {
  static const bool on = [] { const char* e = getenv("MFVI_POINTWISE"); return e == nullptr || e[0] != '0'; }();
  return on;
}

static bool shape_ok(const MfviConvDesc* d) {
  return d->KH == 1 && d->KW == 1 && d->stride == 1 && d->Cin <= 32 && d->Cout <= 4 && d->Hin == d->Hout && d->Win == d->Wout;
}

}  // namespace pw
}  // namespace mfvi

using namespace mfvi;

extern "C" {

// Return 0 on success, -1 when the shape is not taken (the caller goes on to the tensor-core / generic kernels), > 0 on error.
int mfvi_conv2d_fwd_pw(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                       double* stats, mfvi_stream_t st) {
  if (!pw::enabled() || !pw::shape_ok(d) || !pw::ok_view(x, d->Cin) || !pw::ok_view(y, d->Cout)) return -1;
  return pw::launch(d, false, x, y, w, bias, w_sstride, stats, 0, st, "conv2d_fwd_pw");
}

int mfvi_conv2d_dgrad_pw(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                         mfvi_stream_t st) {
  if (!pw::enabled() || !pw::shape_ok(d) || !pw::ok_view(dy, d->Cout) || !pw::ok_view(dx, d->Cin)) return -1;
  return pw::launch(d, true, dy, dx, w, nullptr, w_sstride, nullptr, accumulate, st, "conv2d_dgrad_pw");
}

// dw (+)= x^T dy, dbias (+)= sum dy in ONE launch (both accumulate: zero them first, as for every weight-gradient kernel)
int mfvi_conv2d_wgrad_pw(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                         mfvi_stream_t st) {
  const int G = d->Cin / 4;
  if (!pw::enabled() || !pw::shape_ok(d) || d->Cin % 4 != 0 || !(G == 1 || G == 2 || G == 4 || G == 8) || !pw::ok_view(x, d->Cin) ||
      (dy.sstride == 0 && d->S > 1))
    return -1;
  pw::WArgs p{};
  p.x = x; p.dy = dy; p.dw = dw; p.dbias = dbias; p.w_sstride = w_sstride;
  p.Cin = d->Cin; p.Cout = d->Cout; p.HW = d->Hout * d->Wout; p.W = d->Wout; p.G = G;
  const int PPB = pw::kThreads / G;
  int chunks = (p.HW + 4 * PPB - 1) / (4 * PPB);
  chunks = std::max(1, std::min(chunks, std::min(64, (2 * kNumSMs + d->S - 1) / d->S)));
  dry_detail("G=%d", G);
  launch_k(pw::k_wgrad_pointwise, dim3(chunks, d->S), pw::kThreads, 0, as_stream(st), p);
  return check_launch("conv2d_wgrad_pw");
}

}  // extern "C"
