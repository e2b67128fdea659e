// BatchNorm + LeakyReLU + reflection-pad BACKWARD without the intermediate gradient buffer (EXPERIMENTAL, off by default:
// MFVI_FUSED_BN_BWD=1 in the engine).
//
// The standard pair is   k_pad_act_bwd  : g  = fold_reflect(dxp) * act'(bn(y)),  red += (sum g, sum g*xhat)     [writes g]
//                        k_bn_bwd_apply : dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))                  [reads g]
// and it is 35 % of the step's main chain.  The reduction in the middle forces two passes, but not the round trip of g through
// HBM: pass 1 here only reduces, pass 2 recomputes g from dxp and y (the same loads pass 1 made, the same arithmetic in the same
// order as the standard pair) and applies the BatchNorm backward.  Traffic per element drops
// from 6 tensor passes to 5 (fp32) and the g buffers disappear.
#include "elementwise.cuh"

namespace mfvi {

// g of one pixel: the gradient of the padded tensor folded back onto its source pixel, times the LeakyReLU derivative.
// Identical to the body of k_pad_act_bwd (elementwise.cu).
template <int V>
__device__ __forceinline__ void folded_grad(const float* __restrict__ dbase, const MfviView& dxp, int h, int w, int H, int W, int pad,
                                            const Vec<V>& yy, const BnRegs<V>& bn, int act, Vec<V>& a, float (&xhat)[V]) {
  a.load(dbase + (size_t)(h + pad) * dxp.hstride + (size_t)(w + pad) * dxp.wstride);
  const bool edge = pad > 0 && (h <= pad || w <= pad || h >= H - 1 - pad || w >= W - 1 - pad);
  if (edge) {          // reflected border positions fold back onto this pixel
    int qh[3], qw[3];
    const int nh = fold_sources(h, H, pad, qh), nw = fold_sources(w, W, pad, qw);
    for (int ih = 0; ih < nh; ++ih)
      for (int iw = 0; iw < nw; ++iw) {
        if (ih == 0 && iw == 0) continue;
        Vec<V> t;
        t.load(dbase + (size_t)qh[ih] * dxp.hstride + (size_t)qw[iw] * dxp.wstride);
#pragma unroll
        for (int j = 0; j < V; ++j) a.v[j] += t.v[j];
      }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const float z = fmaf(yy.v[j], bn.sc[j], bn.sh[j]);
    if (act && z <= 0.f) a.v[j] *= kLreluSlope;
    xhat[j] = (yy.v[j] - bn.mean[j]) * bn.invstd[j];
  }
}

// pass 1: red[S][C][2] += (sum g, sum g*xhat); nothing is stored.            grid = (chunks, S)
template <int V>
__global__ void __launch_bounds__(kEwThreads, 4)
k_pad_act_reduce(MfviView dxp, int H, int W, int C, int pad, MfviView y, const double* __restrict__ sums,
                 const float* __restrict__ gamma, const float* __restrict__ beta, int act, double* __restrict__ red, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  const int s = blockIdx.y;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  Acc2<V> acc(sm_red);
  __shared__ BnTable tab;
  tab.fill(sums, gamma, beta, s, C, 1.0 / ((double)H * W));
  __syncthreads();
  if (active) {
    BnRegs<V> bn;
    bn.load(tab, c0, C);
    const float* dbase = dxp.ptr + (size_t)s * dxp.sstride + c0;
    const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
    for (PixIter it(H * W, W, PPB, slot); it.valid(); it.next()) {
      Vec<V> a, yy;
      float xhat[V];
      yy.load(ybase + (size_t)it.h * y.hstride + (size_t)it.w * y.wstride);
      folded_grad<V>(dbase, dxp, it.h, it.w, H, W, pad, yy, bn, act, a, xhat);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        acc.fa[j] += a.v[j];
        acc.fb[j] = fmaf(a.v[j], xhat[j], acc.fb[j]);
      }
      acc.tick();
    }
    acc.flush();
  }
  cta_reduce_cells<V>(sm_red, G, PPB, C, red + (size_t)s * C * 2);
}

// pass 2: dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)) with g recomputed; block (0,0) also writes dgamma / dbeta.
template <int V, bool OBF>       // OBF: dy is a bf16 view (strides in bf16 elements)
__global__ void __launch_bounds__(kEwThreads, 4)
k_bn_bwd_from_dxp(MfviView dxp, MfviView y, int S, int H, int W, int C, int pad, const double* __restrict__ sums,
                  const double* __restrict__ red, const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                  MfviView dy, float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.y;
  const double inv_count = 1.0 / ((double)H * W);
  if (blockIdx.x == 0 && blockIdx.y == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double dg = 0.0, db = 0.0;
      for (int ss = 0; ss < S; ++ss) {
        db += red[((size_t)ss * C + c) * 2 + 0];
        dg += red[((size_t)ss * C + c) * 2 + 1];
      }
      dgamma[c] = (float)dg;
      dbeta[c] = (float)db;
    }
  }
  __shared__ BnTable tab;
  __shared__ float sm_m1[kMaxC], sm_m2[kMaxC];
  tab.fill(sums, gamma, beta, s, C, inv_count);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sm_m1[c] = (float)(red[((size_t)s * C + c) * 2 + 0] * inv_count);
    sm_m2[c] = (float)(red[((size_t)s * C + c) * 2 + 1] * inv_count);
  }
  __syncthreads();
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  if (slot >= PPB) return;
  const int c0 = group * V;
  BnRegs<V> bn;
  bn.load(tab, c0, C);
  float m1[V], m2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const bool ok = c0 + j < C;
    m1[j] = ok ? sm_m1[c0 + j] : 0.f;
    m2[j] = ok ? sm_m2[c0 + j] : 0.f;
  }
  const float* dbase = dxp.ptr + (size_t)s * dxp.sstride + c0;
  const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
  float* obase = dy.ptr + (size_t)s * dy.sstride + c0;
  __nv_bfloat16* obase16 = reinterpret_cast<__nv_bfloat16*>(dy.ptr) + (size_t)s * dy.sstride + c0;
  for (PixIter it(H * W, W, PPB, slot); it.valid(); it.next()) {
    Vec<V> a, yy;
    float xhat[V];
    yy.load(ybase + (size_t)it.h * y.hstride + (size_t)it.w * y.wstride);
    folded_grad<V>(dbase, dxp, it.h, it.w, H, W, pad, yy, bn, act, a, xhat);
    // k = gamma*invstd is BnRegs::sc (BnTable::fill stores gamma * invstd there)
#pragma unroll
    for (int j = 0; j < V; ++j) a.v[j] = bn.sc[j] * (a.v[j] - m1[j] - xhat[j] * m2[j]);
    if (!OBF) a.store(obase + (size_t)it.h * dy.hstride + (size_t)it.w * dy.wstride);
    else store_bf16<V>(obase16 + (size_t)it.h * dy.hstride + (size_t)it.w * dy.wstride, a.v);
  }
}

}  // namespace mfvi

using namespace mfvi;

#define MFVI_EWF_DISPATCH(GEOM, KERNEL, GRID, ...)                                                 \
  do {                                                                                             \
    if ((GEOM).V == 4)                                                                             \
      launch_k(KERNEL<4>, GRID, kEwThreads, 0, as_stream(st), __VA_ARGS__, (GEOM).G, (GEOM).PPB);  \
    else                                                                                           \
      launch_k(KERNEL<1>, GRID, kEwThreads, 0, as_stream(st), __VA_ARGS__, (GEOM).G, (GEOM).PPB);  \
  } while (0)

extern "C" {

int mfvi_pad_act_bwd_reduce(MfviView dxp, int S, int H, int W, int C, int pad, MfviView y, const double* sums, const float* gamma,
                            const float* beta, int act, double* red, mfvi_stream_t st) {
  MFVI_REQUIRE(dxp.ptr && y.ptr && red && sums, "pad_act_bwd_reduce: null pointer");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "pad_act_bwd_reduce: C out of range");
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "pad_act_bwd_reduce: pad must be smaller than the image");
  const EwGeom ge = ew_geom(C, view_vec_ok(dxp) && view_vec_ok(y));
  MFVI_REQUIRE(ge.G <= kEwThreads, "pad_act_bwd_reduce: too many channel groups");
  dim3 grid(ew_grid(H * W, ge.PPB, S), S);
  MFVI_EWF_DISPATCH(ge, k_pad_act_reduce, grid, dxp, H, W, C, pad, y, sums, gamma, beta, act, red);
  return check_launch("pad_act_bwd_reduce");
}

static int bn_bwd_from_dxp(MfviView dxp, MfviView y, int S, int H, int W, int C, int pad, const double* sums, const double* red,
                           const float* gamma, const float* beta, int act, MfviView dy, float* dgamma, float* dbeta, bool bf16_out,
                           mfvi_stream_t st, const char* what) {
  MFVI_REQUIRE(dxp.ptr && y.ptr && dy.ptr && sums && red, "%s: null pointer", what);
  MFVI_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "%s: dgamma/dbeta must both be set or NULL", what);
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "%s: C out of range", what);
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "%s: pad must be smaller than the image", what);
  const EwGeom ge = ew_geom(C, view_vec_ok(dxp) && view_vec_ok(y) && view_vec_ok(dy));
  MFVI_REQUIRE(ge.G <= kEwThreads, "%s: too many channel groups", what);
  dim3 grid(ew_grid(H * W, ge.PPB, S), S);
  if (bf16_out) {
    if (ge.V == 4) launch_k(k_bn_bwd_from_dxp<4, true>, grid, kEwThreads, 0, as_stream(st), dxp, y, S, H, W, C, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta, ge.G, ge.PPB);
    else launch_k(k_bn_bwd_from_dxp<1, true>, grid, kEwThreads, 0, as_stream(st), dxp, y, S, H, W, C, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta, ge.G, ge.PPB);
  } else {
    if (ge.V == 4) launch_k(k_bn_bwd_from_dxp<4, false>, grid, kEwThreads, 0, as_stream(st), dxp, y, S, H, W, C, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta, ge.G, ge.PPB);
    else launch_k(k_bn_bwd_from_dxp<1, false>, grid, kEwThreads, 0, as_stream(st), dxp, y, S, H, W, C, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta, ge.G, ge.PPB);
  }
  return check_launch(what);
}

int mfvi_bn_bwd_apply_from_dxp(MfviView dxp, MfviView y, int S, int H, int W, int C, int pad, const double* sums, const double* red,
                               const float* gamma, const float* beta, int act, MfviView dy, float* dgamma, float* dbeta,
                               mfvi_stream_t st) {
  return bn_bwd_from_dxp(dxp, y, S, H, W, C, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta, false, st, "bn_bwd_apply_from_dxp");
}

int mfvi_bn_bwd_apply_from_dxp_bf16(MfviView dxp, MfviView y, int S, int H, int W, int C, int pad, const double* sums,
                                    const double* red, const float* gamma, const float* beta, int act, MfviView dy, float* dgamma,
                                    float* dbeta, mfvi_stream_t st) {
  return bn_bwd_from_dxp(dxp, y, S, H, W, C, pad, sums, red, gamma, beta, act, dy, dgamma, dbeta, true, st,
                         "bn_bwd_apply_from_dxp_bf16");
}

}  // extern "C"
