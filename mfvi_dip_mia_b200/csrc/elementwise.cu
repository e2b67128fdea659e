// Skip-net elementwise path (models/common.py:77-135, models/skip.py:68,102 of the reference):
// BatchNorm (train mode, per-sample statistics) + LeakyReLU(0.2) + ReflectionPad2d + x2 upsample + concat,
// forward and backward.  NHWC fp32, HBM-bound: each thread owns one channel quad of one pixel (float4 when
// C % 4 == 0), threads of a CTA are laid out [pixel-slot][channel-group] so that a warp touches contiguous
// memory; per-(sample,channel) reductions go registers -> shared -> one double atomic per channel per CTA.
#include "elementwise.cuh"

namespace mfvi {

// F1: xp = reflect_pad(act(bn(y)))            grid = (chunks, S)
template <int V, bool OBF = false>       // OBF: xp is a bf16 view (strides in bf16 elements)
__global__ void __launch_bounds__(kEwThreads)
k_bn_act_pad_fwd(MfviView y, int H, int W, int C, const double* __restrict__ sums, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int act, int pad, MfviView xp, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ BnTable tab;
  const int s = blockIdx.y;
  tab.fill(sums, gamma, beta, s, C, 1.0 / ((double)H * W));
  __syncthreads();
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  if (slot >= PPB) return;
  const int c0 = group * V;
  BnRegs<V> bn;
  bn.load(tab, c0, C);
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
  float* xbase = xp.ptr + (size_t)s * xp.sstride + c0;
  __nv_bfloat16* xbase16 = reinterpret_cast<__nv_bfloat16*>(xp.ptr) + (size_t)s * xp.sstride + c0;
  for (PixIter it(Hp * Wp, Wp, PPB, slot); it.valid(); it.next()) {
    const int h = reflect_idx(it.h - pad, H), w = reflect_idx(it.w - pad, W);
    Vec<V> t;
    t.load(ybase + (size_t)h * y.hstride + (size_t)w * y.wstride);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float z = fmaf(t.v[j], bn.sc[j], bn.sh[j]);
      if (act) z = z > 0.f ? z : kLreluSlope * z;
      t.v[j] = z;
    }
    if (!OBF) t.store(xbase + (size_t)it.h * xp.hstride + (size_t)it.w * xp.wstride);
    else store_bf16<V>(xbase16 + (size_t)it.h * xp.hstride + (size_t)it.w * xp.wstride, t.v);
  }
}

// x2 upsample source taps along one dimension (align_corners=False).  i: hi-res index, n: low-res size.
__device__ __forceinline__ void up_taps(int i, int n, int mode, int& i0, int& i1, float& w0, float& w1) {
  if (mode == 1) {  // nearest
    i0 = i >> 1; i1 = i0; w0 = 1.f; w1 = 0.f;
    return;
  }
  float src = 0.5f * (float)i - 0.25f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  const float f = src - (float)i0;
  i1 = i0 + 1 < n ? i0 + 1 : n - 1;
  w0 = 1.f - f;
  w1 = f;
}

// F2: A = cat(lrelu(bn(ys)), up2x(lrelu(bn(yd)))), sumsA += (sum, sumsq)      grid = (chunks, S)
// The x2 upsample is evaluated per 2x2 output QUAD: quad (a, b), a in [0, H/2], b in [0, W/2], covers output rows
// {2a-1, 2a} and columns {2b-1, 2b} (those inside the image), which all read the same four low-resolution pixels
// rows {max(a-1,0), min(a,h2-1)} x columns {max(b-1,0), min(b,w2-1)} — 4 loads and 4 BatchNorm+LeakyReLU evaluations
// for 4 outputs instead of 16.  Weights per output row (align_corners=False, as up_taps): odd row 2a-1 -> (0.75, 0.25),
// even row 2a -> (0.25, 0.75), row 0 -> (1, 0); nearest: (1, 0) / (0, 1).  Same arithmetic per output as the per-pixel form.
__device__ __forceinline__ void quad_weights(int a, int mode, float (&wodd)[2], float (&wevn)[2]) {
  if (mode == 1) {
    wodd[0] = 1.f; wodd[1] = 0.f; wevn[0] = 0.f; wevn[1] = 1.f;
  } else {
    wodd[0] = 0.75f; wodd[1] = 0.25f;
    wevn[0] = a == 0 ? 1.f : 0.25f;
    wevn[1] = a == 0 ? 0.f : 0.75f;
  }
}

template <int V>
__global__ void __launch_bounds__(kEwThreads, 3)
k_cat_up_fwd(MfviView ys, int Cs, const double* __restrict__ sums_s, const float* __restrict__ gamma_s,
             const float* __restrict__ beta_s, MfviView yd, int Cd, const double* __restrict__ sums_d,
             const float* __restrict__ gamma_d, const float* __restrict__ beta_d, int H, int W, int mode, MfviView A,
             double* __restrict__ sumsA, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  __shared__ BnTable tab;
  const int s = blockIdx.y;
  const int C = Cs + Cd;
  const int h2 = H / 2, w2 = W / 2;
  if (Cs > 0) tab.fill(sums_s, gamma_s, beta_s, s, Cs, 1.0 / ((double)H * W));
  tab.fill(sums_d, gamma_d, beta_d, s, Cd, 1.0 / ((double)h2 * w2), Cs);
  __syncthreads();
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  const bool skip = c0 < Cs;          // Cs % V == 0 is guaranteed by the host (V falls back to 1 otherwise)
  Acc2<V> acc(sm_red);
  if (active) {
    BnRegs<V> bn;
    bn.load(tab, c0, C);
    const float* sbase = skip ? ys.ptr + (size_t)s * ys.sstride + c0 : nullptr;
    const float* dbase = yd.ptr + (size_t)s * yd.sstride + (c0 - Cs);
    float* abase = A.ptr + (size_t)s * A.sstride + c0;
    for (PixIter it((h2 + 1) * (w2 + 1), w2 + 1, PPB, slot); it.valid(); it.next()) {
      const int a = it.h, b = it.w;
      const int rows[2] = {2 * a - 1, 2 * a}, cols[2] = {2 * b - 1, 2 * b};
      const bool rok[2] = {a >= 1, a < h2}, cok[2] = {b >= 1, b < w2};
      Vec<V> z[2][2];
      float wr[2][2], wc[2][2];     // [output row/col parity: 0 = odd (2a-1), 1 = even (2a)][tap]
      if (!skip) {
        const int r0 = a >= 1 ? a - 1 : 0, r1 = a < h2 ? a : h2 - 1;
        const int q0 = b >= 1 ? b - 1 : 0, q1 = b < w2 ? b : w2 - 1;
        z[0][0].load(dbase + (size_t)r0 * yd.hstride + (size_t)q0 * yd.wstride);
        z[0][1].load(dbase + (size_t)r0 * yd.hstride + (size_t)q1 * yd.wstride);
        z[1][0].load(dbase + (size_t)r1 * yd.hstride + (size_t)q0 * yd.wstride);
        z[1][1].load(dbase + (size_t)r1 * yd.hstride + (size_t)q1 * yd.wstride);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const float t = fmaf(z[u][v].v[j], bn.sc[j], bn.sh[j]);
              z[u][v].v[j] = t > 0.f ? t : kLreluSlope * t;
            }
        quad_weights(a, mode, wr[0], wr[1]);
        quad_weights(b, mode, wc[0], wc[1]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!rok[u]) continue;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (!cok[v]) continue;
          const int h = rows[u], w = cols[v];
          Vec<V> o;
          if (skip) {
            o.load(sbase + (size_t)h * ys.hstride + (size_t)w * ys.wstride);
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const float t = fmaf(o.v[j], bn.sc[j], bn.sh[j]);
              o.v[j] = t > 0.f ? t : kLreluSlope * t;
            }
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j)
              o.v[j] = wr[u][0] * (wc[v][0] * z[0][0].v[j] + wc[v][1] * z[0][1].v[j]) +
                       wr[u][1] * (wc[v][0] * z[1][0].v[j] + wc[v][1] * z[1][1].v[j]);
          }
          o.store(abase + (size_t)h * A.hstride + (size_t)w * A.wstride);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            acc.fa[j] += o.v[j];
            acc.fb[j] = fmaf(o.v[j], o.v[j], acc.fb[j]);
          }
        }
      }
      acc.tick();
    }
    acc.flush();
  }
  cta_reduce_cells<V>(sm_red, G, PPB, C, sumsA + (size_t)s * C * 2);
}


// B1: g = fold_reflect(dxp) * act'(bn(y)), red += (sum g, sum g*xhat)      grid = (chunks, S)
template <int V>
__global__ void __launch_bounds__(kEwThreads, 4)
k_pad_act_bwd(MfviView dxp, int H, int W, int C, int pad, MfviView y, const double* __restrict__ sums,
              const float* __restrict__ gamma, const float* __restrict__ beta, int act, MfviView g,
              double* __restrict__ red, int G, int PPB) {
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
    float* gbase = g.ptr + (size_t)s * g.sstride + c0;
    for (PixIter it(H * W, W, PPB, slot); it.valid(); it.next()) {
      const int h = it.h, w = it.w;
      Vec<V> a;
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
      Vec<V> yy;
      yy.load(ybase + (size_t)h * y.hstride + (size_t)w * y.wstride);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float z = fmaf(yy.v[j], bn.sc[j], bn.sh[j]);
        float gg = a.v[j];
        if (act && z <= 0.f) gg *= kLreluSlope;
        const float xhat = (yy.v[j] - bn.mean[j]) * bn.invstd[j];
        a.v[j] = gg;
        acc.fa[j] += gg;
        acc.fb[j] = fmaf(gg, xhat, acc.fb[j]);
      }
      a.store(gbase + (size_t)h * g.hstride + (size_t)w * g.wstride);
      acc.tick();
    }
    acc.flush();
  }
  cta_reduce_cells<V>(sm_red, G, PPB, C, red + (size_t)s * C * 2);
}

// B2: dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); block (0,0) also writes dgamma/dbeta.
template <int V, bool OBF = false>       // OBF: dy is a bf16 view (strides in bf16 elements)
__global__ void __launch_bounds__(kEwThreads)
k_bn_bwd_apply(MfviView g, MfviView y, int S, int H, int W, int C, const double* __restrict__ sums,
               const double* __restrict__ red, const float* __restrict__ gamma, MfviView dy,
               float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int PPB) {
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
  __shared__ float sm_k[kMaxC], sm_mean[kMaxC], sm_invstd[kMaxC], sm_m1[kMaxC], sm_m2[kMaxC];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean_c, invstd_c;
    bn_mean_invstd(sums + ((size_t)s * C + c) * 2, inv_count, mean_c, invstd_c);
    sm_mean[c] = mean_c;
    sm_invstd[c] = invstd_c;
    sm_k[c] = (gamma != nullptr ? gamma[c] : 1.f) * invstd_c;
    sm_m1[c] = (float)(red[((size_t)s * C + c) * 2 + 0] * inv_count);
    sm_m2[c] = (float)(red[((size_t)s * C + c) * 2 + 1] * inv_count);
  }
  __syncthreads();
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  if (slot >= PPB) return;
  const int c0 = group * V;
  float k[V], mean[V], invstd[V], m1[V], m2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const bool ok = c0 + j < C;
    mean[j] = ok ? sm_mean[c0 + j] : 0.f;
    invstd[j] = ok ? sm_invstd[c0 + j] : 1.f;
    k[j] = ok ? sm_k[c0 + j] : 0.f;
    m1[j] = ok ? sm_m1[c0 + j] : 0.f;
    m2[j] = ok ? sm_m2[c0 + j] : 0.f;
  }
  const float* gbase = g.ptr + (size_t)s * g.sstride + c0;
  const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
  float* obase = dy.ptr + (size_t)s * dy.sstride + c0;
  __nv_bfloat16* obase16 = reinterpret_cast<__nv_bfloat16*>(dy.ptr) + (size_t)s * dy.sstride + c0;
  for (PixIter it(H * W, W, PPB, slot); it.valid(); it.next()) {
    const int h = it.h, w = it.w;
    Vec<V> gg, yy;
    gg.load(gbase + (size_t)h * g.hstride + (size_t)w * g.wstride);
    yy.load(ybase + (size_t)h * y.hstride + (size_t)w * y.wstride);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xhat = (yy.v[j] - mean[j]) * invstd[j];
      gg.v[j] = k[j] * (gg.v[j] - m1[j] - xhat * m2[j]);
    }
    if (!OBF) gg.store(obase + (size_t)h * dy.hstride + (size_t)w * dy.wstride);
    else store_bf16<V>(obase16 + (size_t)h * dy.hstride + (size_t)w * dy.wstride, gg.v);
  }
}

// B3a: skip branch of the concat: gs = dA[:, :Cs] * lrelu'(bn(ys)), red_s += …
template <int V>
__global__ void __launch_bounds__(kEwThreads, 4)
k_cat_bwd_skip(MfviView dA, int H, int W, MfviView ys, int Cs, const double* __restrict__ sums_s,
               const float* __restrict__ gamma_s, const float* __restrict__ beta_s, MfviView gs,
               double* __restrict__ red_s, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  const int s = blockIdx.y;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  Acc2<V> acc(sm_red);
  __shared__ BnTable tab;
  tab.fill(sums_s, gamma_s, beta_s, s, Cs, 1.0 / ((double)H * W));
  __syncthreads();
  if (active) {
    BnRegs<V> bn;
    bn.load(tab, c0, Cs);
    const float* dbase = dA.ptr + (size_t)s * dA.sstride + c0;
    const float* ybase = ys.ptr + (size_t)s * ys.sstride + c0;
    float* gbase = gs.ptr + (size_t)s * gs.sstride + c0;
    for (PixIter it(H * W, W, PPB, slot); it.valid(); it.next()) {
      const int h = it.h, w = it.w;
      Vec<V> d, yy;
      d.load(dbase + (size_t)h * dA.hstride + (size_t)w * dA.wstride);
      yy.load(ybase + (size_t)h * ys.hstride + (size_t)w * ys.wstride);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float z = fmaf(yy.v[j], bn.sc[j], bn.sh[j]);
        const float gg = z > 0.f ? d.v[j] : kLreluSlope * d.v[j];
        const float xhat = (yy.v[j] - bn.mean[j]) * bn.invstd[j];
        d.v[j] = gg;
        acc.fa[j] += gg;
        acc.fb[j] = fmaf(gg, xhat, acc.fb[j]);
      }
      d.store(gbase + (size_t)h * gs.hstride + (size_t)w * gs.wstride);
      acc.tick();
    }
    acc.flush();
  }
  cta_reduce_cells<V>(sm_red, G, PPB, Cs, red_s + (size_t)s * Cs * 2);
}

// weight with which low-res index k enters hi-res index i (0 if not a tap)
__device__ __forceinline__ float up_weight_of(int i, int k, int n, int mode) {
  if (i < 0 || i >= 2 * n) return 0.f;
  int i0, i1;
  float w0, w1;
  up_taps(i, n, mode, i0, i1, w0, w1);
  float wgt = 0.f;
  if (i0 == k) wgt += w0;
  if (i1 == k) wgt += w1;
  return wgt;
}

// B3b: deeper branch: gd = up2x^T(dA[:, Cs:]) * lrelu'(bn(yd)), red_d += …   (iterates low-res pixels)
template <int V>
__global__ void __launch_bounds__(kEwThreads, 4)
k_cat_bwd_up(MfviView dA, int H, int W, int mode, int Cs, MfviView yd, int Cd, const double* __restrict__ sums_d,
             const float* __restrict__ gamma_d, const float* __restrict__ beta_d, MfviView gd,
             double* __restrict__ red_d, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  const int s = blockIdx.y;
  const int h2 = H / 2, w2 = W / 2;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  Acc2<V> acc(sm_red);
  __shared__ BnTable tab;
  tab.fill(sums_d, gamma_d, beta_d, s, Cd, 1.0 / ((double)h2 * w2));
  __syncthreads();
  if (active) {
    BnRegs<V> bn;
    bn.load(tab, c0, Cd);
    const float* dbase = dA.ptr + (size_t)s * dA.sstride + Cs + c0;
    const float* ybase = yd.ptr + (size_t)s * yd.sstride + c0;
    float* gbase = gd.ptr + (size_t)s * gd.sstride + c0;
    for (PixIter it(h2 * w2, w2, PPB, slot); it.valid(); it.next()) {
      const int kh = it.h, kw = it.w;
      // hi-res rows 2kh-1 .. 2kh+2 receive low-res row kh with these weights (0 outside the image / when not a tap)
      float wh[4], ww[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        wh[d] = up_weight_of(2 * kh + d - 1, kh, h2, mode);
        ww[d] = up_weight_of(2 * kw + d - 1, kw, w2, mode);
      }
      Vec<V> a;
#pragma unroll
      for (int j = 0; j < V; ++j) a.v[j] = 0.f;
#pragma unroll
      for (int dh = 0; dh < 4; ++dh) {
        if (wh[dh] == 0.f) continue;
        const float* row = dbase + (size_t)(2 * kh + dh - 1) * dA.hstride;
#pragma unroll
        for (int dw = 0; dw < 4; ++dw) {
          if (ww[dw] == 0.f) continue;
          Vec<V> t;
          t.load(row + (size_t)(2 * kw + dw - 1) * dA.wstride);
          const float wgt = wh[dh] * ww[dw];
#pragma unroll
          for (int j = 0; j < V; ++j) a.v[j] = fmaf(wgt, t.v[j], a.v[j]);
        }
      }
      Vec<V> yy;
      yy.load(ybase + (size_t)kh * yd.hstride + (size_t)kw * yd.wstride);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float z = fmaf(yy.v[j], bn.sc[j], bn.sh[j]);
        const float gg = z > 0.f ? a.v[j] : kLreluSlope * a.v[j];
        const float xhat = (yy.v[j] - bn.mean[j]) * bn.invstd[j];
        a.v[j] = gg;
        acc.fa[j] += gg;
        acc.fb[j] = fmaf(gg, xhat, acc.fb[j]);
      }
      a.store(gbase + (size_t)kh * gd.hstride + (size_t)kw * gd.wstride);
      acc.tick();
    }
    acc.flush();
  }
  cta_reduce_cells<V>(sm_red, G, PPB, Cd, red_d + (size_t)s * Cd * 2);
}

// running stats of all BatchNorms (one thread per channel)
__global__ void k_bn_running(const double* __restrict__ arena, const int* __restrict__ ch_off,
                             const long long* __restrict__ sums_off, const int* __restrict__ Cs,
                             const int* __restrict__ count, int n_bn, int S, float momentum,
                             float* __restrict__ running_mean, float* __restrict__ running_var) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  if (b >= n_bn) return;
  const int C = Cs[b];
  const double n = (double)count[b];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float rm = running_mean[ch_off[b] + c], rv = running_var[ch_off[b] + c];
    for (int s = 0; s < S; ++s) {
      const double* sm = arena + sums_off[b] + ((size_t)s * C + c) * 2;
      const double mean = sm[0] / n;
      double var = sm[1] / n - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      rm = (1.f - momentum) * rm + momentum * (float)mean;
      rv = (1.f - momentum) * rv + momentum * (float)unbiased;
    }
    running_mean[ch_off[b] + c] = rm;
    running_var[ch_off[b] + c] = rv;
  }
}


}  // namespace mfvi

using namespace mfvi;

#define MFVI_EW_DISPATCH(GEOM, KERNEL, GRID, ...)                                              \
  do {                                                                                         \
    if ((GEOM).V == 4)                                                                         \
      launch_k(KERNEL<4>, GRID, kEwThreads, 0, as_stream(st), __VA_ARGS__, (GEOM).G, (GEOM).PPB);  \
    else                                                                                       \
      launch_k(KERNEL<1>, GRID, kEwThreads, 0, as_stream(st), __VA_ARGS__, (GEOM).G, (GEOM).PPB);  \
  } while (0)

extern "C" {

int mfvi_bn_act_pad_fwd(MfviView y, int S, int H, int W, int C, const double* sums, const float* gamma,
                        const float* beta, int act, int pad, MfviView xp, mfvi_stream_t st) {
  MFVI_REQUIRE(y.ptr && xp.ptr, "bn_act_pad_fwd: null pointer");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "bn_act_pad_fwd: C=%d out of range (1..%d)", C, kMaxC);
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "bn_act_pad_fwd: pad must be smaller than the image");
  const EwGeom ge = ew_geom(C, view_vec_ok(y) && view_vec_ok(xp));
  MFVI_REQUIRE(ge.G <= kEwThreads, "bn_act_pad_fwd: too many channel groups");
  dim3 grid(ew_grid((H + 2 * pad) * (W + 2 * pad), ge.PPB, S), S);
  MFVI_EW_DISPATCH(ge, k_bn_act_pad_fwd, grid, y, H, W, C, sums, gamma, beta, act, pad, xp);
  return check_launch("bn_act_pad_fwd");
}

int mfvi_cat_up_fwd(MfviView ys, int Cs, const double* sums_s, const float* gamma_s, const float* beta_s,
                    MfviView yd, int Cd, const double* sums_d, const float* gamma_d, const float* beta_d,
                    int S, int H, int W, int mode, MfviView A, double* sumsA, mfvi_stream_t st) {
  MFVI_REQUIRE(yd.ptr && A.ptr && sumsA, "cat_up_fwd: null pointer");
  MFVI_REQUIRE(Cs == 0 || ys.ptr, "cat_up_fwd: null skip branch");
  MFVI_REQUIRE(H % 2 == 0 && W % 2 == 0, "cat_up_fwd: H,W must be even (centre-crop concat is not supported)");
  MFVI_REQUIRE(Cs + Cd <= kMaxC && Cd >= 1, "cat_up_fwd: channel count out of range");
  MFVI_REQUIRE(mode == 0 || mode == 1, "cat_up_fwd: mode must be 0 (bilinear) or 1 (nearest)");
  const bool al = view_vec_ok(yd) && view_vec_ok(A) && (Cs == 0 || view_vec_ok(ys)) && Cs % 4 == 0 && Cd % 4 == 0;
  const EwGeom ge = ew_geom(Cs + Cd, al);
  MFVI_REQUIRE(ge.G <= kEwThreads, "cat_up_fwd: too many channel groups");
  dim3 grid(ew_grid((H / 2 + 1) * (W / 2 + 1), ge.PPB, S), S);
  MFVI_EW_DISPATCH(ge, k_cat_up_fwd, grid, ys, Cs, sums_s, gamma_s, beta_s, yd, Cd, sums_d, gamma_d, beta_d, H, W, mode,
                   A, sumsA);
  return check_launch("cat_up_fwd");
}

int mfvi_pad_act_bwd(MfviView dxp, int S, int H, int W, int C, int pad, MfviView y, const double* sums,
                     const float* gamma, const float* beta, int act, MfviView g, double* red, mfvi_stream_t st) {
  MFVI_REQUIRE(dxp.ptr && y.ptr && g.ptr && red && sums, "pad_act_bwd: null pointer");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "pad_act_bwd: C out of range");
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "pad_act_bwd: pad must be smaller than the image");
  const EwGeom ge = ew_geom(C, view_vec_ok(dxp) && view_vec_ok(y) && view_vec_ok(g));
  MFVI_REQUIRE(ge.G <= kEwThreads, "pad_act_bwd: too many channel groups");
  dim3 grid(ew_grid(H * W, ge.PPB, S), S);
  MFVI_EW_DISPATCH(ge, k_pad_act_bwd, grid, dxp, H, W, C, pad, y, sums, gamma, beta, act, g, red);
  return check_launch("pad_act_bwd");
}

int mfvi_bn_bwd_apply(MfviView g, MfviView y, int S, int H, int W, int C, const double* sums, const double* red,
                      const float* gamma, MfviView dy, float* dgamma, float* dbeta, mfvi_stream_t st) {
  MFVI_REQUIRE(g.ptr && y.ptr && dy.ptr && sums && red, "bn_bwd_apply: null pointer");
  MFVI_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "bn_bwd_apply: dgamma/dbeta must both be set or NULL");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "bn_bwd_apply: C out of range");
  const EwGeom ge = ew_geom(C, view_vec_ok(g) && view_vec_ok(y) && view_vec_ok(dy));
  MFVI_REQUIRE(ge.G <= kEwThreads, "bn_bwd_apply: too many channel groups");
  dim3 grid(ew_grid(H * W, ge.PPB, S), S);
  MFVI_EW_DISPATCH(ge, k_bn_bwd_apply, grid, g, y, S, H, W, C, sums, red, gamma, dy, dgamma, dbeta);
  return check_launch("bn_bwd_apply");
}

int mfvi_cat_up_bwd(MfviView dA, int S, int H, int W, int mode, MfviView ys, int Cs, const double* sums_s,
                    const float* gamma_s, const float* beta_s, MfviView gs, double* red_s, MfviView yd, int Cd,
                    const double* sums_d, const float* gamma_d, const float* beta_d, MfviView gd, double* red_d, int part,
                    mfvi_stream_t st) {
  MFVI_REQUIRE(part >= 0 && part <= 2, "cat_up_bwd: part must be 0 (both branches), 1 (skip branch) or 2 (upsampled branch)");
  MFVI_REQUIRE(dA.ptr, "cat_up_bwd: null pointer");
  MFVI_REQUIRE(H % 2 == 0 && W % 2 == 0, "cat_up_bwd: H,W must be even");
  MFVI_REQUIRE(Cs + Cd <= kMaxC && Cd >= 1, "cat_up_bwd: channel count out of range");
  if (Cs > 0 && part != 2) {
    MFVI_REQUIRE(ys.ptr && gs.ptr && red_s, "cat_up_bwd: null skip branch");
    const EwGeom ge = ew_geom(Cs, view_vec_ok(dA) && view_vec_ok(ys) && view_vec_ok(gs));
    dim3 grid(ew_grid(H * W, ge.PPB, S), S);
    MFVI_EW_DISPATCH(ge, k_cat_bwd_skip, grid, dA, H, W, ys, Cs, sums_s, gamma_s, beta_s, gs, red_s);
    if (int rc = check_launch("cat_up_bwd(skip)")) return rc;
  }
  if (part == 1) return 0;
  MFVI_REQUIRE(yd.ptr && gd.ptr && red_d, "cat_up_bwd: null upsampled branch");
  const EwGeom ge = ew_geom(Cd, view_vec_ok(dA) && view_vec_ok(yd) && view_vec_ok(gd) && Cs % 4 == 0);
  MFVI_REQUIRE(ge.G <= kEwThreads, "cat_up_bwd: too many channel groups");
  dim3 grid(ew_grid((H / 2) * (W / 2), ge.PPB, S), S);
  MFVI_EW_DISPATCH(ge, k_cat_bwd_up, grid, dA, H, W, mode, Cs, yd, Cd, sums_d, gamma_d, beta_d, gd, red_d);
  return check_launch("cat_up_bwd(up)");
}

int mfvi_bn_running_update(const double* arena, const int* ch_off, const long long* sums_off, const int* C,
                           const int* count, int n_bn, int S, float momentum, float* running_mean,
                           float* running_var, mfvi_stream_t st) {
  MFVI_REQUIRE(arena && ch_off && sums_off && C && count && running_mean && running_var, "bn_running_update: null pointer");
  if (n_bn == 0) return 0;
  launch_k(k_bn_running, n_bn, 128, 0, as_stream(st), arena, ch_off, sums_off, C, count, n_bn, S, momentum, running_mean,
                                                 running_var);
  return check_launch("bn_running_update");
}

}  // extern "C"
