// Bodies of the skip-net elementwise kernels as __device__ functions over a VIRTUAL block index: elementwise.cu wraps each in a
// __global__ kernel (virtual block = real block), mega.cu runs the same bodies as stages of its persistent kernel (a CTA loops
// over virtual blocks).  Arithmetic is therefore identical in both.
#pragma once
#include "elementwise.cuh"

namespace mfvi {


// Every body issues the loads of its first pixels BEFORE it computes the BatchNorm tables (both depend only on the producer
// kernels, not on each other), so the two L2 round trips of a small launch overlap instead of adding up.

// F1: xp = reflect_pad(act(bn(y)))            grid = (chunks, S)
template <int V, bool OBF = false, bool PIPE = false>       // OBF: xp is a bf16 view (strides in bf16 elements); PIPE: cp.async slots
__device__ __forceinline__ void body_bn_act_pad_fwd(const VGrid& vg, EwSmem sm, MfviView y, int H, int W, int C, const double* __restrict__ sums, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int act, int pad, MfviView xp, int G, int PPB) {
  BnTable& tab = *sm.tab;
  const int s = vg.by;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
  float* xbase = xp.ptr + (size_t)s * xp.sstride + c0;
  __nv_bfloat16* xbase16 = reinterpret_cast<__nv_bfloat16*>(xp.ptr) + (size_t)s * xp.sstride + c0;
  PixIter it(vg, active ? Hp * Wp : 0, Wp, PPB, slot);
  BnRegs<V> bn;
  auto emit = [&](int hp, int wp, Vec<V>& t) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float z = fmaf(t.v[j], bn.sc[j], bn.sh[j]);
      if (act) z = z > 0.f ? z : kLreluSlope * z;
      t.v[j] = z;
    }
    if (!OBF) t.store(xbase + (size_t)hp * xp.hstride + (size_t)wp * xp.wstride);
    else store_bf16<V>(xbase16 + (size_t)hp * xp.hstride + (size_t)wp * xp.wstride, t.v);
  };
  if constexpr (PIPE) {
    using Pipe = LdPipe<V, 1>;
    const Pipe pp(sm.pipe);
    PixIter ld = it;
    auto issue = [&](int d) {
      if (ld.valid()) {
        const int h = reflect_idx(ld.h - pad, H), w = reflect_idx(ld.w - pad, W);
        ew_cp_async<V>(pp.slot(d, 0), ybase + (size_t)h * y.hstride + (size_t)w * y.wstride);
        ld.next();
      }
      ew_cp_commit();
    };
#pragma unroll
    for (int d = 0; d < Pipe::D; ++d) issue(d);
    tab.fill(sums, gamma, beta, s, C, 1.0 / ((double)H * W));
    __syncthreads();
    bn.load(tab, c0, C);
    for (int d = 0; it.valid(); it.next()) {
      ew_cp_wait<Pipe::D - 1>();
      Vec<V> t;
      t.load(pp.slot(d, 0));
      emit(it.h, it.w, t);
      issue(d);
      d = d + 1 == Pipe::D ? 0 : d + 1;
    }
    ew_cp_wait<0>();
  } else {
    constexpr int U = 4;
    PixBatch<U> b;
    Vec<V> t[U];
    auto fetch = [&]() {
      if (!b.fill(it)) return false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int h = reflect_idx(b.h[u] - pad, H), w = reflect_idx(b.w[u] - pad, W);
        t[u].load(ybase + (size_t)h * y.hstride + (size_t)w * y.wstride);
      }
      return true;
    };
    bool have = fetch();
    tab.fill(sums, gamma, beta, s, C, 1.0 / ((double)H * W));
    __syncthreads();
    bn.load(tab, c0, C);
    while (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u >= b.n) break;
        emit(b.h[u], b.w[u], t[u]);
      }
      have = fetch();
    }
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
__device__ __forceinline__ void body_cat_up_fwd(const VGrid& vg, EwSmem sm, MfviView ys, int Cs, const double* __restrict__ sums_s, const float* __restrict__ gamma_s,
             const float* __restrict__ beta_s, MfviView yd, int Cd, const double* __restrict__ sums_d,
             const float* __restrict__ gamma_d, const float* __restrict__ beta_d, int H, int W, int mode, MfviView A,
             double* __restrict__ sumsA, int G, int PPB) {
  double* sm_red = sm.red;
  BnTable& tab = *sm.tab;
  const int s = vg.by;
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
    for (PixIter it(vg, (h2 + 1) * (w2 + 1), w2 + 1, PPB, slot); it.valid(); it.next()) {
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
template <int V, bool PIPE = false>     // PIPE: loads through the cp.async slots of sm.pipe (stand-alone kernel only)
__device__ __forceinline__ void body_pad_act_bwd(const VGrid& vg, EwSmem sm, MfviView dxp, int H, int W, int C, int pad, MfviView y, const double* __restrict__ sums,
              const float* __restrict__ gamma, const float* __restrict__ beta, int act, MfviView g,
              double* __restrict__ red, int G, int PPB) {
  double* sm_red = sm.red;
  const int s = vg.by;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  Acc2<V> acc(sm_red);
  BnTable& tab = *sm.tab;
  BnRegs<V> bn;
  const float* dbase = dxp.ptr + (size_t)s * dxp.sstride + c0;
  const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
  float* gbase = g.ptr + (size_t)s * g.sstride + c0;
  auto finish = [&](int h, int w, Vec<V>& a, const Vec<V>& yy) {
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
      float gg = a.v[j];
      if (act && z <= 0.f) gg *= kLreluSlope;
      const float xhat = (yy.v[j] - bn.mean[j]) * bn.invstd[j];
      a.v[j] = gg;
      acc.fa[j] += gg;
      acc.fb[j] = fmaf(gg, xhat, acc.fb[j]);
    }
    a.store(gbase + (size_t)h * g.hstride + (size_t)w * g.wstride);
    acc.tick();
  };
  PixIter it(vg, active ? H * W : 0, W, PPB, slot);
  if constexpr (PIPE) {
    using Pipe = LdPipe<V, 2>;
    const Pipe pp(sm.pipe);
    PixIter ld = it;
    auto issue = [&](int d) {
      if (ld.valid()) {
        ew_cp_async<V>(pp.slot(d, 0), dbase + (size_t)(ld.h + pad) * dxp.hstride + (size_t)(ld.w + pad) * dxp.wstride);
        ew_cp_async<V>(pp.slot(d, 1), ybase + (size_t)ld.h * y.hstride + (size_t)ld.w * y.wstride);
        ld.next();
      }
      ew_cp_commit();
    };
#pragma unroll
    for (int d = 0; d < Pipe::D; ++d) issue(d);
    tab.fill(sums, gamma, beta, s, C, 1.0 / ((double)H * W));
    __syncthreads();
    bn.load(tab, c0, C);
    for (int d = 0; it.valid(); it.next()) {
      ew_cp_wait<Pipe::D - 1>();
      Vec<V> a, yy;
      a.load(pp.slot(d, 0));
      yy.load(pp.slot(d, 1));
      finish(it.h, it.w, a, yy);
      issue(d);                      // after the slot's values were consumed
      d = d + 1 == Pipe::D ? 0 : d + 1;
    }
    ew_cp_wait<0>();
  } else {
    constexpr int U = 2;
    PixBatch<U> b;
    Vec<V> a[U], yy[U];
    auto fetch = [&]() {
      if (!b.fill(it)) return false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u].load(dbase + (size_t)(b.h[u] + pad) * dxp.hstride + (size_t)(b.w[u] + pad) * dxp.wstride);
        yy[u].load(ybase + (size_t)b.h[u] * y.hstride + (size_t)b.w[u] * y.wstride);
      }
      return true;
    };
    bool have = fetch();
    tab.fill(sums, gamma, beta, s, C, 1.0 / ((double)H * W));
    __syncthreads();
    bn.load(tab, c0, C);
    while (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u >= b.n) break;
        finish(b.h[u], b.w[u], a[u], yy[u]);
      }
      have = fetch();
    }
  }
  acc.flush();
  cta_reduce_cells<V>(sm_red, G, PPB, C, red + (size_t)s * C * 2);
}

// B2: dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)); block (0,0) also writes dgamma/dbeta.
template <int V, bool OBF = false, bool PIPE = false>       // OBF: dy is a bf16 view (strides in bf16 elements); PIPE: cp.async slots
__device__ __forceinline__ void body_bn_bwd_apply(const VGrid& vg, EwSmem sm, MfviView g, MfviView y, int S, int H, int W, int C, const double* __restrict__ sums,
               const double* __restrict__ red, const float* __restrict__ gamma, MfviView dy,
               float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int PPB) {
  const int s = vg.by;
  const double inv_count = 1.0 / ((double)H * W);
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  const float* gbase = g.ptr + (size_t)s * g.sstride + c0;
  const float* ybase = y.ptr + (size_t)s * y.sstride + c0;
  float* obase = dy.ptr + (size_t)s * dy.sstride + c0;
  __nv_bfloat16* obase16 = reinterpret_cast<__nv_bfloat16*>(dy.ptr) + (size_t)s * dy.sstride + c0;
  float *sm_k = sm.misc, *sm_mean = sm.misc + kMaxC, *sm_invstd = sm.misc + 2 * kMaxC, *sm_m1 = sm.misc + 3 * kMaxC,
        *sm_m2 = sm.misc + 4 * kMaxC;
  float k[V], mean[V], invstd[V], m1[V], m2[V];
  auto tables = [&]() {            // per-channel constants of this sample: shared memory, then the V channels of this thread
    if (vg.bx == 0 && vg.by == 0 && dgamma != nullptr) {
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
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const bool ok = c0 + j < C;
      mean[j] = ok ? sm_mean[c0 + j] : 0.f;
      invstd[j] = ok ? sm_invstd[c0 + j] : 1.f;
      k[j] = ok ? sm_k[c0 + j] : 0.f;
      m1[j] = ok ? sm_m1[c0 + j] : 0.f;
      m2[j] = ok ? sm_m2[c0 + j] : 0.f;
    }
  };
  auto emit = [&](int h, int w, Vec<V>& gg, const Vec<V>& yy) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float xhat = (yy.v[j] - mean[j]) * invstd[j];
      gg.v[j] = k[j] * (gg.v[j] - m1[j] - xhat * m2[j]);
    }
    if (!OBF) gg.store(obase + (size_t)h * dy.hstride + (size_t)w * dy.wstride);
    else store_bf16<V>(obase16 + (size_t)h * dy.hstride + (size_t)w * dy.wstride, gg.v);
  };
  PixIter it(vg, active ? H * W : 0, W, PPB, slot);
  if constexpr (PIPE) {
    using Pipe = LdPipe<V, 2>;
    const Pipe pp(sm.pipe);
    PixIter ld = it;
    auto issue = [&](int d) {
      if (ld.valid()) {
        ew_cp_async<V>(pp.slot(d, 0), gbase + (size_t)ld.h * g.hstride + (size_t)ld.w * g.wstride);
        ew_cp_async<V>(pp.slot(d, 1), ybase + (size_t)ld.h * y.hstride + (size_t)ld.w * y.wstride);
        ld.next();
      }
      ew_cp_commit();
    };
#pragma unroll
    for (int d = 0; d < Pipe::D; ++d) issue(d);
    tables();
    for (int d = 0; it.valid(); it.next()) {
      ew_cp_wait<Pipe::D - 1>();
      Vec<V> gg, yy;
      gg.load(pp.slot(d, 0));
      yy.load(pp.slot(d, 1));
      emit(it.h, it.w, gg, yy);
      issue(d);
      d = d + 1 == Pipe::D ? 0 : d + 1;
    }
    ew_cp_wait<0>();
  } else {
    constexpr int U = 2;
    PixBatch<U> b;
    Vec<V> gg[U], yy[U];
    auto fetch = [&]() {
      if (!b.fill(it)) return false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        gg[u].load(gbase + (size_t)b.h[u] * g.hstride + (size_t)b.w[u] * g.wstride);
        yy[u].load(ybase + (size_t)b.h[u] * y.hstride + (size_t)b.w[u] * y.wstride);
      }
      return true;
    };
    bool have = fetch();
    tables();
    while (have) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u >= b.n) break;
        emit(b.h[u], b.w[u], gg[u], yy[u]);
      }
      have = fetch();
    }
  }
}

// B3a: skip branch of the concat: gs = dA[:, :Cs] * lrelu'(bn(ys)), red_s += …
template <int V>
__device__ __forceinline__ void body_cat_bwd_skip(const VGrid& vg, EwSmem sm, MfviView dA, int H, int W, MfviView ys, int Cs, const double* __restrict__ sums_s,
               const float* __restrict__ gamma_s, const float* __restrict__ beta_s, MfviView gs,
               double* __restrict__ red_s, int G, int PPB) {
  double* sm_red = sm.red;
  const int s = vg.by;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  Acc2<V> acc(sm_red);
  BnTable& tab = *sm.tab;
  BnRegs<V> bn;
  const float* dbase = dA.ptr + (size_t)s * dA.sstride + c0;
  const float* ybase = ys.ptr + (size_t)s * ys.sstride + c0;
  float* gbase = gs.ptr + (size_t)s * gs.sstride + c0;
  constexpr int U = 2;
  PixIter it(vg, active ? H * W : 0, W, PPB, slot);
  PixBatch<U> b;
  Vec<V> d[U], yy[U];
  auto fetch = [&]() {
    if (!b.fill(it)) return false;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      d[u].load(dbase + (size_t)b.h[u] * dA.hstride + (size_t)b.w[u] * dA.wstride);
      yy[u].load(ybase + (size_t)b.h[u] * ys.hstride + (size_t)b.w[u] * ys.wstride);
    }
    return true;
  };
  bool have = fetch();
  tab.fill(sums_s, gamma_s, beta_s, s, Cs, 1.0 / ((double)H * W));
  __syncthreads();
  bn.load(tab, c0, Cs);
  while (have) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (u >= b.n) break;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float z = fmaf(yy[u].v[j], bn.sc[j], bn.sh[j]);
        const float gg = z > 0.f ? d[u].v[j] : kLreluSlope * d[u].v[j];
        const float xhat = (yy[u].v[j] - bn.mean[j]) * bn.invstd[j];
        d[u].v[j] = gg;
        acc.fa[j] += gg;
        acc.fb[j] = fmaf(gg, xhat, acc.fb[j]);
      }
      d[u].store(gbase + (size_t)b.h[u] * gs.hstride + (size_t)b.w[u] * gs.wstride);
      acc.tick();
    }
    have = fetch();
  }
  acc.flush();
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

// B3b: deeper branch: gd = up2x^T(dA[:, Cs:]) * lrelu'(bn(yd)), red_d += …
// A thread owns a 2x2 block of low-resolution pixels (rows kh0, kh0+1 x columns kw0, kw0+1) and one channel group.  The block
// receives the 6x6 high-resolution patch rows 2kh0-1..2kh0+4 x columns 2kw0-1..2kw0+4: 36 loads for 4 outputs instead of the
// 16 per output of the per-pixel form, evaluated separably (each patch row is first combined along w for the two columns,
// then added to the two rows with its vertical weights).  Weights come from up_weight_of, i.e. from the forward's up_taps,
// and are 0 outside the image and for non-taps.
template <int V>
__device__ __forceinline__ void body_cat_bwd_up(const VGrid& vg, EwSmem sm, MfviView dA, int H, int W, int mode, int Cs, MfviView yd, int Cd, const double* __restrict__ sums_d,
             const float* __restrict__ gamma_d, const float* __restrict__ beta_d, MfviView gd,
             double* __restrict__ red_d, int G, int PPB) {
  double* sm_red = sm.red;
  const int s = vg.by;
  const int h2 = H / 2, w2 = W / 2;
  const int qh = (h2 + 1) / 2, qw = (w2 + 1) / 2;
  const int group = threadIdx.x % G, slot = threadIdx.x / G;
  const bool active = slot < PPB;
  const int c0 = group * V;
  Acc2<V> acc(sm_red);
  BnTable& tab = *sm.tab;
  tab.fill(sums_d, gamma_d, beta_d, s, Cd, 1.0 / ((double)h2 * w2));
  __syncthreads();
  if (active) {
    BnRegs<V> bn;
    bn.load(tab, c0, Cd);
    const float* dbase = dA.ptr + (size_t)s * dA.sstride + Cs + c0;
    const float* ybase = yd.ptr + (size_t)s * yd.sstride + c0;
    float* gbase = gd.ptr + (size_t)s * gd.sstride + c0;
    for (PixIter it(vg, qh * qw, qw, PPB, slot); it.valid(); it.next()) {
      const int kh0 = 2 * it.h, kw0 = 2 * it.w;
      const bool rin[2] = {true, kh0 + 1 < h2}, cin[2] = {true, kw0 + 1 < w2};
      float wc[2][6];
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        wc[0][d] = up_weight_of(2 * kw0 - 1 + d, kw0, w2, mode);
        wc[1][d] = cin[1] ? up_weight_of(2 * kw0 - 1 + d, kw0 + 1, w2, mode) : 0.f;
      }
      Vec<V> a[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
          for (int j = 0; j < V; ++j) a[u][v].v[j] = 0.f;
#pragma unroll
      for (int dr = 0; dr < 6; ++dr) {
        const int hr = 2 * kh0 - 1 + dr;
        const float wr0 = up_weight_of(hr, kh0, h2, mode), wr1 = rin[1] ? up_weight_of(hr, kh0 + 1, h2, mode) : 0.f;
        if (wr0 == 0.f && wr1 == 0.f) continue;          // also every row outside the image
        const float* row = dbase + (size_t)hr * dA.hstride;
        Vec<V> t[6];
#pragma unroll
        for (int dc = 0; dc < 6; ++dc) {
          const bool on = wc[0][dc] != 0.f || wc[1][dc] != 0.f;
          if (on) t[dc].load(row + (size_t)(2 * kw0 - 1 + dc) * dA.wstride);
          else
#pragma unroll
            for (int j = 0; j < V; ++j) t[dc].v[j] = 0.f;
        }
        float t0[V], t1[V];
#pragma unroll
        for (int j = 0; j < V; ++j) t0[j] = t1[j] = 0.f;
#pragma unroll
        for (int dc = 0; dc < 6; ++dc)
#pragma unroll
          for (int j = 0; j < V; ++j) {
            t0[j] = fmaf(wc[0][dc], t[dc].v[j], t0[j]);
            t1[j] = fmaf(wc[1][dc], t[dc].v[j], t1[j]);
          }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          a[0][0].v[j] = fmaf(wr0, t0[j], a[0][0].v[j]);
          a[0][1].v[j] = fmaf(wr0, t1[j], a[0][1].v[j]);
          a[1][0].v[j] = fmaf(wr1, t0[j], a[1][0].v[j]);
          a[1][1].v[j] = fmaf(wr1, t1[j], a[1][1].v[j]);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!rin[u]) continue;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (!cin[v]) continue;
          const int kh = kh0 + u, kw = kw0 + v;
          Vec<V> yy;
          yy.load(ybase + (size_t)kh * yd.hstride + (size_t)kw * yd.wstride);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float z = fmaf(yy.v[j], bn.sc[j], bn.sh[j]);
            const float gg = z > 0.f ? a[u][v].v[j] : kLreluSlope * a[u][v].v[j];
            const float xhat = (yy.v[j] - bn.mean[j]) * bn.invstd[j];
            a[u][v].v[j] = gg;
            acc.fa[j] += gg;
            acc.fb[j] = fmaf(gg, xhat, acc.fb[j]);
          }
          a[u][v].store(gbase + (size_t)kh * gd.hstride + (size_t)kw * gd.wstride);
        }
      }
      acc.tick();
    }
    acc.flush();
  }
  cta_reduce_cells<V>(sm_red, G, PPB, Cd, red_d + (size_t)s * Cd * 2);
}

}  // namespace mfvi

