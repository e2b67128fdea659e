// Flat / elementwise kernels of the MFVI-DIP step: RNG fills, weight sampling, tempered KL + reparam chain,
// losses, AdamW, input jitter.  All HBM-bound: vectorised (float4) coalesced access, grid sized in multiples
// of the SM count, warp-shuffle reductions, one double atomic per CTA.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"
#include "mega.cuh"

namespace mfvi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MFVI_PDL"); return e == nullptr || e[0] != '0'; }();
  return on;
}

static thread_local DryRunInfo* g_dry = nullptr;

DryRunInfo* dry_run() { return g_dry; }
void set_dry_run(DryRunInfo* info) { g_dry = info; }

void dry_note(dim3 grid, dim3 block, size_t smem) {
  if (g_dry == nullptr) return;
  if (g_dry->launches == 0) {          // the first launch of a call is the convolution; later ones are helpers (bias grad)
    g_dry->grid[0] = grid.x; g_dry->grid[1] = grid.y; g_dry->grid[2] = grid.z;
    g_dry->block = block.x * block.y * block.z;
    g_dry->smem = smem;
  }
  ++g_dry->launches;
}

void dry_detail(const char* fmt, ...) {
  if (g_dry == nullptr || g_dry->launches != 0) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_dry->detail, sizeof(g_dry->detail), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  if (g_dry != nullptr) return 0;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

__device__ __forceinline__ uint32_t eff_step(const MfviPhiloxKey& key) {
  return key.step + (key.step_dev != nullptr ? *key.step_dev : 0u);
}

static inline int grid_for(size_t work_items, int threads, int max_waves = 8) {
  size_t blocks = (work_items + threads - 1) / threads;
  size_t cap = static_cast<size_t>(kNumSMs) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---------------------------------------------------------------------------------------------
__global__ void k_philox_raw(uint32_t* __restrict__ out, size_t n_words, MfviPhiloxKey key, uint32_t stream_id) {
  const size_t nb = (n_words + 3) / 4;
  const uint32_t step = eff_step(key);
  for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const uint4 x = philox4x32_10((uint32_t)b, stream_id, key.sample0, step, (uint32_t)key.seed,
                                  (uint32_t)(key.seed >> 32));
    const uint32_t v[4] = {x.x, x.y, x.z, x.w};
    for (int j = 0; j < 4; ++j)
      if (b * 4 + j < n_words) out[b * 4 + j] = v[j];
  }
}

__global__ void k_philox_normal(float* __restrict__ out, size_t n, MfviPhiloxKey key, uint32_t stream_id) {
  const size_t nb = (n + 3) / 4;
  const uint32_t step = eff_step(key);
  for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const float4 z = philox_normal4((uint32_t)b, stream_id, key.sample0, step, key.seed);
    const float v[4] = {z.x, z.y, z.z, z.w};
    for (int j = 0; j < 4; ++j)
      if (b * 4 + j < n) out[b * 4 + j] = v[j];
  }
}

// w_s[i] = mu[i] + softplus(rho[i]) * eps_s[i]; one thread per Philox block (4 params) per sample.
__global__ void k_sample_weights(const float* __restrict__ mu, const float* __restrict__ rho, size_t n, int S,
                                 const float* __restrict__ eps, long long eps_sstride, MfviPhiloxKey key,
                                 float* __restrict__ w_out, long long w_sstride) {
  pdl_trigger();
  pdl_wait();
  const size_t nb = (n + 3) / 4;
  const uint32_t step = eff_step(key);
  // vector path per 4-element block: strides keep 16-byte alignment; only the last (partial) block of n goes scalar
  const bool vec_ok = (w_sstride % 4 == 0) && (eps == nullptr || eps_sstride % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(rho) | reinterpret_cast<uintptr_t>(w_out) |
                        reinterpret_cast<uintptr_t>(eps)) % 16 == 0);
  for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const bool vec = vec_ok && (b * 4 + 4 <= n);
    float m[4], sg[4];
    if (vec) {
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu) + b);
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(rho) + b);
      m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
      sg[0] = softplus_f(r4.x); sg[1] = softplus_f(r4.y); sg[2] = softplus_f(r4.z); sg[3] = softplus_f(r4.w);
    } else {
      for (int j = 0; j < 4; ++j) {
        const size_t i = b * 4 + j;
        m[j] = i < n ? mu[i] : 0.f;
        sg[j] = i < n ? softplus_f(rho[i]) : 0.f;
      }
    }
    for (int s = 0; s < S; ++s) {
      float e[4];
      if (eps != nullptr) {
        if (vec) {
          const float4 e4 = __ldg(reinterpret_cast<const float4*>(eps + (size_t)s * eps_sstride) + b);
          e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
        } else {
          for (int j = 0; j < 4; ++j) {
            const size_t i = b * 4 + j;
            e[j] = i < n ? eps[(size_t)s * eps_sstride + i] : 0.f;
          }
        }
      } else {
        const float4 z = philox_normal4((uint32_t)b, MFVI_STREAM_WEIGHTS, key.sample0 + s, step, key.seed);
        e[0] = z.x; e[1] = z.y; e[2] = z.z; e[3] = z.w;
      }
      float* dst = w_out + (size_t)s * w_sstride + b * 4;
      if (vec) {
        *reinterpret_cast<float4*>(dst) =
            make_float4(fmaf(sg[0], e[0], m[0]), fmaf(sg[1], e[1], m[1]), fmaf(sg[2], e[2], m[2]), fmaf(sg[3], e[3], m[3]));
      } else {
        for (int j = 0; j < 4; ++j)
          if (b * 4 + j < n) dst[j] = fmaf(sg[j], e[j], m[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// KL + reparameterisation chain, flat.
// reverse (direction 0): KL(N(mp,sp) || N(mu,sig)) = 0.5*(r + t - 1 - ln r), r=(sp/sig)^2, t=((mp-mu)/sig)^2
//    dKL/dmu  = (mu-mp)/sig^2 ;  dKL/dsig = (sig^2 - sp^2 - (mu-mp)^2)/sig^3
// forward (direction 1): KL(N(mu,sig) || N(mp,sp)): r=(sig/sp)^2, t=((mu-mp)/sp)^2
//    dKL/dmu = (mu-mp)/sp^2 ; dKL/dsig = sig/sp^2 - 1/sig
// drho = dsig * sigmoid(rho)
__global__ void __launch_bounds__(256)
k_kl_reparam(const float* __restrict__ mu, const float* __restrict__ rho, size_t n, float mp, float sp, int direction,
             float kscale, const float* __restrict__ kscale_dev, const float* __restrict__ dw, long long dw_sstride, int S, const float* __restrict__ eps,
             long long eps_sstride, MfviPhiloxKey key, float gscale, double* __restrict__ kl_out,
             float* __restrict__ grad_mu, float* __restrict__ grad_rho, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const size_t nb = (n + 3) / 4;
  const uint32_t step = eff_step(key);
  if (kscale_dev != nullptr) kscale *= *kscale_dev;
  double kl_acc = 0.0;
  const bool vec_ok = (dw == nullptr || dw_sstride % 4 == 0) && (eps == nullptr || eps_sstride % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(rho) | reinterpret_cast<uintptr_t>(dw) |
                        reinterpret_cast<uintptr_t>(eps) | reinterpret_cast<uintptr_t>(grad_mu) |
                        reinterpret_cast<uintptr_t>(grad_rho)) % 16 == 0);
  for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const bool vec = vec_ok && (b * 4 + 4 <= n);
    float gm[4] = {0.f, 0.f, 0.f, 0.f}, ge[4] = {0.f, 0.f, 0.f, 0.f};
    float mr[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float gold[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    if (vec) {
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu) + b), r4 = __ldg(reinterpret_cast<const float4*>(rho) + b);
      mr[0][0] = m4.x; mr[0][1] = m4.y; mr[0][2] = m4.z; mr[0][3] = m4.w;
      mr[1][0] = r4.x; mr[1][1] = r4.y; mr[1][2] = r4.z; mr[1][3] = r4.w;
      if (accumulate && grad_mu != nullptr) {
        const float4 a4 = reinterpret_cast<const float4*>(grad_mu)[b], c4 = reinterpret_cast<const float4*>(grad_rho)[b];
        gold[0][0] = a4.x; gold[0][1] = a4.y; gold[0][2] = a4.z; gold[0][3] = a4.w;
        gold[1][0] = c4.x; gold[1][1] = c4.y; gold[1][2] = c4.z; gold[1][3] = c4.w;
      }
    } else {
      for (int j = 0; j < 4; ++j) {
        const size_t i = b * 4 + j;
        if (i < n) {
          mr[0][j] = mu[i];
          mr[1][j] = rho[i];
          if (accumulate && grad_mu != nullptr) {
            gold[0][j] = grad_mu[i];
            gold[1][j] = grad_rho[i];
          }
        }
      }
    }
    if (dw != nullptr) {
      if (vec) {
        // the S gradient rows are independent streams: issue the loads of a group of 4 samples before the Philox work
        for (int s0 = 0; s0 < S; s0 += 4) {
          float4 d4[4], e4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int s = s0 + u < S ? s0 + u : S - 1;
            d4[u] = __ldg(reinterpret_cast<const float4*>(dw + (size_t)s * dw_sstride) + b);
            if (eps != nullptr) e4[u] = __ldg(reinterpret_cast<const float4*>(eps + (size_t)s * eps_sstride) + b);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (s0 + u >= S) break;
            if (eps == nullptr) e4[u] = philox_normal4((uint32_t)b, MFVI_STREAM_WEIGHTS, key.sample0 + s0 + u, step, key.seed);
            gm[0] += d4[u].x; gm[1] += d4[u].y; gm[2] += d4[u].z; gm[3] += d4[u].w;
            ge[0] = fmaf(d4[u].x, e4[u].x, ge[0]); ge[1] = fmaf(d4[u].y, e4[u].y, ge[1]);
            ge[2] = fmaf(d4[u].z, e4[u].z, ge[2]); ge[3] = fmaf(d4[u].w, e4[u].w, ge[3]);
          }
        }
      } else {
        for (int s = 0; s < S; ++s) {
          float e[4];
          if (eps == nullptr) {
            const float4 z = philox_normal4((uint32_t)b, MFVI_STREAM_WEIGHTS, key.sample0 + s, step, key.seed);
            e[0] = z.x; e[1] = z.y; e[2] = z.z; e[3] = z.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const size_t i = b * 4 + j;
            if (i < n) {
              const float d = dw[(size_t)s * dw_sstride + i];
              const float ee = eps != nullptr ? eps[(size_t)s * eps_sstride + i] : e[j];
              gm[j] += d;
              ge[j] = fmaf(d, ee, ge[j]);
            }
          }
        }
      }
    }
    float ga[4], gc[4];
    float kl_local = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = b * 4 + j;
      if (i >= n) continue;
      const float m = mr[0][j], r = mr[1][j];
      const float sig = softplus_f(r);
      const float sgm = sigmoid_f(r);
      const float dm = m - mp;
      float kl, dkl_dmu, dkl_dsig;
      if (direction == 0) {
        const float inv = 1.f / sig;
        const float rr = sp * inv, tt = dm * inv;
        const float r2 = rr * rr;
        kl = 0.5f * (r2 + tt * tt - 1.f - __logf(r2));
        dkl_dmu = dm * inv * inv;
        dkl_dsig = (1.f - r2 - tt * tt) * inv;
      } else {
        const float inv = 1.f / sp;
        const float rr = sig * inv, tt = dm * inv;
        const float r2 = rr * rr;
        kl = 0.5f * (r2 + tt * tt - 1.f - __logf(r2));
        dkl_dmu = dm * inv * inv;
        dkl_dsig = sig * inv * inv - 1.f / sig;
      }
      kl_local += kl;
      ga[j] = gscale * gm[j] + kscale * dkl_dmu + gold[0][j];
      gc[j] = (gscale * ge[j] + kscale * dkl_dsig) * sgm + gold[1][j];
      if (grad_mu != nullptr && !vec) {
        grad_mu[i] = ga[j];
        grad_rho[i] = gc[j];
      }
    }
    if (grad_mu != nullptr && vec) {
      reinterpret_cast<float4*>(grad_mu)[b] = make_float4(ga[0], ga[1], ga[2], ga[3]);
      reinterpret_cast<float4*>(grad_rho)[b] = make_float4(gc[0], gc[1], gc[2], gc[3]);
    }
    kl_acc += (double)kl_local;
  }
  if (kl_out != nullptr) {
    __shared__ double part[8];
    kl_acc = warp_sum(kl_acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = kl_acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
      v = warp_sum(v);
      if (threadIdx.x == 0) atomicAdd(kl_out, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian NLL forward + backward.  One thread per contributing pixel.
//   mode 0: loss = exp(clamp(s)) * (t-mu)^2 - clamp(s);  d/dmu = -2 e (t-mu); d/ds = (e (t-mu)^2 - 1)*[|s|<=20]
//   mode 1: inpainting: mu_c = sigmoid(o_c), c<3; s = o_3 shared; loss_c = (e (t_c-mu_c)^2 - s) * mask
//   mode 3: as mode 1 with mu_c = o_c (sigmoid already applied by the caller)
__global__ void __launch_bounds__(256)
k_nll(int mode, MfviView out, int S, int H, int W, int C, int sub, const float* __restrict__ target,
      const float* __restrict__ mask, double* __restrict__ loss_out, MfviView dout, float inv_count) {
  pdl_trigger();
  pdl_wait();
  const int Hs = H / sub, Ws = W / sub;
  const size_t per_s = (size_t)H * W;
  const size_t total = per_s * S;
  double acc = 0.0;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(idx / per_s);
    const int rem = (int)(idx % per_s);
    const int h = rem / W, w = rem % W;
    const float* o = out.ptr + view_off(out, s, h, w);
    float* d = dout.ptr + view_off(dout, s, h, w);
    if (mode == 0) {
      const bool on = (h % sub == 0) && (w % sub == 0) && (h / sub < Hs) && (w / sub < Ws);
      float dmu = 0.f, ds = 0.f;
      if (on) {
        const float t = target[(size_t)(h / sub) * Ws + (w / sub)];
        const float mu = o[0], sraw = o[1];
        const float sc = fminf(fmaxf(sraw, -20.f), 20.f);
        const float e = __expf(sc);
        const float diff = t - mu;
        acc += (double)(e * diff * diff - sc);
        dmu = -2.f * e * diff * inv_count;
        ds = (sraw >= -20.f && sraw <= 20.f) ? (e * diff * diff - 1.f) * inv_count : 0.f;
      }
      d[0] = dmu;
      d[1] = ds;
      for (int c = 2; c < C; ++c) d[c] = 0.f;
    } else {
      const float mk = mask[(size_t)h * W + w];
      const float sraw = o[3];
      const float sc = fminf(fmaxf(sraw, -20.f), 20.f);
      const float e = __expf(sc);
      const bool pass = (sraw >= -20.f && sraw <= 20.f);
      float ds = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float t = target[((size_t)h * W + w) * 3 + c];
        const float p = mode == 1 ? sigmoid_f(o[c]) : o[c];
        const float diff = t - p;
        acc += (double)((e * diff * diff - sc) * mk);
        d[c] = -2.f * e * diff * mk * (mode == 1 ? p * (1.f - p) : 1.f) * inv_count;
        ds += pass ? (e * diff * diff - 1.f) * mk : 0.f;
      }
      d[3] = ds * inv_count;
    }
  }
  __shared__ double part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0 && loss_out != nullptr) atomicAdd(loss_out, v * (double)inv_count);
  }
}

__global__ void __launch_bounds__(256)
k_mse(const float* __restrict__ a, long long a_sstride, const float* __restrict__ b, size_t n, int S,
      double* __restrict__ loss_out, float* __restrict__ da, float inv_count) {
  pdl_trigger();
  pdl_wait();
  double acc = 0.0;
  const size_t total = n * S;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int s = (int)(idx / n);
    const size_t i = idx % n;
    const float diff = a[(size_t)s * a_sstride + i] - b[i];
    acc += (double)(diff * diff);
    if (da != nullptr) da[(size_t)s * a_sstride + i] = 2.f * diff * inv_count;
  }
  __shared__ double part[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0 && loss_out != nullptr) atomicAdd(loss_out, v * (double)inv_count);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
        float lr, float b1, float b2, float eps, float wd, int step, const uint32_t* __restrict__ step_dev,
        const float* __restrict__ skip_if_nonfinite) {
  pdl_trigger();
  pdl_wait();
  if (skip_if_nonfinite != nullptr) {
    const float l = *skip_if_nonfinite;
    if (!(l == l) || fabsf(l) > 3.0e38f) return;
  }
  __shared__ float sm_bc[2];
  if (threadIdx.x == 0) {
    const double t = (double)step + (step_dev != nullptr ? (double)*step_dev : 0.0);
    sm_bc[0] = (float)(1.0 - pow((double)b1, t));
    sm_bc[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float bc1 = sm_bc[0], bc2_sqrt = sm_bc[1];
  const float step_size = lr / bc1;
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define MFVI_ADAM1(X)                                         \
  pp.X *= (1.f - lr * wd);                                    \
  mm.X = b1 * mm.X + (1.f - b1) * gg.X;                       \
  vv.X = b2 * vv.X + (1.f - b2) * gg.X * gg.X;                \
  pp.X -= step_size * mm.X / (sqrtf(vv.X) / bc2_sqrt + eps);
    MFVI_ADAM1(x) MFVI_ADAM1(y) MFVI_ADAM1(z) MFVI_ADAM1(w)
#undef MFVI_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = n4 * 4 + threadIdx.x;
    float pp = p[i] * (1.f - lr * wd);
    const float gg = g[i];
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    pp -= step_size * mm / (sqrtf(vv) / bc2_sqrt + eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

// xp[hp][wp][c] = saved[h][w][c] + std * z(c,h,w);  (h,w) = reflect(hp-pad, wp-pad)
// padded positions whose reflection is source index h (per dimension): fills q[0..cnt)
__device__ __forceinline__ int jitter_targets(int h, int n, int pad, int (&q)[3]) {
  int cnt = 0;
  q[cnt++] = h + pad;
  if (h >= 1 && h <= pad) q[cnt++] = pad - h;
  if (h <= n - 2 && h >= n - 1 - pad) q[cnt++] = 2 * (n - 1) - h + pad;
  return cnt;
}

// One thread per (h, 4 consecutive w, c): ONE Philox block yields the four normals of NCHW flat indices (c,h,w..w+3)
// (W % 4 == 0), each written to its interior position and to the border positions that reflect onto it.
__global__ void __launch_bounds__(256)
k_input_jitter_pad4(const float* __restrict__ saved, int H, int W, int C, float stdv, int pad, MfviPhiloxKey key, MfviView xp) {
  pdl_trigger();
  pdl_wait();
  const int W4 = W >> 2;
  const int total = H * W4 * C;
  const uint32_t step = eff_step(key);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % C;
    const int g = idx / C;
    const int w0 = (g % W4) << 2, h = g / W4;
    const uint32_t flat = (static_cast<uint32_t>(c) * H + h) * W + w0;
    const float4 z4 = philox_normal4(flat >> 2, MFVI_STREAM_INPUT_JITTER, key.sample0, step, key.seed);
    const float z[4] = {z4.x, z4.y, z4.z, z4.w};
    int qh[3];
    const int nh = jitter_targets(h, H, pad, qh);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = w0 + j;
      const float v = fmaf(stdv, z[j], __ldg(saved + (static_cast<size_t>(h) * W + w) * C + c));
      int qw[3];
      const int nw = jitter_targets(w, W, pad, qw);
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b)
          xp.ptr[static_cast<size_t>(qh[a]) * xp.hstride + static_cast<size_t>(qw[b]) * xp.wstride + c] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
k_input_jitter_pad(const float* __restrict__ saved, const float* __restrict__ noise, int H, int W, int C, float stdv,
                   int pad, MfviPhiloxKey key, MfviView xp) {
  pdl_trigger();
  pdl_wait();
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const size_t total = (size_t)Hp * Wp * C;
  const uint32_t step = eff_step(key);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const size_t pix = idx / C;
    const int wp = (int)(pix % Wp), hp = (int)(pix / Wp);
    const int h = reflect_idx(hp - pad, H), w = reflect_idx(wp - pad, W);
    const size_t src = ((size_t)h * W + w) * C + c;
    float z;
    if (noise != nullptr) {
      z = noise[src];
    } else {
      const size_t flat = ((size_t)c * H + h) * W + w;  // NCHW flat index of the reference's noise tensor
      const float4 z4 = philox_normal4((uint32_t)(flat >> 2), MFVI_STREAM_INPUT_JITTER, key.sample0, step, key.seed);
      const int j = (int)(flat & 3);
      z = j == 0 ? z4.x : j == 1 ? z4.y : j == 2 ? z4.z : z4.w;
    }
    xp.ptr[(size_t)hp * xp.hstride + (size_t)wp * xp.wstride + c] = fmaf(stdv, z, saved[src]);
  }
}

__global__ void k_counter_add(uint32_t* ctr, uint32_t inc, const float* __restrict__ only_if_finite) {
  pdl_trigger();
  pdl_wait();
  if (only_if_finite != nullptr) {
    const float l = *only_if_finite;
    if (!(l == l) || fabsf(l) > 3.0e38f) return;
  }
  *ctr += inc;
}

// flag = (float)(data loss + temp * kl): the scalar the reference's NaN guard tests (bayesian_optimization.py:577,581)
__global__ void k_loss_flag(const double* __restrict__ kl, const double* __restrict__ nll, float temp, float* __restrict__ flag) {
  pdl_trigger();
  pdl_wait();
  *flag = static_cast<float>(*nll + static_cast<double>(temp) * *kl);
}

__global__ void k_fill(float* __restrict__ p, size_t n, float v) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// dst NHWC <- src NCHW (to_nhwc) or the inverse; simple gather, used at setup / for returning `out`.
__global__ void k_layout(const float* __restrict__ src, float* __restrict__ dst, int N, int C, int H, int W,
                         int to_nhwc) {
  const size_t total = (size_t)N * C * H * W;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    // idx enumerates dst
    if (to_nhwc) {
      const int c = (int)(idx % C);
      size_t r = idx / C;
      const int w = (int)(r % W); r /= W;
      const int h = (int)(r % H);
      const int n = (int)(r / H);
      dst[idx] = src[(((size_t)n * C + c) * H + h) * W + w];
    } else {
      const int w = (int)(idx % W);
      size_t r = idx / W;
      const int h = (int)(r % H); r /= H;
      const int c = (int)(r % C);
      const int n = (int)(r / C);
      dst[idx] = src[(((size_t)n * H + h) * W + w) * C + c];
    }
  }
}

}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_abi_version(void) { return MFVI_ABI_VERSION; }
const char* mfvi_last_error(void) { return mfvi::g_err; }

int mfvi_philox_raw_fill(uint32_t* out, size_t n_words, MfviPhiloxKey key, uint32_t stream_id, mfvi_stream_t st) {
  MFVI_REQUIRE(out != nullptr, "philox_raw_fill: null out");
  if (n_words == 0) return 0;
  k_philox_raw<<<grid_for((n_words + 3) / 4, 256), 256, 0, as_stream(st)>>>(out, n_words, key, stream_id);
  return check_launch("philox_raw_fill");
}

int mfvi_philox_normal_fill(float* out, size_t n, MfviPhiloxKey key, uint32_t stream_id, mfvi_stream_t st) {
  MFVI_REQUIRE(out != nullptr, "philox_normal_fill: null out");
  if (n == 0) return 0;
  k_philox_normal<<<grid_for((n + 3) / 4, 256), 256, 0, as_stream(st)>>>(out, n, key, stream_id);
  return check_launch("philox_normal_fill");
}

int mfvi_sample_weights(const float* mu, const float* rho, size_t n, int S, const float* eps, long long eps_sstride,
                        MfviPhiloxKey key, float* w_out, long long w_sstride, mfvi_stream_t st) {
  MFVI_REQUIRE(mu && rho && w_out, "sample_weights: null pointer");
  MFVI_REQUIRE(S >= 1, "sample_weights: S must be >= 1");
  if (n == 0) return 0;
  launch_k(k_sample_weights, grid_for((n + 3) / 4, 256), 256, 0, as_stream(st), mu, rho, n, S, eps, eps_sstride, key, w_out,
                                                                          w_sstride);
  return check_launch("sample_weights");
}

int mfvi_kl_reparam_fwd_bwd(const float* mu, const float* rho, size_t n, float prior_mu, double prior_sigma_plus_eps,
                            int direction, float kscale, const float* kscale_dev, const float* dw,
                            long long dw_sstride, int S, const float* eps, long long eps_sstride, MfviPhiloxKey key,
                            float gscale, double* kl_out, float* grad_mu, float* grad_rho, int accumulate,
                            mfvi_stream_t st) {
  MFVI_REQUIRE(mu && rho, "kl_reparam: null parameters");
  MFVI_REQUIRE((grad_mu == nullptr) == (grad_rho == nullptr), "kl_reparam: grad_mu and grad_rho must both be set or both NULL");
  MFVI_REQUIRE(prior_sigma_plus_eps > 0.0, "kl_reparam: prior scale must be positive");
  MFVI_REQUIRE(direction == 0 || direction == 1, "kl_reparam: direction must be 0 (reverse) or 1 (forward)");
  if (n == 0) return 0;
  if (S <= 0) dw = nullptr;
  launch_k(k_kl_reparam, grid_for((n + 3) / 4, 256, 8), 256, 0, as_stream(st), 
      mu, rho, n, prior_mu, (float)prior_sigma_plus_eps, direction, kscale, kscale_dev, dw, dw_sstride, S, eps, eps_sstride, key,
      gscale, kl_out, grad_mu, grad_rho, accumulate);
  return check_launch("kl_reparam");
}

int mfvi_gauss_nll_fwd_bwd(int mode, MfviView out, int S, int H, int W, int C, int sub, const float* target,
                           const float* mask, double* loss_out, MfviView dout, mfvi_stream_t st) {
  MFVI_REQUIRE(mode == 0 || mode == 1 || mode == 3, "gauss_nll: mode must be 0, 1 or 3");
  MFVI_REQUIRE(out.ptr && dout.ptr && target, "gauss_nll: null pointer");
  MFVI_REQUIRE(sub >= 1 && H % sub == 0 && W % sub == 0, "gauss_nll: H,W must be multiples of sub");
  MFVI_REQUIRE(mode == 0 ? C >= 2 : (C == 4 && mask != nullptr && sub == 1), "gauss_nll: bad channel count / mask");
  const double count = mode == 0 ? (double)(H / sub) * (W / sub) * S : (double)H * W * 3 * S;
  launch_k(k_nll, grid_for((size_t)H * W * S, 256), 256, 0, as_stream(st), mode, out, S, H, W, C, sub, target, mask, loss_out,
                                                                    dout, (float)(1.0 / count));
  return check_launch("gauss_nll");
}

int mfvi_mse_fwd_bwd(const float* a, long long a_sstride, const float* b, size_t n, int S, double* loss_out, float* da,
                     mfvi_stream_t st) {
  MFVI_REQUIRE(a && b, "mse: null pointer");
  launch_k(k_mse, grid_for(n * S, 256), 256, 0, as_stream(st), a, a_sstride, b, n, S, loss_out, da,
                                                         (float)(1.0 / ((double)n * S)));
  return check_launch("mse");
}

int mfvi_adamw_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, const uint32_t* step_dev, const float* skip_if_nonfinite,
                    mfvi_stream_t st) {
  MFVI_REQUIRE(p && g && m && v, "adamw: null pointer");
  MFVI_REQUIRE(step >= 1 || (step >= 0 && step_dev != nullptr), "adamw: step counts from 1");
  MFVI_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) % 16 == 0, "adamw: buffers must be 16-byte aligned");
  launch_k(k_adamw, grid_for(n / 4 + 1, 256), 256, 0, as_stream(st), p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                step, step_dev, skip_if_nonfinite);
  return check_launch("adamw");
}

int mfvi_counter_add(uint32_t* ctr, uint32_t inc, mfvi_stream_t st) {
  MFVI_REQUIRE(ctr != nullptr, "counter_add: null pointer");
  launch_k(k_counter_add, 1, 1, 0, as_stream(st), ctr, inc, static_cast<const float*>(nullptr));
  return check_launch("counter_add");
}

int mfvi_counter_add_if_finite(uint32_t* ctr, uint32_t inc, const float* flag, mfvi_stream_t st) {
  MFVI_REQUIRE(ctr != nullptr && flag != nullptr, "counter_add_if_finite: null pointer");
  launch_k(k_counter_add, 1, 1, 0, as_stream(st), ctr, inc, flag);
  return check_launch("counter_add_if_finite");
}

int mfvi_loss_flag(const double* kl, const double* nll, float temp, float* flag, mfvi_stream_t st) {
  MFVI_REQUIRE(kl && nll && flag, "loss_flag: null pointer");
  launch_k(k_loss_flag, 1, 1, 0, as_stream(st), kl, nll, temp, flag);
  return check_launch("loss_flag");
}

int mfvi_input_jitter_pad(const float* saved, const float* noise, int H, int W, int C, float stdv, int pad,
                          MfviPhiloxKey key, MfviView xp, mfvi_stream_t st) {
  MFVI_REQUIRE(saved && xp.ptr, "input_jitter_pad: null pointer");
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "input_jitter_pad: pad must be smaller than the image");
  if (noise == nullptr && W % 4 == 0 && 2 * pad < H && 2 * pad < W && (size_t)H * W * C < (1u << 31))
    launch_k(k_input_jitter_pad4, grid_for((size_t)H * (W / 4) * C, 256), 256, 0, as_stream(st), saved, H, W, C, stdv, pad, key, xp);
  else
    launch_k(k_input_jitter_pad, grid_for((size_t)(H + 2 * pad) * (W + 2 * pad) * C, 256), 256, 0, as_stream(st),
             saved, noise, H, W, C, stdv, pad, key, xp);
  return check_launch("input_jitter_pad");
}

int mfvi_fill_f32(float* p, size_t n, float v, mfvi_stream_t st) {
  if (n == 0) return 0;
  MFVI_REQUIRE(p, "fill: null pointer");
  if (mega::Stage* ms = mega::append(mega::OP_FILL)) {
    ms->fill_ptr = p; ms->fill_n = n; ms->fill_v = v;
    return 0;
  }
  launch_k(k_fill, grid_for(n, 256), 256, 0, as_stream(st), p, n, v);
  return check_launch("fill");
}

int mfvi_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, mfvi_stream_t st) {
  MFVI_REQUIRE(src && dst, "nchw_to_nhwc: null pointer");
  k_layout<<<grid_for((size_t)N * C * H * W, 256), 256, 0, as_stream(st)>>>(src, dst, N, C, H, W, 1);
  return check_launch("nchw_to_nhwc");
}

int mfvi_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, mfvi_stream_t st) {
  MFVI_REQUIRE(src && dst, "nhwc_to_nchw: null pointer");
  k_layout<<<grid_for((size_t)N * C * H * W, 256), 256, 0, as_stream(st)>>>(src, dst, N, C, H, W, 0);
  return check_launch("nhwc_to_nchw");
}

}  // extern "C"
