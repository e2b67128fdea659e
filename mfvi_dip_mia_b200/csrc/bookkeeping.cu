// Per-iteration bookkeeping of the runners on the device (reference
// bayesian_optimization.py:1374-1416, utils/common_utils.py:297-353): exp(-s), exponential moving average of the
// output, clipping, the ring buffers behind the epistemic / aleatoric uncertainty maps, the squared errors behind the
// PSNR values, and the 11x11-Gaussian SSIM.  The reference does this with ~40 ATen launches and >= 8 host
// synchronisations per iteration; here it is ONE HBM-bound launch per iteration with no synchronisation, graph
// capturable (the iteration index is read from the device-side step counter), plus on-demand SSIM / uncertainty kernels.
#include "common.cuh"

namespace mfvi {

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

__device__ __forceinline__ void block_atomic_add(double* __restrict__ dst, double (&v)[5], int n) {
  __shared__ double part[5][8];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double w = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0) part[k][threadIdx.x >> 5] = w;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double w = threadIdx.x < (blockDim.x >> 5) ? part[k][threadIdx.x] : 0.0;
      w = warp_sum(w);
      if (threadIdx.x == 0 && k < n) atomicAdd(dst + k, w);
    }
  }
}

// One pixel per thread (grid-stride).  Cm image channels (1: denoising / SR / CT, 3: inpainting), optionally squashed by a
// sigmoid (inpainting: out[:, :3].sigmoid(), bayesian_optimization.py:3034) and optionally followed by the channel
// s = -log sigma^2 whose exp(-s) is the aleatoric variance (:1375; the CT net has no such channel, :533).
//   cur_c = mean_s f(out[s,p,c]);  cur_var = mean_s exp(-out[s,p,Cm])
// Layouts: gt / noisy (Cm,H,W); mask (H,W) or NULL; out_avg (Cm [+1],H,W); ring_epi (Cm,R,H,W); ring_ale (R,H,W).
// `mask` (inpainting, :3064-3065) multiplies BOTH images of the gt comparisons acc[1], acc[2].
constexpr int kBkSigmoid = 1, kBkAleatoric = 2;

__global__ void __launch_bounds__(256)
k_bookkeep(MfviView out, int S, int HW, int W, int Cm, int flags, float exp_weight, const float* __restrict__ gt,
           const float* __restrict__ noisy, const float* __restrict__ mask, float* __restrict__ out_avg,
           float* __restrict__ ring_epi, float* __restrict__ ring_ale, int R, const uint32_t* __restrict__ iter_dev,
           int iter_offset, double* __restrict__ acc) {
  pdl_trigger();
  pdl_wait();
  const int it = iter_offset + (iter_dev != nullptr ? static_cast<int>(*iter_dev) : 0);
  const int slot = R > 0 ? it % R : 0;
  const float inv_s = 1.f / static_cast<float>(S);
  const bool sig = flags & kBkSigmoid, ale = flags & kBkAleatoric;
  double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int h = p / W, w = p - h * W;
    float m[3] = {0.f, 0.f, 0.f}, v = 0.f;
    for (int s = 0; s < S; ++s) {
      const float* o = out.ptr + view_off(out, s, h, w);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c < Cm) m[c] += sig ? 1.f / (1.f + expf(-o[c])) : o[c];
      if (ale) v += expf(-o[Cm]);                                // aleatoric variance (:1375)
    }
    const float mk = mask != nullptr ? mask[p] : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c >= Cm) break;
      const float mc_raw = m[c] * inv_s;
      const size_t q = static_cast<size_t>(c) * HW + p;
      const float am = it == 0 ? mc_raw : out_avg[q] * exp_weight + mc_raw * (1.f - exp_weight);   // (:1378-1381)
      out_avg[q] = am;
      const float mc = clip01(mc_raw), amc = clip01(am);
      if (R > 0) ring_epi[(static_cast<size_t>(c) * R + slot) * HW + p] = mc;                      // (:1396)
      if (noisy != nullptr) {
        const float t = noisy[q];
        a[0] += static_cast<double>((t - mc) * (t - mc));          // PSNR_noisy / psnr_corrupted
        a[3] += static_cast<double>((t - am) * (t - am));          // mse_corrupted (unclipped, :1388)
      }
      if (gt != nullptr) {
        const float t = gt[q];
        a[1] += static_cast<double>((t * mk - mc * mk) * (t * mk - mc * mk));      // PSNR_gt
        a[2] += static_cast<double>((t * mk - amc * mk) * (t * mk - amc * mk));    // PSNR_gt_sm
        a[4] += static_cast<double>((t - am) * (t - am));          // mse_gt (:1389)
      }
    }
    if (ale) {
      v *= inv_s;
      const size_t q = static_cast<size_t>(Cm) * HW + p;
      out_avg[q] = it == 0 ? v : out_avg[q] * exp_weight + v * (1.f - exp_weight);
      if (R > 0) ring_ale[static_cast<size_t>(slot) * HW + p] = clip01(v);                         // (:1397)
    }
  }
  block_atomic_add(acc, a, 5);
}

// SSIM (utils/common_utils.py:308-353): 11-tap Gaussian (sigma 1.5) window, zero padding, C1 = 0.01^2, C2 = 0.03^2.
// 32x32 output tile per CTA; the five windowed moments are computed separably from a 42x42 staged input tile.
constexpr int kSsimT = 32, kSsimR = 5, kSsimIn = kSsimT + 2 * kSsimR;

__global__ void __launch_bounds__(256)
k_ssim(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int clip_b, double* __restrict__ out_sum) {
  __shared__ float ta[kSsimIn][kSsimIn + 1], tb[kSsimIn][kSsimIn + 1];
  __shared__ float hz[5][kSsimIn][kSsimT + 1];
  __shared__ float g[2 * kSsimR + 1];
  pdl_trigger();
  if (threadIdx.x < 2 * kSsimR + 1) {
    float s = 0.f;
    for (int k = 0; k < 2 * kSsimR + 1; ++k) s += expf(-(k - kSsimR) * (k - kSsimR) / (2.f * 1.5f * 1.5f));
    g[threadIdx.x] = expf(-((int)threadIdx.x - kSsimR) * ((int)threadIdx.x - kSsimR) / (2.f * 1.5f * 1.5f)) / s;
  }
  pdl_wait();
  const int x0 = blockIdx.x * kSsimT - kSsimR, y0 = blockIdx.y * kSsimT - kSsimR;
  for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += blockDim.x) {
    const int r = i / kSsimIn, c = i - r * kSsimIn;
    const int y = y0 + r, x = x0 + c;
    float va = 0.f, vb = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      va = a[(size_t)y * W + x];
      vb = b[(size_t)y * W + x];
      if (clip_b) vb = clip01(vb);
    }
    ta[r][c] = va;
    tb[r][c] = vb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimIn * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, c = i - r * kSsimT;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kSsimR + 1; ++k) {
      const float wa = ta[r][c + k], wb = tb[r][c + k], gk = g[k];
      m1 += gk * wa;
      m2 += gk * wb;
      s11 += gk * wa * wa;
      s22 += gk * wb * wb;
      s12 += gk * wa * wb;
    }
    hz[0][r][c] = m1; hz[1][r][c] = m2; hz[2][r][c] = s11; hz[3][r][c] = s22; hz[4][r][c] = s12;
  }
  __syncthreads();
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < kSsimT * kSsimT; i += blockDim.x) {
    const int r = i / kSsimT, c = i - r * kSsimT;
    const int y = blockIdx.y * kSsimT + r, x = blockIdx.x * kSsimT + c;
    if (y >= H || x >= W) continue;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * kSsimR + 1; ++k) {
      const float gk = g[k];
      m1 += gk * hz[0][r + k][c];
      m2 += gk * hz[1][r + k][c];
      s11 += gk * hz[2][r + k][c];
      s22 += gk * hz[3][r + k][c];
      s12 += gk * hz[4][r + k][c];
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float v1 = s11 - m1 * m1, v2 = s22 - m2 * m2, v12 = s12 - m1 * m2;
    acc[0] += static_cast<double>(((2.f * m1 * m2 + C1) * (2.f * v12 + C2)) / ((m1 * m1 + m2 * m2 + C1) * (v1 + v2 + C2)));
  }
  block_atomic_add(out_sum, acc, 1);
}

// Uncertainty maps from the ring buffers (:1410-1411; UCE recipe eval_denoising.ipynb:467-482):
//   epi[p] = unbiased variance over the n ring entries of the clipped means, ale[p] = mean of the clipped variances,
//   err2[p] = mean_n (mean_n[p] - gt[p])^2 (optional).
__global__ void __launch_bounds__(256)
k_ring_uncertainty(const float* __restrict__ ring_epi, const float* __restrict__ ring_ale, int n, int HW,
                   const float* __restrict__ gt, float* __restrict__ epi, float* __restrict__ ale, float* __restrict__ err2) {
  pdl_trigger();
  pdl_wait();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float mean = 0.f, am = 0.f;
    for (int k = 0; k < n; ++k) {
      mean += ring_epi[(size_t)k * HW + p];
      am += ring_ale[(size_t)k * HW + p];
    }
    mean /= n;
    float var = 0.f, e2 = 0.f;
    const float t = gt != nullptr ? gt[p] : 0.f;
    for (int k = 0; k < n; ++k) {
      const float v = ring_epi[(size_t)k * HW + p];
      var += (v - mean) * (v - mean);
      e2 += (v - t) * (v - t);
    }
    epi[p] = n > 1 ? var / (n - 1) : 0.f;
    ale[p] = am / n;
    if (err2 != nullptr) err2[p] = e2 / n;
  }
}

static inline int grid_1d(size_t items, int threads) {
  size_t blocks = (items + threads - 1) / threads;
  const size_t cap = (size_t)kNumSMs * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_bookkeep_step_ex(MfviView out, int S, int H, int W, int Cm, int flags, float exp_weight, const float* gt,
                          const float* noisy, const float* mask, float* out_avg, float* ring_epi, float* ring_ale, int ring,
                          const uint32_t* iter_dev, int iter_offset, double* acc, mfvi_stream_t st) {
  MFVI_REQUIRE(out.ptr && out_avg && acc, "bookkeep_step: null pointer");
  MFVI_REQUIRE(S >= 1 && H >= 1 && W >= 1, "bookkeep_step: empty output");
  MFVI_REQUIRE(Cm == 1 || Cm == 3, "bookkeep_step: 1 or 3 image channels");
  MFVI_REQUIRE(ring == 0 || (ring_epi && (ring_ale || !(flags & kBkAleatoric))), "bookkeep_step: ring buffers missing");
  launch_k(k_bookkeep, grid_1d((size_t)H * W, 256), 256, 0, as_stream(st), out, S, H * W, W, Cm, flags, exp_weight, gt, noisy, mask,
           out_avg, ring_epi, ring_ale, ring, iter_dev, iter_offset, acc);
  return check_launch("bookkeep_step");
}

int mfvi_bookkeep_step(MfviView out, int S, int H, int W, float exp_weight, const float* gt, const float* noisy,
                       float* out_avg, float* ring_epi, float* ring_ale, int ring, const uint32_t* iter_dev,
                       int iter_offset, double* acc, mfvi_stream_t st) {
  return mfvi_bookkeep_step_ex(out, S, H, W, 1, kBkAleatoric, exp_weight, gt, noisy, nullptr, out_avg, ring_epi, ring_ale, ring,
                               iter_dev, iter_offset, acc, st);
}

int mfvi_ssim(const float* a, const float* b, int H, int W, int clip_b, double* out_sum, mfvi_stream_t st) {
  MFVI_REQUIRE(a && b && out_sum, "ssim: null pointer");
  MFVI_REQUIRE(H >= 1 && W >= 1, "ssim: empty image");
  dim3 grid((W + kSsimT - 1) / kSsimT, (H + kSsimT - 1) / kSsimT);
  launch_k(k_ssim, grid, 256, 0, as_stream(st), a, b, H, W, clip_b, out_sum);
  return check_launch("ssim");
}

int mfvi_ring_uncertainty(const float* ring_epi, const float* ring_ale, int n, int H, int W, const float* gt, float* epi,
                          float* ale, float* err2, mfvi_stream_t st) {
  MFVI_REQUIRE(ring_epi && ring_ale && epi && ale, "ring_uncertainty: null pointer");
  MFVI_REQUIRE(n >= 1, "ring_uncertainty: needs at least one ring entry");
  launch_k(k_ring_uncertainty, grid_1d((size_t)H * W, 256), 256, 0, as_stream(st), ring_epi, ring_ale, n, H * W, gt, epi, ale,
           err2);
  return check_launch("ring_uncertainty");
}

}  // extern "C"
