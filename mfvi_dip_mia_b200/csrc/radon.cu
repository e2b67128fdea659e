// CT forward projector (radon/radon.py:32-55 of the reference) and its adjoint.
// The reference materialises T rotated copies with grid_sample over a precomputed (T,H,W,2) grid and sums rows;
// here the coordinates are analytic (same fp32 formula as affine_grid/grid_sample, align_corners=False) so the only
// traffic is the image (L2-resident, <= 1 MB) and the sinogram.
//   forward : one warp-lane per detector column j of one (sample, channel, angle); loops over the H ray samples
//             (neighbouring lanes hit neighbouring pixels -> coalesced / L1-hit gathers).
//   backward: gather form (no atomics, deterministic): one thread per image pixel loops over angles, inverts the
//             rotation to find the <=4x4 candidate ray samples and re-evaluates the forward weights exactly.
#include "common.cuh"

namespace mfvi {

// forward sample position of ray sample (i, j) at angle (sn, cs): same operation order as the reference grid
__device__ __forceinline__ void radon_pos(int i, int j, int H, int W, float sn, float cs, float& ix, float& iy) {
  const float xj = (2.f * (float)j + 1.f) / (float)W - 1.f;
  const float yi = (2.f * (float)i + 1.f) / (float)H - 1.f;
  const float gx = cs * xj - sn * yi;
  const float gy = sn * xj + cs * yi;
  ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f;
  iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
}

__global__ void __launch_bounds__(128)
k_radon_fwd(MfviView img, int C, int H, int W, const float* __restrict__ theta, int T, float* __restrict__ sino) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  const int s = blockIdx.z / C, c = blockIdx.z % C;
  if (j >= W) return;
  float sn, cs;
  sincosf(theta[t], &sn, &cs);
  const float* base = img.ptr + (size_t)s * img.sstride + c;
  float acc = 0.f;
  for (int i = 0; i < H; ++i) {
    float ix, iy;
    radon_pos(i, j, H, W, sn, cs, ix, iy);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float fx = ix - fx0, fy = iy - fy0;
    const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W;
    const bool ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
    float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
    if (ya && xa) v00 = __ldg(base + (size_t)y0 * img.hstride + (size_t)x0 * img.wstride);
    if (ya && xb) v01 = __ldg(base + (size_t)y0 * img.hstride + (size_t)(x0 + 1) * img.wstride);
    if (yb && xa) v10 = __ldg(base + (size_t)(y0 + 1) * img.hstride + (size_t)x0 * img.wstride);
    if (yb && xb) v11 = __ldg(base + (size_t)(y0 + 1) * img.hstride + (size_t)(x0 + 1) * img.wstride);
    acc += (1.f - fy) * ((1.f - fx) * v00 + fx * v01) + fy * ((1.f - fx) * v10 + fx * v11);
  }
  sino[(((size_t)s * C + c) * T + t) * W + j] = acc;
}

__global__ void __launch_bounds__(128)
k_radon_bwd(const float* __restrict__ dsino, int C, int H, int W, const float* __restrict__ theta, int T,
            MfviView dimg) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int s = blockIdx.z / C, c = blockIdx.z % C;
  if (x >= W) return;
  const float ctr = 0.5f * (float)(W - 1);
  const float a = (float)x - ctr, b = (float)y - ctr;
  const float* ds = dsino + ((size_t)s * C + c) * T * W;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    float sn, cs;
    sincosf(theta[t], &sn, &cs);
    // inverse rotation about the image centre (H == W): continuous (j*, i*) of this pixel
    const float js = cs * a + sn * b + ctr;
    const float is = -sn * a + cs * b + ctr;
    const int j0 = (int)floorf(js) - 1, i0 = (int)floorf(is) - 1;
    const float* dst = ds + (size_t)t * W;
    for (int di = 0; di < 4; ++di) {
      const int i = i0 + di;
      if (i < 0 || i >= H) continue;
      for (int dj = 0; dj < 4; ++dj) {
        const int j = j0 + dj;
        if (j < 0 || j >= W) continue;
        float ix, iy;
        radon_pos(i, j, H, W, sn, cs, ix, iy);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float fx = ix - fx0, fy = iy - fy0;
        float wx = 0.f, wy = 0.f;
        if (x == x0) wx = 1.f - fx; else if (x == x0 + 1) wx = fx;
        if (y == y0) wy = 1.f - fy; else if (y == y0 + 1) wy = fy;
        const float wgt = wx * wy;
        if (wgt != 0.f) acc = fmaf(wgt, __ldg(dst + j), acc);
      }
    }
  }
  dimg.ptr[(size_t)s * dimg.sstride + (size_t)y * dimg.hstride + (size_t)x * dimg.wstride + c] = acc;
}

}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_radon_fwd(MfviView img, int S, int C, int H, int W, const float* theta_rad, int T, float* sino,
                   mfvi_stream_t st) {
  MFVI_REQUIRE(img.ptr && theta_rad && sino, "radon_fwd: null pointer");
  MFVI_REQUIRE(H == W, "radon_fwd: image must be square (FastRadonTransform asserts the same)");
  MFVI_REQUIRE(S >= 1 && C >= 1 && T >= 1 && (long long)S * C <= 65535 && T <= 65535, "radon_fwd: bad sizes");
  dim3 grid((W + 127) / 128, T, S * C);
  k_radon_fwd<<<grid, 128, 0, as_stream(st)>>>(img, C, H, W, theta_rad, T, sino);
  return check_launch("radon_fwd");
}

int mfvi_radon_bwd(const float* dsino, int S, int C, int H, int W, const float* theta_rad, int T, MfviView dimg,
                   mfvi_stream_t st) {
  MFVI_REQUIRE(dimg.ptr && theta_rad && dsino, "radon_bwd: null pointer");
  MFVI_REQUIRE(H == W, "radon_bwd: image must be square");
  MFVI_REQUIRE(S >= 1 && C >= 1 && T >= 1 && (long long)S * C <= 65535 && H <= 65535, "radon_bwd: bad sizes");
  dim3 grid((W + 127) / 128, H, S * C);
  k_radon_bwd<<<grid, 128, 0, as_stream(st)>>>(dsino, C, H, W, theta_rad, T, dimg);
  return check_launch("radon_bwd");
}

}  // extern "C"
