// CT forward projector (radon/radon.py:32-55 of the reference) and its adjoint.
// The reference materialises T rotated copies with grid_sample over a precomputed (T,H,W,2) grid and sums rows;
// here the coordinates are analytic (same fp32 formula as affine_grid/grid_sample, align_corners=False) so the only
// traffic is the image (L2-resident, <= 1 MB) and the sinogram.
//   forward : shared-memory staged gather-reduce (k_radon_fwd): a CTA per (sample, channel, angle, 32 detector columns) stages
//             the rotated bounding box of 32 x 32 ray samples in shared memory and reduces along the ray.
//   backward: gather form (no atomics, deterministic): one thread per image pixel loops over angles, inverts the
//             rotation to find the <=3x3 candidate ray samples and re-evaluates the forward weights exactly.
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

// one bilinear tap set of the reference's grid_sample (zeros padding): value at (ix, iy) read from global memory
__device__ __forceinline__ float radon_sample_global(const float* __restrict__ base, const MfviView& img, int H, int W, float ix,
                                                     float iy) {
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
  return (1.f - fy) * ((1.f - fx) * v00 + fx * v01) + fy * ((1.f - fx) * v10 + fx * v11);
}

// Forward projector, shared-memory staged.  A CTA owns kRJ = 32 detector columns of one (sample, channel, angle) and walks the
// ray in blocks of kRI = 32 samples.  The 32 x 32 sample positions of a block are a rotated square of the image, so they fall
// inside a bounding box of at most 32 (|cos| + |sin|) + 2 <= 48 pixels per side: that box (zero-filled outside the image =
// grid_sample's zeros padding) is staged in shared memory with coalesced row reads, and the 4 taps of every sample are read
// from there.  256 threads = 32 columns x 8 ray lanes; the 8 partial sums of a column are added in a fixed order
// (deterministic).  Global traffic per block: <= 52 x 52 pixels for 1024 samples (2.6 loads per sample instead of 4 scattered
// gathers through a 4-float channel pitch).
constexpr int kRJ = 32, kRI = 32, kRT = 52, kRLanes = 8;

__global__ void __launch_bounds__(kRJ * kRLanes)
k_radon_fwd(MfviView img, int C, int H, int W, const float* __restrict__ theta, int T, float* __restrict__ sino) {
  __shared__ float tile[kRT][kRT + 1];
  __shared__ float part[kRLanes][kRJ];
  pdl_trigger();
  pdl_wait();
  const int j0 = blockIdx.x * kRJ;
  const int t = blockIdx.y;
  const int s = blockIdx.z / C, c = blockIdx.z % C;
  const int lj = threadIdx.x & (kRJ - 1), li = threadIdx.x / kRJ;
  const int j = j0 + lj;
  float sn, cs;
  sincosf(theta[t], &sn, &cs);
  const float* base = img.ptr + (size_t)s * img.sstride + c;
  const int j1 = min(j0 + kRJ, W) - 1;
  float acc = 0.f;
  for (int i0 = 0; i0 < H; i0 += kRI) {
    const int i1 = min(i0 + kRI, H) - 1;
    // bounding box of the block's sample positions (the map is affine: the corners are extreme); one pixel of margin each
    // side absorbs fp32 rounding of interior samples, one more on the high side holds the +1 taps
    float xa, ya, xb, yb, xc, yc, xd, yd;
    radon_pos(i0, j0, H, W, sn, cs, xa, ya);
    radon_pos(i0, j1, H, W, sn, cs, xb, yb);
    radon_pos(i1, j0, H, W, sn, cs, xc, yc);
    radon_pos(i1, j1, H, W, sn, cs, xd, yd);
    const int bx0 = (int)floorf(fminf(fminf(xa, xb), fminf(xc, xd))) - 1;
    const int by0 = (int)floorf(fminf(fminf(ya, yb), fminf(yc, yd))) - 1;
    __syncthreads();                                   // the previous block's readers are done with the tile
    for (int idx = threadIdx.x; idx < kRT * kRT; idx += blockDim.x) {
      const int ty = idx / kRT, tx = idx - ty * kRT;
      const int gy = by0 + ty, gx = bx0 + tx;
      float v = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(base + (size_t)gy * img.hstride + (size_t)gx * img.wstride);
      tile[ty][tx] = v;
    }
    __syncthreads();
    if (j < W) {
#pragma unroll
      for (int k = 0; k < kRI / kRLanes; ++k) {
        const int i = i0 + li + k * kRLanes;
        if (i > i1) break;
        float ix, iy;
        radon_pos(i, j, H, W, sn, cs, ix, iy);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int tx = (int)fx0 - bx0, ty = (int)fy0 - by0;
        if ((unsigned)tx < (unsigned)(kRT - 1) && (unsigned)ty < (unsigned)(kRT - 1)) {
          const float fx = ix - fx0, fy = iy - fy0;
          acc += (1.f - fy) * ((1.f - fx) * tile[ty][tx] + fx * tile[ty][tx + 1]) +
                 fy * ((1.f - fx) * tile[ty + 1][tx] + fx * tile[ty + 1][tx + 1]);
        } else {
          acc += radon_sample_global(base, img, H, W, ix, iy);      // never taken if the box bound holds; keeps the result exact
        }
      }
    }
  }
  part[li][lj] = acc;
  __syncthreads();
  if (li == 0 && j < W) {
    float v = part[0][lj];
#pragma unroll
    for (int k = 1; k < kRLanes; ++k) v += part[k][lj];
    sino[(((size_t)s * C + c) * T + t) * W + j] = v;
  }
}

// Adjoint, gather form (no atomics, deterministic): one thread per image pixel loops over the angles.  The ray samples (i, j)
// that touch pixel (x, y) at angle t are those whose position lies within one pixel of it; under the inverse rotation that
// unit square becomes a rotated square around the continuous (j*, i*), so |j - j*| and |i - i*| are below r = |cos| + |sin|
// (<= 1.415): at most 3 x 3 candidates (2 x 2 at axis-aligned angles), each re-evaluated with the forward's own arithmetic so
// that the weights are exactly the forward's.  sin / cos of all angles are tabulated once per CTA.
constexpr int kRadonMaxT = 1024;

__global__ void __launch_bounds__(128)
k_radon_bwd(const float* __restrict__ dsino, int C, int H, int W, const float* __restrict__ theta, int T,
            MfviView dimg) {
  __shared__ float s_sn[kRadonMaxT], s_cs[kRadonMaxT];
  pdl_trigger();
  pdl_wait();
  for (int t = threadIdx.x; t < T; t += blockDim.x) sincosf(theta[t], &s_sn[t], &s_cs[t]);
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int s = blockIdx.z / C, c = blockIdx.z % C;
  if (x >= W) return;
  const float ctr = 0.5f * (float)(W - 1);
  const float a = (float)x - ctr, b = (float)y - ctr;
  const float* ds = dsino + ((size_t)s * C + c) * T * W;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    const float sn = s_sn[t], cs = s_cs[t];
    // inverse rotation about the image centre (H == W): continuous (j*, i*) of this pixel
    const float js = cs * a + sn * b + ctr;
    const float is = -sn * a + cs * b + ctr;
    const float r = fabsf(cs) + fabsf(sn) + 1e-3f;
    const int ja = max((int)ceilf(js - r), 0), jb = min((int)floorf(js + r), W - 1);
    const int ia = max((int)ceilf(is - r), 0), ib = min((int)floorf(is + r), H - 1);
    const float* dst = ds + (size_t)t * W;
    for (int i = ia; i <= ib; ++i) {
      for (int j = ja; j <= jb; ++j) {
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
  dim3 grid((W + kRJ - 1) / kRJ, T, S * C);
  launch_k(k_radon_fwd, grid, kRJ * kRLanes, 0, as_stream(st), img, C, H, W, theta_rad, T, sino);
  return check_launch("radon_fwd");
}

int mfvi_radon_bwd(const float* dsino, int S, int C, int H, int W, const float* theta_rad, int T, MfviView dimg,
                   mfvi_stream_t st) {
  MFVI_REQUIRE(dimg.ptr && theta_rad && dsino, "radon_bwd: null pointer");
  MFVI_REQUIRE(H == W, "radon_bwd: image must be square");
  MFVI_REQUIRE(S >= 1 && C >= 1 && T >= 1 && T <= kRadonMaxT && (long long)S * C <= 65535 && H <= 65535, "radon_bwd: bad sizes");
  dim3 grid((W + 127) / 128, H, S * C);
  launch_k(k_radon_bwd, grid, 128, 0, as_stream(st), dsino, C, H, W, theta_rad, T, dimg);
  return check_launch("radon_bwd");
}

}  // extern "C"
