// bf16-operand mode (EXPERIMENTAL, DESIGN.md section 8, stage C): the two conversions the engine needs besides the
// bf16-writing elementwise kernels — an NHWC view (network input, loss gradient) and the sampled weights.  Both are small and
// HBM-bound; round-to-nearest-even like torch's .to(torch.bfloat16).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mfvi {

// dst[s][h][w][c] = bf16(src[s][h][w][c]); one thread per element, channels fastest.  grid-stride.
__global__ void __launch_bounds__(256) k_view_to_bf16(MfviView src, int S, int H, int W, int C, MfviView dst) {
  pdl_trigger();
  pdl_wait();
  __nv_bfloat16* d16 = reinterpret_cast<__nv_bfloat16*>(dst.ptr);
  const size_t n = static_cast<size_t>(S) * H * W * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    size_t r = i / C;
    const int w = static_cast<int>(r % W);
    r /= W;
    const int h = static_cast<int>(r % H), s = static_cast<int>(r / H);
    d16[view_off(dst, s, h, w) + c] = __float2bfloat16_rn(src.ptr[view_off(src, s, h, w) + c]);
  }
}

// Weight blocks of all layers at once.  fp32 storage: w[s][w_off[l] + row*cin[l] + ci] (row = tap*Cout + co);
// bf16 storage: w16[s][w16_off[l] + row*cpitch[l] + ci], cpitch = cin rounded up to 8 (16-byte rows for TMA); the padding
// channels are written as zero.  One thread per 8-channel group of a row; `grp_start[l]` = first group index of layer l
// (grp_start[n_layers] = total).
struct PackTable {
  static constexpr int kMaxLayers = 40;
  int n_layers;
  long long w_off[kMaxLayers], w16_off[kMaxLayers];
  int cin[kMaxLayers], cpitch[kMaxLayers];
  long long grp_start[kMaxLayers + 1];
};

__global__ void __launch_bounds__(256) k_pack_weights_bf16(const float* __restrict__ w, long long w_sstride, int S,
                                                           const __grid_constant__ PackTable t, void* __restrict__ w16v,
                                                           long long w16_sstride) {
  pdl_trigger();
  pdl_wait();
  __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(w16v);
  const long long total = t.grp_start[t.n_layers];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total * S;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int s = static_cast<int>(i / total);
    const long long g = i - static_cast<long long>(s) * total;
    int l = 0;
    while (l + 1 < t.n_layers && g >= t.grp_start[l + 1]) ++l;
    const long long gl = g - t.grp_start[l];
    const int gpr = t.cpitch[l] / 8;                     // groups per row
    const long long row = gl / gpr;
    const int c0 = static_cast<int>(gl - row * gpr) * 8;
    const float* src = w + static_cast<size_t>(s) * w_sstride + t.w_off[l] + row * t.cin[l] + c0;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16_rn(c0 + j < t.cin[l] ? src[j] : 0.f);
    // w16_off, cpitch and w16_sstride are multiples of 8 elements: one 16-byte store
    *reinterpret_cast<uint4*>(w16 + static_cast<size_t>(s) * w16_sstride + t.w16_off[l] + row * t.cpitch[l] + c0) =
        *reinterpret_cast<const uint4*>(v);
  }
}

}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_view_f32_to_bf16(MfviView src, int S, int H, int W, int C, MfviView dst, mfvi_stream_t st) {
  MFVI_REQUIRE(src.ptr && dst.ptr && S >= 1 && H >= 1 && W >= 1 && C >= 1, "view_f32_to_bf16: bad argument");
  const size_t n = static_cast<size_t>(S) * H * W * C;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(kNumSMs) * 8));
  launch_k(k_view_to_bf16, blocks, 256, 0, as_stream(st), src, S, H, W, C, dst);
  return check_launch("view_f32_to_bf16");
}

// HOST arrays of n_layers entries: w_off / w16_off (element offsets of a layer's block inside one sample's fp32 / bf16
// storage), rows (= KH*KW*Cout), cin.  w16_off and w16_sstride must be multiples of 8; a bf16 row holds cin rounded up to 8.
int mfvi_pack_weights_bf16(const float* w, long long w_sstride, int S, int n_layers, const long long* w_off,
                           const long long* w16_off, const int* rows, const int* cin, void* w16, long long w16_sstride,
                           mfvi_stream_t st) {
  MFVI_REQUIRE(w && w16 && w_off && w16_off && rows && cin, "pack_weights_bf16: null pointer");
  MFVI_REQUIRE(n_layers >= 1 && n_layers <= PackTable::kMaxLayers, "pack_weights_bf16: 1..%d layers", PackTable::kMaxLayers);
  MFVI_REQUIRE(w16_sstride % 8 == 0 && reinterpret_cast<uintptr_t>(w16) % 16 == 0, "pack_weights_bf16: unaligned bf16 storage");
  PackTable t{};
  t.n_layers = n_layers;
  long long g = 0;
  for (int l = 0; l < n_layers; ++l) {
    MFVI_REQUIRE(w16_off[l] % 8 == 0 && rows[l] >= 1 && cin[l] >= 1, "pack_weights_bf16: bad layer %d", l);
    t.w_off[l] = w_off[l]; t.w16_off[l] = w16_off[l]; t.cin[l] = cin[l];
    t.cpitch[l] = (cin[l] + 7) / 8 * 8;
    t.grp_start[l] = g;
    g += static_cast<long long>(rows[l]) * (t.cpitch[l] / 8);
  }
  t.grp_start[n_layers] = g;
  const long long n = g * S;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(kNumSMs) * 8));
  launch_k(k_pack_weights_bf16, blocks, 256, 0, as_stream(st), w, w_sstride, S, t, w16, w16_sstride);
  return check_launch("pack_weights_bf16");
}

}  // extern "C"
