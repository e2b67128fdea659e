// Local-reparameterisation layers (reference BayTorch/modules/reparam_layers.py:39-72): the two
// convolutions (x * mu and x^2 * sigma^2) run on the library's conv kernels; these are the flat elementwise pieces around
// them — sigma^2 = softplus(rho)^2, x^2, the output-space noise  out = act_mu + sqrt(1e-16 + act_var) * eps, and their
// backward chains.  All HBM-bound, float4 where alignment allows.
#include "common.cuh"

namespace mfvi {

static inline int lrt_grid(size_t n) {
  size_t blocks = (n / 4 + 256) / 256;
  const size_t cap = (size_t)kNumSMs * 8;
  return (int)(blocks > cap ? cap : blocks);
}

// op 0: out = softplus(a)^2           op 1: out = a*a
// op 2: out (+)= b * 2*softplus(a)*sigmoid(a)      (d softplus^2 / d rho)        op 3: out (+)= b * 2*a   (d x^2 / dx)
// op 4: out = a + sqrt(1e-16 + b) * c               (a = act_mu, b = act_var, c = eps)
// op 5: out = a * c / (2*sqrt(1e-16 + b))           (a = dout,   b = act_var, c = eps -> d act_var)
template <int OP>
__device__ __forceinline__ float lrt_apply(float a, float b, float c, float old) {
  if (OP == 0) { const float s = softplus_f(a); return s * s; }
  if (OP == 1) return a * a;
  if (OP == 2) return old + b * 2.f * softplus_f(a) * sigmoid_f(a);
  if (OP == 3) return old + b * 2.f * a;
  if (OP == 4) return a + sqrtf(1e-16f + b) * c;
  return a * c / (2.f * sqrtf(1e-16f + b));
}

template <int OP>
__global__ void __launch_bounds__(256)
k_lrt(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, float* __restrict__ out,
      size_t n, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                     reinterpret_cast<uintptr_t>(out)) % 16) == 0;
  const size_t nv = vec ? n / 4 : 0;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < nv; i += nt) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 va = __ldg(reinterpret_cast<const float4*>(a) + i);
    const float4 vb = b != nullptr ? __ldg(reinterpret_cast<const float4*>(b) + i) : z;
    const float4 vc = c != nullptr ? __ldg(reinterpret_cast<const float4*>(c) + i) : z;
    const float4 vo = ((OP == 2 || OP == 3) && accumulate) ? reinterpret_cast<const float4*>(out)[i] : z;
    reinterpret_cast<float4*>(out)[i] = make_float4(lrt_apply<OP>(va.x, vb.x, vc.x, vo.x), lrt_apply<OP>(va.y, vb.y, vc.y, vo.y),
                                                    lrt_apply<OP>(va.z, vb.z, vc.z, vo.z), lrt_apply<OP>(va.w, vb.w, vc.w, vo.w));
  }
  for (size_t i = nv * 4 + tid; i < n; i += nt) {
    const float old = ((OP == 2 || OP == 3) && accumulate) ? out[i] : 0.f;
    out[i] = lrt_apply<OP>(a[i], b != nullptr ? b[i] : 0.f, c != nullptr ? c[i] : 0.f, old);
  }
}

template <int OP>
static int lrt_launch(const char* what, const float* a, const float* b, const float* c, float* out, size_t n, int accumulate,
                      mfvi_stream_t st) {
  if (n == 0) return 0;
  launch_k(k_lrt<OP>, lrt_grid(n), 256, 0, as_stream(st), a, b, c, out, n, accumulate);
  return check_launch(what);
}

}  // namespace mfvi

using namespace mfvi;

extern "C" {

int mfvi_softplus_sq_fwd(const float* rho, size_t n, float* sigma2, mfvi_stream_t st) {
  MFVI_REQUIRE(rho && sigma2, "softplus_sq_fwd: null pointer");
  return lrt_launch<0>("softplus_sq_fwd", rho, nullptr, nullptr, sigma2, n, 0, st);
}

int mfvi_softplus_sq_bwd(const float* rho, const float* dsigma2, size_t n, float* drho, int accumulate, mfvi_stream_t st) {
  MFVI_REQUIRE(rho && dsigma2 && drho, "softplus_sq_bwd: null pointer");
  return lrt_launch<2>("softplus_sq_bwd", rho, dsigma2, nullptr, drho, n, accumulate, st);
}

int mfvi_square_fwd(const float* x, size_t n, float* x2, mfvi_stream_t st) {
  MFVI_REQUIRE(x && x2, "square_fwd: null pointer");
  return lrt_launch<1>("square_fwd", x, nullptr, nullptr, x2, n, 0, st);
}

int mfvi_square_bwd(const float* x, const float* dx2, size_t n, float* dx, int accumulate, mfvi_stream_t st) {
  MFVI_REQUIRE(x && dx2 && dx, "square_bwd: null pointer");
  return lrt_launch<3>("square_bwd", x, dx2, nullptr, dx, n, accumulate, st);
}

int mfvi_lrt_noise_fwd(const float* act_mu, const float* act_var, const float* eps, size_t n, float* out, mfvi_stream_t st) {
  MFVI_REQUIRE(act_mu && act_var && eps && out, "lrt_noise_fwd: null pointer");
  return lrt_launch<4>("lrt_noise_fwd", act_mu, act_var, eps, out, n, 0, st);
}

int mfvi_lrt_noise_bwd(const float* dout, const float* act_var, const float* eps, size_t n, float* dvar, mfvi_stream_t st) {
  MFVI_REQUIRE(dout && act_var && eps && dvar, "lrt_noise_bwd: null pointer");
  return lrt_launch<5>("lrt_noise_bwd", dout, act_var, eps, dvar, n, 0, st);
}

}  // extern "C"
