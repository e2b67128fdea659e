// Public conv entry points: with MFVI_MATH_TF32 the pointwise kernel (conv_pointwise.cu: small 1x1 layers, exact fp32), then
// the tcgen05 kernels (conv_tc2.cu, conv_tc.cu) when the shape qualifies (each returns -1 otherwise), else the exact-fp32
// CUDA-core kernels (conv_simt.cu).  mfvi_conv2d_plan walks the same chains in planning-only mode (common.cuh: dry_run).
#include "common.cuh"
#include "mega.cuh"

namespace mfvi {
namespace mega {
int record_conv(int op, const MfviConvDesc* d, MfviView act_in, MfviView act_out, const float* w, const float* bias,
                long long w_sstride, float* dw, double* stats, int accumulate);
int launch_conv_mma(int op, const MfviConvDesc* d, MfviView act_in, MfviView act_out, const float* w, const float* bias,
                    long long w_sstride, float* dw, double* stats, int accumulate, mfvi_stream_t stream, const char* what);
}
}  // namespace mfvi

extern "C" {
int mfvi_conv2d_fwd_simt(const MfviConvDesc*, MfviView, const float*, const float*, long long, MfviView, double*, mfvi_stream_t);
int mfvi_conv2d_dgrad_simt(const MfviConvDesc*, MfviView, const float*, long long, MfviView, int, mfvi_stream_t);
int mfvi_conv2d_wgrad_simt(const MfviConvDesc*, MfviView, MfviView, float*, float*, long long, mfvi_stream_t);
int mfvi_conv2d_fwd_tc(const MfviConvDesc*, MfviView, const float*, const float*, long long, MfviView, double*, mfvi_stream_t);
int mfvi_conv2d_dgrad_tc(const MfviConvDesc*, MfviView, const float*, long long, MfviView, int, mfvi_stream_t);
int mfvi_conv2d_wgrad_tc(const MfviConvDesc*, MfviView, MfviView, float*, float*, long long, mfvi_stream_t);
int mfvi_conv2d_fwd_tc2(const MfviConvDesc*, MfviView, const float*, const float*, long long, MfviView, double*, mfvi_stream_t);
int mfvi_conv2d_dgrad_tc2(const MfviConvDesc*, MfviView, const float*, long long, MfviView, int, mfvi_stream_t);
int mfvi_conv2d_wgrad_tc2(const MfviConvDesc*, MfviView, MfviView, float*, long long, mfvi_stream_t);
int mfvi_conv2d_wgrad_pw(const MfviConvDesc*, MfviView, MfviView, float*, float*, long long, mfvi_stream_t);
int mfvi_conv2d_bias_grad_tc(const MfviConvDesc*, MfviView, float*, long long, mfvi_stream_t);
int mfvi_conv2d_fwd_pw(const MfviConvDesc*, MfviView, const float*, const float*, long long, MfviView, double*, mfvi_stream_t);
int mfvi_conv2d_dgrad_pw(const MfviConvDesc*, MfviView, const float*, long long, MfviView, int, mfvi_stream_t);


// The chains name the kernel family that took the shape: "pointwise", "halo" (conv_tc2.cu), "alias" (conv_wgrad2.cu),
// "tc" (conv_tc.cu) or "simt".
// Exact-fp32 mode on 3xTF32 tensor-core tiles instead of the CUDA-core kernels: OFF by default — measured on a B200 the
// mma.sync tiles of mega.cu run the 256x256 MC=8 step at 76 steps/s against 154 steps/s for conv_simt.cu
// (profiles/r02_fp32_mma_negative.txt); MFVI_FP32_MMA=1 switches it on.
static bool fp32_mma_on() {
  static const bool on = [] { const char* e = getenv("MFVI_FP32_MMA"); return e != nullptr && e[0] == '1'; }();
  return on;
}

static int fwd_chain(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                     double* stats, mfvi_stream_t st, const char** family) {
  if (d != nullptr && d->math == MFVI_MATH_TF32) {
    // small 1x1 layers are streaming work: the CUDA-core pointwise kernel beats the epilogue-bound tensor-core path
    int rc = mfvi_conv2d_fwd_pw(d, x, w, bias, w_sstride, y, stats, st);
    *family = "pointwise";
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_fwd_tc2(d, x, w, bias, w_sstride, y, stats, st);
    *family = "halo";
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_fwd_tc(d, x, w, bias, w_sstride, y, stats, st);
    *family = "tc";
    if (rc >= 0) return rc;
  }
  if (d != nullptr && d->math == MFVI_MATH_FP32 && fp32_mma_on()) {
    // exact-fp32 mode: 3xTF32 tensor-core tiles (mega.cu) — fp32 accuracy, several times the CUDA-core rate
    const int rc = mfvi::mega::launch_conv_mma(mfvi::mega::OP_CONV_FWD, d, x, y, w, bias, w_sstride, nullptr, stats, 0, st, "conv2d_fwd_mma");
    *family = "mma3";
    if (rc >= 0) return rc;
  }
  *family = "simt";
  return mfvi_conv2d_fwd_simt(d, x, w, bias, w_sstride, y, stats, st);
}

static int dgrad_chain(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                       mfvi_stream_t st, const char** family) {
  if (d != nullptr && d->math == MFVI_MATH_TF32) {
    int rc = mfvi_conv2d_dgrad_pw(d, dy, w, w_sstride, dx, accumulate, st);
    *family = "pointwise";
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_dgrad_tc2(d, dy, w, w_sstride, dx, accumulate, st);
    *family = "halo";
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_dgrad_tc(d, dy, w, w_sstride, dx, accumulate, st);
    *family = "tc";
    if (rc >= 0) return rc;
  }
  if (d != nullptr && d->math == MFVI_MATH_FP32 && fp32_mma_on()) {
    const int rc = mfvi::mega::launch_conv_mma(mfvi::mega::OP_CONV_DGRAD, d, dy, dx, w, nullptr, w_sstride, nullptr, nullptr, accumulate, st,
                                               "conv2d_dgrad_mma");
    *family = "mma3";
    if (rc >= 0) return rc;
  }
  *family = "simt";
  return mfvi_conv2d_dgrad_simt(d, dy, w, w_sstride, dx, accumulate, st);
}

static int wgrad_chain(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                       mfvi_stream_t st, const char** family) {
  if (d != nullptr && d->math == MFVI_MATH_TF32) {
    // 1x1 layers with <= 4 output channels: a streaming reduction (weight and bias gradient in one launch)
    int rc = mfvi_conv2d_wgrad_pw(d, x, dy, dw, dbias, w_sstride, st);
    *family = "pointwise";
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_wgrad_tc2(d, x, dy, dw, w_sstride, st);
    if (rc == 0 && dbias != nullptr) rc = mfvi_conv2d_bias_grad_tc(d, dy, dbias, w_sstride, st);
    *family = "alias";
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_wgrad_tc(d, x, dy, dw, dbias, w_sstride, st);
    *family = "tc";
    if (rc >= 0) return rc;
  }
  if (d != nullptr && d->math == MFVI_MATH_FP32 && fp32_mma_on()) {
    int rc = mfvi::mega::launch_conv_mma(mfvi::mega::OP_CONV_WGRAD, d, x, dy, nullptr, nullptr, w_sstride, dw, nullptr, 0, st,
                                         "conv2d_wgrad_mma");
    if (rc == 0 && dbias != nullptr) rc = mfvi_conv2d_bias_grad_tc(d, dy, dbias, w_sstride, st);      // exact fp32 reduction of dy
    *family = "mma3";
    if (rc >= 0) return rc;
  }
  *family = "simt";
  return mfvi_conv2d_wgrad_simt(d, x, dy, dw, dbias, w_sstride, st);
}

int mfvi_conv2d_fwd(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride,
                    MfviView y, double* stats, mfvi_stream_t st) {
  if (mfvi::mega::recording()) return mfvi::mega::record_conv(mfvi::mega::OP_CONV_FWD, d, x, y, w, bias, w_sstride, nullptr, stats, 0);
  const char* family;
  return fwd_chain(d, x, w, bias, w_sstride, y, stats, st, &family);
}

int mfvi_conv2d_dgrad(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx,
                      int accumulate, mfvi_stream_t st) {
  if (mfvi::mega::recording())
    return mfvi::mega::record_conv(mfvi::mega::OP_CONV_DGRAD, d, dy, dx, w, nullptr, w_sstride, nullptr, nullptr, accumulate);
  const char* family;
  return dgrad_chain(d, dy, w, w_sstride, dx, accumulate, st, &family);
}

int mfvi_conv2d_wgrad(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                      mfvi_stream_t st) {
  if (mfvi::mega::recording()) {
    MFVI_REQUIRE(dbias == nullptr, "mega: a bias gradient is not part of the fused stages");
    return mfvi::mega::record_conv(mfvi::mega::OP_CONV_WGRAD, d, x, dy, nullptr, nullptr, w_sstride, dw, nullptr, 0);
  }
  const char* family;
  return wgrad_chain(d, x, dy, dw, dbias, w_sstride, st, &family);
}

// Which kernel family would take this convolution, and with which launch geometry / tile plan.  Host-only: no device is
// touched, so it also runs on a machine without a GPU.  `a`, `b` are the two activation views of the pass (only their
// alignment and strides are looked at; the pointers are never dereferenced): pass 0 = forward (x, y), 1 = data gradient
// (dy, dx), 2 = weight gradient (x, dy).
int mfvi_conv2d_plan(const MfviConvDesc* d, int pass, MfviView a, MfviView b, long long w_sstride, int accumulate, int with_bias,
                     MfviPlanInfo* out) {
  MFVI_REQUIRE(d != nullptr && out != nullptr, "conv2d_plan: null argument");
  MFVI_REQUIRE(pass >= 0 && pass <= 2, "conv2d_plan: pass must be 0 (fwd), 1 (dgrad) or 2 (wgrad)");
  mfvi::DryRunInfo info{};
  // any 16-byte aligned non-null address stands in for the weight / bias / statistics buffers
  float* const fake = reinterpret_cast<float*>(static_cast<uintptr_t>(0x1000));
  // a planning-only host (tensors on torch's "meta" device) numbers its buffers from address 0: shift both views by 1 MiB,
  // which keeps every alignment property and passes the kernels' null-pointer checks
  a.ptr = reinterpret_cast<float*>(reinterpret_cast<uintptr_t>(a.ptr) + (1u << 20));
  b.ptr = reinterpret_cast<float*>(reinterpret_cast<uintptr_t>(b.ptr) + (1u << 20));
  const char* family = "";
  mfvi::set_dry_run(&info);
  int rc;
  if (pass == 0)
    rc = fwd_chain(d, a, fake, with_bias ? fake : nullptr, w_sstride, b, reinterpret_cast<double*>(fake), nullptr, &family);
  else if (pass == 1)
    rc = dgrad_chain(d, a, fake, w_sstride, b, accumulate, nullptr, &family);
  else
    rc = wgrad_chain(d, a, b, fake, with_bias ? fake : nullptr, w_sstride, nullptr, &family);
  mfvi::set_dry_run(nullptr);
  if (rc != 0) return rc;
  memset(out, 0, sizeof(*out));
  strncpy(out->family, family, sizeof(out->family) - 1);
  for (int i = 0; i < 3; ++i) out->grid[i] = info.grid[i];
  out->block = info.block;
  out->smem_bytes = info.smem;
  out->launches = info.launches;
  strncpy(out->detail, info.detail, sizeof(out->detail) - 1);
  return 0;
}

// explicit entry points of the mma.sync tile kernels (mega.cu) as stand-alone launches, whatever the dispatch default
int mfvi_conv2d_fwd_mma(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                        double* stats, mfvi_stream_t st) {
  const int rc = mfvi::mega::launch_conv_mma(mfvi::mega::OP_CONV_FWD, d, x, y, w, bias, w_sstride, nullptr, stats, 0, st, "conv2d_fwd_mma");
  MFVI_REQUIRE(rc >= 0, "conv2d_fwd_mma: shape not taken (stride must be 1 or 2)");
  return rc;
}
int mfvi_conv2d_dgrad_mma(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                          mfvi_stream_t st) {
  const int rc = mfvi::mega::launch_conv_mma(mfvi::mega::OP_CONV_DGRAD, d, dy, dx, w, nullptr, w_sstride, nullptr, nullptr, accumulate, st,
                                             "conv2d_dgrad_mma");
  MFVI_REQUIRE(rc >= 0, "conv2d_dgrad_mma: shape not taken (stride must be 1 or 2)");
  return rc;
}
int mfvi_conv2d_wgrad_mma(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                          mfvi_stream_t st) {
  int rc = mfvi::mega::launch_conv_mma(mfvi::mega::OP_CONV_WGRAD, d, x, dy, nullptr, nullptr, w_sstride, dw, nullptr, 0, st,
                                       "conv2d_wgrad_mma");
  MFVI_REQUIRE(rc >= 0, "conv2d_wgrad_mma: shape not taken (stride must be 1 or 2)");
  if (rc == 0 && dbias != nullptr) rc = mfvi_conv2d_bias_grad_tc(d, dy, dbias, w_sstride, st);
  return rc;
}
}
