// Public conv entry points: with MFVI_MATH_TF32 the pointwise kernel (conv_pointwise.cu: small 1x1 layers, exact fp32), then
// the tcgen05 kernels (conv_tc2.cu, conv_tc.cu) when the shape qualifies (each returns -1 otherwise), else the exact-fp32
// CUDA-core kernels (conv_simt.cu).
#include "common.cuh"

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
int mfvi_conv2d_bias_grad_tc(const MfviConvDesc*, MfviView, float*, long long, mfvi_stream_t);
int mfvi_conv2d_fwd_pw(const MfviConvDesc*, MfviView, const float*, const float*, long long, MfviView, double*, mfvi_stream_t);
int mfvi_conv2d_dgrad_pw(const MfviConvDesc*, MfviView, const float*, long long, MfviView, int, mfvi_stream_t);

int mfvi_conv2d_fwd(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride,
                    MfviView y, double* stats, mfvi_stream_t st) {
  if (d != nullptr && d->math == MFVI_MATH_TF32) {
    // small 1x1 layers are streaming work: the CUDA-core pointwise kernel beats the epilogue-bound tensor-core path
    int rc = mfvi_conv2d_fwd_pw(d, x, w, bias, w_sstride, y, stats, st);
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_fwd_tc2(d, x, w, bias, w_sstride, y, stats, st);
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_fwd_tc(d, x, w, bias, w_sstride, y, stats, st);
    if (rc >= 0) return rc;
  }
  return mfvi_conv2d_fwd_simt(d, x, w, bias, w_sstride, y, stats, st);
}

int mfvi_conv2d_dgrad(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx,
                      int accumulate, mfvi_stream_t st) {
  if (d != nullptr && d->math == MFVI_MATH_TF32) {
    int rc = mfvi_conv2d_dgrad_pw(d, dy, w, w_sstride, dx, accumulate, st);
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_dgrad_tc2(d, dy, w, w_sstride, dx, accumulate, st);
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_dgrad_tc(d, dy, w, w_sstride, dx, accumulate, st);
    if (rc >= 0) return rc;
  }
  return mfvi_conv2d_dgrad_simt(d, dy, w, w_sstride, dx, accumulate, st);
}

int mfvi_conv2d_wgrad(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                      mfvi_stream_t st) {
  if (d != nullptr && d->math == MFVI_MATH_TF32) {
    int rc = mfvi_conv2d_wgrad_tc2(d, x, dy, dw, w_sstride, st);
    if (rc == 0 && dbias != nullptr) rc = mfvi_conv2d_bias_grad_tc(d, dy, dbias, w_sstride, st);
    if (rc >= 0) return rc;
    rc = mfvi_conv2d_wgrad_tc(d, x, dy, dw, dbias, w_sstride, st);
    if (rc >= 0) return rc;
  }
  return mfvi_conv2d_wgrad_simt(d, x, dy, dw, dbias, w_sstride, st);
}
}
