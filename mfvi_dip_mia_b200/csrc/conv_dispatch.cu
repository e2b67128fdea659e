// Public conv entry points: pick the tcgen05 kernel when the descriptor asks for it and the shape qualifies,
// otherwise the fp32 CUDA-core kernels.
#include "common.cuh"

extern "C" {
int mfvi_conv2d_fwd_simt(const MfviConvDesc*, MfviView, const float*, const float*, long long, MfviView, double*, mfvi_stream_t);
int mfvi_conv2d_dgrad_simt(const MfviConvDesc*, MfviView, const float*, long long, MfviView, int, mfvi_stream_t);
int mfvi_conv2d_wgrad_simt(const MfviConvDesc*, MfviView, MfviView, float*, float*, long long, mfvi_stream_t);

int mfvi_conv2d_fwd(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride,
                    MfviView y, double* stats, mfvi_stream_t st) {
  return mfvi_conv2d_fwd_simt(d, x, w, bias, w_sstride, y, stats, st);
}

int mfvi_conv2d_dgrad(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx,
                      int accumulate, mfvi_stream_t st) {
  return mfvi_conv2d_dgrad_simt(d, dy, w, w_sstride, dx, accumulate, st);
}

int mfvi_conv2d_wgrad(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                      mfvi_stream_t st) {
  return mfvi_conv2d_wgrad_simt(d, x, dy, dw, dbias, w_sstride, st);
}
}
