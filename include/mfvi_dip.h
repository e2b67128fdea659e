/*
 * libmfvidip — C ABI of the B200-native MFVI-DIP training-step kernels.
 *
 * The reference (Cardio-AI/mfvi-dip-mia) has no FFI: its "plugin API" for this path is Python
 * (BayTorch modules, MeanFieldVI, models/skip.py, radon/radon.py, utils/bayesian_utils.py) calling
 * PyTorch/ATen.  Each entry point below names the reference call it replaces (file:line relative to
 * the reference root).  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; mfvi_last_error() gives the message
 *    (thread-local);
 *  - the caller owns and allocates every buffer (device pointers unless stated otherwise); no hidden
 *    allocation, no global mutable state; every call is asynchronous on `stream` and CUDA-graph
 *    capturable;
 *  - activations are NHWC fp32 "views": base pointer + sample/row/pixel strides in ELEMENTS, so a conv
 *    can read the interior of a reflection-padded buffer or broadcast one input to all MC samples
 *    (sample stride 0);
 *  - variational parameters live in flat fp32 buffers mu[P], rho[P]; a layer's weight block is stored
 *    tap-major [KH][KW][Cout][Cin] (the tcgen05 K-major operand order); the host side places all weight
 *    blocks first and all bias vectors [Cout] behind them, and exposes reference-shaped (Cout,Cin,KH,KW)
 *    views of that storage;
 *  - eps is a pure function of (seed, step, global sample id, flat parameter index): Philox4x32-10 +
 *    Box-Muller (oracle/philox.py restates it).  Passing `eps != NULL` injects eps instead ("eps
 *    injected identically" parity tests).
 */
#ifndef MFVI_DIP_H_
#define MFVI_DIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mfvi_stream_t; /* cudaStream_t */

#define MFVI_ABI_VERSION 1

/* Philox stream ids (counter word c1) */
#define MFVI_STREAM_WEIGHTS 0u
#define MFVI_STREAM_INPUT_JITTER 1u

typedef struct {
  uint64_t seed;            /* Philox key */
  uint32_t step;            /* optimiser step (counter word c3 = step + *step_dev) */
  uint32_t sample0;         /* global id of local sample 0 (counter word c2 = sample0 + s) */
  const uint32_t* step_dev; /* optional DEVICE counter added to `step` when the kernel runs, so that a captured
                               CUDA graph draws fresh eps on every replay (see mfvi_counter_add); may be NULL */
} MfviPhiloxKey;

/* NHWC fp32 tensor view: element (s,h,w,c) at ptr[s*sstride + h*hstride + w*wstride + c]. */
typedef struct {
  float* ptr;
  long long sstride; /* 0 = broadcast one image to all samples */
  int hstride;
  int wstride;
} MfviView;

/* Geometry of one sampled-weight convolution, padding=0 semantics on an already padded input
 * (the reference pads with nn.ReflectionPad2d and calls F.conv2d(padding=0): models/common.py:117-123,
 *  BayTorch/modules/reparam_layers.py:37). */
typedef struct {
  int S;              /* MC samples in this launch */
  int Cin, Cout;
  int KH, KW, stride;
  int Hin, Win;       /* padded input size  */
  int Hout, Wout;     /* (Hin-KH)/stride+1 … */
  int math;           /* MFVI_MATH_FP32: fp32 CUDA-core path (exact parity mode);
                         MFVI_MATH_TF32: tcgen05 kind::tf32 implicit GEMM, fp32 accumulate in TMEM
                         (shapes the tensor-core kernel does not take fall back to fp32 CUDA cores) */
} MfviConvDesc;

#define MFVI_MATH_FP32 0
#define MFVI_MATH_TF32 1

int mfvi_abi_version(void);
const char* mfvi_last_error(void);

/* ---- RNG (replaces torch.randn_like in VIModule.rsample, BayTorch/modules/module.py:82-85, and
 *      noise.normal_() in the runners, bayesian_optimization.py:1363-1364) ------------------------ */
int mfvi_philox_raw_fill(uint32_t* out, size_t n_words, MfviPhiloxKey key, uint32_t stream_id, mfvi_stream_t st);
int mfvi_philox_normal_fill(float* out, size_t n, MfviPhiloxKey key, uint32_t stream_id, mfvi_stream_t st);

/* ---- a2: w_s = mu + softplus(rho)*eps_s for all layers at once (flat, storage layout) ---------------
 * BayTorch/modules/module.py:82-85 + reparam_layers.py:28-30.  w_out[s*w_sstride + i], s<S, i<n.
 * eps (optional) is indexed eps[s*eps_sstride + i]. */
int mfvi_sample_weights(const float* mu, const float* rho, size_t n, int S, const float* eps, long long eps_sstride,
                        MfviPhiloxKey key, float* w_out, long long w_sstride, mfvi_stream_t st);

/* ---- a2: the convolution itself (F.conv2d in RTLayer.forward, reparam_layers.py:37) -----------------
 * w: sampled weights of this layer, [S][KH][KW][Cout][Cin] with sample stride w_sstride; bias [S][Cout]
 * (sample stride w_sstride, may be NULL).  stats (optional): double[S][Cout][2] += (sum y, sum y^2) for
 * the BatchNorm that follows (models/common.py:96-97). */
int mfvi_conv2d_fwd(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride,
                    MfviView y, double* stats, mfvi_stream_t st);
/* dx (padded-input-sized) = conv_transpose(dy, w); accumulate!=0 adds into dx. */
int mfvi_conv2d_dgrad(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx,
                      int accumulate, mfvi_stream_t st);
/* dw[s] (+)= x (*) dy in storage layout [KH][KW][Cout][Cin]; dbias[s][Cout] (+)= sum_pixels dy.
 * Both are accumulated with atomics: zero them first. */
int mfvi_conv2d_wgrad(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                      mfvi_stream_t st);

/* Explicit kernel families behind mfvi_conv2d_{fwd,dgrad,wgrad} (same arguments; used by the tests and for benchmarking one
 * family against another): *_simt = the exact-fp32 CUDA-core kernels (conv_simt.cu), the fp32 mode's default; *_mma = mma.sync
 * tf32 tensor-core tiles (mega.cu) as stand-alone launches — 3xTF32 error-compensated (fp32 accuracy) when desc.math ==
 * MFVI_MATH_FP32.  Measured slower than *_simt on the metric shape, hence not the default. */
int mfvi_conv2d_fwd_simt(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                         double* stats, mfvi_stream_t st);
int mfvi_conv2d_dgrad_simt(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                           mfvi_stream_t st);
int mfvi_conv2d_wgrad_simt(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                           mfvi_stream_t st);
int mfvi_conv2d_fwd_mma(const MfviConvDesc* d, MfviView x, const float* w, const float* bias, long long w_sstride, MfviView y,
                        double* stats, mfvi_stream_t st);
int mfvi_conv2d_dgrad_mma(const MfviConvDesc* d, MfviView dy, const float* w, long long w_sstride, MfviView dx, int accumulate,
                          mfvi_stream_t st);
int mfvi_conv2d_wgrad_mma(const MfviConvDesc* d, MfviView x, MfviView dy, float* dw, float* dbias, long long w_sstride,
                          mfvi_stream_t st);

/* ---- Persistent multi-stage execution of a sub-list of the plan (no reference counterpart: the reference launches one ATen
 * kernel per op; models/skip.py:58-132 is the sub-network concerned).  Between mfvi_mega_begin and mfvi_mega_end the entry points
 * mfvi_conv2d_{fwd,dgrad,wgrad}, mfvi_bn_act_pad_fwd, mfvi_cat_up_fwd, mfvi_pad_act_bwd, mfvi_bn_bwd_apply, mfvi_cat_up_bwd and
 * mfvi_fill_f32 RECORD a stage (same arguments, nothing is launched; thread-local).  mfvi_mega_mark_nosync declares the next
 * recorded stage independent of the one before it (no barrier in between).  mfvi_mega_end copies the program into caller-owned
 * device memory (n_stages * mfvi_mega_stage_bytes() bytes; synchronous copy at plan-build time).  mfvi_mega_run executes the whole
 * program for S MC samples in ONE launch: one 8-CTA thread-block cluster per sample, cluster barriers instead of kernel
 * boundaries (every recorded op must be per-sample independent, which all ops of the network are; a recorded mfvi_fill_f32
 * must cover S equal per-sample slices).  Convolution stages run on mma.sync tf32 tensor-core tiles sized for the <= 16x16
 * scales; with desc.math == MFVI_MATH_FP32 they use 3xTF32 error-compensated products.  A recorded mfvi_bn_bwd_apply does NOT
 * write dgamma / dbeta (they sum over all samples): call mfvi_bn_param_grads after the run with DEVICE tables of the n layers'
 * reduction tables (red[S][C][2]), dgamma / dbeta addresses (int64) and channel counts.
 * stage_times: NULL, or n_stages + 1 device int64 receiving %globaltimer (ns) at the start of every stage of sample 0 and at the
 * end, as seen by CTA 0 — a profiling aid (scripts/mega_profile.py). */
int mfvi_mega_begin(void);
int mfvi_mega_mark_nosync(void);
size_t mfvi_mega_stage_bytes(void);
int mfvi_mega_end(void* program_dev, size_t capacity_bytes, int* n_stages);
int mfvi_mega_run(const void* program_dev, int n_stages, int S, long long* stage_times, mfvi_stream_t st);
int mfvi_bn_param_grads(const long long* red_ptrs, const long long* dgamma_ptrs, const long long* dbeta_ptrs, const int* Cs, int n,
                        int S, mfvi_stream_t st);

/* Planning-only query (no reference counterpart; host-only, touches no device, works without a GPU): which kernel family
 * mfvi_conv2d_{fwd,dgrad,wgrad} would run for this geometry and these views — "pointwise", "halo", "alias", "tc" (tcgen05
 * paths) or "simt" (fp32 CUDA cores) — with its launch geometry and a one-line tile plan.  pass: 0 = forward (a = x, b = y),
 * 1 = data gradient (a = dy, b = dx), 2 = weight gradient (a = x, b = dy).  Only the alignment and strides of the views
 * are looked at.  tests/test_host_cpu.py pins the dispatch table of the four task networks with it. */
typedef struct {
  char family[24];
  unsigned grid[3];
  unsigned block;
  unsigned long long smem_bytes;
  int launches;      /* kernels one call enqueues (the convolution itself plus e.g. the bias-gradient reduction) */
  char detail[200];  /* tile plan of the convolution kernel, "key=value ..." */
} MfviPlanInfo;
int mfvi_conv2d_plan(const MfviConvDesc* d, int pass, MfviView a, MfviView b, long long w_sstride, int accumulate,
                     int with_bias, MfviPlanInfo* out);

/* ---- a7 + a8: tempered KL and the reparameterisation chain, one flat pass ----------------------------
 * VIModule._kl / kl_divergence (BayTorch/modules/module.py:64-80), MeanFieldVI.kl (freq_to_bayes.py:43-48)
 * and the autograd of  w = mu + softplus(rho)*eps  (module.py:82-85):
 *   kl_out[0]   += sum_i KL_i                        (double accumulator, caller zeroes it)
 *   grad_mu[i]   = [accum? grad_mu[i]:0]  + gscale*sum_s dw[s][i]           + kscale*dKL_i/dmu
 *   grad_rho[i]  = [accum? grad_rho[i]:0] + gscale*sigmoid(rho_i)*sum_s eps_s[i]*dw[s][i] + kscale*dKL_i/drho
 * direction 0 = reference 'reverse' = KL(prior || posterior); 1 = KL(posterior || prior).
 * dw == NULL (or S==0) skips the data term (pure KL forward/backward); grad_* == NULL skips gradients.
 * kscale_dev (optional, device): kscale is multiplied by *kscale_dev (upstream gradient of a kl() tensor). */
int mfvi_kl_reparam_fwd_bwd(const float* mu, const float* rho, size_t n, float prior_mu, double prior_sigma_plus_eps,
                            int direction, float kscale, const float* kscale_dev, const float* dw, long long dw_sstride, int S,
                            const float* eps, long long eps_sstride, MfviPhiloxKey key, float gscale,
                            double* kl_out, float* grad_mu, float* grad_rho, int accumulate, mfvi_stream_t st);

/* ---- a5: skip-net elementwise path (models/common.py:77-135, models/skip.py:68,102) ------------------
 * BatchNorm in training mode with per-sample statistics (the reference always has N=1), LeakyReLU(0.2),
 * ReflectionPad2d, bilinear/nearest x2 upsample, channel concat.  `sums` = double[S][C][2] (sum, sumsq)
 * produced by the conv epilogue / cat kernel; mean/invstd are derived on the fly (eps 1e-5, biased var). */
/* xp[s, reflect-padded by `pad`] = act(bn(y)) ; act: 0 none, 1 LeakyReLU(0.2); gamma==NULL -> identity BN. */
int mfvi_bn_act_pad_fwd(MfviView y, int S, int H, int W, int C, const double* sums, const float* gamma,
                        const float* beta, int act, int pad, MfviView xp, mfvi_stream_t st);
/* A = cat(lrelu(bn(ys)), up2x(lrelu(bn(yd)))) and sumsA += (sum A, sum A^2).  Cs may be 0 (no skip branch).
 * mode: 0 bilinear (align_corners=False), 1 nearest.  ys is (S,H,W,Cs), yd is (S,H/2,W/2,Cd). */
int mfvi_cat_up_fwd(MfviView ys, int Cs, const double* sums_s, const float* gamma_s, const float* beta_s,
                    MfviView yd, int Cd, const double* sums_d, const float* gamma_d, const float* beta_d,
                    int S, int H, int W, int mode, MfviView A, double* sumsA, mfvi_stream_t st);
/* g = fold_reflect(dxp) * act'(bn(y)) ; red[S][C][2] += (sum g, sum g*xhat). */
int mfvi_pad_act_bwd(MfviView dxp, int S, int H, int W, int C, int pad, MfviView y, const double* sums,
                     const float* gamma, const float* beta, int act, MfviView g, double* red, mfvi_stream_t st);
/* dy = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat))  (in place allowed: dy.ptr == g.ptr);
 * dgamma[c] = sum_s red[s][c][1], dbeta[c] = sum_s red[s][c][0] (written, not accumulated). */
int mfvi_bn_bwd_apply(MfviView g, MfviView y, int S, int H, int W, int C, const double* sums, const double* red,
                      const float* gamma, MfviView dy, float* dgamma, float* dbeta, mfvi_stream_t st);
/* backward of mfvi_cat_up_fwd w.r.t. the two activated branches, folded through their LeakyReLU (part 0: both, 1: only the
 * skip branch, 2: only the upsampled branch — the two are independent kernels and may run on different streams):
 *   gs = dA[:, :Cs] * lrelu'(bn(ys)), red_s += …;  gd = up2x^T(dA[:, Cs:]) * lrelu'(bn(yd)), red_d += … */
int mfvi_cat_up_bwd(MfviView dA, int S, int H, int W, int mode, MfviView ys, int Cs, const double* sums_s,
                    const float* gamma_s, const float* beta_s, MfviView gs, double* red_s, MfviView yd, int Cd,
                    const double* sums_d, const float* gamma_d, const float* beta_d, MfviView gd, double* red_d, int part,
                    mfvi_stream_t st);
/* running_mean/var update of every BatchNorm of the net in one launch (momentum 0.1, unbiased var), applied
 * once per MC sample in order, as S sequential reference forwards would. sums arena double[..], per-BN
 * tables on device: off[b] (channel offset into running arrays), sums_off[b] (offset into arena, in doubles),
 * C[b], count[b] (= H*W). */
int mfvi_bn_running_update(const double* arena, const int* ch_off, const long long* sums_off, const int* C,
                           const int* count, int n_bn, int S, float momentum, float* running_mean,
                           float* running_var, mfvi_stream_t st);

/* ---- a6: losses (utils/bayesian_utils.py:29-39; task variants bayesian_optimization.py:2095-2099,3033-3036)
 * out: (S,H,W,C) network output.  mode 0: gaussian_nll on (ch0=mu, ch1=s), every `sub`-th pixel (sub=1
 * denoising, sub=4 super-resolution, target is (H/sub, W/sub));  mode 1: inpainting, C=4, sigmoid on ch0..2,
 * ch3 = s, mask (H,W), target (H,W,3) NHWC;  mode 3: as mode 1 but ch0..2 already hold the sigmoid-ed means;
 * loss_out[0] += mean_s nll_s ; dout = d(mean_s nll_s)/dout.
 * mode 2: plain MSE between a(S,n) and target b(n) (CT sinogram space, bayesian_optimization.py:576). */
int mfvi_gauss_nll_fwd_bwd(int mode, MfviView out, int S, int H, int W, int C, int sub, const float* target,
                           const float* mask, double* loss_out, MfviView dout, mfvi_stream_t st);
int mfvi_mse_fwd_bwd(const float* a, long long a_sstride, const float* b, size_t n, int S, double* loss_out,
                     float* da, mfvi_stream_t st);

/* ---- a10: CT forward projector (radon/radon.py:32-55): img (S,H,W,C) view -> sino [S][C][T][W] -------- */
int mfvi_radon_fwd(MfviView img, int S, int C, int H, int W, const float* theta_rad, int T, float* sino,
                   mfvi_stream_t st);
int mfvi_radon_bwd(const float* dsino, int S, int C, int H, int W, const float* theta_rad, int T, MfviView dimg,
                   mfvi_stream_t st);

/* ---- a9: input jitter (bayesian_optimization.py:1363-1364) + reflection pad ---------------------------
 * xp[reflect-pad(saved + std*N(0,1))]; saved is NHWC (H,W,C); noise (optional, NHWC (H,W,C)) injects the
 * normals, otherwise Philox stream MFVI_STREAM_INPUT_JITTER indexed by the NCHW flat index. */
int mfvi_input_jitter_pad(const float* saved, const float* noise, int H, int W, int C, float std, int pad,
                          MfviPhiloxKey key, MfviView xp, mfvi_stream_t st);

/* ---- a8: optimiser (torch.optim.AdamW, bayesian_optimization.py:1357,1372) over one flat buffer --------
 * skip_if_nonfinite: device float* holding the step's loss nll + temp*kl; when it is NaN/Inf the update is skipped (CT runner
 * :577-582: `if not torch.isnan(loss): optimizer.step()`).  The flag is written by mfvi_loss_flag into a slot that rides at the
 * tail of the all-reduced gradient buffer, so every rank of a sharded run takes the same decision; the optimiser's own step
 * count (mfvi_counter_add_if_finite) does not advance on a skipped update, as torch's does not. */
int mfvi_adamw_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, const uint32_t* step_dev, const float* skip_if_nonfinite,
                    mfvi_stream_t st);
int mfvi_loss_flag(const double* kl, const double* nll, float temp, float* flag, mfvi_stream_t st);
int mfvi_counter_add_if_finite(uint32_t* ctr, uint32_t inc, const float* flag, mfvi_stream_t st);
/* *ctr += inc (one thread): the device-side step counter read through MfviPhiloxKey.step_dev / adamw step_dev
 * (effective AdamW step = step + *step_dev). */
int mfvi_counter_add(uint32_t* ctr, uint32_t inc, mfvi_stream_t st);

/* ---- f1: per-iteration bookkeeping of the runners (bayesian_optimization.py:1374-1416) ----------------------
 * One launch per iteration, no host synchronisation, CUDA-graph capturable.  it = iter_offset + *iter_dev.
 *   cur = (mean_s out[s,.,0], mean_s exp(-out[s,.,1]));  out_avg[2][H][W] = it==0 ? cur : out_avg*w + cur*(1-w);
 *   ring_epi[it % ring] = clip01(cur_mean), ring_ale[it % ring] = clip01(cur_var)   (ring may be 0);
 *   acc[0] += sum (noisy - clip(mean))^2     acc[1] += sum (gt - clip(mean))^2    acc[2] += sum (gt - clip(avg_mean))^2
 *   acc[3] += sum (noisy - avg_mean)^2       acc[4] += sum (gt - avg_mean)^2      (caller zeroes acc; gt/noisy optional)
 * PSNR = 10*log10(H*W/acc[k])  (utils/common_utils.py:297-305). */
int mfvi_bookkeep_step(MfviView out, int S, int H, int W, float exp_weight, const float* gt, const float* noisy,
                       float* out_avg, float* ring_epi, float* ring_ale, int ring, const uint32_t* iter_dev,
                       int iter_offset, double* acc, mfvi_stream_t st);
/* The same for the other runners: Cm image channels (1 or 3), flags bit 0 = sigmoid on the image channels (inpainting,
 * bayesian_optimization.py:3034), bit 1 = channel Cm is s = -log sigma^2 (absent in the CT net, :533); `mask` (H,W) or NULL
 * multiplies both images of the gt comparisons acc[1], acc[2] (inpainting, :3064-3065).  gt / noisy are (Cm,H,W); out_avg is
 * (Cm [+1],H,W); ring_epi is (Cm,ring,H,W) so that each channel's ring feeds mfvi_ring_uncertainty; ring_ale is (ring,H,W).
 * mfvi_bookkeep_step(...) == mfvi_bookkeep_step_ex(..., Cm = 1, flags = 2, mask = NULL, ...). */
int mfvi_bookkeep_step_ex(MfviView out, int S, int H, int W, int Cm, int flags, float exp_weight, const float* gt,
                          const float* noisy, const float* mask, float* out_avg, float* ring_epi, float* ring_ale, int ring,
                          const uint32_t* iter_dev, int iter_offset, double* acc, mfvi_stream_t st);
/* out_sum[0] += sum over pixels of the SSIM map of (a, b) — 11x11 Gaussian window sigma 1.5, zero padding
 * (utils/common_utils.py:308-353); clip_b != 0 clips b to [0,1] first.  SSIM = out_sum/(H*W). */
int mfvi_ssim(const float* a, const float* b, int H, int W, int clip_b, double* out_sum, mfvi_stream_t st);
/* epi = unbiased variance over the first n ring entries, ale = their mean variance (bayesian_optimization.py:1410-1411),
 * err2 (optional, needs gt) = mean_n (ring_epi - gt)^2 for the UCE calibration curve (eval_denoising.ipynb:467-482). */
int mfvi_ring_uncertainty(const float* ring_epi, const float* ring_ale, int n, int H, int W, const float* gt, float* epi,
                          float* ale, float* err2, mfvi_stream_t st);

/* ---- f3: local-reparameterisation layers (BayTorch/modules/reparam_layers.py:39-72) ------------------------------
 * LRTLayer.forward = conv(x, W_mu, b_mu) + sqrt(1e-16 + conv(x^2, softplus(W_rho)^2, softplus(b_rho)^2)) * eps.  The two
 * convolutions are mfvi_conv2d_{fwd,dgrad,wgrad} with one weight set for the whole batch (w_sstride 0); these are the
 * flat elementwise pieces and their backward chains (n contiguous floats each):
 *   softplus_sq_fwd : sigma2 = softplus(rho)^2           softplus_sq_bwd : drho (+)= dsigma2 * 2*softplus(rho)*sigmoid(rho)
 *   square_fwd      : x2 = x*x                           square_bwd      : dx   (+)= dx2 * 2*x
 *   lrt_noise_fwd   : out = act_mu + sqrt(1e-16+act_var)*eps
 *   lrt_noise_bwd   : dvar = dout * eps / (2*sqrt(1e-16+act_var))           (d act_mu = dout) */
int mfvi_softplus_sq_fwd(const float* rho, size_t n, float* sigma2, mfvi_stream_t st);
int mfvi_softplus_sq_bwd(const float* rho, const float* dsigma2, size_t n, float* drho, int accumulate, mfvi_stream_t st);
int mfvi_square_fwd(const float* x, size_t n, float* x2, mfvi_stream_t st);
int mfvi_square_bwd(const float* x, const float* dx2, size_t n, float* dx, int accumulate, mfvi_stream_t st);
int mfvi_lrt_noise_fwd(const float* act_mu, const float* act_var, const float* eps, size_t n, float* out, mfvi_stream_t st);
int mfvi_lrt_noise_bwd(const float* dout, const float* act_var, const float* eps, size_t n, float* dvar, mfvi_stream_t st);

/* small utilities used by the host side */
int mfvi_fill_f32(float* p, size_t n, float v, mfvi_stream_t st);
int mfvi_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, mfvi_stream_t st);
int mfvi_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, mfvi_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* MFVI_DIP_H_ */
