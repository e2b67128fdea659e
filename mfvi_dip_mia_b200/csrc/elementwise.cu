// Skip-net elementwise path (models/common.py:77-135, models/skip.py:68,102 of the reference):
// BatchNorm (train mode, per-sample statistics) + LeakyReLU(0.2) + ReflectionPad2d + x2 upsample + concat,
// forward and backward.  NHWC fp32, HBM-bound: each thread owns one channel quad of one pixel (float4 when
// C % 4 == 0), threads of a CTA are laid out [pixel-slot][channel-group] so that a warp touches contiguous
// memory; per-(sample,channel) reductions go registers -> shared -> one double atomic per channel per CTA.
// The kernel bodies live in elementwise_body.cuh (shared with the persistent multi-stage kernel of mega.cu).
#include "elementwise_body.cuh"
#include "mega.cuh"

namespace mfvi {

template <int V, bool OBF = false>       // OBF: xp is a bf16 view (strides in bf16 elements)
__global__ void __launch_bounds__(kEwThreads)
k_bn_act_pad_fwd(MfviView y, int H, int W, int C, const double* __restrict__ sums, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int act, int pad, MfviView xp, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ BnTable tab;
  body_bn_act_pad_fwd<V, OBF>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{nullptr, &tab, nullptr}, y, H, W, C, sums, gamma, beta, act, pad, xp, G, PPB);
}

template <int V>       // the same body behind the cp.async slots (kPipeBytes of dynamic shared memory)
__global__ void __launch_bounds__(kEwThreads)
k_bn_act_pad_fwd_p(MfviView y, int H, int W, int C, const double* __restrict__ sums, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int act, int pad, MfviView xp, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float ew_pipe[];
  __shared__ BnTable tab;
  body_bn_act_pad_fwd<V, false, true>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{nullptr, &tab, nullptr, ew_pipe}, y, H, W, C, sums, gamma, beta, act, pad, xp, G, PPB);
}

template <int V>
__global__ void __launch_bounds__(kEwThreads, 3)
k_cat_up_fwd(MfviView ys, int Cs, const double* __restrict__ sums_s, const float* __restrict__ gamma_s,
             const float* __restrict__ beta_s, MfviView yd, int Cd, const double* __restrict__ sums_d,
             const float* __restrict__ gamma_d, const float* __restrict__ beta_d, int H, int W, int mode, MfviView A,
             double* __restrict__ sumsA, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  __shared__ BnTable tab;
  body_cat_up_fwd<V>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{sm_red, &tab, nullptr}, ys, Cs, sums_s, gamma_s, beta_s, yd, Cd, sums_d, gamma_d, beta_d, H, W, mode, A, sumsA, G, PPB);
}

template <int V>
__global__ void __launch_bounds__(kEwThreads, 2)
k_pad_act_bwd(MfviView dxp, int H, int W, int C, int pad, MfviView y, const double* __restrict__ sums,
              const float* __restrict__ gamma, const float* __restrict__ beta, int act, MfviView g,
              double* __restrict__ red, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  __shared__ BnTable tab;
  body_pad_act_bwd<V>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{sm_red, &tab, nullptr}, dxp, H, W, C, pad, y, sums, gamma, beta, act, g, red, G, PPB);
}
template <int V>       // the same body behind the cp.async slots (kPipeBytes of dynamic shared memory)
__global__ void __launch_bounds__(kEwThreads, 3)
k_pad_act_bwd_p(MfviView dxp, int H, int W, int C, int pad, MfviView y, const double* __restrict__ sums,
                const float* __restrict__ gamma, const float* __restrict__ beta, int act, MfviView g,
                double* __restrict__ red, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float ew_pipe[];
  __shared__ double sm_red[2 * kEwThreads * 4];
  __shared__ BnTable tab;
  body_pad_act_bwd<V, true>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{sm_red, &tab, nullptr, ew_pipe}, dxp, H, W, C, pad, y, sums, gamma, beta, act, g, red, G, PPB);
}

template <int V, bool OBF = false>       // OBF: dy is a bf16 view (strides in bf16 elements)
__global__ void __launch_bounds__(kEwThreads)
k_bn_bwd_apply(MfviView g, MfviView y, int S, int H, int W, int C, const double* __restrict__ sums,
               const double* __restrict__ red, const float* __restrict__ gamma, MfviView dy,
               float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm_misc[5 * kMaxC];
  body_bn_bwd_apply<V, OBF>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{nullptr, nullptr, sm_misc}, g, y, S, H, W, C, sums, red, gamma, dy, dgamma, dbeta, G, PPB);
}

template <int V>       // the same body behind the cp.async slots
__global__ void __launch_bounds__(kEwThreads)
k_bn_bwd_apply_p(MfviView g, MfviView y, int S, int H, int W, int C, const double* __restrict__ sums,
                 const double* __restrict__ red, const float* __restrict__ gamma, MfviView dy,
                 float* __restrict__ dgamma, float* __restrict__ dbeta, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float ew_pipe[];
  __shared__ float sm_misc[5 * kMaxC];
  body_bn_bwd_apply<V, false, true>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{nullptr, nullptr, sm_misc, ew_pipe}, g, y, S, H, W, C, sums, red, gamma, dy, dgamma, dbeta, G, PPB);
}

template <int V>
__global__ void __launch_bounds__(kEwThreads, 3)
k_cat_bwd_skip(MfviView dA, int H, int W, MfviView ys, int Cs, const double* __restrict__ sums_s,
               const float* __restrict__ gamma_s, const float* __restrict__ beta_s, MfviView gs,
               double* __restrict__ red_s, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  __shared__ BnTable tab;
  body_cat_bwd_skip<V>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{sm_red, &tab, nullptr}, dA, H, W, ys, Cs, sums_s, gamma_s, beta_s, gs, red_s, G, PPB);
}

template <int V>
__global__ void __launch_bounds__(kEwThreads, 2)
k_cat_bwd_up(MfviView dA, int H, int W, int mode, int Cs, MfviView yd, int Cd, const double* __restrict__ sums_d,
             const float* __restrict__ gamma_d, const float* __restrict__ beta_d, MfviView gd,
             double* __restrict__ red_d, int G, int PPB) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sm_red[2 * kEwThreads * 4];
  __shared__ BnTable tab;
  body_cat_bwd_up<V>(VGrid{(int)blockIdx.x, (int)blockIdx.y, (int)gridDim.x}, EwSmem{sm_red, &tab, nullptr}, dA, H, W, mode, Cs, yd, Cd, sums_d, gamma_d, beta_d, gd, red_d, G, PPB);
}

// running stats of all BatchNorms (one thread per channel)
__global__ void k_bn_running(const double* __restrict__ arena, const int* __restrict__ ch_off,
                             const long long* __restrict__ sums_off, const int* __restrict__ Cs,
                             const int* __restrict__ count, int n_bn, int S, float momentum,
                             float* __restrict__ running_mean, float* __restrict__ running_var) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  if (b >= n_bn) return;
  const int C = Cs[b];
  const double n = (double)count[b];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float rm = running_mean[ch_off[b] + c], rv = running_var[ch_off[b] + c];
    for (int s = 0; s < S; ++s) {
      const double* sm = arena + sums_off[b] + ((size_t)s * C + c) * 2;
      const double mean = sm[0] / n;
      double var = sm[1] / n - mean * mean;
      var = var < 0.0 ? 0.0 : var;
      const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
      rm = (1.f - momentum) * rm + momentum * (float)mean;
      rv = (1.f - momentum) * rv + momentum * (float)unbiased;
    }
    running_mean[ch_off[b] + c] = rm;
    running_var[ch_off[b] + c] = rv;
  }
}

}  // namespace mfvi

using namespace mfvi;

#define MFVI_EW_DISPATCH_SMEM(GEOM, KERNEL, GRID, SMEM, ...)                                   \
  do {                                                                                         \
    if ((GEOM).V == 4)                                                                         \
      launch_k(KERNEL<4>, GRID, kEwThreads, SMEM, as_stream(st), __VA_ARGS__, (GEOM).G, (GEOM).PPB);  \
    else                                                                                       \
      launch_k(KERNEL<1>, GRID, kEwThreads, SMEM, as_stream(st), __VA_ARGS__, (GEOM).G, (GEOM).PPB);  \
  } while (0)
#define MFVI_EW_DISPATCH(GEOM, KERNEL, GRID, ...) MFVI_EW_DISPATCH_SMEM(GEOM, KERNEL, GRID, 0, __VA_ARGS__)

// resident CTAs per SM of one elementwise kernel (both vector widths), asked of the runtime once per process; a kernel with
// dynamic shared memory (the cp.async slots) is first allowed to exceed the 48 KB default together with its static arrays
template <typename K>
static void ew_allow_smem(K kernel, size_t dyn_smem, unsigned long long* done_mask) {      // once per device and kernel
  allow_dyn_smem(kernel, static_cast<int>(dyn_smem), done_mask);
}
template <typename K>
static int ew_occupancy_of(K kernel, size_t dyn_smem = 0) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kEwThreads, dyn_smem) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 1;
  }
  return n;
}
#define MFVI_EW_OCC_SMEM(GEOM, KERNEL, SMEM)                                                            \
  ([&]() -> int {                                                                                       \
    static unsigned long long allowed4 = 0, allowed1 = 0;                                               \
    if ((SMEM) > 0) {                                                                                   \
      if ((GEOM).V == 4) ew_allow_smem(KERNEL<4>, SMEM, &allowed4);                                     \
      else ew_allow_smem(KERNEL<1>, SMEM, &allowed1);                                                   \
    }                                                                                                   \
    static const int occ4 = ew_occupancy_of(KERNEL<4>, SMEM), occ1 = ew_occupancy_of(KERNEL<1>, SMEM);  \
    return (GEOM).V == 4 ? occ4 : occ1;                                                                 \
  }())
#define MFVI_EW_OCC(GEOM, KERNEL) MFVI_EW_OCC_SMEM(GEOM, KERNEL, 0)
// The cp.async kernels pay a longer prologue (D stages issued before the first pixel is used), so a launch takes them only
// when its threads sweep at least kPipeMinIter pixels; below that the register-batched kernels are quicker
// (profiles/r02_ew_load_batching.txt).  MFVI_EW_PIPE = bit mask of the kernels allowed to (1 pad_act_bwd, 2 bn_bwd_apply,
// 4 bn_act_pad_fwd; measurement knob), MFVI_EW_PIPE_MINITER overrides the threshold.
constexpr int kPipeMinIter = 8;
static inline bool ew_piped(int bit, int npix, int PPB, int gx) {
  static const int mask = [] {
    const char* e = getenv("MFVI_EW_PIPE");
    return e == nullptr ? 7 : atoi(e);
  }();
  static const int miniter = ew_knob("MFVI_EW_PIPE_MINITER", kPipeMinIter);
  if (!(mask & bit)) return false;
  const int blocks = (npix + PPB - 1) / PPB;
  return blocks >= miniter * gx;
}

extern "C" {

int mfvi_bn_act_pad_fwd(MfviView y, int S, int H, int W, int C, const double* sums, const float* gamma,
                        const float* beta, int act, int pad, MfviView xp, mfvi_stream_t st) {
  MFVI_REQUIRE(y.ptr && xp.ptr, "bn_act_pad_fwd: null pointer");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "bn_act_pad_fwd: C=%d out of range (1..%d)", C, kMaxC);
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "bn_act_pad_fwd: pad must be smaller than the image");
  const EwGeom ge = ew_geom(C, view_vec_ok(y) && view_vec_ok(xp));
  MFVI_REQUIRE(ge.G <= kEwThreads, "bn_act_pad_fwd: too many channel groups");
  const int npix_p = (H + 2 * pad) * (W + 2 * pad);
  dim3 grid(ew_grid(npix_p, ge.PPB, S, MFVI_EW_OCC_SMEM(ge, k_bn_act_pad_fwd_p, kPipeBytes)), S);
  const bool piped = ew_piped(4, npix_p, ge.PPB, grid.x);
  if (!piped) grid.x = ew_grid(npix_p, ge.PPB, S, MFVI_EW_OCC(ge, k_bn_act_pad_fwd));
  if (mega::Stage* ms = mega::append(mega::OP_BN_ACT_PAD_FWD)) {
    ms->V = ge.V; ms->G = ge.G; ms->PPB = ge.PPB; ms->gx = grid.x < 8u ? grid.x : 8u;      /* one virtual block per CTA of the sample's cluster */ ms->S = S;
    ms->a = y; ms->b = xp; ms->H = H; ms->W = W; ms->C = C; ms->sums = sums; ms->gamma = gamma; ms->beta = beta; ms->act = act;
    ms->pad = pad;
    return 0;
  }
  if (piped)
    MFVI_EW_DISPATCH_SMEM(ge, k_bn_act_pad_fwd_p, grid, kPipeBytes, y, H, W, C, sums, gamma, beta, act, pad, xp);
  else
    MFVI_EW_DISPATCH(ge, k_bn_act_pad_fwd, grid, y, H, W, C, sums, gamma, beta, act, pad, xp);
  return check_launch("bn_act_pad_fwd");
}

int mfvi_cat_up_fwd(MfviView ys, int Cs, const double* sums_s, const float* gamma_s, const float* beta_s,
                    MfviView yd, int Cd, const double* sums_d, const float* gamma_d, const float* beta_d,
                    int S, int H, int W, int mode, MfviView A, double* sumsA, mfvi_stream_t st) {
  MFVI_REQUIRE(yd.ptr && A.ptr && sumsA, "cat_up_fwd: null pointer");
  MFVI_REQUIRE(Cs == 0 || ys.ptr, "cat_up_fwd: null skip branch");
  MFVI_REQUIRE(H % 2 == 0 && W % 2 == 0, "cat_up_fwd: H,W must be even (centre-crop concat is not supported)");
  MFVI_REQUIRE(Cs + Cd <= kMaxC && Cd >= 1, "cat_up_fwd: channel count out of range");
  MFVI_REQUIRE(mode == 0 || mode == 1, "cat_up_fwd: mode must be 0 (bilinear) or 1 (nearest)");
  const bool al = view_vec_ok(yd) && view_vec_ok(A) && (Cs == 0 || view_vec_ok(ys)) && Cs % 4 == 0 && Cd % 4 == 0;
  const EwGeom ge = ew_geom(Cs + Cd, al);
  MFVI_REQUIRE(ge.G <= kEwThreads, "cat_up_fwd: too many channel groups");
  dim3 grid(ew_grid((H / 2 + 1) * (W / 2 + 1), ge.PPB, S, MFVI_EW_OCC(ge, k_cat_up_fwd), true), S);
  if (mega::Stage* ms = mega::append(mega::OP_CAT_UP_FWD)) {
    ms->V = ge.V; ms->G = ge.G; ms->PPB = ge.PPB; ms->gx = grid.x < 8u ? grid.x : 8u;      /* one virtual block per CTA of the sample's cluster */ ms->S = S;
    ms->a = ys; ms->C = Cs; ms->sums = sums_s; ms->gamma = gamma_s; ms->beta = beta_s;
    ms->b = yd; ms->C2 = Cd; ms->sums2 = sums_d; ms->gamma2 = gamma_d; ms->beta2 = beta_d;
    ms->H = H; ms->W = W; ms->mode = mode; ms->c = A; ms->red = sumsA;
    return 0;
  }
  MFVI_EW_DISPATCH(ge, k_cat_up_fwd, grid, ys, Cs, sums_s, gamma_s, beta_s, yd, Cd, sums_d, gamma_d, beta_d, H, W, mode,
                   A, sumsA);
  return check_launch("cat_up_fwd");
}

int mfvi_pad_act_bwd(MfviView dxp, int S, int H, int W, int C, int pad, MfviView y, const double* sums,
                     const float* gamma, const float* beta, int act, MfviView g, double* red, mfvi_stream_t st) {
  MFVI_REQUIRE(dxp.ptr && y.ptr && g.ptr && red && sums, "pad_act_bwd: null pointer");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "pad_act_bwd: C out of range");
  MFVI_REQUIRE(pad >= 0 && pad < H && pad < W, "pad_act_bwd: pad must be smaller than the image");
  const EwGeom ge = ew_geom(C, view_vec_ok(dxp) && view_vec_ok(y) && view_vec_ok(g));
  MFVI_REQUIRE(ge.G <= kEwThreads, "pad_act_bwd: too many channel groups");
  const int occ_p = MFVI_EW_OCC_SMEM(ge, k_pad_act_bwd_p, kPipeBytes);
  dim3 grid(ew_grid(H * W, ge.PPB, S, occ_p, true), S);
  const bool piped = ew_piped(1, H * W, ge.PPB, grid.x);
  if (!piped) grid.x = ew_grid(H * W, ge.PPB, S, MFVI_EW_OCC(ge, k_pad_act_bwd), true);
  if (mega::Stage* ms = mega::append(mega::OP_PAD_ACT_BWD)) {
    ms->V = ge.V; ms->G = ge.G; ms->PPB = ge.PPB; ms->gx = grid.x < 8u ? grid.x : 8u;      /* one virtual block per CTA of the sample's cluster */ ms->S = S;
    ms->a = dxp; ms->b = y; ms->c = g; ms->H = H; ms->W = W; ms->C = C; ms->pad = pad; ms->sums = sums; ms->gamma = gamma;
    ms->beta = beta; ms->act = act; ms->red = red;
    return 0;
  }
  if (piped)
    MFVI_EW_DISPATCH_SMEM(ge, k_pad_act_bwd_p, grid, kPipeBytes, dxp, H, W, C, pad, y, sums, gamma, beta, act, g, red);
  else
    MFVI_EW_DISPATCH(ge, k_pad_act_bwd, grid, dxp, H, W, C, pad, y, sums, gamma, beta, act, g, red);
  return check_launch("pad_act_bwd");
}

int mfvi_bn_bwd_apply(MfviView g, MfviView y, int S, int H, int W, int C, const double* sums, const double* red,
                      const float* gamma, MfviView dy, float* dgamma, float* dbeta, mfvi_stream_t st) {
  MFVI_REQUIRE(g.ptr && y.ptr && dy.ptr && sums && red, "bn_bwd_apply: null pointer");
  MFVI_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "bn_bwd_apply: dgamma/dbeta must both be set or NULL");
  MFVI_REQUIRE(C >= 1 && C <= kMaxC, "bn_bwd_apply: C out of range");
  const EwGeom ge = ew_geom(C, view_vec_ok(g) && view_vec_ok(y) && view_vec_ok(dy));
  MFVI_REQUIRE(ge.G <= kEwThreads, "bn_bwd_apply: too many channel groups");
  dim3 grid(ew_grid(H * W, ge.PPB, S, MFVI_EW_OCC_SMEM(ge, k_bn_bwd_apply_p, kPipeBytes)), S);
  const bool piped = ew_piped(2, H * W, ge.PPB, grid.x);
  if (!piped) grid.x = ew_grid(H * W, ge.PPB, S, MFVI_EW_OCC(ge, k_bn_bwd_apply));
  if (mega::Stage* ms = mega::append(mega::OP_BN_BWD_APPLY)) {
    ms->V = ge.V; ms->G = ge.G; ms->PPB = ge.PPB; ms->gx = grid.x < 8u ? grid.x : 8u;      /* one virtual block per CTA of the sample's cluster */ ms->S = S;
    ms->a = g; ms->b = y; ms->c = dy; ms->H = H; ms->W = W; ms->C = C; ms->sums = sums; ms->red = const_cast<double*>(red);
    ms->gamma = gamma; ms->dgamma = dgamma; ms->dbeta = dbeta;
    return 0;
  }
  if (piped)
    MFVI_EW_DISPATCH_SMEM(ge, k_bn_bwd_apply_p, grid, kPipeBytes, g, y, S, H, W, C, sums, red, gamma, dy, dgamma, dbeta);
  else
    MFVI_EW_DISPATCH(ge, k_bn_bwd_apply, grid, g, y, S, H, W, C, sums, red, gamma, dy, dgamma, dbeta);
  return check_launch("bn_bwd_apply");
}

int mfvi_cat_up_bwd(MfviView dA, int S, int H, int W, int mode, MfviView ys, int Cs, const double* sums_s,
                    const float* gamma_s, const float* beta_s, MfviView gs, double* red_s, MfviView yd, int Cd,
                    const double* sums_d, const float* gamma_d, const float* beta_d, MfviView gd, double* red_d, int part,
                    mfvi_stream_t st) {
  MFVI_REQUIRE(part >= 0 && part <= 2, "cat_up_bwd: part must be 0 (both branches), 1 (skip branch) or 2 (upsampled branch)");
  MFVI_REQUIRE(dA.ptr, "cat_up_bwd: null pointer");
  MFVI_REQUIRE(H % 2 == 0 && W % 2 == 0, "cat_up_bwd: H,W must be even");
  MFVI_REQUIRE(Cs + Cd <= kMaxC && Cd >= 1, "cat_up_bwd: channel count out of range");
  if (Cs > 0 && part != 2) {
    MFVI_REQUIRE(ys.ptr && gs.ptr && red_s, "cat_up_bwd: null skip branch");
    const EwGeom ge = ew_geom(Cs, view_vec_ok(dA) && view_vec_ok(ys) && view_vec_ok(gs));
    dim3 grid(ew_grid(H * W, ge.PPB, S, MFVI_EW_OCC(ge, k_cat_bwd_skip), true), S);
    if (mega::Stage* ms = mega::append(mega::OP_CAT_BWD_SKIP)) {
      ms->V = ge.V; ms->G = ge.G; ms->PPB = ge.PPB; ms->gx = grid.x < 8u ? grid.x : 8u;      /* one virtual block per CTA of the sample's cluster */ ms->S = S;
      ms->a = dA; ms->b = ys; ms->c = gs; ms->H = H; ms->W = W; ms->C = Cs; ms->sums = sums_s; ms->gamma = gamma_s;
      ms->beta = beta_s; ms->red = red_s;
    } else {
      MFVI_EW_DISPATCH(ge, k_cat_bwd_skip, grid, dA, H, W, ys, Cs, sums_s, gamma_s, beta_s, gs, red_s);
      if (int rc = check_launch("cat_up_bwd(skip)")) return rc;
    }
  }
  if (part == 1) return 0;
  MFVI_REQUIRE(yd.ptr && gd.ptr && red_d, "cat_up_bwd: null upsampled branch");
  const EwGeom ge = ew_geom(Cd, view_vec_ok(dA) && view_vec_ok(yd) && view_vec_ok(gd) && Cs % 4 == 0);
  MFVI_REQUIRE(ge.G <= kEwThreads, "cat_up_bwd: too many channel groups");
  dim3 grid(ew_grid(((H / 2 + 1) / 2) * ((W / 2 + 1) / 2), ge.PPB, S, MFVI_EW_OCC(ge, k_cat_bwd_up), true), S);      // 2x2 low-res blocks
  if (mega::Stage* ms = mega::append(mega::OP_CAT_BWD_UP)) {
    ms->V = ge.V; ms->G = ge.G; ms->PPB = ge.PPB; ms->gx = grid.x < 8u ? grid.x : 8u;      /* one virtual block per CTA of the sample's cluster */ ms->S = S;
    ms->a = dA; ms->b = yd; ms->c = gd; ms->H = H; ms->W = W; ms->mode = mode; ms->C = Cs; ms->C2 = Cd; ms->sums = sums_d;
    ms->gamma = gamma_d; ms->beta = beta_d; ms->red = red_d;
    return 0;
  }
  MFVI_EW_DISPATCH(ge, k_cat_bwd_up, grid, dA, H, W, mode, Cs, yd, Cd, sums_d, gamma_d, beta_d, gd, red_d);
  return check_launch("cat_up_bwd(up)");
}

int mfvi_bn_running_update(const double* arena, const int* ch_off, const long long* sums_off, const int* C,
                           const int* count, int n_bn, int S, float momentum, float* running_mean,
                           float* running_var, mfvi_stream_t st) {
  MFVI_REQUIRE(arena && ch_off && sums_off && C && count && running_mean && running_var, "bn_running_update: null pointer");
  if (n_bn == 0) return 0;
  launch_k(k_bn_running, n_bn, 128, 0, as_stream(st), arena, ch_off, sums_off, C, count, n_bn, S, momentum, running_mean,
                                                 running_var);
  return check_launch("bn_running_update");
}

}  // extern "C"
