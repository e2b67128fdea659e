// Persistent multi-stage kernel for the coarse scales of the hour-glass net (mega.cu): program representation and the
// recording hook the stand-alone entry points call.
//
// A "program" is a list of stages; each stage is one operation of the engine's plan (a convolution pass or an elementwise
// kernel) with the arguments the stand-alone entry point would have launched it with.  While recording is on
// (mfvi_mega_begin .. mfvi_mega_end) the C-ABI entry points append a stage instead of launching, so the host replays the very
// same op list it would otherwise launch kernel by kernel; mfvi_mega_run then executes the whole list in ONE launch, with grid
// barriers where kernel boundaries used to be.
#pragma once
#include "elementwise.cuh"

namespace mfvi {
namespace mega {

enum Op : int {
  OP_CONV_FWD = 1,
  OP_CONV_DGRAD = 2,
  OP_CONV_WGRAD = 3,
  OP_BN_ACT_PAD_FWD = 4,
  OP_CAT_UP_FWD = 5,
  OP_PAD_ACT_BWD = 6,
  OP_BN_BWD_APPLY = 7,
  OP_CAT_BWD_SKIP = 8,
  OP_CAT_BWD_UP = 9,
  OP_FILL = 10,
};

struct Stage {
  int op;
  int nosync;      // 1: independent of the stage before it — no grid barrier in between
  // ---- elementwise stages: geometry of the stand-alone kernel (ew_geom / ew_grid) and its arguments
  int V, G, PPB, gx, S;
  int H, W, C, C2, pad, act, mode;
  MfviView a, b, c;
  const double* sums;
  const double* sums2;
  double* red;
  const float* gamma;
  const float* beta;
  const float* gamma2;
  const float* beta2;
  float* dgamma;
  float* dbeta;
  // ---- convolution stages
  MfviConvDesc d;
  const float* w;           // sampled weights of the layer, sample stride w_sstride
  const float* bias;
  long long w_sstride;
  float* dw;                // weight-gradient block of the layer (wgrad), sample stride w_sstride
  double* stats;            // forward: per-sample (sum, sumsq) of the output channels, or NULL
  int accumulate;           // dgrad: dx += instead of dx =
  int nt;                   // 8-column MMA tiles per warp: the CTA tile is 64 x (16 * nt)
  int m_tiles, n_tiles, k_splits, k_len, items;
  int split3;               // 3xTF32 error-compensated products (the exact-fp32 mode)
  int vec_a, vec_b;         // operand rows may be read as float4
  // ---- fill
  float* fill_ptr;
  unsigned long long fill_n;
  float fill_v;
};

bool recording();
// a zero-initialised stage appended to the program being recorded (NULL when not recording)
Stage* append(int op);

}  // namespace mega
}  // namespace mfvi
