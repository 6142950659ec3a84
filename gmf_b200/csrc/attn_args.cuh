// Kernel arguments of the fusion cross-attention (fusion_layer.py:82-94): operands arrive as bf16 tiles already in the
// 128B-swizzled K-major image (common.cuh) written by the projection epilogues; 1/sqrt(64) * log2(e) is folded into the Q projection.
#pragma once
#include "common.cuh"

namespace gmf {

struct AttnArgs {
  const __nv_bfloat16* q_t;   // [pairs][q_tiles][128*D]
  const __nv_bfloat16* k_t;   // [pairs][k_tiles][128*D]
  const __nv_bfloat16* vt_t;  // [pairs][k_tiles][D*128]
  float* out;                 // [pairs][Lq][D] fp32
  int Lq, Lk, q_tiles, k_tiles;   // tiles of 128 rows
  // fused to_out (fusion_layer.py:94) + residual.  wo_packed != NULL switches it on:
  //   xout[B, Lq, 128] = softmax(..) v . Wo^T + bo + resid        (`out` is then unused)
  const float* wo_packed;     // [128 x 64] tf32, swizzled K-major image (pack_linear(wo, 128, 64, 64, 128))
  const float* bo;            // [128]
  const float* resid;         // [B, Lq, 128]
  float* xout;                // [B, Lq, 128]
};

}  // namespace gmf
