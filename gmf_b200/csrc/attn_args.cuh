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
  // fused query side (fusion_layer.py:172-183).  xq != NULL switches it on: the kernel itself computes
  //   x0 = xq + dwconv(xq) (conditional position encoding, k = 3 along the row axis; skipped when cpe_w == NULL),  Q = LN_q(x0) . Wq^T
  // for its two row tiles (q_t unused).  With the fused output projection it also writes x0 to `x0` (the caller passes resid == x0, or
  // resid == xq and x0 == NULL when there is no position encoding): the rows come back out of L2 in the epilogue of the same CTA.
  const float* xq;            // [B, Lq, 128]
  float* x0;                  // [B, Lq, 128] or NULL
#ifdef GMF_FFN_TRACE
  long long* trace;           // development build: clock64 stamps of one CTA (tools/fa_trace.py)
#endif
  const float* cpe_w;         // [128][3] or NULL
  const float* cpe_b;         // [128]
  const float* lnq_g;
  const float* lnq_b;
  const float* wq16;          // to_q.weight x (log2 e / 8) as one fp16 image: 2 atoms of [64 rows x 64 k] (pack_linear_f16(wq, 64, 128, 64))
};

}  // namespace gmf
