// Shared pieces of the spatial-consistency attention path: kernel arguments, the bf16 3-term coordinate split and the per-point
// distance-feature rows that turn  |s_i - s_j|^2 / sigma^2  and  1 - |t_i - t_j|^2 / sigma^2  into K = 32 bf16 MMAs, and the FMA-pipe
// exp2 used by the (optional) polynomial softmax variants.
#pragma once
#include "common.cuh"

namespace gmf {

struct ScAttnArgs {
  const __nv_bfloat16* q_t;   // [pairs][tiles][128*128]   (scale * log2e folded into the projection)
  const __nv_bfloat16* k_t;   // [pairs][tiles][128*128]
  const __nv_bfloat16* vt_t;  // [pairs][tiles][128*128]   V^T tiles
  const __nv_bfloat16* aq_t;  // [pairs][tiles][128*64]    query-side distance features (s-part | t-part)
  const __nv_bfloat16* bd_t;  // [pairs][tiles][128*64]    key-side distance features
  float* out;                 // [pairs][N][128] fp32
  int N, tiles;               // keys (and queries unless Nq is set)
  // gen 9 only: fused head of fc_message (PointDSC.py:13-21,65).  fc1_w != NULL switches it on: instead of msg the kernel
  // writes m2 = ReLU(BN(conv64x64(ReLU(BN(conv128x64(msg))))))  [pairs][N][64]  (BN folded into the packed weights / biases)
  const float* fc1_w;         // pack_linear(W, 64, 128, 32, 64)
  const float* fc1_b;
  const float* fc2_w;         // pack_linear(W, 64, 64, 64, 64)
  const float* fc2_b;
  float* m2_out;
  // cross-attention use (DGR head): Nq != 0 gives the query side its own length / tile count (q_t, aq_t, out are indexed with it)
  int Nq, q_tiles;
  // split-key mode (SPLIT instantiation, one pair): CTA (q tile, split) takes key tiles [split * tiles_per_split, ...) and writes the
  // UNnormalised partial output part_o [splits][Nq][128] plus (row sum, softmax reference in log2 units) part_l [splits][Nq][2]; the
  // consumer combines the splits (flash-decoding style)
  int tiles_per_split;
  float* part_o;
  float* part_l;
#ifdef GMF_SC_TRACE
  long long* trace;           // development build only: clock64 timeline of one CTA, [role 0..6][tile 0..63][stamp 0..7]
#endif
};


__device__ __forceinline__ void split3(float x, __nv_bfloat16& h0, __nv_bfloat16& h1, __nv_bfloat16& h2) {
  h0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h0);
  h1 = __float2bfloat16_rn(r1);
  h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
}


// 2^x on the FMA/ALU pipes for x <= ~100: round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax of 2^f, exponent add.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float fl = x + 12582912.f;                         // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (fl - 12582912.f);
  float p = fmaf(0.0551716648f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(fl) << 23));
}


// distance-feature tiles: coordinates are pre-divided by sigma_d and the t-part carries a constant so that the accumulators are
//   DA = |ds|^2 / sigma^2  and  Y = |dt|^2 / sigma^2 - 1   (the softmax needs DA (Y + 1) and DA + Y: no negations, packed fp32x2 friendly).
//   A (query side): per coord (-2u0,-2u0,-2u1,-2u1,-2u0,-2u2), |u|^2 split (n0,n1,n2), (1,1,1) [, -1 in the t-part]
//   B (key side):   per coord (  w0,  w1,  w0,  w1,  w2,  w0), (1,1,1), |w|^2 split             [,  1 in the t-part]
__global__ void dist_feature_scaled_kernel(const float* __restrict__ kpts, int Np, float inv_sigma, __nv_bfloat16* __restrict__ aq_t,
                                           __nv_bfloat16* __restrict__ bd_t) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // padded point index
  if (i >= Np) return;
  const float4 s4 = *reinterpret_cast<const float4*>(kpts + ((size_t)pair * Np + i) * 8);
  const float4 t4 = *reinterpret_cast<const float4*>(kpts + ((size_t)pair * Np + i) * 8 + 4);
  __align__(16) __nv_bfloat16 A[64];
  __align__(16) __nv_bfloat16 B[64];
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f), mone = __float2bfloat16_rn(-1.f);
#pragma unroll
  for (int k = 0; k < 64; ++k) { A[k] = zero; B[k] = zero; }
  float pts[2][4] = {{s4.x * inv_sigma, s4.y * inv_sigma, s4.z * inv_sigma, 0.f}, {t4.x * inv_sigma, t4.y * inv_sigma, t4.z * inv_sigma, 0.f}};
#pragma unroll
  for (int part = 0; part < 2; ++part) pts[part][3] = fmaf(pts[part][0], pts[part][0], fmaf(pts[part][1], pts[part][1], pts[part][2] * pts[part][2]));
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const int o = part * 32;
    const float sg = -2.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      __nv_bfloat16 h0, h1, h2;
      split3(pts[part][c], h0, h1, h2);
      const __nv_bfloat16 m0 = __float2bfloat16_rn(sg * __bfloat162float(h0)), m1 = __float2bfloat16_rn(sg * __bfloat162float(h1)),
                          m2 = __float2bfloat16_rn(sg * __bfloat162float(h2));
      A[o + 6 * c + 0] = m0; B[o + 6 * c + 0] = h0;
      A[o + 6 * c + 1] = m0; B[o + 6 * c + 1] = h1;
      A[o + 6 * c + 2] = m1; B[o + 6 * c + 2] = h0;
      A[o + 6 * c + 3] = m1; B[o + 6 * c + 3] = h1;
      A[o + 6 * c + 4] = m0; B[o + 6 * c + 4] = h2;
      A[o + 6 * c + 5] = m2; B[o + 6 * c + 5] = h0;
    }
    __nv_bfloat16 n0, n1, n2;
    split3(pts[part][3], n0, n1, n2);                                   // A side: |u|^2
    __nv_bfloat16 w0, w1, w2;
    split3(pts[part][3], w0, w1, w2);                                   // B side: |w|^2, multiplied by +-1 from the A side
    const __nv_bfloat16 sgn1 = one;
    A[o + 18] = n0; A[o + 19] = n1; A[o + 20] = n2; B[o + 18] = one; B[o + 19] = one; B[o + 20] = one;
    A[o + 21] = sgn1; A[o + 22] = sgn1; A[o + 23] = sgn1; B[o + 21] = w0; B[o + 22] = w1; B[o + 23] = w2;
    if (part) { A[o + 24] = mone; B[o + 24] = one; }
  }
  const int tile = i >> 7, r = i & 127;
  const size_t tbase = ((size_t)pair * (Np >> 7) + tile) * (128 * 64);
  uint8_t* ad = (uint8_t*)(aq_t + tbase);
  uint8_t* bd = (uint8_t*)(bd_t + tbase);
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    *reinterpret_cast<uint4*>(ad + swz_off(r, ch)) = *reinterpret_cast<const uint4*>(&A[ch * 8]);
    *reinterpret_cast<uint4*>(bd + swz_off(r, ch)) = *reinterpret_cast<const uint4*>(&B[ch * 8]);
  }
}


}  // namespace gmf
