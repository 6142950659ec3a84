// Classical spectral-matching baseline `SM` (GMF_PointDSC/baseline_scripts/baseline_3DMatch.py:19-53), SURVEY.md section 8f N4: a by-product of the
// on-the-fly length-consistency machinery at full N x N.
//     M_ij = max(0, 4.5 - (|s_i - s_j| - |t_i - t_j|)^2 / (2 sigma^2)),  sigma = inlier_threshold / 3,  M_ii = 0          (:21-35)
//     v <- M v / (|M v| + 1e-6), 10 times from v = 1                                                                      (:38-42)
//     labels = top int(N * top_ratio) of v; pose = rigid_transform_3d(src, tgt, v * labels)                               (:45-52)
// The N x N matrix is never stored: every power iteration recomputes M_ij from the points (2 MUFU.SQRT per pair) inside a fused mat-vec.
#pragma once
#include "common.cuh"

namespace gmf {

// y_i = sum_j M_ij v_j.  CTA = 128 rows x 4 column lanes (512 threads); the j tile (points + v) is staged in shared memory; the four partial sums of
// a row are combined with shuffles in a fixed order (deterministic).
__global__ void __launch_bounds__(512) sm_matvec_kernel(const float4* __restrict__ src4, const float4* __restrict__ tgt4, const float* __restrict__ v,
                                                        int N, float inv_2s2, float* __restrict__ y) {
  __shared__ float4 ss[128], st[128];
  __shared__ float sv[128];
  const int pair = blockIdx.y;
  const int row = blockIdx.x * 128 + (threadIdx.x >> 2), part = threadIdx.x & 3;
  const float4* S = src4 + (size_t)pair * N;
  const float4* T = tgt4 + (size_t)pair * N;
  const float* V = v ? v + (size_t)pair * N : nullptr;
  float4 si = make_float4(0.f, 0.f, 0.f, 0.f), ti = si;
  if (row < N) { si = S[row]; ti = T[row]; }
  float acc = 0.f;
  for (int j0 = 0; j0 < N; j0 += 128) {
    __syncthreads();
    if (threadIdx.x < 128) {
      const int j = j0 + threadIdx.x;
      ss[threadIdx.x] = j < N ? S[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      st[threadIdx.x] = j < N ? T[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      sv[threadIdx.x] = j < N ? (V ? V[j] : 1.0f) : 0.f;          // v == nullptr: first iteration, v = 1; padding columns contribute 0
    }
    __syncthreads();
#pragma unroll 8
    for (int jj = part; jj < 128; jj += 4) {
      const float4 a = ss[jj], b = st[jj];
      const float dx = si.x - a.x, dy = si.y - a.y, dz = si.z - a.z;
      const float ex = ti.x - b.x, ey = ti.y - b.y, ez = ti.z - b.z;
      const float ds = sqrt_approx(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      const float dt = sqrt_approx(fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
      const float d = ds - dt;
      float m = fmaxf(fmaf(-d * d, inv_2s2, 4.5f), 0.f);
      if (j0 + jj == row) m = 0.f;                                  // zero diagonal (:35)
      acc = fmaf(m, sv[jj], acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (part == 0 && row < N) y[(size_t)pair * N + row] = acc;
}

// v = y / (|y| + 1e-6) per pair (one CTA per pair, fixed-order block reduction)
__global__ void __launch_bounds__(1024) sm_normalize_kernel(const float* __restrict__ y, int N, float* __restrict__ v) {
  __shared__ float red[32];
  const int pair = blockIdx.x;
  const float* Y = y + (size_t)pair * N;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s = fmaf(Y[i], Y[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = 1.0f / (sqrtf(t) + 1e-6f);
  }
  __syncthreads();
  const float inv = red[0];
  for (int i = threadIdx.x; i < N; i += blockDim.x) v[(size_t)pair * N + i] = Y[i] * inv;
}

// labels = 1 at the top-S indices, weights = v * labels (:45-49)
__global__ void sm_labels_kernel(const float* __restrict__ v, const int* __restrict__ top, int N, int S, float* __restrict__ labels, float* __restrict__ weights) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { labels[(size_t)pair * N + i] = 0.f; weights[(size_t)pair * N + i] = 0.f; }
}
__global__ void sm_scatter_kernel(const float* __restrict__ v, const int* __restrict__ top, int N, int S, float* __restrict__ labels, float* __restrict__ weights) {
  const int pair = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) {
    const int i = top[(size_t)pair * S + s];
    labels[(size_t)pair * N + i] = 1.f;
    weights[(size_t)pair * N + i] = v[(size_t)pair * N + i];
  }
}

}  // namespace gmf
