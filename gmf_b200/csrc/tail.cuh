// SIMT kernels of the GMF-PointDSC path that are HBM/latency-bound rather than tensor-bound:
// point preparation, layer0, classifier+normalise, NMS seed picking, seed kNN, per-seed spectral matching
// (40x40 compatibility + power iteration), per-seed weighted Kabsch, hypothesis scoring and post-refinement.
// Reference: GMF_PointDSC/models/PointDSC.py:229-528, models/common.py:10-75, utils/SE3.py:43-96.
#pragma once
#include "common.cuh"

namespace gmf {

// ------------------------------------------------------------------------------------------------
// 3x3 helpers (fp64 on one thread; the work is O(100) flops per seed)
// ------------------------------------------------------------------------------------------------
// Rotation of the weighted Kabsch problem from the covariance H = sum w (a-ca)(b-cb)^T  (common.py:33-46):
// with H = U S V^T,  R = V diag(1,1,det(V U^T)) U^T  ==  v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T  (sign-convention free).
__device__ inline void kabsch_rotation(const double H[9], double R[9]) {
  // A = H^T H, symmetric; cyclic Jacobi for eigenvectors V
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i * 3 + j] = H[0 * 3 + i] * H[0 * 3 + j] + H[1 * 3 + i] * H[1 * 3 + j] + H[2 * 3 + i] * H[2 * 3 + j];
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
    const double dia = fabs(A[0]) + fabs(A[4]) + fabs(A[8]);
    if (off <= 1e-30 || off <= 1e-17 * dia) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq;
          A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk;
          A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq;
          V[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  // order eigenpairs by descending eigenvalue
  int i0 = 0, i1 = 1, i2 = 2;
  double e0 = A[0], e1 = A[4], e2 = A[8];
  if (e0 < e1) { double t = e0; e0 = e1; e1 = t; int ti = i0; i0 = i1; i1 = ti; }
  if (e0 < e2) { double t = e0; e0 = e2; e2 = t; int ti = i0; i0 = i2; i2 = ti; }
  if (e1 < e2) { double t = e1; e1 = e2; e2 = t; int ti = i1; i1 = i2; i2 = ti; }
  double v1[3] = {V[0 * 3 + i0], V[1 * 3 + i0], V[2 * 3 + i0]};
  double v2[3] = {V[0 * 3 + i1], V[1 * 3 + i1], V[2 * 3 + i1]};
  if (!(e0 > 1e-60)) {  // H == 0: LAPACK returns U = V = I -> R = I
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double u1[3], u2[3];
  for (int i = 0; i < 3; ++i) u1[i] = H[i * 3 + 0] * v1[0] + H[i * 3 + 1] * v1[1] + H[i * 3 + 2] * v1[2];
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int i = 0; i < 3; ++i) u1[i] /= n1;
  for (int i = 0; i < 3; ++i) u2[i] = H[i * 3 + 0] * v2[0] + H[i * 3 + 1] * v2[1] + H[i * 3 + 2] * v2[2];
  double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  if (!(n2 > 1e-12 * n1)) {
    // rank-1 covariance: rotation about u1 is undetermined (as in the reference); pick a deterministic frame
    double ax[3] = {fabs(u1[0]) < 0.9 ? 1.0 : 0.0, fabs(u1[0]) < 0.9 ? 0.0 : 1.0, 0.0};
    double d = ax[0] * u1[0] + ax[1] * u1[1] + ax[2] * u1[2];
    for (int i = 0; i < 3; ++i) u2[i] = ax[i] - d * u1[i];
    n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  }
  for (int i = 0; i < 3; ++i) u2[i] /= n2;
  const double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
  const double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
}

__device__ inline void write_trans(float* T, const double R[9], const double ca[3], const double cb[3]) {
  // t = cb - R ca (common.py:46), T = [R t; 0 0 0 1] (SE3.py:73-96)
  for (int i = 0; i < 3; ++i) {
    const float r0 = (float)R[i * 3 + 0], r1 = (float)R[i * 3 + 1], r2 = (float)R[i * 3 + 2];
    T[i * 4 + 0] = r0; T[i * 4 + 1] = r1; T[i * 4 + 2] = r2;
    T[i * 4 + 3] = (float)(cb[i] - ((double)r0 * ca[0] + (double)r1 * ca[1] + (double)r2 * ca[2]));
  }
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// ------------------------------------------------------------------------------------------------
// prep: centred key points (sx,sy,sz,|s|^2,tx,ty,tz,|t|^2) padded to tiles*128 rows, float4 point copies
// ------------------------------------------------------------------------------------------------
__global__ void prep_points_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int N, int Np,
                                   float* __restrict__ kpts, float4* __restrict__ src4, float4* __restrict__ tgt4) {
  const int pair = blockIdx.x;
  const float* s = src + (size_t)pair * N * 3;
  const float* t = tgt + (size_t)pair * N * 3;
  __shared__ double red[6][32];
  __shared__ float mean[6];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    acc[0] += s[i * 3]; acc[1] += s[i * 3 + 1]; acc[2] += s[i * 3 + 2];
    acc[3] += t[i * 3]; acc[4] += t[i * 3 + 1]; acc[5] += t[i * 3 + 2];
  }
  for (int k = 0; k < 6; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
    mean[threadIdx.x] = (float)(v / N);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Np; i += blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (i < N) {
      const float sx = s[i * 3], sy = s[i * 3 + 1], sz = s[i * 3 + 2];
      const float tx = t[i * 3], ty = t[i * 3 + 1], tz = t[i * 3 + 2];
      src4[(size_t)pair * N + i] = make_float4(sx, sy, sz, 0.f);
      tgt4[(size_t)pair * N + i] = make_float4(tx, ty, tz, 0.f);
      a.x = sx - mean[0]; a.y = sy - mean[1]; a.z = sz - mean[2];
      b.x = tx - mean[3]; b.y = ty - mean[4]; b.z = tz - mean[5];
      a.w = a.x * a.x + a.y * a.y + a.z * a.z;
      b.w = b.x * b.x + b.y * b.y + b.z * b.z;
    }
    float4* o = reinterpret_cast<float4*>(kpts + ((size_t)pair * Np + i) * 8);
    o[0] = a; o[1] = b;
  }
}

// layer0: Conv1d(6 -> 128, k=1)  (PointDSC.py:88,139).  One thread = one token x 4 channels.
// One warp per row (lane = 4 output channels), weights and bias held in registers across a grid-stride loop over rows, four rows in
// flight per iteration: the kernel is a pure 512-byte-per-row store stream (HBM bound) instead of 30 scattered loads per output float4.
// img != NULL (forward path): the rows are written as the split fp16 tile image [B][tiles][hi 2 x 16 KB | lo 2 x 16 KB] the persistent PointCN / QKV
// kernel loads with one bulk copy (pcn_qkv.cuh); `rows` then counts PADDED rows (B x tiles x 128) and rows >= L of a pair's last tile are zeros.
__global__ void __launch_bounds__(256) layer0_kernel(const float* __restrict__ corr, const float* __restrict__ w, const float* __restrict__ b,
                                                     float* __restrict__ out, long long rows, int in_dim, float* __restrict__ img = nullptr, int L = 0,
                                                     int tiles = 0) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int c4 = lane * 4;
  float wr[4][8], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    br[j] = b[c4 + j];
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[j][k] = k < in_dim ? w[(c4 + j) * in_dim + k] : 0.f;
  }
  for (long long row = gw; row < rows; row += 4 * nw) {
    float xv[4];
    long long src_row[4];                                        // row of `corr` (-1: padding row of an image tile)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long rr = row + u * nw;
      src_row[u] = rr;
      if (img) {
        const long long pair = rr / ((long long)tiles * 128);
        const int r = (int)(rr - pair * tiles * 128);
        src_row[u] = r < L ? pair * L + r : -1;
      }
      xv[u] = (rr < rows && src_row[u] >= 0 && lane < in_dim) ? corr[src_row[u] * in_dim + lane] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long rr = row + u * nw;
      float o[4] = {br[0], br[1], br[2], br[3]};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < in_dim) {
          const float x = __shfl_sync(0xffffffffu, xv[u], k);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = fmaf(wr[j][k], x, o[j]);
        }
      }
      if (rr < rows && !img) *reinterpret_cast<float4*>(out + rr * 128 + c4) = make_float4(o[0], o[1], o[2], o[3]);
      if (rr < rows && img) {
        if (src_row[u] < 0) o[0] = o[1] = o[2] = o[3] = 0.f;
        uint2 H, Lw;
        split_f16x2(o[0], o[1], H.x, Lw.x);
        split_f16x2(o[2], o[3], H.y, Lw.y);
        // channels 4 lane .. 4 lane + 3 -> fp16 atom lane / 16, 16-byte chunk (lane & 15) / 2, 8-byte half (lane & 1)
        uint8_t* base = (uint8_t*)(img + (rr >> 7) * (128 * 128)) + (lane >> 4) * 16384 + swz_off((int)(rr & 127), (lane & 15) >> 1) + (lane & 1) * 8;
        *reinterpret_cast<uint2*>(base) = H;
        *reinterpret_cast<uint2*>(base + 32768) = Lw;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// classifier (fp32, accuracy critical) + L2 normalise   (PointDSC.py:229, 175-181, 241)
// ------------------------------------------------------------------------------------------------
struct ClsWeights { const float *w1, *b1, *w2, *b2, *w3, *b3; };   // [32][128],[32],[32][32],[32],[32],[1]

__global__ void __launch_bounds__(256) classify_normalize_kernel(const float* __restrict__ feat, long long rows, ClsWeights cw,
                                                                 float* __restrict__ normed, float* __restrict__ conf) {
  extern __shared__ float sm[];
  float* tile = sm;                 // [64][132]
  float* w1t = tile + 64 * 132;     // [128][32]
  float* w2t = w1t + 128 * 32;      // [32][32]
  float* h1 = w2t + 32 * 32;        // [64][33]
  float* h2 = h1 + 64 * 33;         // [64][33]
  float* inv = h2 + 64 * 33;        // [64]
  const int tid = threadIdx.x;
  // weights are staged (transposed) once per block; the block then walks over 64-row tiles (grid-stride), so the 20 KB of weight
  // reads and the bank-conflicted transposing stores are paid once per block instead of once per 64 rows
  for (int i = tid; i < 32 * 128; i += 256) w1t[(i & 127) * 32 + (i >> 7)] = cw.w1[i];
  for (int i = tid; i < 32 * 32; i += 256) w2t[(i & 31) * 32 + (i >> 5)] = cw.w2[i];
  for (long long row0 = (long long)blockIdx.x * 64; row0 < rows; row0 += (long long)gridDim.x * 64) {
  __syncthreads();                                           // the previous tile's readers are done with tile / h1 / h2 / inv
  for (int i = tid; i < 64 * 32; i += 256) {
    const int r = i >> 5, c4 = (i & 31) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows) v = *reinterpret_cast<const float4*>(feat + (row0 + r) * 128 + c4);
    *reinterpret_cast<float4*>(tile + r * 132 + c4) = v;
  }
  __syncthreads();
  const int r = tid >> 2, part = tid & 3;
  {
    float acc[8];
    for (int u = 0; u < 8; ++u) acc[u] = cw.b1[part * 8 + u];
    float ss = 0.f;
    for (int c = 0; c < 128; ++c) {
      const float f = tile[r * 132 + c];
      if ((c >> 5) == part) ss = fmaf(f, f, ss);
      const float4 wa = *reinterpret_cast<const float4*>(w1t + c * 32 + part * 8);
      const float4 wb = *reinterpret_cast<const float4*>(w1t + c * 32 + part * 8 + 4);
      acc[0] = fmaf(wa.x, f, acc[0]); acc[1] = fmaf(wa.y, f, acc[1]); acc[2] = fmaf(wa.z, f, acc[2]); acc[3] = fmaf(wa.w, f, acc[3]);
      acc[4] = fmaf(wb.x, f, acc[4]); acc[5] = fmaf(wb.y, f, acc[5]); acc[6] = fmaf(wb.z, f, acc[6]); acc[7] = fmaf(wb.w, f, acc[7]);
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    if (part == 0) inv[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);     // F.normalize eps
    for (int u = 0; u < 8; ++u) h1[r * 33 + part * 8 + u] = fmaxf(acc[u], 0.f);
  }
  __syncthreads();
  {
    float acc[8];
    for (int u = 0; u < 8; ++u) acc[u] = cw.b2[part * 8 + u];
    for (int c = 0; c < 32; ++c) {
      const float f = h1[r * 33 + c];
      for (int u = 0; u < 8; ++u) acc[u] = fmaf(w2t[c * 32 + part * 8 + u], f, acc[u]);
    }
    for (int u = 0; u < 8; ++u) h2[r * 33 + part * 8 + u] = fmaxf(acc[u], 0.f);
  }
  __syncthreads();
  if (tid < 64 && row0 + tid < rows) {
    float acc = cw.b3[0];
    for (int c = 0; c < 32; ++c) acc = fmaf(cw.w3[c], h2[tid * 33 + c], acc);
    conf[row0 + tid] = acc;
  }
  for (int i = tid; i < 64 * 32; i += 256) {
    const int rr = i >> 5, c4 = (i & 31) * 4;
    if (row0 + rr < rows) {
      const float s = inv[rr];
      const float4 v = *reinterpret_cast<const float4*>(tile + rr * 132 + c4);
      *reinterpret_cast<float4*>(normed + (row0 + rr) * 128 + c4) = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
    }
  }
  }
}
constexpr int kClsSmem = (64 * 132 + 128 * 32 + 32 * 32 + 2 * 64 * 33 + 64) * 4;

// ------------------------------------------------------------------------------------------------
// seed picking: NMS (PointDSC.py:268-286) on the fly (no N x N distance matrix), then a stable descending sort
// ------------------------------------------------------------------------------------------------
// key_i = score_i * is_local_max_i, is_local_max_i = AND_j (score_i >= score_j  OR  |s_i - s_j| >= R)
// The reference compares sqrt(d2) >= R in fp32.  sqrtf is monotone and correctly rounded, so that test equals d2 >= T2 with T2 the
// smallest float whose square root reaches R (found by stepping a few ulps around R*R): no square root in the N^2 loop, same bits.
__device__ __forceinline__ float sqrt_threshold(float radius) {
  float t = radius * radius;
  if (!(t > 0.f) || !isfinite(t)) return t;
  for (int n = 0; n < 8 && sqrtf(t) >= radius; ++n) t = __uint_as_float(__float_as_uint(t) - 1u);   // now sqrtf(t) < radius (or 8 ulps below)
  for (int n = 0; n < 16 && sqrtf(t) < radius; ++n) t = __uint_as_float(__float_as_uint(t) + 1u);   // first float with sqrtf(t) >= radius
  return t;
}

// perm (optional): the points of each pair in descending-x order.  Thread i then owns point perm[i] and tile j0 holds points
// perm[j0 .. j0+255], so both the block and the tile cover a narrow x interval and a tile whose interval is at least the radius away
// from the block's is skipped as a whole (|dx| alone already gives d2 >= T2 for every pair in it: fl is monotone, the other two squares
// only add).  The AND over j is order independent, so the keys are bit-identical to the unsorted sweep.
__global__ void __launch_bounds__(256) nms_key_kernel(const float4* __restrict__ src4, const float* __restrict__ score, int N,
                                                      float radius, int use_nms, const int* __restrict__ perm, float* __restrict__ key) {
  __shared__ float4 tp[256];
  __shared__ float xr[2];
  const int pair = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float4* P = src4 + (size_t)pair * N;
  const float* S = score + (size_t)pair * N;
  const int* Q = perm ? perm + (size_t)pair * N : nullptr;
  const int pi = (i < N && Q) ? Q[i] : i;
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  float ms = 0.f;
  if (i < N) { me = P[pi]; ms = S[pi]; }
  bool ismax = true;
  if (use_nms) {
    const float t2 = sqrt_threshold(radius);
    float bmax = 0.f, bmin = 0.f;
    if (Q) {                                                 // x interval of this block's points (descending order)
      if (threadIdx.x == 0) xr[0] = me.x;
      if (i == min(blockIdx.x * 256 + 255, N - 1)) xr[1] = me.x;
      __syncthreads();
      bmax = xr[0]; bmin = xr[1];
    }
    for (int j0 = 0; j0 < N; j0 += 256) {
      if (Q) {
        const float tmax = P[Q[j0]].x, tmin = P[Q[min(j0 + 255, N - 1)]].x;      // block-uniform loads
        const float gap = fmaxf(__fsub_rn(tmin, bmax), __fsub_rn(bmin, tmax));     // > 0: the intervals are disjoint
        if (gap > 0.f && __fmul_rn(gap, gap) >= t2) continue;
      }
      const int j = j0 + threadIdx.x;
      float4 v = make_float4(0.f, 0.f, 0.f, -INFINITY);     // padding never suppresses: its score is -inf
      if (j < N) { const int pj = Q ? Q[j] : j; v = P[pj]; v.w = S[pj]; }
      __syncthreads();
      tp[threadIdx.x] = v;
      __syncthreads();
#pragma unroll 8
      for (int jj = 0; jj < 256; ++jj) {
        const float4 o = tp[jj];
        const float dx = me.x - o.x, dy = me.y - o.y, dz = me.z - o.z;
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        ismax = ismax && ((ms >= o.w) || (d2 >= t2));
      }
    }
  }
  if (i < N) key[(size_t)pair * N + pi] = __fmul_rn(ms, ismax ? 1.0f : 0.0f) + 0.0f;   // +0 canonicalises -0
}

// descending stable sort of key (ties -> lower index first) in one CTA per pair; writes the first S indices
__global__ void __launch_bounds__(1024) topk_sort_kernel(const float* __restrict__ key, int stride, int N, int npow2, int S, int* __restrict__ seeds) {
  extern __shared__ unsigned long long sk[];
  const int pair = blockIdx.x;
  const float* K = key + (size_t)pair * N * stride;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    unsigned long long v = ~0ull;
    if (i < N) {
      unsigned u = __float_as_uint(K[(size_t)i * stride]);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending-orderable
      v = ((unsigned long long)(~u) << 32) | (unsigned)i; // descending value, then ascending index
    }
    sk[i] = v;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < npow2; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const unsigned long long a = sk[t], b = sk[p];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { sk[t] = b; sk[p] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < S; i += blockDim.x) seeds[(size_t)pair * S + i] = (int)(sk[i] & 0xffffffffu);
}

// ------------------------------------------------------------------------------------------------
// seed kNN in feature space (common.py:53-75 restricted to the seed rows; PointDSC.py:325-329):
// d_j = 2 - 2 <f_seed, f_j>, (k+1) smallest, rank 0 dropped.  Ties -> lower index first.
// ------------------------------------------------------------------------------------------------
// Kernel 1 is the distance GEMM on the tensor pipe (dgr_head.cuh: knn_operand_kernel + img_gemm_kernel<128, DE_DIST>).
// Kernel 2: per seed, the (k+1) smallest distances in ascending order (ties -> lower index), rank 0 dropped.  One warp per seed,
// the seed's distance row in shared memory.  Two scans instead of k+1: (1) every lane finds its two smallest values; the (k+1)-th
// smallest of those 64 is an upper bound T of the row's (k+1)-th smallest; (2) all entries <= T (usually 41..100) are compacted
// into a candidate list, from which the k+1 smallest are extracted by repeated (value, index) arg-min.  Rows with too many
// candidates (massive ties) or fewer than 64 entries take the plain k+1 full scans.
constexpr int kSelCap = 256;
// SPC seeds (warps) per CTA.  The distance row is streamed from global memory twice (float4 per lane; the second pass hits L1/L2)
// instead of being staged in shared memory: 2 KB of shared memory per warp (the candidate list) keeps the SM fully occupied.
template <int SPC>
__global__ void __launch_bounds__(SPC * 32) seed_select_kernel(const float* __restrict__ dist, int N, int S, int k, int* __restrict__ knn_idx) {
  __shared__ float s_cv[SPC][kSelCap];
  __shared__ int s_ci[SPC][kSelCap];
  const int pair = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * SPC + warp;
  if (s >= S) return;
  float* cv = s_cv[warp];
  int* ci = s_ci[warp];
  const float* src = dist + ((size_t)pair * S + s) * N;
  const bool vec = (N & 3) == 0;                                // rows are 16-byte aligned when N % 4 == 0
  const int nv = vec ? N >> 2 : 0;
  float m1 = INFINITY, m2 = INFINITY;
  auto upd = [&](float v) { if (v < m1) { m2 = m1; m1 = v; } else if (v < m2) m2 = v; };
  for (int q = lane; q < nv; q += 32) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
    upd(v.x); upd(v.y); upd(v.z); upd(v.w);
  }
  for (int j = nv * 4 + lane; j < N; j += 32) upd(__ldg(src + j));
  int ncand = -1;
  if (N >= 64 && k + 1 <= 64) {
    // rank of m1 / m2 among the 64 lane minima (ties broken by slot number): the value of rank k is the threshold
    int r1 = 0, r2 = 0;
    for (int l = 0; l < 32; ++l) {
      const float o1 = __shfl_sync(0xffffffffu, m1, l), o2 = __shfl_sync(0xffffffffu, m2, l);
      r1 += (o1 < m1 || (o1 == m1 && 2 * l < 2 * lane)) + (o2 < m1 || (o2 == m1 && 2 * l + 1 < 2 * lane));
      r2 += (o1 < m2 || (o1 == m2 && 2 * l < 2 * lane + 1)) + (o2 < m2 || (o2 == m2 && 2 * l + 1 < 2 * lane + 1));
    }
    const unsigned h1 = __ballot_sync(0xffffffffu, r1 == k), h2 = __ballot_sync(0xffffffffu, r2 == k);
    const float T = h1 ? __shfl_sync(0xffffffffu, m1, __ffs(h1) - 1) : __shfl_sync(0xffffffffu, m2, __ffs(h2) - 1);
    // second pass: compact every entry <= T (index order is not needed: the extraction below orders by (value, index))
    ncand = 0;
    auto push = [&](float v, int j, bool ok) {
      const unsigned mk = __ballot_sync(0xffffffffu, ok);
      const int pos = ncand + __popc(mk & ((1u << lane) - 1u));
      if (ok && pos < kSelCap) { cv[pos] = v; ci[pos] = j; }
      ncand += __popc(mk);
    };
    for (int q0 = 0; q0 < nv; q0 += 32) {
      const int q = q0 + lane;
      float4 v = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
      if (q < nv) v = __ldg(reinterpret_cast<const float4*>(src) + q);
      const bool any4 = fminf(fminf(v.x, v.y), fminf(v.z, v.w)) <= T;
      if (__any_sync(0xffffffffu, any4)) {
        push(v.x, 4 * q, v.x <= T); push(v.y, 4 * q + 1, v.y <= T); push(v.z, 4 * q + 2, v.z <= T); push(v.w, 4 * q + 3, v.w <= T);
      }
    }
    for (int j0 = nv * 4; j0 < N; j0 += 32) {
      const int j = j0 + lane;
      const float v = j < N ? __ldg(src + j) : INFINITY;
      push(v, j, v <= T);
    }
    __syncwarp();
    if (ncand > kSelCap) ncand = -1;
  }
  if (ncand >= 0) {
    for (int rnk = 0; rnk <= k; ++rnk) {
      float bv = INFINITY;
      int bi = 0x7fffffff, bp = -1;
      for (int c = lane; c < ncand; c += 32) {
        const float v = cv[c];
        const int i = ci[c];
        if (v < bv || (v == bv && i < bi)) { bv = v; bi = i; bp = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o), op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bp = op; }
      }
      if (lane == 0) {
        if (bp >= 0) { cv[bp] = INFINITY; ci[bp] = 0x7fffffff; }
        if (rnk > 0) knn_idx[((size_t)pair * S + s) * k + rnk - 1] = bi < N ? bi : 0;
      }
      __syncwarp();
    }
    return;
  }
  // fallback (tiny rows or massive ties): k + 1 full scans of the global row, excluding the indices already taken
  float last_v = -INFINITY;
  int last_i = -1;
  for (int rnk = 0; rnk <= k; ++rnk) {
    float bv = INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < N; j += 32) {
      const float v = __ldg(src + j);
      const bool after = v > last_v || (v == last_v && j > last_i);   // strictly after the previous pick in (value, index) order
      if (after && (v < bv || (v == bv && j < bi))) { bv = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    last_v = bv; last_i = bi;
    if (lane == 0 && rnk > 0) knn_idx[((size_t)pair * S + s) * k + rnk - 1] = bi < N ? bi : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// per-seed spectral matching + Kabsch (PointDSC.py:335-407, 429-448; common.py:10-50)
// PASS 0: run all iterations, record at which iterations this seed satisfies allclose(v_new, v_old) and AND the bitmask into
//         the pair's word (the reference's early exit is global over all seeds of the pair).
// PASS 1: run exactly the reference's iteration count, then weights -> weighted Kabsch -> seed transform.
// ------------------------------------------------------------------------------------------------
template <int PASS>
__global__ void __launch_bounds__(128) seed_spectral_kernel(const float* __restrict__ normed, const float* __restrict__ src,
                                                            const float* __restrict__ tgt, const int* __restrict__ knn_idx, int N, int S,
                                                            int k, float sigma, float sigma_spat, int iters,
                                                            unsigned* __restrict__ pair_mask, float* __restrict__ seed_w,
                                                            float* __restrict__ seed_trans, float* __restrict__ seedM) {
  constexpr int KM = 40;
  __shared__ float kf[KM][129];
  __shared__ float ks[KM][3], kt[KM][3];
  __shared__ float M[KM][KM + 1];
  __shared__ float v[KM], vn[KM];
  __shared__ float red[4];
  __shared__ int idx[KM];
  const int sidx = blockIdx.x, pair = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t sg = (size_t)pair * S + sidx;
  for (int e = tid; e < k; e += blockDim.x) idx[e] = knn_idx[sg * k + e];
  __syncthreads();
  for (int e = tid; e < k * 3; e += blockDim.x) {
    const int r = e / 3, c = e % 3;
    ks[r][c] = src[((size_t)pair * N + idx[r]) * 3 + c];
    kt[r][c] = tgt[((size_t)pair * N + idx[r]) * 3 + c];
  }
  // pass 0 publishes the eigenvector estimate after every iteration; pass 1 (one warp per seed) only picks the iterate at which
  // the reference's global allclose test stopped, so neither M nor the iterations are recomputed
  float* vh = seedM + sg * (KM * KM);
  if (PASS == 0) {
    const float* F = normed + (size_t)pair * N * 128;
    for (int r = warp; r < k; r += 4) {
      const float4 f = *reinterpret_cast<const float4*>(F + (size_t)idx[r] * 128 + lane * 4);
      kf[r][lane * 4] = f.x; kf[r][lane * 4 + 1] = f.y; kf[r][lane * 4 + 2] = f.z; kf[r][lane * 4 + 3] = f.w;
    }
    __syncthreads();
    const float inv_s2 = 1.0f / (sigma * sigma), inv_d2 = 1.0f / (sigma_spat * sigma_spat);
    // M is symmetric: only the 4 x 4 blocks on and above the diagonal are computed (55 of 100 for k = 40), one block per
    // thread: 8 shared-memory loads feed 16 FMAs per channel (the 1 x 4 version was shared-memory bound: 5 loads per 4 FMAs).
    const int nb = (k + 3) >> 2;
    int t = tid, bi = 0;
    while (bi < nb && t >= nb - bi) { t -= nb - bi; ++bi; }
    if (bi < nb) {
      const int bj = bi + t, i0 = bi * 4, j0 = bj * 4;
      int ri[4], rj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { ri[u] = min(i0 + u, k - 1); rj[u] = min(j0 + u, k - 1); }
      float dot[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) dot[u][w] = 0.f;
#pragma unroll 4
      for (int c = 0; c < 128; ++c) {
        float av[4], bv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { av[u] = kf[ri[u]][c]; bv[u] = kf[rj[u]][c]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) dot[u][w] = fmaf(av[u], bv[w], dot[u][w]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u;
        if (i >= k) break;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int j = j0 + w;
          if (j >= k || (bi == bj && j < i)) continue;
          const float mf = fmaxf(1.0f - (1.0f - dot[u][w]) * inv_s2, 0.f);                 // PointDSC.py:338
          const float ax = ks[i][0] - ks[j][0], ay = ks[i][1] - ks[j][1], az = ks[i][2] - ks[j][2];
          const float bx = kt[i][0] - kt[j][0], by = kt[i][1] - kt[j][1], bz = kt[i][2] - kt[j][2];
          const float dd = sqrtf(ax * ax + ay * ay + az * az) - sqrtf(bx * bx + by * by + bz * bz);
          const float ms = fmaxf(1.0f - dd * dd * inv_d2, 0.f);                          // :351
          const float m = (i == j) ? 0.f : mf * ms;                                      // :360-361
          M[i][j] = m; M[j][i] = m;
        }
      }
    }
  }
  if (PASS == 1) {
    int n_iter = iters;
    const unsigned m = pair_mask[pair] & ((iters >= 32) ? 0xffffffffu : ((1u << iters) - 1u));
    if (m) n_iter = __ffs(m);       // first iteration (1-based) at which every seed was close -> break after it
    for (int e = tid; e < k; e += blockDim.x) v[e] = vh[(n_iter - 1) * KM + e];
    __syncthreads();
  } else {
  if (tid < k) v[tid] = 1.0f;
  __syncthreads();
  const int n_iter = iters;
  unsigned close_mask = 0;
  for (int it = 0; it < n_iter; ++it) {
    float acc = 0.f;
    if (tid < k) {
      for (int j = 0; j < k; ++j) acc = fmaf(M[tid][j], v[j], acc);
    }
    float sq = (tid < k) ? acc * acc : 0.f;
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    const float nrm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    float nv = acc / (nrm + 1e-6f);                                                  // :443
    int bad = 0;
    if (tid < k) {
      vn[tid] = nv;
      vh[it * KM + tid] = nv;
      bad = !(fabsf(nv - v[tid]) <= 1e-8f + 1e-5f * fabsf(v[tid]));                  // torch.allclose defaults (:444)
    }
    const int any_bad = __syncthreads_or(bad);
    if (!any_bad) close_mask |= 1u << it;
    if (tid < k) v[tid] = vn[tid];
    __syncthreads();
  }
  if (tid == 0) atomicAnd(&pair_mask[pair], close_mask);
  return;
  }
  // weights (:365) and weighted Kabsch on the k neighbours
  if (warp == 0) {
    float w0 = lane < k ? v[lane] : 0.f, w1 = (lane + 32) < k ? v[lane + 32] : 0.f;
    const float tot = warp_sum(w0 + w1) + 1e-6f;
    w0 /= tot; w1 /= tot;
    if (seed_w) {
      if (lane < k) seed_w[sg * k + lane] = w0;
      if (lane + 32 < k) seed_w[sg * k + lane + 32] = w1;
    }
    // common.py:20 weights[weights < 0] = 0
    w0 = fmaxf(w0, 0.f); w1 = fmaxf(w1, 0.f);
    const float wsum = warp_sum(w0 + w1) + 1e-6f;
    float ca[3], cb[3];
    for (int c = 0; c < 3; ++c) {
      const float a0 = lane < k ? ks[lane][c] : 0.f, a1 = lane + 32 < k ? ks[lane + 32][c] : 0.f;
      const float b0 = lane < k ? kt[lane][c] : 0.f, b1 = lane + 32 < k ? kt[lane + 32][c] : 0.f;
      ca[c] = warp_sum(w0 * a0 + w1 * a1) / wsum;
      cb[c] = warp_sum(w0 * b0 + w1 * b1) / wsum;
    }
    double H[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float h = 0.f;
        if (lane < k) h += w0 * (ks[lane][i] - ca[i]) * (kt[lane][j] - cb[j]);
        if (lane + 32 < k) h += w1 * (ks[lane + 32][i] - ca[i]) * (kt[lane + 32][j] - cb[j]);
        H[i * 3 + j] = (double)warp_sum(h);
      }
    if (lane == 0) {
      double R[9];
      kabsch_rotation(H, R);
      const double cad[3] = {ca[0], ca[1], ca[2]}, cbd[3] = {cb[0], cb[1], cb[2]};
      write_trans(seed_trans + sg * 16, R, cad, cbd);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hypothesis scoring: inlier counts of every seed transform over all points (PointDSC.py:413-419)
// ------------------------------------------------------------------------------------------------
constexpr int kScoreSeeds = 32;
constexpr int kScorePPT = 4;   // points per thread (kept in registers across the seed loop)
__global__ void __launch_bounds__(256) score_kernel(const float4* __restrict__ src4, const float4* __restrict__ tgt4,
                                                    const float* __restrict__ seed_trans, int N, int S, float tau, int* __restrict__ counts) {
  __shared__ float4 T[kScoreSeeds][3];
  __shared__ int cnt_s[kScoreSeeds];
  const int pair = blockIdx.z, s0 = blockIdx.x * kScoreSeeds, tid = threadIdx.x;
  if (tid < kScoreSeeds * 3) {
    const int s = tid / 3, r = tid % 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s0 + s < S) v = *reinterpret_cast<const float4*>(seed_trans + ((size_t)pair * S + s0 + s) * 16 + r * 4);
    T[s][r] = v;
  }
  if (tid < kScoreSeeds) cnt_s[tid] = 0;
  __syncthreads();
  const float4* A = src4 + (size_t)pair * N;
  const float4* B = tgt4 + (size_t)pair * N;
  float4 x[kScorePPT], y[kScorePPT];
  bool ok[kScorePPT];
#pragma unroll
  for (int p = 0; p < kScorePPT; ++p) {
    const int i = blockIdx.y * (256 * kScorePPT) + p * 256 + tid;      // coalesced float4 streams
    ok[p] = i < N;
    x[p] = ok[p] ? A[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    y[p] = ok[p] ? B[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int ns = min(kScoreSeeds, S - s0);
#pragma unroll 2
  for (int s = 0; s < ns; ++s) {
    const float4 r0 = T[s][0], r1 = T[s][1], r2 = T[s][2];
    int c = 0;
#pragma unroll
    for (int p = 0; p < kScorePPT; ++p) {
      const float dx = fmaf(r0.x, x[p].x, fmaf(r0.y, x[p].y, fmaf(r0.z, x[p].z, r0.w))) - y[p].x;
      const float dy = fmaf(r1.x, x[p].x, fmaf(r1.y, x[p].y, fmaf(r1.z, x[p].z, r1.w))) - y[p].y;
      const float dz = fmaf(r2.x, x[p].x, fmaf(r2.y, x[p].y, fmaf(r2.z, x[p].z, r2.w))) - y[p].z;
      c += (ok[p] && sqrtf(dx * dx + dy * dy + dz * dz) < tau) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0 && c) atomicAdd(&cnt_s[s], c);
  }
  __syncthreads();
  if (tid < ns && cnt_s[tid]) atomicAdd(&counts[(size_t)pair * S + s0 + tid], cnt_s[tid]);
}

// argmax fitness (first max, :421), final transform + labels (:423-425), then post-refinement (:493-528) on device.
__global__ void __launch_bounds__(1024) select_refine_kernel(const float4* __restrict__ src4, const float4* __restrict__ tgt4,
                                                             const float* __restrict__ seed_trans, const int* __restrict__ counts, int N,
                                                             int S, float tau, float refine_tau, int refine_iters, int do_refine,
                                                             float* __restrict__ pre_refine, float* __restrict__ final_trans,
                                                             float* __restrict__ labels, int* __restrict__ best_out) {
  __shared__ int s_best;
  __shared__ float T[16];
  __shared__ double red[13][32];
  __shared__ double tot[13];
  __shared__ int s_stop;
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* A = src4 + (size_t)pair * N;
  const float4* B = tgt4 + (size_t)pair * N;
  // --- argmax over seeds (count, lowest index wins ties)
  long long best = -1;
  for (int s = tid; s < S; s += 1024) {
    const long long key = ((long long)counts[(size_t)pair * S + s] << 32) | (unsigned)(0x7fffffff - s);
    best = key > best ? key : best;
  }
  for (int o = 16; o > 0; o >>= 1) { const long long ot = __shfl_xor_sync(0xffffffffu, best, o); best = ot > best ? ot : best; }
  __shared__ long long wbest[32];
  if (lane == 0) wbest[warp] = best;
  __syncthreads();
  if (tid == 0) {
    long long b = wbest[0];
    for (int w = 1; w < 32; ++w) b = wbest[w] > b ? wbest[w] : b;
    s_best = 0x7fffffff - (int)(b & 0xffffffffll);
    if (best_out) best_out[pair] = s_best;
  }
  __syncthreads();
  if (tid < 16) {
    T[tid] = seed_trans[((size_t)pair * S + s_best) * 16 + tid];
    if (pre_refine) pre_refine[pair * 16 + tid] = T[tid];
  }
  __syncthreads();
  // --- labels from the best hypothesis
  for (int i = tid; i < N; i += 1024) {
    const float4 x = A[i], y = B[i];
    const float dx = fmaf(T[0], x.x, fmaf(T[1], x.y, fmaf(T[2], x.z, T[3]))) - y.x;
    const float dy = fmaf(T[4], x.x, fmaf(T[5], x.y, fmaf(T[6], x.z, T[7]))) - y.y;
    const float dz = fmaf(T[8], x.x, fmaf(T[9], x.y, fmaf(T[10], x.z, T[11]))) - y.z;
    labels[(size_t)pair * N + i] = (sqrtf(dx * dx + dy * dy + dz * dz) < tau) ? 1.f : 0.f;
  }
  // --- post refinement
  long long prev = 0;
  if (do_refine) {
    for (int it = 0; it < refine_iters; ++it) {
      double acc[13];
      for (int q = 0; q < 13; ++q) acc[q] = 0.0;
      // pass 1: inlier count, sum w, sum w a, sum w b
      for (int i = tid; i < N; i += 1024) {
        const float4 x = A[i], y = B[i];
        const float dx = fmaf(T[0], x.x, fmaf(T[1], x.y, fmaf(T[2], x.z, T[3]))) - y.x;
        const float dy = fmaf(T[4], x.x, fmaf(T[5], x.y, fmaf(T[6], x.z, T[7]))) - y.y;
        const float dz = fmaf(T[8], x.x, fmaf(T[9], x.y, fmaf(T[10], x.z, T[11]))) - y.z;
        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
        if (d < refine_tau) {
          const float q = d / refine_tau;
          const float w = 1.0f / (1.0f + q * q);                                      // :525
          acc[0] += 1.0; acc[1] += w;
          acc[2] += w * x.x; acc[3] += w * x.y; acc[4] += w * x.z;
          acc[5] += w * y.x; acc[6] += w * y.y; acc[7] += w * y.z;
        }
      }
      for (int q = 0; q < 8; ++q) {
        double vq = acc[q];
        for (int o = 16; o > 0; o >>= 1) vq += __shfl_xor_sync(0xffffffffu, vq, o);
        if (lane == 0) red[q][warp] = vq;
      }
      __syncthreads();
      if (tid < 8) {
        double vq = 0;
        for (int w = 0; w < 32; ++w) vq += red[tid][w];
        tot[tid] = vq;
      }
      __syncthreads();
      if (tid == 0) {
        const long long num = (long long)(tot[0] + 0.5);
        s_stop = (num - prev == 0) ? 1 : 0;                                           // :516 abs(int(diff)) < 1
      }
      __syncthreads();
      if (s_stop) break;
      prev = (long long)(tot[0] + 0.5);
      const double wsum = tot[1] + 1e-6;
      const float ca0 = (float)(tot[2] / wsum), ca1 = (float)(tot[3] / wsum), ca2 = (float)(tot[4] / wsum);
      const float cb0 = (float)(tot[5] / wsum), cb1 = (float)(tot[6] / wsum), cb2 = (float)(tot[7] / wsum);
      for (int q = 0; q < 9; ++q) acc[q] = 0.0;
      for (int i = tid; i < N; i += 1024) {
        const float4 x = A[i], y = B[i];
        const float dx = fmaf(T[0], x.x, fmaf(T[1], x.y, fmaf(T[2], x.z, T[3]))) - y.x;
        const float dy = fmaf(T[4], x.x, fmaf(T[5], x.y, fmaf(T[6], x.z, T[7]))) - y.y;
        const float dz = fmaf(T[8], x.x, fmaf(T[9], x.y, fmaf(T[10], x.z, T[11]))) - y.z;
        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
        if (d < refine_tau) {
          const float q = d / refine_tau;
          const float w = 1.0f / (1.0f + q * q);
          const float a0 = x.x - ca0, a1 = x.y - ca1, a2 = x.z - ca2;
          const float b0 = w * (y.x - cb0), b1 = w * (y.y - cb1), b2 = w * (y.z - cb2);
          acc[0] += a0 * b0; acc[1] += a0 * b1; acc[2] += a0 * b2;
          acc[3] += a1 * b0; acc[4] += a1 * b1; acc[5] += a1 * b2;
          acc[6] += a2 * b0; acc[7] += a2 * b1; acc[8] += a2 * b2;
        }
      }
      __syncthreads();
      for (int q = 0; q < 9; ++q) {
        double vq = acc[q];
        for (int o = 16; o > 0; o >>= 1) vq += __shfl_xor_sync(0xffffffffu, vq, o);
        if (lane == 0) red[q][warp] = vq;
      }
      __syncthreads();
      if (tid == 0) {
        double H[9], R[9];
        for (int q = 0; q < 9; ++q) {
          double vq = 0;
          for (int w = 0; w < 32; ++w) vq += red[q][w];
          H[q] = vq;
        }
        kabsch_rotation(H, R);
        const double cad[3] = {ca0, ca1, ca2}, cbd[3] = {cb0, cb1, cb2};
        write_trans(T, R, cad, cbd);
      }
      __syncthreads();
    }
  }
  if (tid < 16) final_trans[pair * 16 + tid] = T[tid];
}

// standalone batched weighted Kabsch: C-ABI mirror of models/common.py:10-50 (one warp per problem)
// ------------------------------------------------------------------------------------------------
// DGR weighted Procrustes (GMF_DeepGlobalRegistration_fcgf/core/registration.py:91-113): one pose per pair from ALL its correspondences.
//   w_norm = w / (sum |w| + eps); mu_x = sum w_norm x; mu_y = sum w_norm y; Sxy = (Y - mu_y)^T diag(w_norm) (X - mu_x)
//   R = U diag(1, 1, sign) V^T  (SVD in double precision in the reference, on the host),  t = mu_y - R mu_x
// One CTA per pair, two sweeps over the points (means, then the 3 x 3 moment), block reductions; the 3 x 3 algebra in fp64 on one lane.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weighted_procrustes_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ W,
                                                                  int N, float eps, float* __restrict__ R_out, float* __restrict__ t_out) {
  __shared__ float red[8][10];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = X + (size_t)pair * N * 3;
  const float* y = Y + (size_t)pair * N * 3;
  const float* w = W + (size_t)pair * N;
  auto block_sum = [&](float (&v)[9], int n, float* out) {       // all threads get the n sums
    for (int c = 0; c < n; ++c) v[c] = warp_sum(v[c]);
    __syncthreads();
    if (lane == 0) for (int c = 0; c < n; ++c) red[warp][c] = v[c];
    __syncthreads();
    for (int c = 0; c < n; ++c) {
      float s = 0.f;
      for (int k = 0; k < 8; ++k) s += red[k][c];
      out[c] = s;
    }
  };
  float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tot[9];
  for (int i = tid; i < N; i += 256) {
    const float wi = w[i];
    acc[0] += fabsf(wi);
    for (int c = 0; c < 3; ++c) { acc[1 + c] += wi * x[i * 3 + c]; acc[4 + c] += wi * y[i * 3 + c]; }
  }
  block_sum(acc, 7, tot);
  const float inv = 1.0f / (tot[0] + eps);
  float mx[3], my[3];
  for (int c = 0; c < 3; ++c) { mx[c] = tot[1 + c] * inv; my[c] = tot[4 + c] * inv; }
  for (int q = 0; q < 9; ++q) acc[q] = 0.f;
  for (int i = tid; i < N; i += 256) {
    const float wn = w[i] * inv;
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 3; ++q) acc[p * 3 + q] += wn * (x[i * 3 + p] - mx[p]) * (y[i * 3 + q] - my[q]);   // H = Sxy^T, as rigid_transform_3d builds it
  }
  block_sum(acc, 9, tot);
  if (tid == 0) {
    double H[9], R[9];
    for (int q = 0; q < 9; ++q) H[q] = (double)tot[q];
    kabsch_rotation(H, R);
    for (int q = 0; q < 9; ++q) R_out[(size_t)pair * 9 + q] = (float)R[q];
    for (int p = 0; p < 3; ++p)
      t_out[(size_t)pair * 3 + p] = (float)((double)my[p] - (R[p * 3] * mx[0] + R[p * 3 + 1] * mx[1] + R[p * 3 + 2] * mx[2]));
  }
}

__global__ void rigid_transform_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ W, int M, int k,
                                       float* __restrict__ out) {
  const int prob = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (prob >= M) return;
  const float* a = A + (size_t)prob * k * 3;
  const float* b = B + (size_t)prob * k * 3;
  float sw = 0.f, sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0};
  for (int i = lane; i < k; i += 32) {
    const float w = W ? fmaxf(W[(size_t)prob * k + i], 0.f) : 1.f;
    sw += w;
    for (int c = 0; c < 3; ++c) { sa[c] += w * a[i * 3 + c]; sb[c] += w * b[i * 3 + c]; }
  }
  sw = warp_sum(sw) + 1e-6f;
  float ca[3], cb[3];
  for (int c = 0; c < 3; ++c) { ca[c] = warp_sum(sa[c]) / sw; cb[c] = warp_sum(sb[c]) / sw; }
  double H[9];
  float h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < k; i += 32) {
    const float w = W ? fmaxf(W[(size_t)prob * k + i], 0.f) : 1.f;
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 3; ++q) h[p * 3 + q] += w * (a[i * 3 + p] - ca[p]) * (b[i * 3 + q] - cb[q]);
  }
  for (int q = 0; q < 9; ++q) H[q] = (double)warp_sum(h[q]);
  if (lane == 0) {
    double R[9];
    kabsch_rotation(H, R);
    const double cad[3] = {ca[0], ca[1], ca[2]}, cbd[3] = {cb[0], cb[1], cb[2]};
    write_trans(out + (size_t)prob * 16, R, cad, cbd);
  }
}

}  // namespace gmf
