// Training step of the DGR bottleneck fusion head (BASELINE.json configs[4]: "forward + one training step with NCCL gradient allreduce";
// reference: GMF_DeepGlobalRegistration_fcgf/core/trainer.py:226-300 - forward :236, loss.backward() :271, optimizer.step() :300 - around
// model/perceiver_io.py:187-221).  Forward with saved activations, analytic backward and SGD for
//     x0 = x + dwconv(x), c0 = ctx + dwconv(ctx)                   (ConvPosEnc :105-136, when pe)
//     q = LN(x0) Wq^T, [k | v] = LN(c0) Wkv^T, P = softmax(q k^T / sqrt(128)), a = P v
//     x1 = x0 + a Wo^T + bo;  u = LN(x1) W1^T + b1;  out = x1 + (u_val * gelu(u_gate)) W2^T + b2
// Every matrix product (7 forward, 14 backward) runs on the tensor pipe through ONE generic path: mat_to_img_kernel turns a row-major matrix (or
// its transpose) into the K-chunked tf32 tile image and img_gemm_kernel<128, DE_STORE> (dgr_head.cuh) multiplies two images.  The head is
// ~9 GFLOP forward at M = 2048, T = 4800, so the step is bound by ~80 small launches, not by a roofline; it is built for completeness of
// cfg#5 (gradient parity against autograd, SGD, one flat NCCL allreduce), not tuned like the inference path.
#pragma once
#include "dgr_head.cuh"

namespace gmf {

// src: row-major; !trans: element (r, k) = src[r * ld + k]; trans: src[k * ld + r].  img [tiles][kch][128 x 32] tf32, zero padded.
__global__ void __launch_bounds__(256) mat_to_img_kernel(const float* __restrict__ src, int ld, int rows, int K, int trans, int kch, float* __restrict__ img) {
  const int tile = blockIdx.x, kc = blockIdx.y;
  uint8_t* chunk = (uint8_t*)(img + ((size_t)tile * kch + kc) * 4096);
  for (int idx = threadIdx.x; idx < 1024; idx += 256) {
    int rr, g;
    if (trans) { rr = idx & 127; g = idx >> 7; } else { rr = idx >> 3; g = idx & 7; }
    const int r = tile * 128 + rr, k0 = kc * 32 + g * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      if (!trans) {
        const float* p = src + (size_t)r * ld + k0;
        if (k0 + 3 < K && (ld & 3) == 0 && ((uintptr_t)src & 15) == 0) v = *reinterpret_cast<const float4*>(p);
        else { if (k0 < K) v.x = p[0]; if (k0 + 1 < K) v.y = p[1]; if (k0 + 2 < K) v.z = p[2]; if (k0 + 3 < K) v.w = p[3]; }
      } else {
        if (k0 < K) v.x = src[(size_t)k0 * ld + r];
        if (k0 + 1 < K) v.y = src[(size_t)(k0 + 1) * ld + r];
        if (k0 + 2 < K) v.z = src[(size_t)(k0 + 2) * ld + r];
        if (k0 + 3 < K) v.w = src[(size_t)(k0 + 3) * ld + r];
      }
    }
    *reinterpret_cast<float4*>(chunk + swz_off(rr, g)) = to_tf32(v);
  }
}

// ConvPosEnc forward: y[t] = x[t] + w0 x[t-1] + w1 x[t] + w2 x[t+1] + b (depthwise, zero padding)
__global__ void cpe_fwd_kernel(const float* __restrict__ x, int L, int C, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)L * C) return;
  const int t = (int)(i / C), c = (int)(i % C);
  const float p = t > 0 ? x[i - C] : 0.f, n = t + 1 < L ? x[i + C] : 0.f;
  y[i] = x[i] + fmaf(w[c * 3], p, fmaf(w[c * 3 + 1], x[i], fmaf(w[c * 3 + 2], n, b[c])));
}
// backward: dx[t] = dy[t] + w0 dy[t+1] + w1 dy[t] + w2 dy[t-1]; dw_k, db accumulated with atomics (grads zeroed by the caller)
__global__ void cpe_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int L, int C, const float* __restrict__ w, float* __restrict__ dx,
                               float* __restrict__ dw, float* __restrict__ db) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int t0 = blockIdx.y * 64, ty = threadIdx.x >> 5;      // block = 32 channels x 8 row lanes, 64 rows per block
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
  if (c < C) {
    const float w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2];
    for (int t = t0 + ty; t < min(t0 + 64, L); t += 8) {
      const size_t i = (size_t)t * C + c;
      const float g = dy[i], gn = t + 1 < L ? dy[i + C] : 0.f, gp = t > 0 ? dy[i - C] : 0.f;
      if (dx) dx[i] = g + fmaf(w0, gn, fmaf(w1, g, w2 * gp));
      a0 = fmaf(g, t > 0 ? x[i - C] : 0.f, a0); a1 = fmaf(g, x[i], a1); a2 = fmaf(g, t + 1 < L ? x[i + C] : 0.f, a2); ab += g;
    }
  }
  __shared__ float red[8][32][4];
  red[ty][threadIdx.x & 31][0] = a0; red[ty][threadIdx.x & 31][1] = a1; red[ty][threadIdx.x & 31][2] = a2; red[ty][threadIdx.x & 31][3] = ab;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s[4] = {0, 0, 0, 0};
    for (int k = 0; k < 8; ++k) for (int j = 0; j < 4; ++j) s[j] += red[k][threadIdx.x][j];
    atomicAdd(dw + c * 3, s[0]); atomicAdd(dw + c * 3 + 1, s[1]); atomicAdd(dw + c * 3 + 2, s[2]); atomicAdd(db + c, s[3]);
  }
}

// LayerNorm forward (eps 1e-5), one warp per row; saves mean / rstd
template <int C>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, int L, const float* __restrict__ g, const float* __restrict__ b,
                                                     float* __restrict__ y, float* __restrict__ stat) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= L) return;
  float v[C / 32], s = 0.f;
#pragma unroll
  for (int i = 0; i < C / 32; ++i) { v[i] = x[(size_t)r * C + i * 32 + lane]; s += v[i]; }
  const float mean = warp_sum(s) * (1.0f / C);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < C / 32; ++i) { v[i] -= mean; sq = fmaf(v[i], v[i], sq); }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / C) + 1e-5f);
#pragma unroll
  for (int i = 0; i < C / 32; ++i) y[(size_t)r * C + i * 32 + lane] = fmaf(v[i] * rstd, g[i * 32 + lane], b[i * 32 + lane]);
  if (lane == 0) { stat[2 * r] = mean; stat[2 * r + 1] = rstd; }
}
// backward: dx (+= add if given) = rstd (dy g - mean(dy g) - xhat mean(dy g xhat)); dgamma / dbeta by atomics over row blocks
template <int C>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stat, int L,
                                                     const float* __restrict__ g, const float* __restrict__ add, float* __restrict__ dx,
                                                     float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float sg[8][C], sb[8][C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[C / 32], ab[C / 32];
#pragma unroll
  for (int i = 0; i < C / 32; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  for (int r = blockIdx.x * 64 + warp; r < min(blockIdx.x * 64 + 64, L); r += 8) {
    const float mean = stat[2 * r], rstd = stat[2 * r + 1];
    float xh[C / 32], dg_[C / 32], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < C / 32; ++i) {
      const int c = i * 32 + lane;
      const float d = dy[(size_t)r * C + c];
      xh[i] = (x[(size_t)r * C + c] - mean) * rstd;
      dg_[i] = d * g[c];
      s1 += dg_[i]; s2 = fmaf(dg_[i], xh[i], s2);
      ag[i] = fmaf(d, xh[i], ag[i]); ab[i] += d;
    }
    s1 = warp_sum(s1) * (1.0f / C); s2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
    for (int i = 0; i < C / 32; ++i) {
      const size_t o = (size_t)r * C + i * 32 + lane;
      const float v = rstd * (dg_[i] - s1 - xh[i] * s2);
      dx[o] = add ? add[o] + v : v;
    }
  }
#pragma unroll
  for (int i = 0; i < C / 32; ++i) { sg[warp][i * 32 + lane] = ag[i]; sb[warp][i * 32 + lane] = ab[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f, bsum = 0.f;
    for (int k = 0; k < 8; ++k) { a += sg[k][c]; bsum += sb[k][c]; }
    atomicAdd(dg + c, a); atomicAdd(db + c, bsum);
  }
}

// row softmax in place: P = softmax(S) (S already scaled).  One CTA per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ S, int T) {
  __shared__ float red[8];
  float* row = S + (size_t)blockIdx.x * T;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < T; j += 256) m = fmaxf(m, row[j]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
  __syncthreads();
  float s = 0.f;
  for (int j = threadIdx.x; j < T; j += 256) { const float e = __expf(row[j] - m); row[j] = e; s += e; }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int k = 0; k < 8; ++k) s += red[k];
  const float inv = 1.0f / s;
  for (int j = threadIdx.x; j < T; j += 256) row[j] *= inv;
}
// dS = scale * P (dP - sum_j dP P), in place on dP
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int T, float scale) {
  __shared__ float red[8];
  const float* p = P + (size_t)blockIdx.x * T;
  float* d = dP + (size_t)blockIdx.x * T;
  float s = 0.f;
  for (int j = threadIdx.x; j < T; j += 256) s = fmaf(p[j], d[j], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int k = 0; k < 8; ++k) s += red[k];
  for (int j = threadIdx.x; j < T; j += 256) d[j] = scale * p[j] * (d[j] - s);
}

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// u [M][2 H] = (value | gate) -> g [M][H] = value * gelu(gate)   (GEGLU, perceiver_io.py:53-56)
__global__ void geglu_fwd_kernel(const float* __restrict__ u, long long n, int H, float* __restrict__ g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / H;
  const int c = (int)(i % H);
  g[i] = u[r * 2 * H + c] * gelu_exact(u[r * 2 * H + H + c]);
}
__global__ void geglu_bwd_kernel(const float* __restrict__ u, const float* __restrict__ dg, long long n, int H, float* __restrict__ du) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / H;
  const int c = (int)(i % H);
  const float v = u[r * 2 * H + c], x = u[r * 2 * H + H + c], d = dg[i];
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f)), pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  du[r * 2 * H + c] = d * x * cdf;
  du[r * 2 * H + H + c] = d * v * (cdf + x * pdf);
}

// out[c] += sum_r X[r][c]  (bias gradients; out zeroed by the caller)
__global__ void col_sum_kernel(const float* __restrict__ X, int rows, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = blockIdx.y * 256 + ty; r < min(blockIdx.y * 256 + 256, rows); r += 8) s += X[(size_t)r * cols + c];
  __shared__ float red[8][32];
  red[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// torch.optim.SGD(momentum, weight_decay), dampening 0, nesterov False: g = grad * grad_scale + wd p; buf = mu buf + g (buf = g on the first step);
// p -= lr buf
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ buf, long long n, float lr, float mu, float wd,
                                float grad_scale, int first) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = fmaf(wd, p[i], grad[i] * grad_scale);
  const float b = first ? g : fmaf(mu, buf[i], g);
  buf[i] = b;
  p[i] -= lr * b;
}

}  // namespace gmf
