// Training step of the GMF-PointDSC path itself (reference: GMF_PointDSC/libs/trainer.py:123-166 - forward :134,
// losses :137-146, loss.backward() :160, optimizer.step() :168 - libs/loss.py:66-139, models/PointDSC.py:191-266 in training mode).
// Kernels of the training-mode forward (batch-statistics BatchNorm, SC-guided softmax with the compatibility matrix as a multiplicative
// factor), the loss head (BCE-with-logits + a FUSED spectral-matching loss: the N x N feature-compatibility matrix M and its gradient are
// produced tile by tile in shared memory and consumed on the spot, never written to HBM) and of the analytic backward.  Every matrix product
// of the trunk runs on the tensor pipe through the generic operand-image path of the DGR training step (mat_to_img -> img_gemm_kernel<128,
// DE_STORE>, dgr_head.cuh), here batched over the pairs of a step (blockIdx.z).  Like that step this is built for completeness of the
// training capability (gradient parity against autograd of the unmodified reference), not tuned like the inference path.
#pragma once
#include "dgr_train.cuh"

namespace gmf {

// Batched operand image: item z of `src` (element stride `stride`) -> img + z * tiles * kch * per * 4096.  Layout as mat_to_img_kernel.
// x3 (error-compensated products, "3xTF32"): every 32-wide K chunk becomes a (hi, lo) pair of chunks, hi = tf32(v), lo = tf32(v - hi);
// img_gemm_kernel<128, DE_STORE_X3> loads the pair of both operands once per K chunk and accumulates hi hi + hi lo + lo hi in fp32:
// products at ~2^-22 relative error instead of 2^-11, for three times the tensor work and twice the operand traffic.
__global__ void __launch_bounds__(256) mat_to_img_b_kernel(const float* __restrict__ src0, size_t stride, int ld, int rows, int K, int trans, int kch,
                                                           int tiles, int x3, float* __restrict__ img0) {
  const int tile = blockIdx.x, kc = blockIdx.y;
  const float* src = src0 + (size_t)blockIdx.z * stride;
  const int per = x3 ? 2 : 1;
  float* chunk0 = img0 + ((size_t)blockIdx.z * tiles * kch * per + ((size_t)tile * kch + kc) * per) * 4096;
  const bool vec = (ld & 3) == 0 && ((uintptr_t)src & 15) == 0;
  __shared__ float tbuf[32][129];                          // transposed operands: coalesced reads along r, then the row-major write pattern
  if (trans) {
    for (int idx = threadIdx.x; idx < 32 * 128; idx += 256) {
      const int kk = idx >> 7, rr = idx & 127, r = tile * 128 + rr, k = kc * 32 + kk;
      tbuf[kk][rr] = (r < rows && k < K) ? src[(size_t)k * ld + r] : 0.f;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < 1024; idx += 256) {
    const int rr = idx >> 3, g = idx & 7;
    const int r = tile * 128 + rr, k0 = kc * 32 + g * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (trans) {
      v = make_float4(tbuf[g * 4][rr], tbuf[g * 4 + 1][rr], tbuf[g * 4 + 2][rr], tbuf[g * 4 + 3][rr]);
    } else if (r < rows) {
      const float* p = src + (size_t)r * ld + k0;
      if (k0 + 3 < K && vec) v = *reinterpret_cast<const float4*>(p);
      else { if (k0 < K) v.x = p[0]; if (k0 + 1 < K) v.y = p[1]; if (k0 + 2 < K) v.z = p[2]; if (k0 + 3 < K) v.w = p[3]; }
    }
    const float4 hi = to_tf32(v);
    const uint32_t off = swz_off(rr, g);
    *reinterpret_cast<float4*>((uint8_t*)chunk0 + off) = hi;
    if (x3) {
      *reinterpret_cast<float4*>((uint8_t*)(chunk0 + 4096) + off) = to_tf32(make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
    }
  }
}

// ConvPosEnc over a batch of sequences of length L stored back to back (rows = B * L; fusion_layer.py:118-128): the stencil stops at the
// ends of every sequence.
__global__ void cpe_seq_fwd_kernel(const float* __restrict__ x, long long rows, int L, int C, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int t = (int)((i / C) % L), c = (int)(i % C);
  const float p = t > 0 ? x[i - C] : 0.f, n = t + 1 < L ? x[i + C] : 0.f;
  y[i] = x[i] + fmaf(w[c * 3], p, fmaf(w[c * 3 + 1], x[i], fmaf(w[c * 3 + 2], n, b[c])));
}
// dx[t] (+)= dy[t] + w0 dy[t+1] + w1 dy[t] + w2 dy[t-1]; dw, db by atomics.  Block = 32 channels x 8 row lanes over 64 rows.
__global__ void __launch_bounds__(256) cpe_seq_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, long long rows, int L, int C,
                                                          const float* __restrict__ w, float* __restrict__ dx, int accumulate, float* __restrict__ dw,
                                                          float* __restrict__ db) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const long long r0 = (long long)blockIdx.y * 64;
  const int ty = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
  if (c < C) {
    const float w0 = w[c * 3], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2];
    for (long long r = r0 + ty; r < min(r0 + 64, rows); r += 8) {
      const int t = (int)(r % L);
      const size_t i = (size_t)r * C + c;
      const float g = dy[i], gn = t + 1 < L ? dy[i + C] : 0.f, gp = t > 0 ? dy[i - C] : 0.f;
      if (dx) { const float v = g + fmaf(w0, gn, fmaf(w1, g, w2 * gp)); dx[i] = accumulate ? dx[i] + v : v; }
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

// out[c] += sum_r X[r * ld + c]  (bias gradients of column blocks of a wider matrix; out zeroed by the caller)
__global__ void __launch_bounds__(256) col_sum_ld_kernel(const float* __restrict__ X, int ld, int rows, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = blockIdx.y * 256 + ty; r < min(blockIdx.y * 256 + 256, rows); r += 8) s += X[(size_t)r * ld + c];
  __shared__ float red[8][32];
  red[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// out[i] = sum_z part[z * n + i]  (split-K reduction of the weight-gradient products)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int slices, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < slices; ++z) s += part[(size_t)z * n + i];
  out[i] = s;
}

// y += x (n elements)
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}
__global__ void relu_inplace_kernel(float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(y[i], 0.f);
}
// dx = dy * (y > 0), in place on dy
__global__ void relu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(y[i] > 0.f)) dy[i] = 0.f;
}

// ---- BatchNorm1d in training mode over all rows of a step (nn.BatchNorm1d, eps 1e-5, momentum 0.1; PointDSC.py:15,18,107) + ReLU ----
// acc[0..C) += sum_r z, acc[C..2C) += sum_r z^2 (double; zeroed by the caller).  Block = 32 channels x 8 row lanes over 256 rows.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ z, int rows, int C, double* __restrict__ acc) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  float s = 0.f, q = 0.f;
  if (c < C)
    for (int r = blockIdx.y * 256 + ty; r < min(blockIdx.y * 256 + 256, rows); r += 8) { const float v = z[(size_t)r * C + c]; s += v; q = fmaf(v, v, q); }
  __shared__ float red[8][32][2];
  red[ty][threadIdx.x & 31][0] = s; red[ty][threadIdx.x & 31][1] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 8; ++k) { a += red[k][threadIdx.x][0]; b += red[k][threadIdx.x][1]; }
    atomicAdd(acc + c, (double)a); atomicAdd(acc + C + c, (double)b);
  }
}
// stat[c] = mean, stat[C + c] = rstd (biased variance); running statistics updated like torch (unbiased variance, momentum 0.1)
__global__ void bn_finalize_kernel(const double* __restrict__ acc, int rows, int C, float* __restrict__ stat, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = acc[c] / rows, var = fmax(acc[C + c] / rows - mean * mean, 0.0);
  stat[c] = (float)mean;
  stat[C + c] = (float)(1.0 / sqrt(var + 1e-5));
  if (running_mean) {
    running_mean[c] = 0.9f * running_mean[c] + 0.1f * (float)mean;
    running_var[c] = 0.9f * running_var[c] + 0.1f * (float)(rows > 1 ? var * rows / (rows - 1) : var);
  }
}
// a = relu(gamma (z - mean) rstd + beta)
__global__ void bn_relu_fwd_kernel(const float* __restrict__ z, long long n, int C, const float* __restrict__ stat, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  a[i] = fmaxf(fmaf((z[i] - stat[c]) * stat[C + c], gamma[c], beta[c]), 0.f);
}
// dy = da (a > 0); acc[c] += sum dy, acc[C + c] += sum dy xhat (double; zeroed by the caller)
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ a, const float* __restrict__ z, int rows,
                                                                 int C, const float* __restrict__ stat, double* __restrict__ acc) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
  float s = 0.f, q = 0.f;
  if (c < C) {
    const float mean = stat[c], rstd = stat[C + c];
    for (int r = blockIdx.y * 256 + ty; r < min(blockIdx.y * 256 + 256, rows); r += 8) {
      const size_t i = (size_t)r * C + c;
      const float dy = a[i] > 0.f ? da[i] : 0.f;
      s += dy; q = fmaf(dy, (z[i] - mean) * rstd, q);
    }
  }
  __shared__ float red[8][32][2];
  red[ty][threadIdx.x & 31][0] = s; red[ty][threadIdx.x & 31][1] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    float u = 0.f, v = 0.f;
    for (int k = 0; k < 8; ++k) { u += red[k][threadIdx.x][0]; v += red[k][threadIdx.x][1]; }
    atomicAdd(acc + c, (double)u); atomicAdd(acc + C + c, (double)v);
  }
}
// dz = gamma rstd (dy - mean(dy) - xhat mean(dy xhat)); dgamma = sum dy xhat, dbeta = sum dy (written by the first block)
__global__ void bn_relu_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ a, const float* __restrict__ z, long long n, int rows, int C,
                                         const float* __restrict__ stat, const float* __restrict__ gamma, const double* __restrict__ acc,
                                         float* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { dbeta[i] = (float)acc[i]; dgamma[i] = (float)acc[C + i]; }
  if (i >= n) return;
  const int c = (int)(i % C);
  const float rstd = stat[C + c], xh = (z[i] - stat[c]) * rstd;
  const float dy = a[i] > 0.f ? da[i] : 0.f;
  const float m1 = (float)(acc[c] / rows), m2 = (float)(acc[C + c] / rows);
  dz[i] = gamma[c] * rstd * (dy - m1 - xh * m2);
}

// ---- spatial-consistency matrix (PointDSC.py:216-221), materialised for the training step only (B N^2 floats at the training sizes) ----
__global__ void compat_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int N, const float* __restrict__ sigma_spat, float* __restrict__ c) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
  if (j >= N) return;
  const float* s = src + (size_t)b * N * 3;
  const float* t = tgt + (size_t)b * N * 3;
  const float sx = s[i * 3] - s[j * 3], sy = s[i * 3 + 1] - s[j * 3 + 1], sz = s[i * 3 + 2] - s[j * 3 + 2];
  const float tx = t[i * 3] - t[j * 3], ty = t[i * 3 + 1] - t[j * 3 + 1], tz = t[i * 3 + 2] - t[j * 3 + 2];
  const float d = sqrtf(sx * sx + sy * sy + sz * sz) - sqrtf(tx * tx + ty * ty + tz * tz);
  const float sg = sigma_spat[0];
  c[((size_t)b * N + i) * N + j] = fmaxf(1.0f - d * d / (sg * sg), 0.f);
}
// row softmax in place with an optional multiplicative factor: P = softmax(c * S) (PointDSC.py:62).  One CTA per row.
__global__ void __launch_bounds__(256) softmax_mul_rows_kernel(float* __restrict__ S, const float* __restrict__ c, int T) {
  __shared__ float red[8];
  float* row = S + (size_t)blockIdx.x * T;
  const float* cr = c ? c + (size_t)blockIdx.x * T : nullptr;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < T; j += 256) { const float v = cr ? cr[j] * row[j] : row[j]; row[j] = v; m = fmaxf(m, v); }
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
// dS = scale * c * P (dP - sum_j dP P), in place on dP (c optional)
__global__ void __launch_bounds__(256) softmax_mul_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, const float* __restrict__ c, int T, float scale) {
  __shared__ float red[8];
  const float* p = P + (size_t)blockIdx.x * T;
  const float* cr = c ? c + (size_t)blockIdx.x * T : nullptr;
  float* d = dP + (size_t)blockIdx.x * T;
  float s = 0.f;
  for (int j = threadIdx.x; j < T; j += 256) s = fmaf(p[j], d[j], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int k = 0; k < 8; ++k) s += red[k];
  for (int j = threadIdx.x; j < T; j += 256) d[j] = scale * (cr ? cr[j] : 1.0f) * p[j] * (d[j] - s);
}

// ---- loss head ----
// F.normalize(p=2, dim=-1, eps 1e-12) (PointDSC.py:229): fh [R][128], fhT [B][128][Np] (zero padded to Np = multiple of 64), inv[r] = 1 / max(|f|, eps)
__global__ void __launch_bounds__(256) normalize_fwd_kernel(const float* __restrict__ f, int B, int N, int Np, float* __restrict__ fh, float* __restrict__ fhT,
                                                            float* __restrict__ inv) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= B * N) return;
  const float4 v = *reinterpret_cast<const float4*>(f + (size_t)r * 128 + lane * 4);
  const float nrm = sqrtf(warp_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w));
  const float s = 1.0f / fmaxf(nrm, 1e-12f);
  const float4 o = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
  *reinterpret_cast<float4*>(fh + (size_t)r * 128 + lane * 4) = o;
  const int b = r / N, i = r % N;
  float* t = fhT + ((size_t)b * 128 + lane * 4) * Np + i;
  t[0] = o.x; t[Np] = o.y; t[2 * (size_t)Np] = o.z; t[3 * (size_t)Np] = o.w;
  if (lane == 0) inv[r] = s;
}
// df (+)= inv (dfh - fh <fh, dfh>)
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ dfh, const float* __restrict__ fh, const float* __restrict__ inv, int rows,
                                                            float* __restrict__ df, int accumulate) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float4 d = *reinterpret_cast<const float4*>(dfh + (size_t)r * 128 + lane * 4);
  const float4 h = *reinterpret_cast<const float4*>(fh + (size_t)r * 128 + lane * 4);
  const float dot = warp_sum(d.x * h.x + d.y * h.y + d.z * h.z + d.w * h.w), s = inv[r];
  float4 o = make_float4(s * (d.x - h.x * dot), s * (d.y - h.y * dot), s * (d.z - h.z * dot), s * (d.w - h.w * dot));
  float4* dst = reinterpret_cast<float4*>(df + (size_t)r * 128 + lane * 4);
  if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
  *dst = o;
}
// cnt[b] = number of positive labels of pair b, cnt[B] = total (double, zeroed by the caller)
__global__ void label_count_kernel(const float* __restrict__ gt, int B, int N, double* __restrict__ cnt) {
  const int b = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s += gt[(size_t)b * N + i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0 && s != 0.f) { atomicAdd(cnt + b, (double)s); atomicAdd(cnt + B, (double)s); }
}
// ClassificationLoss (libs/loss.py:66-98): mean BCE-with-logits, pos_weight = num_neg / num_pos over the whole batch when balanced.
// loss_acc[0] += sum of the per-element losses / n; dlogit = weight * d(mean loss) / d logit
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ logit, const float* __restrict__ gt, long long n, int balanced, const double* __restrict__ cnt_total,
                                                  float weight, double* __restrict__ loss_acc, float* __restrict__ dlogit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float pw = 1.0f;
  if (balanced) {
    const double pos = cnt_total[0], neg = (double)n - pos;
    pw = (float)((fmax(neg - 1.0, 0.0) + 1.0) / (fmax(pos - 1.0, 0.0) + 1.0));
  }
  float l = 0.f;
  if (i < n) {
    const float x = logit[i], y = gt[i];
    const float sp = fmaxf(-x, 0.f) + log1pf(__expf(-fabsf(x)));        // softplus(-x)
    const float wy = 1.0f + (pw - 1.0f) * y;
    l = (1.0f - y) * x + wy * sp;
    const float sig = 1.0f / (1.0f + __expf(-x));
    dlogit[i] = weight * ((1.0f - y) - wy * (1.0f - sig)) / (float)n;
  }
  __shared__ float red[8];
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(loss_acc, (double)s / (double)n);
  }
}

// Fused SpectralMatchingLoss (libs/loss.py:118-139) on M = clamp(1 - (1 - fh fh^T) / sigma^2, 0, 1) with a zero diagonal (PointDSC.py:231-234):
// one CTA owns 64 rows i of a pair and walks the 64-column tiles j; per tile S = fh_i fh_j^T (fp32 FMA from shared memory), the loss terms, G =
// d loss / d S, and dfh_i += 2 G fh_j (M and gt_M are symmetric, so the column contribution equals the row contribution).  Nothing N x N is stored.
//   balanced: loss_b = 0.5 sum (M-1)^2 gtM / np + 0.5 sum M^2 (1-gtM) / nn, np = relu(k(k-1) - 1) + 1, nn = relu(N^2 - k(k-1) - 1) + 1; mean over pairs
//   else:     MSE mean over B N^2
// acc[0] += loss, acc[1] += d loss / d sigma (doubles).  dfh gets weight * d loss / d fh.  Thread (ty, tx) = (tid / 16, tid % 16).
constexpr int kSmlSmem = (128 * 64 * 2 + 64 * 128 + 64 * 64) * 4;
__global__ void __launch_bounds__(256) sm_loss_fused_kernel(const float* __restrict__ fh, const float* __restrict__ fhT, const float* __restrict__ gt, int B, int N, int Np,
                                                            const float* __restrict__ sigma, int balanced, const double* __restrict__ cnt, float weight,
                                                            double* __restrict__ acc, float* __restrict__ dfh) {
  extern __shared__ __align__(16) float sml[];
  float* sAT = sml;                    // [128][64] rows i, k-major
  float* sBT = sAT + 128 * 64;         // [128][64] rows j, k-major
  float* sB = sBT + 128 * 64;          // [64][128] rows j, row-major
  float* sG = sB + 64 * 128;           // [64][64]
  __shared__ float sgi[64], sgj[64];
  __shared__ double redd[16];
  const int b = blockIdx.y, i0 = blockIdx.x * 64, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const float sg = sigma[0], is2 = 1.0f / (sg * sg), dsc = 2.0f * is2 / sg;          // d M / d sigma = dsc (1 - S)
  const double k = cnt[b], kk = k * (k - 1.0);
  float cp, cn;                        // d loss / d M = cp (M - 1) on positive pairs, cn M on the others (before `weight`)
  if (balanced) {
    cp = (float)(1.0 / ((fmax(kk - 1.0, 0.0) + 1.0) * B));
    cn = (float)(1.0 / ((fmax((double)N * N - kk - 1.0, 0.0) + 1.0) * B));
  } else {
    cp = cn = (float)(2.0 / ((double)B * N * N));
  }
  const float* fT = fhT + (size_t)b * 128 * Np;
  for (int idx = tid; idx < 128 * 16; idx += 256) {
    const int kq = idx >> 4, c4 = idx & 15;
    *reinterpret_cast<float4*>(sAT + kq * 64 + c4 * 4) = *reinterpret_cast<const float4*>(fT + (size_t)kq * Np + i0 + c4 * 4);
  }
  if (tid < 64) sgi[tid] = i0 + tid < N ? gt[(size_t)b * N + i0 + tid] : 0.f;
  float dacc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) dacc[r][c] = 0.f;
  float lsum = 0.f, dsig = 0.f;
  for (int j0 = 0; j0 < Np; j0 += 64) {
    __syncthreads();
    for (int idx = tid; idx < 128 * 16; idx += 256) {
      const int kq = idx >> 4, c4 = idx & 15;
      *reinterpret_cast<float4*>(sBT + kq * 64 + c4 * 4) = *reinterpret_cast<const float4*>(fT + (size_t)kq * Np + j0 + c4 * 4);
    }
    for (int idx = tid; idx < 64 * 32; idx += 256) {
      const int r = idx >> 5, c4 = idx & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + r < N) v = *reinterpret_cast<const float4*>(fh + ((size_t)b * N + j0 + r) * 128 + c4 * 4);
      *reinterpret_cast<float4*>(sB + r * 128 + c4 * 4) = v;
    }
    if (tid < 64) sgj[tid] = j0 + tid < N ? gt[(size_t)b * N + j0 + tid] : 0.f;
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
#pragma unroll 8
    for (int kq = 0; kq < 128; ++kq) {
      const float4 a = *reinterpret_cast<const float4*>(sAT + kq * 64 + ty * 4);
      const float4 bb = *reinterpret_cast<const float4*>(sBT + kq * 64 + tx * 4);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[r][c] = fmaf(av[r], bv[c], s[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty * 4 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        float g = 0.f;
        if (i < N && j < N && i != j) {
          const float pre = 1.0f - (1.0f - s[r][c]) * is2;
          const float m = fminf(fmaxf(pre, 0.f), 1.0f);
          const bool pos = sgi[ty * 4 + r] + sgj[tx * 4 + c] == 2.0f;
          const float e = pos ? m - 1.0f : m;
          const float dm = (pos ? cp : cn) * e;                        // d loss / d M
          lsum = fmaf(0.5f * dm, e, lsum);                             // balanced: 0.5 c e^2; MSE: (2 / (B N^2)) / 2 * e^2
          if (pre >= 0.f && pre <= 1.0f) { g = dm * is2; dsig = fmaf(dm * dsc, 1.0f - s[r][c], dsig); }
        }
        sG[(ty * 4 + r) * 64 + tx * 4 + c] = g;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < 64; ++j) {
      const float4 b0 = *reinterpret_cast<const float4*>(sB + j * 128 + tx * 8), b1 = *reinterpret_cast<const float4*>(sB + j * 128 + tx * 8 + 4);
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float g = sG[(ty * 4 + r) * 64 + j];
#pragma unroll
        for (int c = 0; c < 8; ++c) dacc[r][c] = fmaf(g, bv[c], dacc[r][c]);
      }
    }
  }
  const float w2 = 2.0f * weight;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i < N) {
      float* d = dfh + ((size_t)b * N + i) * 128 + tx * 8;
      *reinterpret_cast<float4*>(d) = make_float4(w2 * dacc[r][0], w2 * dacc[r][1], w2 * dacc[r][2], w2 * dacc[r][3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(w2 * dacc[r][4], w2 * dacc[r][5], w2 * dacc[r][6], w2 * dacc[r][7]);
    }
  }
  lsum = warp_sum(lsum); dsig = warp_sum(dsig);
  __syncthreads();
  if ((tid & 31) == 0) { redd[tid >> 5] = (double)lsum; redd[8 + (tid >> 5)] = (double)dsig; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, d = 0.0;
    for (int q = 0; q < 8; ++q) { a += redd[q]; d += redd[8 + q]; }
    atomicAdd(acc, a); atomicAdd(acc + 1, d);
  }
}
// out[0] = class loss, out[1] = SM loss, out[2] = w_c class + w_sm sm; grad of sigma (a learnable scalar, PointDSC.py:164)
__global__ void loss_finalize_kernel(const double* __restrict__ acc, float w_class, float w_sm, float* __restrict__ out, float* __restrict__ dsigma) {
  if (out) { out[0] = (float)acc[2]; out[1] = (float)acc[0]; out[2] = (float)(w_class * acc[2] + w_sm * acc[0]); }
  if (dsigma) dsigma[0] = (float)(w_sm * acc[1]);
}

// Training-mode output M of PointDSC.forward (PointDSC.py:231-234) from S = fh fh^T, in place: clamp(1 - (1 - S) / sigma^2, 0, 1), zero diagonal
__global__ void m_from_s_kernel(float* __restrict__ S, int N, long long n, const float* __restrict__ sigma) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long e = i % ((long long)N * N);
  const int r = (int)(e / N), c = (int)(e % N);
  const float sg = sigma[0];
  S[i] = r == c ? 0.f : fminf(fmaxf(1.0f - (1.0f - S[i]) / (sg * sg), 0.f), 1.0f);
}
// autograd entry (d loss / d M given by the caller): S -> G = d loss / d S in place (clamp pass-through inside [0, 1], zero diagonal);
// acc[0] += d loss / d sigma (double)
__global__ void __launch_bounds__(256) ds_from_dm_kernel(float* __restrict__ S, const float* __restrict__ dM, int N, long long n, const float* __restrict__ sigma,
                                                         double* __restrict__ acc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float sg = sigma[0], is2 = 1.0f / (sg * sg);
  float ds = 0.f;
  if (i < n) {
    const long long e = i % ((long long)N * N);
    const int r = (int)(e / N), c = (int)(e % N);
    const float s = S[i], pre = 1.0f - (1.0f - s) * is2;
    float g = 0.f;
    if (r != c && pre >= 0.f && pre <= 1.0f) { g = dM[i] * is2; ds = dM[i] * 2.0f * (1.0f - s) * is2 / sg; }
    S[i] = g;
  }
  __shared__ float red[8];
  ds = warp_sum(ds);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    if (t != 0.f) atomicAdd(acc, (double)t);
  }
}
__global__ void store_double_as_float_kernel(const double* __restrict__ src, float* __restrict__ dst) { dst[0] = (float)src[0]; }

// torch.optim.Adam (amsgrad False): g = grad_scale * grad + weight_decay * p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          (train_3DMatch.py:52-58: lr 1e-4, weight_decay 1e-6)
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                 const unsigned char* __restrict__ mask, long long n, float lr, float b1, float b2, float eps, float wd, float grad_scale,
                                 float bc1, float bc2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (mask && !mask[i])) return;
  const float g = fmaf(wd, p[i], grad[i] * grad_scale);
  const float mi = fmaf(b1, m[i], (1.0f - b1) * g), vi = fmaf(b2, v[i], (1.0f - b2) * g * g);
  m[i] = mi; v[i] = vi;
  p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
}

}  // namespace gmf
