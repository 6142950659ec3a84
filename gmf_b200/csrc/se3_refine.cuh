// DGR's robust SE(3) refinement (GMF_DeepGlobalRegistration_fcgf/core/registration.py:116-194 `Transformation` + `GlobalRegistration`, loss
// core/loss.py:42-61 `HighDimSmoothL1Loss`), SURVEY.md section 8f N3 second half.  The reference runs up to 1000 Adam iterations on the HOST, each with a
// forward, an autograd backward and several `.item()` round trips; here ONE persistent CTA per pair runs the whole loop on the device:
//   z_i = R(rot6d) x_i + t - y_i,  s_i = |z_i|^2 / q^2,  l_i = s_i < 1 ? 0.5 s_i : 0.5 (sqrt(s_i + eps) - 0.5),  L = sum w_i l_i / sum w_i    (loss.py:51-61)
//   R = ortho2rotation(rot6d): Gram-Schmidt of the two 3-vectors + cross product                                                          (:16-66)
//   Adam(lr 0.1, betas 0.9 / 0.999, eps 1e-8) with ExponentialLR(0.999); stop when L < 1e-7, or after `max_break` (cumulative) iterations whose
//   loss changed by less than `break_ratio` x previous loss, or after max_iter                                                             (:165-187)
// Per iteration every thread streams its share of the points and accumulates L, dL/dt (3) and dL/dR (3 x 3); the 13 sums are reduced in double
// precision in a fixed order (deterministic); thread 0 back-propagates through the Gram-Schmidt step analytically and applies the Adam update
// in fp32, exactly as torch does.
#pragma once
#include "common.cuh"

namespace gmf {

struct Se3Rot { float x[3], y[3], z[3], xn, yn, s; float a[3], b[3], yp[3]; };   // forward intermediates of ortho2rotation

__device__ __forceinline__ void se3_ortho(const float* p6, Se3Rot& r) {
  for (int c = 0; c < 3; ++c) { r.a[c] = p6[c]; r.b[c] = p6[3 + c]; }
  r.xn = fmaxf(sqrtf(r.a[0] * r.a[0] + r.a[1] * r.a[1] + r.a[2] * r.a[2]), 1e-8f);                  // normalize_vector (:20-27)
  for (int c = 0; c < 3; ++c) r.x[c] = r.a[c] / r.xn;
  const float ip = r.x[0] * r.b[0] + r.x[1] * r.b[1] + r.x[2] * r.b[2];                            // proj_u2a (:44-52)
  const float n2 = fmaxf(r.x[0] * r.x[0] + r.x[1] * r.x[1] + r.x[2] * r.x[2], 1e-8f);
  r.s = ip / n2;
  for (int c = 0; c < 3; ++c) r.yp[c] = r.b[c] - r.s * r.x[c];
  r.yn = fmaxf(sqrtf(r.yp[0] * r.yp[0] + r.yp[1] * r.yp[1] + r.yp[2] * r.yp[2]), 1e-8f);
  for (int c = 0; c < 3; ++c) r.y[c] = r.yp[c] / r.yn;
  r.z[0] = r.x[1] * r.y[2] - r.x[2] * r.y[1];                                                      // cross_product (:29-41)
  r.z[1] = r.x[2] * r.y[0] - r.x[0] * r.y[2];
  r.z[2] = r.x[0] * r.y[1] - r.x[1] * r.y[0];
}

// G = dL/dR (row-major, R = [x | y | z] as columns) -> gradient w.r.t. the 6 parameters
__device__ __forceinline__ void se3_ortho_backward(const Se3Rot& r, const float* G, float* g6) {
  float gx[3], gy[3], gz[3];
  for (int c = 0; c < 3; ++c) { gx[c] = G[c * 3 + 0]; gy[c] = G[c * 3 + 1]; gz[c] = G[c * 3 + 2]; }
  // z = x cross y
  gx[0] += r.y[1] * gz[2] - r.y[2] * gz[1]; gx[1] += r.y[2] * gz[0] - r.y[0] * gz[2]; gx[2] += r.y[0] * gz[1] - r.y[1] * gz[0];
  gy[0] += gz[1] * r.x[2] - gz[2] * r.x[1]; gy[1] += gz[2] * r.x[0] - gz[0] * r.x[2]; gy[2] += gz[0] * r.x[1] - gz[1] * r.x[0];
  // y = yp / |yp|
  const float ydg = r.y[0] * gy[0] + r.y[1] * gy[1] + r.y[2] * gy[2];
  float gyp[3];
  for (int c = 0; c < 3; ++c) gyp[c] = (gy[c] - r.y[c] * ydg) / r.yn;
  // yp = b - s x,  s = (x . b) / n2,  n2 = |x|^2 (both x and b carry gradient; the reference's autograd differentiates through n2 as well)
  const float n2 = fmaxf(r.x[0] * r.x[0] + r.x[1] * r.x[1] + r.x[2] * r.x[2], 1e-8f);
  const float xg = r.x[0] * gyp[0] + r.x[1] * gyp[1] + r.x[2] * gyp[2];
  for (int c = 0; c < 3; ++c) {
    g6[3 + c] = gyp[c] - xg * r.x[c] / n2;                                                          // d/db
    gx[c] += -r.s * gyp[c] - xg * (r.b[c] / n2 - 2.f * r.s * r.x[c] / n2);                          // d/dx through s and the explicit x
  }
  // x = a / |a|
  const float xdg = r.x[0] * gx[0] + r.x[1] * gx[1] + r.x[2] * gx[2];
  for (int c = 0; c < 3; ++c) g6[c] = (gx[c] - r.x[c] * xdg) / r.xn;
}

// info[pair] = (iterations = the reference's loop index `i` at exit, final loss, break count)
__global__ void __launch_bounds__(512) se3_refine_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ W, int N,
                                                         float quant, float eps, int max_iter, int max_break, float break_ratio,
                                                         const float* __restrict__ R_init, const float* __restrict__ t_init,
                                                         float* __restrict__ R_out, float* __restrict__ t_out, float* __restrict__ info) {
  __shared__ double red[16][13];
  __shared__ float sP[12];          // R (9, row-major) | t (3) of the current iterate
  __shared__ int sStop;
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = X + (size_t)pair * N * 3;
  const float* y = Y + (size_t)pair * N * 3;
  const float* w = W + (size_t)pair * N;
  // thread-0 state: parameters, Adam moments, loop bookkeeping
  float p[9], m1[9], m2[9];
  Se3Rot rot;
  float lr = 0.1f, loss_prev = 0.f, b1t = 1.f, b2t = 1.f, last_loss = 0.f;
  int brk = 0, it_exit = 0;
  double w1 = 0.0;
  if (tid == 0) {
    const float* R0 = R_init + (size_t)pair * 9;
    for (int c = 0; c < 3; ++c) { p[c] = R0[c * 3 + 0]; p[3 + c] = R0[c * 3 + 1]; p[6 + c] = t_init[(size_t)pair * 3 + c]; }   // Transformation.__init__ (:121-125)
    for (int c = 0; c < 9; ++c) { m1[c] = 0.f; m2[c] = 0.f; }
    se3_ortho(p, rot);
    for (int c = 0; c < 3; ++c) { sP[c * 3 + 0] = rot.x[c]; sP[c * 3 + 1] = rot.y[c]; sP[c * 3 + 2] = rot.z[c]; sP[9 + c] = p[6 + c]; }
    sStop = 0;
  }
  {  // w1 = weights.sum() (loss.py:49)
    double s = 0.0;
    for (int i = tid; i < N; i += blockDim.x) s += (double)w[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp][0] = s;
    __syncthreads();
    if (tid == 0) { for (int k = 0; k < 16; ++k) w1 += red[k][0]; }
  }
  __syncthreads();
  const float iq2 = 1.0f / (quant * quant);
  // iteration -1 evaluates loss_prev (:167); iterations 0 .. max_iter - 1 are the reference's loop
  for (int it = -1; it < max_iter; ++it) {
    float R[9], t[3];
    for (int c = 0; c < 9; ++c) R[c] = sP[c];
    for (int c = 0; c < 3; ++c) t[c] = sP[9 + c];
    double acc[13];
    for (int c = 0; c < 13; ++c) acc[c] = 0.0;
    for (int i = tid; i < N; i += blockDim.x) {
      const float xi[3] = {x[i * 3], x[i * 3 + 1], x[i * 3 + 2]};
      float z[3];
      for (int c = 0; c < 3; ++c) z[c] = fmaf(R[c * 3], xi[0], fmaf(R[c * 3 + 1], xi[1], fmaf(R[c * 3 + 2], xi[2], t[c]))) - y[i * 3 + c];
      const float sq = (z[0] * z[0] + z[1] * z[1] + z[2] * z[2]) * iq2;
      const float wi = w[i];
      float li, gs;                                             // loss and d loss / d sq
      if (sq < 1.f) { li = 0.5f * sq; gs = 0.5f; }
      else { const float rt = sqrtf(sq + eps); li = 0.5f * (rt - 0.5f); gs = 0.25f / rt; }   // (0.5 - use_sq_half) (sqrt(sq + eps) - 0.5), loss.py:55
      acc[0] += (double)(wi * li);
      const float k = wi * gs * 2.f * iq2;                      // d(w l)/dz = k z
      for (int c = 0; c < 3; ++c) {
        const float gz = k * z[c];
        acc[1 + c] += (double)gz;
        acc[4 + c * 3] += (double)(gz * xi[0]); acc[5 + c * 3] += (double)(gz * xi[1]); acc[6 + c * 3] += (double)(gz * xi[2]);
      }
    }
    for (int c = 0; c < 13; ++c)
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    __syncthreads();
    if (lane == 0) for (int c = 0; c < 13; ++c) red[warp][c] = acc[c];
    __syncthreads();
    if (tid == 0) {
      double tot[13];
      for (int c = 0; c < 13; ++c) { double s = 0.0; for (int k = 0; k < 16; ++k) s += red[k][c]; tot[c] = s / w1; }
      const float loss = (float)tot[0];
      last_loss = loss;
      if (it < 0) {
        loss_prev = loss;
      } else {
        it_exit = it;
        if (loss < 1e-7f) {                                     // :172
          sStop = 1;
        } else {
          float G[9], g[9];
          for (int c = 0; c < 9; ++c) G[c] = (float)tot[4 + c];
          se3_ortho_backward(rot, G, g);
          for (int c = 0; c < 3; ++c) g[6 + c] = (float)tot[1 + c];
          b1t *= 0.9f; b2t *= 0.999f;
          const float bc1 = 1.f - b1t, bc2s = sqrtf(1.f - b2t), step = lr / bc1;
          for (int c = 0; c < 9; ++c) {                         // torch.optim.Adam, single-tensor formulation
            m1[c] = 0.9f * m1[c] + 0.1f * g[c];
            m2[c] = 0.999f * m2[c] + 0.001f * g[c] * g[c];
            p[c] -= step * m1[c] / (sqrtf(m2[c]) / bc2s + 1e-8f);
          }
          lr *= 0.999f;                                         // ExponentialLR (:164, :179)
          if (fabsf(loss_prev - loss) < loss_prev * break_ratio) {   // :183-186 (the counter is cumulative, never reset)
            if (++brk >= max_break) sStop = 1;
          }
          loss_prev = loss;
          se3_ortho(p, rot);
          for (int c = 0; c < 3; ++c) { sP[c * 3 + 0] = rot.x[c]; sP[c * 3 + 1] = rot.y[c]; sP[c * 3 + 2] = rot.z[c]; sP[9 + c] = p[6 + c]; }
        }
      }
    }
    __syncthreads();
    if (sStop) break;
  }
  if (tid == 0) {
    for (int c = 0; c < 9; ++c) R_out[(size_t)pair * 9 + c] = sP[c];
    for (int c = 0; c < 3; ++c) t_out[(size_t)pair * 3 + c] = sP[9 + c];
    if (info) { info[(size_t)pair * 3] = (float)it_exit; info[(size_t)pair * 3 + 1] = last_loss; info[(size_t)pair * 3 + 2] = (float)brk; }
  }
}

}  // namespace gmf
