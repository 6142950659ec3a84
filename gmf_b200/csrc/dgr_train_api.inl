// C-ABI of the DGR head's training step (include/gmf_b200.h, "DGR head training"); included at the end of gmf_api.cu after dgr_head_api.inl.
// Parameters, gradients and momentum are FLAT fp32 device buffers in the order of gmf_dgr_head_weight_spec (the reference state_dict order),
// owned by the caller (gmf_b200/dgr_head.py keeps them as torch tensors so that torch.distributed can all-reduce the gradient in one call).

namespace {

struct DgrP {                 // offsets (floats) into the flat parameter / gradient buffer
  size_t cqw, cqb, ccw, ccb, lqg, lqb, lcg, lcb, wq, wkv, wo, bo, lfg, lfb, w1, b1, w2, b2, total;
};
DgrP dgr_offsets(bool pe) {
  DgrP p{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += n; return r; };
  if (pe) { p.cqw = take(kDgrLatent * 3); p.cqb = take(kDgrLatent); p.ccw = take(kDgrCtx * 3); p.ccb = take(kDgrCtx); }
  p.lqg = take(kDgrLatent); p.lqb = take(kDgrLatent); p.lcg = take(kDgrCtx); p.lcb = take(kDgrCtx);
  p.wq = take((size_t)kDgrHead * kDgrLatent); p.wkv = take((size_t)2 * kDgrHead * kDgrCtx); p.wo = take((size_t)kDgrLatent * kDgrHead); p.bo = take(kDgrLatent);
  p.lfg = take(kDgrLatent); p.lfb = take(kDgrLatent);
  p.w1 = take((size_t)2 * kDgrHidden * kDgrLatent); p.b1 = take(2 * kDgrHidden); p.w2 = take((size_t)kDgrLatent * kDgrHidden); p.b2 = take(kDgrLatent);
  p.total = o;
  return p;
}

struct DgrTrainWs {
  float *x0, *xn, *x1, *h, *dx1, *dx0, *dxn, *dh;      // [M, 256]
  float *q, *a, *dq, *da;                               // [M, 128]
  float *c0, *cn, *dc0, *dcn;                           // [T, 128]
  float *kv, *dkv;                                      // [T, 256]
  float *u, *du;                                        // [M, 2048]
  float *g, *dg;                                        // [M, 1024]
  float *P, *dP;                                        // [M, T]
  float *stq, *stc, *stf;                               // LayerNorm (mean, rstd) per row
  float *imgA, *imgB;
  size_t img_cap;
};
size_t dgr_train_carve(DgrTrainWs& w, uint8_t* base, int M, int T) {
  Bump b{base};
  const size_t m = M, t = T;
  for (float** p : {&w.x0, &w.xn, &w.x1, &w.h, &w.dx1, &w.dx0, &w.dxn, &w.dh}) *p = b.take<float>(m * kDgrLatent);
  for (float** p : {&w.q, &w.a, &w.dq, &w.da}) *p = b.take<float>(m * kDgrHead);
  for (float** p : {&w.c0, &w.cn, &w.dc0, &w.dcn}) *p = b.take<float>(t * kDgrCtx);
  w.kv = b.take<float>(t * 2 * kDgrHead); w.dkv = b.take<float>(t * 2 * kDgrHead);
  w.u = b.take<float>(m * 2 * kDgrHidden); w.du = b.take<float>(m * 2 * kDgrHidden);
  w.g = b.take<float>(m * kDgrHidden); w.dg = b.take<float>(m * kDgrHidden);
  w.P = b.take<float>(m * t); w.dP = b.take<float>(m * t);
  w.stq = b.take<float>(2 * m); w.stc = b.take<float>(2 * t); w.stf = b.take<float>(2 * m);
  auto p128 = [](size_t v) { return (v + 127) / 128 * 128; };
  auto p32 = [](size_t v) { return (v + 31) / 32 * 32; };
  size_t cap = 0;
  for (size_t c : {p128(m) * p32(t), p128(t) * p32(m), (size_t)2 * kDgrHidden * p32(m), p128(m) * 2 * kDgrHidden, p128(t) * (size_t)256,
                   (size_t)2 * kDgrHidden * kDgrLatent})
    cap = std::max(cap, c);
  w.img_cap = cap;
  w.imgA = b.take<float>(cap); w.imgB = b.take<float>(cap);
  return b.off + 1024;
}

// out[m, n] (ldo) = scale * X Y^T (+ bias[col]) (+ residual): X is m x K, Y is n x K, each given row-major or as the transpose of a row-major matrix
int tgemm(DgrTrainWs& w, const float* X, int ldx, int m, int K, int tx, const float* Y, int ldy, int n, int ty, float* out, int ldo, float scale,
          const float* bias, const float* residual, cudaStream_t st) {
  const int tm = cdiv(m, 128), tn = cdiv(n, 128), kch = cdiv(K, 32);
  if ((size_t)tm * kch * 4096 > w.img_cap || (size_t)tn * kch * 4096 > w.img_cap) return fail(GMF_ERR_STATE, "dgr training: operand image exceeds the workspace");
  mat_to_img_kernel<<<dim3(tm, kch), 256, 0, st>>>(X, ldx, m, K, tx, kch, w.imgA);
  LAUNCHED();
  mat_to_img_kernel<<<dim3(tn, kch), 256, 0, st>>>(Y, ldy, n, K, ty, kch, w.imgB);
  LAUNCHED();
  ImgGemmArgs a{};
  a.a_img = w.imgA; a.w_packed = w.imgB; a.K = kch * 32; a.L = m; a.tiles = tm; a.out = out; a.ld = ldo; a.ncols = n; a.scale = scale; a.bias = bias;
  a.residual = residual;
  cudaError_t e = launch_img_gemm<128, DE_STORE>(a, tn, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "dgr training GEMM launch");
  return 0;
}

int dgr_train_ws(DgrTrainWs& w, void* ws, size_t bytes, int M, int T) {
  if (!ws) return fail(GMF_ERR_INVALID, "workspace is NULL");
  const size_t need = dgr_train_carve(w, nullptr, M, T) + 1024;
  if (bytes < need) return fail(GMF_ERR_STATE, "workspace too small: need " + std::to_string(need) + " bytes");
  dgr_train_carve(w, (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023), M, T);
  return 0;
}
constexpr float kDgrScale = 0.08838834764831845f;             // 128 ** -0.5 (perceiver_io.py:76)

}  // namespace

extern "C" {

size_t gmf_dgr_head_train_workspace_bytes(int M, int T) {
  if (M < 1 || T < 1) return 0;
  DgrTrainWs w;
  return dgr_train_carve(w, nullptr, M, T) + 2048;
}
int64_t gmf_dgr_head_param_count(int pe) { return (int64_t)dgr_offsets(pe != 0).total; }

int gmf_dgr_head_train_forward(gmf_dgr_head* h, const float* params, const float* latents, const float* image_feat, int M, int T, float* out,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !params || !latents || !image_feat || !out) return fail(GMF_ERR_INVALID, "NULL argument");
  if (M < 1 || T < 1) return fail(GMF_ERR_INVALID, "gmf_dgr_head_train_forward: M and T must be >= 1");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  DgrTrainWs w;
  TRY(dgr_train_ws(w, workspace, workspace_bytes, M, T));
  const DgrP o = dgr_offsets(h->pe);
  const float* p = params;
  const long long nm = (long long)M * kDgrLatent, nt = (long long)T * kDgrCtx;
  if (h->pe) {
    cpe_fwd_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(latents, M, kDgrLatent, p + o.cqw, p + o.cqb, w.x0); LAUNCHED();
    cpe_fwd_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(image_feat, T, kDgrCtx, p + o.ccw, p + o.ccb, w.c0); LAUNCHED();
  } else {
    CU(cudaMemcpyAsync(w.x0, latents, nm * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(w.c0, image_feat, nt * 4, cudaMemcpyDeviceToDevice, st));
  }
  ln_fwd_kernel<kDgrLatent><<<cdiv(M, 8), 256, 0, st>>>(w.x0, M, p + o.lqg, p + o.lqb, w.xn, w.stq); LAUNCHED();
  ln_fwd_kernel<kDgrCtx><<<cdiv(T, 8), 256, 0, st>>>(w.c0, T, p + o.lcg, p + o.lcb, w.cn, w.stc); LAUNCHED();
  TRY(tgemm(w, w.xn, kDgrLatent, M, kDgrLatent, 0, p + o.wq, kDgrLatent, kDgrHead, 0, w.q, kDgrHead, 1.f, nullptr, nullptr, st));
  TRY(tgemm(w, w.cn, kDgrCtx, T, kDgrCtx, 0, p + o.wkv, kDgrCtx, 2 * kDgrHead, 0, w.kv, 2 * kDgrHead, 1.f, nullptr, nullptr, st));
  TRY(tgemm(w, w.q, kDgrHead, M, kDgrHead, 0, w.kv, 2 * kDgrHead, T, 0, w.P, T, kDgrScale, nullptr, nullptr, st));      // S = q k^T / sqrt(d)
  softmax_rows_kernel<<<M, 256, 0, st>>>(w.P, T); LAUNCHED();
  TRY(tgemm(w, w.P, T, M, T, 0, w.kv + kDgrHead, 2 * kDgrHead, kDgrHead, 1, w.a, kDgrHead, 1.f, nullptr, nullptr, st));   // a = P v
  TRY(tgemm(w, w.a, kDgrHead, M, kDgrHead, 0, p + o.wo, kDgrHead, kDgrLatent, 0, w.x1, kDgrLatent, 1.f, p + o.bo, w.x0, st));
  ln_fwd_kernel<kDgrLatent><<<cdiv(M, 8), 256, 0, st>>>(w.x1, M, p + o.lfg, p + o.lfb, w.h, w.stf); LAUNCHED();
  TRY(tgemm(w, w.h, kDgrLatent, M, kDgrLatent, 0, p + o.w1, kDgrLatent, 2 * kDgrHidden, 0, w.u, 2 * kDgrHidden, 1.f, p + o.b1, nullptr, st));
  const long long ng = (long long)M * kDgrHidden;
  geglu_fwd_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(w.u, ng, kDgrHidden, w.g); LAUNCHED();
  // out = x1 + g W2^T + b2 (residual read from x1; `out` has the same [M, 256] layout)
  CU(cudaMemcpyAsync(out, w.x1, nm * 4, cudaMemcpyDeviceToDevice, st));
  TRY(tgemm(w, w.g, kDgrHidden, M, kDgrHidden, 0, p + o.w2, kDgrHidden, kDgrLatent, 0, out, kDgrLatent, 1.f, p + o.b2, out, st));
  return 0;
}

// Uses the activations gmf_dgr_head_train_forward left in `workspace` (same M, T, params).  grads: flat, overwritten.  d_latents [M,256] and
// d_image_feat [T,128] may be NULL.
int gmf_dgr_head_train_backward(gmf_dgr_head* h, const float* params, const float* latents, const float* image_feat, const float* d_out, int M, int T,
                                float* d_latents, float* d_image_feat, float* grads, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !params || !latents || !image_feat || !d_out || !grads) return fail(GMF_ERR_INVALID, "NULL argument");
  if (M < 1 || T < 1) return fail(GMF_ERR_INVALID, "gmf_dgr_head_train_backward: M and T must be >= 1");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  DgrTrainWs w;
  TRY(dgr_train_ws(w, workspace, workspace_bytes, M, T));
  const DgrP o = dgr_offsets(h->pe);
  const float* p = params;
  float* G = grads;
  CU(cudaMemsetAsync(G, 0, o.total * sizeof(float), st));
  auto colsum = [&](const float* X, int rows, int cols, float* out) -> int {
    col_sum_kernel<<<dim3(cdiv(cols, 32), cdiv(rows, 256)), 256, 0, st>>>(X, rows, cols, out);
    LAUNCHED();
    return 0;
  };
  // ---- FFN: out = x1 + g W2^T + b2
  TRY(colsum(d_out, M, kDgrLatent, G + o.b2));
  TRY(tgemm(w, d_out, kDgrLatent, M, kDgrLatent, 0, p + o.w2, kDgrHidden, kDgrHidden, 1, w.dg, kDgrHidden, 1.f, nullptr, nullptr, st));          // dg = dOut W2
  TRY(tgemm(w, d_out, kDgrLatent, kDgrLatent, M, 1, w.g, kDgrHidden, kDgrHidden, 1, G + o.w2, kDgrHidden, 1.f, nullptr, nullptr, st));           // dW2 = dOut^T g
  const long long ng = (long long)M * kDgrHidden;
  geglu_bwd_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(w.u, w.dg, ng, kDgrHidden, w.du); LAUNCHED();
  TRY(colsum(w.du, M, 2 * kDgrHidden, G + o.b1));
  TRY(tgemm(w, w.du, 2 * kDgrHidden, M, 2 * kDgrHidden, 0, p + o.w1, kDgrLatent, kDgrLatent, 1, w.dh, kDgrLatent, 1.f, nullptr, nullptr, st));   // dh = dU W1
  TRY(tgemm(w, w.du, 2 * kDgrHidden, 2 * kDgrHidden, M, 1, w.h, kDgrLatent, kDgrLatent, 1, G + o.w1, kDgrLatent, 1.f, nullptr, nullptr, st));    // dW1 = dU^T h
  ln_bwd_kernel<kDgrLatent><<<cdiv(M, 64), 256, 0, st>>>(w.dh, w.x1, w.stf, M, p + o.lfg, d_out, w.dx1, G + o.lfg, G + o.lfb); LAUNCHED();   // dx1 = dOut + LN'(dh)
  // ---- attention output projection: x1 = x0 + a Wo^T + bo
  TRY(colsum(w.dx1, M, kDgrLatent, G + o.bo));
  TRY(tgemm(w, w.dx1, kDgrLatent, M, kDgrLatent, 0, p + o.wo, kDgrHead, kDgrHead, 1, w.da, kDgrHead, 1.f, nullptr, nullptr, st));                // da = dx1 Wo
  TRY(tgemm(w, w.dx1, kDgrLatent, kDgrLatent, M, 1, w.a, kDgrHead, kDgrHead, 1, G + o.wo, kDgrHead, 1.f, nullptr, nullptr, st));                 // dWo = dx1^T a
  // ---- attention: a = P v, P = softmax(scale q k^T)
  TRY(tgemm(w, w.da, kDgrHead, M, kDgrHead, 0, w.kv + kDgrHead, 2 * kDgrHead, T, 0, w.dP, T, 1.f, nullptr, nullptr, st));                         // dP = da v^T
  TRY(tgemm(w, w.P, T, T, M, 1, w.da, kDgrHead, kDgrHead, 1, w.dkv + kDgrHead, 2 * kDgrHead, 1.f, nullptr, nullptr, st));                         // dv = P^T da
  softmax_bwd_kernel<<<M, 256, 0, st>>>(w.P, w.dP, T, kDgrScale); LAUNCHED();                                                                    // dS (scaled)
  TRY(tgemm(w, w.dP, T, M, T, 0, w.kv, 2 * kDgrHead, kDgrHead, 1, w.dq, kDgrHead, 1.f, nullptr, nullptr, st));                                    // dq = dS k
  TRY(tgemm(w, w.dP, T, T, M, 1, w.q, kDgrHead, kDgrHead, 1, w.dkv, 2 * kDgrHead, 1.f, nullptr, nullptr, st));                                    // dk = dS^T q
  // ---- projections
  TRY(tgemm(w, w.dq, kDgrHead, M, kDgrHead, 0, p + o.wq, kDgrLatent, kDgrLatent, 1, w.dxn, kDgrLatent, 1.f, nullptr, nullptr, st));              // dxn = dq Wq
  TRY(tgemm(w, w.dq, kDgrHead, kDgrHead, M, 1, w.xn, kDgrLatent, kDgrLatent, 1, G + o.wq, kDgrLatent, 1.f, nullptr, nullptr, st));               // dWq = dq^T xn
  TRY(tgemm(w, w.dkv, 2 * kDgrHead, T, 2 * kDgrHead, 0, p + o.wkv, kDgrCtx, kDgrCtx, 1, w.dcn, kDgrCtx, 1.f, nullptr, nullptr, st));             // dcn = dkv Wkv
  TRY(tgemm(w, w.dkv, 2 * kDgrHead, 2 * kDgrHead, T, 1, w.cn, kDgrCtx, kDgrCtx, 1, G + o.wkv, kDgrCtx, 1.f, nullptr, nullptr, st));              // dWkv = dkv^T cn
  ln_bwd_kernel<kDgrLatent><<<cdiv(M, 64), 256, 0, st>>>(w.dxn, w.x0, w.stq, M, p + o.lqg, w.dx1, w.dx0, G + o.lqg, G + o.lqb); LAUNCHED();    // dx0 = dx1 + LN'(dxn)
  ln_bwd_kernel<kDgrCtx><<<cdiv(T, 64), 256, 0, st>>>(w.dcn, w.c0, w.stc, T, p + o.lcg, nullptr, w.dc0, G + o.lcg, G + o.lcb); LAUNCHED();
  // ---- ConvPosEnc
  const long long nm = (long long)M * kDgrLatent, nt = (long long)T * kDgrCtx;
  if (h->pe) {
    cpe_bwd_kernel<<<dim3(cdiv(kDgrLatent, 32), cdiv(M, 64)), 256, 0, st>>>(w.dx0, latents, M, kDgrLatent, p + o.cqw, d_latents, G + o.cqw, G + o.cqb); LAUNCHED();
    cpe_bwd_kernel<<<dim3(cdiv(kDgrCtx, 32), cdiv(T, 64)), 256, 0, st>>>(w.dc0, image_feat, T, kDgrCtx, p + o.ccw, d_image_feat, G + o.ccw, G + o.ccb); LAUNCHED();
  } else {
    if (d_latents) CU(cudaMemcpyAsync(d_latents, w.dx0, nm * 4, cudaMemcpyDeviceToDevice, st));
    if (d_image_feat) CU(cudaMemcpyAsync(d_image_feat, w.dc0, nt * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

// torch.optim.SGD step on flat buffers (core/trainer.py:75-79: lr, momentum, weight_decay from the config; :300 optimizer.step()); grad_scale folds the
// 1 / world_size of an averaged all-reduce.  `first` != 0 initialises the momentum buffer with the gradient (torch's first-step behaviour).
int gmf_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum, float weight_decay, float grad_scale,
                 int first, void* stream) {
  if (!params || !grads || !momentum_buf || n < 1) return fail(GMF_ERR_INVALID, "gmf_sgd_step: bad argument");
  sgd_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, momentum_buf, n, lr, momentum, weight_decay, grad_scale, first);
  LAUNCHED();
  return 0;
}

}  // extern "C"
