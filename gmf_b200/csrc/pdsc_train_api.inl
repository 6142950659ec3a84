// C-ABI of the GMF-PointDSC training step (include/gmf_b200.h, "PointDSC training"); included at the end of gmf_api.cu.
// Parameters and gradients are FLAT fp32 device buffers in gmf_weight_spec order (the reference state_dict order without the backbone), owned
// by the caller (gmf_b200/trainer.py keeps them as torch tensors so that torch.distributed all-reduces the gradient in one NCCL call).
// Reference: GMF_PointDSC/libs/trainer.py:123-168, libs/loss.py:66-139, models/PointDSC.py:40-143,191-266 (training mode: batch-statistics
// BatchNorm, M output, logits as `final_labels`).

#include <unordered_map>

#include "pdsc_train.cuh"

namespace {

struct PtOff {
  std::unordered_map<std::string, size_t> m;
  size_t total = 0;
  explicit PtOff(int L) {
    for (const auto& s : build_spec(L)) { m[s.name] = total; total += (size_t)s.numel; }
  }
  size_t at(const std::string& k) const { return m.at(k); }
};

struct FusP { size_t cqw, cqb, ccw, ccb, lqg, lqb, lcg, lcb, wq, wkv, wo, bo, lfg, lfb, w1, b1, w2, b2; bool pe; };
FusP fus_offsets(const PtOff& o, const std::string& p, bool pe) {
  FusP f{};
  f.pe = pe;
  if (pe) { f.cqw = o.at(p + "cpe.proj_q.weight"); f.cqb = o.at(p + "cpe.proj_q.bias"); f.ccw = o.at(p + "cpe.proj_content.weight"); f.ccb = o.at(p + "cpe.proj_content.bias"); }
  const std::string a = p + "cross_attend_blocks.0.", ff = p + "cross_attend_blocks.1.";
  f.lqg = o.at(a + "norm.weight"); f.lqb = o.at(a + "norm.bias"); f.lcg = o.at(a + "norm_context.weight"); f.lcb = o.at(a + "norm_context.bias");
  f.wq = o.at(a + "fn.to_q.weight"); f.wkv = o.at(a + "fn.to_kv.weight"); f.wo = o.at(a + "fn.to_out.weight"); f.bo = o.at(a + "fn.to_out.bias");
  f.lfg = o.at(ff + "norm.weight"); f.lfb = o.at(ff + "norm.bias");
  f.w1 = o.at(ff + "fn.net.0.weight"); f.b1 = o.at(ff + "fn.net.0.bias"); f.w2 = o.at(ff + "fn.net.2.weight"); f.b2 = o.at(ff + "fn.net.2.bias");
  return f;
}
struct BnP { size_t g, b, rm, rv; };
BnP bn_offsets(const PtOff& o, const std::string& p) { return {o.at(p + "weight"), o.at(p + "bias"), o.at(p + "running_mean"), o.at(p + "running_var")}; }

// saved activations of one fusion layer (Rq = B Lq query rows, Rk = B Lk context rows)
struct FusAct { float *x0, *c0, *xn, *cn, *stq, *stc, *q, *kv, *P, *a, *x1, *stf, *h, *u, *g, *out; };
// ... and of one PointCN + NonLocalBlock
struct BlkAct { float *z, *st, *a, *qkv, *P, *msg, *z1, *st1, *m1, *z2, *st2, *m2, *out; FusAct f; };

constexpr int kSplitMax = 96;                            // split-K slices of a weight-gradient product (K = all rows of the step)

struct PtWs {
  FusAct f1;
  std::vector<BlkAct> blk;
  float *f0, *compat;
  float *fh, *fhT, *inv, *c1, *c2, *logit, *dlogit, *dfh;      // loss head
  double *cnt, *acc, *bnacc;
  float *dg, *du, *dh, *dx1, *da, *dq, *dkv, *dP, *dxn, *dcn, *dx0, *dc0;   // fusion backward scratch
  float *dA, *dB, *d64a, *d64b, *dmsg, *dqkv, *dasc, *dxf, *dimg, *dc1, *dc2;
  float *imgA, *imgB, *part;                              // operand images; split-K partial products
  size_t img_cap;
  int x3;                                                  // error-compensated products (3xTF32)
};

void fus_carve(FusAct& f, Bump& b, size_t Bn, size_t Lq, size_t Lk) {
  const size_t rq = Bn * Lq, rk = Bn * Lk;
  f.x0 = b.take<float>(rq * 128); f.c0 = b.take<float>(rk * 128); f.xn = b.take<float>(rq * 128); f.cn = b.take<float>(rk * 128);
  f.stq = b.take<float>(2 * rq); f.stc = b.take<float>(2 * rk); f.q = b.take<float>(rq * 64); f.kv = b.take<float>(rk * 128);
  f.P = b.take<float>(Bn * Lq * Lk); f.a = b.take<float>(rq * 64); f.x1 = b.take<float>(rq * 128); f.stf = b.take<float>(2 * rq);
  f.h = b.take<float>(rq * 128); f.u = b.take<float>(rq * 1024); f.g = b.take<float>(rq * 512); f.out = b.take<float>(rq * 128);
}
size_t pt_carve(PtWs& w, uint8_t* base, int L, int B, int N, int T, int x3) {
  Bump b{base};
  const size_t Bn = B, n = N, t = T, R = Bn * n, RT = Bn * t, Rm = std::max(R, RT), Lm = std::max(n, t), Np = (n + 63) / 64 * 64;
  fus_carve(w.f1, b, Bn, t, t);
  w.blk.resize(L);
  for (auto& k : w.blk) {
    k.z = b.take<float>(R * 128); k.st = b.take<float>(256); k.a = b.take<float>(R * 128); k.qkv = b.take<float>(R * 384); k.P = b.take<float>(Bn * n * n);
    k.msg = b.take<float>(R * 128); k.z1 = b.take<float>(R * 64); k.st1 = b.take<float>(128); k.m1 = b.take<float>(R * 64);
    k.z2 = b.take<float>(R * 64); k.st2 = b.take<float>(128); k.m2 = b.take<float>(R * 64); k.out = b.take<float>(R * 128);
    fus_carve(k.f, b, Bn, n, t);
  }
  w.f0 = b.take<float>(R * 128); w.compat = b.take<float>(Bn * n * n);
  w.fh = b.take<float>(R * 128); w.fhT = b.take<float>(Bn * 128 * Np); w.inv = b.take<float>(R); w.c1 = b.take<float>(R * 32); w.c2 = b.take<float>(R * 32);
  w.logit = b.take<float>(R); w.dlogit = b.take<float>(R); w.dfh = b.take<float>(R * 128);
  w.cnt = b.take<double>(Bn + 1); w.acc = b.take<double>(4); w.bnacc = b.take<double>(256);
  w.dg = b.take<float>(Rm * 512); w.du = b.take<float>(Rm * 1024); w.dh = b.take<float>(Rm * 128); w.dx1 = b.take<float>(Rm * 128);
  w.da = b.take<float>(Rm * 64); w.dq = b.take<float>(Rm * 64); w.dkv = b.take<float>(RT * 128); w.dP = b.take<float>(Bn * Lm * Lm);
  w.dxn = b.take<float>(Rm * 128); w.dcn = b.take<float>(RT * 128); w.dx0 = b.take<float>(Rm * 128); w.dc0 = b.take<float>(RT * 128);
  w.dA = b.take<float>(R * 128); w.dB = b.take<float>(R * 128); w.d64a = b.take<float>(R * 64); w.d64b = b.take<float>(R * 64);
  w.dmsg = b.take<float>(R * 128); w.dqkv = b.take<float>(R * 384); w.dasc = b.take<float>(R * 128); w.dxf = b.take<float>(R * 128);
  w.dimg = b.take<float>(RT * 128); w.dc1 = b.take<float>(R * 32); w.dc2 = b.take<float>(R * 32);
  auto p128 = [](size_t v) { return (v + 127) / 128 * 128; };
  auto p32 = [](size_t v) { return (v + 31) / 32 * 32; };
  w.x3 = x3 != 0;
  w.img_cap = std::max(std::max(Bn * p128(Lm) * p32(Lm), p128(Rm) * (size_t)1024), (size_t)1024 * p32(Rm)) * (x3 ? 2 : 1);
  w.imgA = b.take<float>(w.img_cap); w.imgB = b.take<float>(w.img_cap);
  w.part = b.take<float>((size_t)(kSplitMax + 1) * 1024 * 128);
  return b.off + 1024;
}
int pt_ws(PtWs& w, void* ws, size_t bytes, int L, int B, int N, int T, int x3) {
  if (!ws) return fail(GMF_ERR_INVALID, "workspace is NULL");
  const size_t need = pt_carve(w, nullptr, L, B, N, T, x3) + 1024;
  if (bytes < need) return fail(GMF_ERR_STATE, "workspace too small: need " + std::to_string(need) + " bytes");
  pt_carve(w, (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023), L, B, N, T, x3);
  return 0;
}

// out[z] (m x n, ldo) = scale * X[z] Y[z]^T (+ bias[col]) (+ residual[row][col], batch 1 only); X[z] is m x K, Y[z] is n x K, each row-major
// (t = 0: element (r, k) = p[r * ld + k]) or the transpose of a row-major matrix (t = 1: p[k * ld + r]); z strides in elements.
int pgemm(PtWs& w, const float* X, int ldx, size_t sx, int m, int K, int tx, const float* Y, int ldy, size_t sy, int n, int ty, float* out, int ldo, size_t so,
          float scale, const float* bias, const float* residual, int batch, cudaStream_t st) {
  const int tm = cdiv(m, 128), tn = cdiv(n, 128), kch = cdiv(K, 32), per = w.x3 ? 2 : 1;
  const size_t ia = (size_t)tm * kch * per * 4096, ib = (size_t)tn * kch * per * 4096;
  if (ia * batch > w.img_cap || ib * batch > w.img_cap) return fail(GMF_ERR_STATE, "pointdsc training: operand image exceeds the workspace");
  if (residual && batch != 1) return fail(GMF_ERR_INVALID, "pointdsc training: residual with a batched product");
  mat_to_img_b_kernel<<<dim3(tm, kch, batch), 256, 0, st>>>(X, sx, ldx, m, K, tx, kch, tm, w.x3, w.imgA);
  LAUNCHED();
  mat_to_img_b_kernel<<<dim3(tn, kch, batch), 256, 0, st>>>(Y, sy, ldy, n, K, ty, kch, tn, w.x3, w.imgB);
  LAUNCHED();
  ImgGemmArgs a{};
  a.a_img = w.imgA; a.w_packed = w.imgB; a.K = kch * 32; a.L = m; a.tiles = tm; a.out = out; a.ld = ldo; a.ncols = n; a.scale = scale; a.bias = bias;
  a.residual = residual; a.a_pair_stride = ia; a.w_pair_stride = ib; a.out_pair_stride = so;
  cudaError_t e = w.x3 ? launch_img_gemm<128, DE_STORE_X3>(a, tn, st, batch) : launch_img_gemm<128, DE_STORE>(a, tn, st, batch);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "pointdsc training GEMM launch");
  return 0;
}
// y [rows, nout] (ldo) = x [rows, nin] (ldx) W[nout, nin]^T + b
int lin_fwd(PtWs& w, const float* x, int ldx, int rows, int nin, const float* W, int nout, const float* b, float* y, int ldo, const float* residual, cudaStream_t st) {
  return pgemm(w, x, ldx, 0, rows, nin, 0, W, nin, 0, nout, 0, y, ldo, 0, 1.f, b, residual, 1, st);
}
// dx [rows, nin] = dy [rows, nout] (ldy) W (+ residual);  dW [nout, nin] = dy^T x;  db += colsum(dy)
int lin_bwd(PtWs& w, const float* dy, int ldy, int rows, int nout, const float* x, int ldx, int nin, const float* W, float* dx, const float* residual, float* dW,
            float* db, cudaStream_t st) {
  if (db) { col_sum_ld_kernel<<<dim3(cdiv(nout, 32), cdiv(rows, 256)), 256, 0, st>>>(dy, ldy, rows, nout, db); LAUNCHED(); }
  if (dW) {
    // dW = dy^T x has K = rows (up to B max(N, T)) and at most 8 output tiles: split K over `z` slices (the batched product does the slices, one
    // more call the remainder), partial products summed by splitk_reduce_kernel - otherwise one CTA streams the whole K range
    const int z = std::min(kSplitMax, rows / 1024);
    if (z < 2 || (size_t)nout * nin > (size_t)1024 * 128) {
      TRY(pgemm(w, dy, ldy, 0, nout, rows, 1, x, ldx, 0, nin, 1, dW, nin, 0, 1.f, nullptr, nullptr, 1, st));
    } else {
      const int kc = rows / z / 32 * 32, rem = rows - z * kc, mn = nout * nin;
      TRY(pgemm(w, dy, ldy, (size_t)kc * ldy, nout, kc, 1, x, ldx, (size_t)kc * ldx, nin, 1, w.part, nin, (size_t)mn, 1.f, nullptr, nullptr, z, st));
      if (rem > 0)
        TRY(pgemm(w, dy + (size_t)z * kc * ldy, ldy, 0, nout, rem, 1, x + (size_t)z * kc * ldx, ldx, 0, nin, 1, w.part + (size_t)z * mn, nin, 0, 1.f, nullptr,
                  nullptr, 1, st));
      splitk_reduce_kernel<<<cdiv(mn, 256), 256, 0, st>>>(w.part, z + (rem > 0 ? 1 : 0), mn, dW); LAUNCHED();
    }
  }
  if (dx) TRY(pgemm(w, dy, ldy, 0, rows, nout, 0, W, nin, 0, nin, 1, dx, nin, 0, 1.f, nullptr, residual, 1, st));
  return 0;
}
inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

// training-mode BatchNorm + ReLU: z -> a, statistics saved in st (mean | rstd), running statistics updated in the parameter buffer
int bn_relu_fwd(PtWs& w, const float* z, int rows, int C, float* params, const BnP& o, float* stat, float* a, cudaStream_t st) {
  CU(cudaMemsetAsync(w.bnacc, 0, 2 * C * sizeof(double), st));
  bn_stats_kernel<<<dim3(cdiv(C, 32), cdiv(rows, 256)), 256, 0, st>>>(z, rows, C, w.bnacc); LAUNCHED();
  bn_finalize_kernel<<<1, 128, 0, st>>>(w.bnacc, rows, C, stat, params + o.rm, params + o.rv); LAUNCHED();
  bn_relu_fwd_kernel<<<nblk((long long)rows * C), 256, 0, st>>>(z, (long long)rows * C, C, stat, params + o.g, params + o.b, a); LAUNCHED();
  return 0;
}
// da -> dz (may alias da); dgamma / dbeta written
int bn_relu_bwd(PtWs& w, const float* da, const float* a, const float* z, int rows, int C, const float* params, const BnP& o, const float* stat, float* dz, float* G,
                cudaStream_t st) {
  CU(cudaMemsetAsync(w.bnacc, 0, 2 * C * sizeof(double), st));
  bn_relu_bwd_reduce_kernel<<<dim3(cdiv(C, 32), cdiv(rows, 256)), 256, 0, st>>>(da, a, z, rows, C, stat, w.bnacc); LAUNCHED();
  bn_relu_bwd_apply_kernel<<<nblk((long long)rows * C), 256, 0, st>>>(da, a, z, (long long)rows * C, rows, C, stat, params + o.g, w.bnacc, dz, G + o.g, G + o.b);
  LAUNCHED();
  return 0;
}

constexpr float kFusScale = 0.125f;                       // 64 ** -0.5 (fusion_layer.py:76, cross_dim_head = 64)
constexpr float kScScale = 0.08838834764831845f;          // 128 ** -0.5 (PointDSC.py:60)

// FusionLayer.forward (fusion_layer.py:172-201, depth 0) on B sequences: x_in [B Lq, 128] queries, ctx_in [B Lk, 128] context -> f.out
int fusion_train_fwd(PtWs& w, const float* p, const FusP& o, FusAct& f, const float* x_in, const float* ctx_in, int B, int Lq, int Lk, cudaStream_t st) {
  const int rq = B * Lq, rk = B * Lk;
  const long long nq = (long long)rq * 128, nk = (long long)rk * 128;
  if (o.pe) {
    cpe_seq_fwd_kernel<<<nblk(nq), 256, 0, st>>>(x_in, rq, Lq, 128, p + o.cqw, p + o.cqb, f.x0); LAUNCHED();
    cpe_seq_fwd_kernel<<<nblk(nk), 256, 0, st>>>(ctx_in, rk, Lk, 128, p + o.ccw, p + o.ccb, f.c0); LAUNCHED();
  } else {
    CU(cudaMemcpyAsync(f.x0, x_in, nq * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(f.c0, ctx_in, nk * 4, cudaMemcpyDeviceToDevice, st));
  }
  ln_fwd_kernel<128><<<cdiv(rq, 8), 256, 0, st>>>(f.x0, rq, p + o.lqg, p + o.lqb, f.xn, f.stq); LAUNCHED();
  ln_fwd_kernel<128><<<cdiv(rk, 8), 256, 0, st>>>(f.c0, rk, p + o.lcg, p + o.lcb, f.cn, f.stc); LAUNCHED();
  TRY(lin_fwd(w, f.xn, 128, rq, 128, p + o.wq, 64, nullptr, f.q, 64, nullptr, st));
  TRY(lin_fwd(w, f.cn, 128, rk, 128, p + o.wkv, 128, nullptr, f.kv, 128, nullptr, st));
  TRY(pgemm(w, f.q, 64, (size_t)Lq * 64, Lq, 64, 0, f.kv, 128, (size_t)Lk * 128, Lk, 0, f.P, Lk, (size_t)Lq * Lk, kFusScale, nullptr, nullptr, B, st));   // q k^T
  softmax_mul_rows_kernel<<<rq, 256, 0, st>>>(f.P, nullptr, Lk); LAUNCHED();
  TRY(pgemm(w, f.P, Lk, (size_t)Lq * Lk, Lq, Lk, 0, f.kv + 64, 128, (size_t)Lk * 128, 64, 1, f.a, 64, (size_t)Lq * 64, 1.f, nullptr, nullptr, B, st));    // P v
  TRY(lin_fwd(w, f.a, 64, rq, 64, p + o.wo, 128, p + o.bo, f.x1, 128, f.x0, st));
  ln_fwd_kernel<128><<<cdiv(rq, 8), 256, 0, st>>>(f.x1, rq, p + o.lfg, p + o.lfb, f.h, f.stf); LAUNCHED();
  TRY(lin_fwd(w, f.h, 128, rq, 128, p + o.w1, 1024, p + o.b1, f.u, 1024, nullptr, st));
  geglu_fwd_kernel<<<nblk((long long)rq * 512), 256, 0, st>>>(f.u, (long long)rq * 512, 512, f.g); LAUNCHED();
  TRY(lin_fwd(w, f.g, 512, rq, 512, p + o.w2, 128, p + o.b2, f.out, 128, f.x1, st));
  return 0;
}
// d_out [B Lq, 128] -> dx [B Lq, 128] (overwritten) and dctx [B Lk, 128] (overwritten or accumulated); weight gradients into G
int fusion_train_bwd(PtWs& w, const float* p, float* G, const FusP& o, const FusAct& f, const float* x_in, const float* ctx_in, const float* d_out, int B, int Lq, int Lk,
                     float* dx, float* dctx, int accumulate_ctx, cudaStream_t st) {
  const int rq = B * Lq, rk = B * Lk;
  TRY(lin_bwd(w, d_out, 128, rq, 128, f.g, 512, 512, p + o.w2, w.dg, nullptr, G + o.w2, G + o.b2, st));
  geglu_bwd_kernel<<<nblk((long long)rq * 512), 256, 0, st>>>(f.u, w.dg, (long long)rq * 512, 512, w.du); LAUNCHED();
  TRY(lin_bwd(w, w.du, 1024, rq, 1024, f.h, 128, 128, p + o.w1, w.dh, nullptr, G + o.w1, G + o.b1, st));
  ln_bwd_kernel<128><<<cdiv(rq, 64), 256, 0, st>>>(w.dh, f.x1, f.stf, rq, p + o.lfg, d_out, w.dx1, G + o.lfg, G + o.lfb); LAUNCHED();      // dx1 = d_out + LN'(dh)
  TRY(lin_bwd(w, w.dx1, 128, rq, 128, f.a, 64, 64, p + o.wo, w.da, nullptr, G + o.wo, G + o.bo, st));
  const size_t sP = (size_t)Lq * Lk;
  TRY(pgemm(w, w.da, 64, (size_t)Lq * 64, Lq, 64, 0, f.kv + 64, 128, (size_t)Lk * 128, Lk, 0, w.dP, Lk, sP, 1.f, nullptr, nullptr, B, st));              // dP = da v^T
  TRY(pgemm(w, f.P, Lk, sP, Lk, Lq, 1, w.da, 64, (size_t)Lq * 64, 64, 1, w.dkv + 64, 128, (size_t)Lk * 128, 1.f, nullptr, nullptr, B, st));                // dv = P^T da
  softmax_mul_bwd_kernel<<<rq, 256, 0, st>>>(f.P, w.dP, nullptr, Lk, kFusScale); LAUNCHED();
  TRY(pgemm(w, w.dP, Lk, sP, Lq, Lk, 0, f.kv, 128, (size_t)Lk * 128, 64, 1, w.dq, 64, (size_t)Lq * 64, 1.f, nullptr, nullptr, B, st));                     // dq = dS k
  TRY(pgemm(w, w.dP, Lk, sP, Lk, Lq, 1, f.q, 64, (size_t)Lq * 64, 64, 1, w.dkv, 128, (size_t)Lk * 128, 1.f, nullptr, nullptr, B, st));                     // dk = dS^T q
  TRY(lin_bwd(w, w.dq, 64, rq, 64, f.xn, 128, 128, p + o.wq, w.dxn, nullptr, G + o.wq, nullptr, st));
  TRY(lin_bwd(w, w.dkv, 128, rk, 128, f.cn, 128, 128, p + o.wkv, w.dcn, nullptr, G + o.wkv, nullptr, st));
  float* dx0 = o.pe ? w.dx0 : dx;
  ln_bwd_kernel<128><<<cdiv(rq, 64), 256, 0, st>>>(w.dxn, f.x0, f.stq, rq, p + o.lqg, w.dx1, dx0, G + o.lqg, G + o.lqb); LAUNCHED();         // dx0 = dx1 + LN'(dxn)
  ln_bwd_kernel<128><<<cdiv(rk, 64), 256, 0, st>>>(w.dcn, f.c0, f.stc, rk, p + o.lcg, nullptr, w.dc0, G + o.lcg, G + o.lcb); LAUNCHED();
  if (o.pe) {
    cpe_seq_bwd_kernel<<<dim3(4, cdiv(rq, 64)), 256, 0, st>>>(w.dx0, x_in, rq, Lq, 128, p + o.cqw, dx, 0, G + o.cqw, G + o.cqb); LAUNCHED();
    cpe_seq_bwd_kernel<<<dim3(4, cdiv(rk, 64)), 256, 0, st>>>(w.dc0, ctx_in, rk, Lk, 128, p + o.ccw, dctx, accumulate_ctx, G + o.ccw, G + o.ccb); LAUNCHED();
  } else if (accumulate_ctx) {
    add_inplace_kernel<<<nblk((long long)rk * 128), 256, 0, st>>>(dctx, w.dc0, (long long)rk * 128); LAUNCHED();
  } else {
    CU(cudaMemcpyAsync(dctx, w.dc0, (size_t)rk * 128 * 4, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

struct BlkP { size_t pw, pb; BnP pbn; size_t m0w, m0b; BnP m1; size_t m3w, m3b; BnP m4; size_t m6w, m6b, qw[3], qb[3]; FusP f; };
BlkP blk_offsets(const PtOff& o, int i) {
  BlkP b{};
  const std::string p = "encoder.blocks.PointCN_layer_" + std::to_string(i) + ".", n = "encoder.blocks.NonLocal_layer_" + std::to_string(i) + ".";
  b.pw = o.at(p + "0.weight"); b.pb = o.at(p + "0.bias"); b.pbn = bn_offsets(o, p + "1.");
  b.m0w = o.at(n + "fc_message.0.weight"); b.m0b = o.at(n + "fc_message.0.bias"); b.m1 = bn_offsets(o, n + "fc_message.1.");
  b.m3w = o.at(n + "fc_message.3.weight"); b.m3b = o.at(n + "fc_message.3.bias"); b.m4 = bn_offsets(o, n + "fc_message.4.");
  b.m6w = o.at(n + "fc_message.6.weight"); b.m6b = o.at(n + "fc_message.6.bias");
  const char* q[3] = {"q", "k", "v"};
  for (int t = 0; t < 3; ++t) { b.qw[t] = o.at(n + "projection_" + q[t] + ".weight"); b.qb[t] = o.at(n + "projection_" + q[t] + ".bias"); }
  b.f = fus_offsets(o, n + "fusion_layer_2.", true);
  return b;
}

int pt_check(int L, int B, int N, int T) {
  if (L < 1 || L > 64 || B < 1 || N < 2 || T < 1) return fail(GMF_ERR_INVALID, "pointdsc training: need 1 <= num_layers <= 64, B >= 1, N >= 2, T >= 1");
  if ((long long)B * std::max(N, T) > (1ll << 24)) return fail(GMF_ERR_INVALID, "pointdsc training: B * max(N, T) too large");
  return 0;
}

}  // namespace

extern "C" {

size_t gmf_pointdsc_train_workspace_bytes(int num_layers, int B, int N, int T, int tf32x3) {
  if (num_layers < 1 || num_layers > 64 || B < 1 || N < 2 || T < 1) return 0;
  PtWs w;
  return pt_carve(w, nullptr, num_layers, B, N, T, tf32x3) + 2048;
}
int64_t gmf_pointdsc_param_count(int num_layers) { return num_layers < 1 ? 0 : (int64_t)PtOff(num_layers).total; }

int gmf_pointdsc_train_forward(int device, int num_layers, float* params, const float* corr_pos, const float* src_keypts, const float* tgt_keypts,
                               const float* p_tokens, const float* q_tokens, const float* gt_labels, int B, int N, int T, int balanced, float w_class,
                               float w_sm, int tf32x3, float* losses, float* logits, float* features, float* M, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (!params || !corr_pos || !src_keypts || !tgt_keypts || !p_tokens || !q_tokens) return fail(GMF_ERR_INVALID, "NULL argument");
  if ((gt_labels == nullptr) != (losses == nullptr)) return fail(GMF_ERR_INVALID, "gt_labels and losses go together (both NULL: no loss head, gradients come from the caller)");
  TRY(pt_check(num_layers, B, N, T));
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  PtWs w;
  TRY(pt_ws(w, workspace, workspace_bytes, num_layers, B, N, T, tf32x3));
  const PtOff o(num_layers);
  float* p = params;
  const int R = B * N;
  // Fusion-1: queries = q-image tokens, context = p-image tokens (PointDSC.py:137)
  TRY(fusion_train_fwd(w, p, fus_offsets(o, "encoder.fusion_layer_1.", false), w.f1, q_tokens, p_tokens, B, T, T, st));
  compat_kernel<<<dim3(cdiv(N, 256), N, B), 256, 0, st>>>(src_keypts, tgt_keypts, N, p + o.at("sigma_spat"), w.compat); LAUNCHED();
  TRY(lin_fwd(w, corr_pos, 6, R, 6, p + o.at("encoder.layer0.weight"), 128, p + o.at("encoder.layer0.bias"), w.f0, 128, nullptr, st));
  const float* fin = w.f0;
  for (int i = 0; i < num_layers; ++i) {
    const BlkP bo = blk_offsets(o, i);
    BlkAct& k = w.blk[i];
    TRY(lin_fwd(w, fin, 128, R, 128, p + bo.pw, 128, p + bo.pb, k.z, 128, nullptr, st));
    TRY(bn_relu_fwd(w, k.z, R, 128, p, bo.pbn, k.st, k.a, st));
    for (int t = 0; t < 3; ++t) TRY(lin_fwd(w, k.a, 128, R, 128, p + bo.qw[t], 128, p + bo.qb[t], k.qkv + t * 128, 384, nullptr, st));
    const size_t sq = (size_t)N * 384, sp = (size_t)N * N;
    TRY(pgemm(w, k.qkv, 384, sq, N, 128, 0, k.qkv + 128, 384, sq, N, 0, k.P, N, sp, kScScale, nullptr, nullptr, B, st));                    // Q K^T / sqrt(C)
    softmax_mul_rows_kernel<<<R, 256, 0, st>>>(k.P, w.compat, N); LAUNCHED();                                                                // softmax(compat * .)
    TRY(pgemm(w, k.P, N, sp, N, N, 0, k.qkv + 256, 384, sq, 128, 1, k.msg, 128, (size_t)N * 128, 1.f, nullptr, nullptr, B, st));             // weight V
    TRY(lin_fwd(w, k.msg, 128, R, 128, p + bo.m0w, 64, p + bo.m0b, k.z1, 64, nullptr, st));
    TRY(bn_relu_fwd(w, k.z1, R, 64, p, bo.m1, k.st1, k.m1, st));
    TRY(lin_fwd(w, k.m1, 64, R, 64, p + bo.m3w, 64, p + bo.m3b, k.z2, 64, nullptr, st));
    TRY(bn_relu_fwd(w, k.z2, R, 64, p, bo.m4, k.st2, k.m2, st));
    TRY(fusion_train_fwd(w, p, bo.f, k.f, k.a, w.f1.out, B, N, T, st));
    TRY(lin_fwd(w, k.m2, 64, R, 64, p + bo.m6w, 128, p + bo.m6b, k.out, 128, k.f.out, st));                                                  // message + fused
    fin = k.out;
  }
  // loss head: normalised features for M, classifier on the raw features (PointDSC.py:229-241), losses (libs/loss.py:66-139)
  const int Np = (N + 63) / 64 * 64;
  CU(cudaMemsetAsync(w.fhT, 0, (size_t)B * 128 * Np * sizeof(float), st));
  normalize_fwd_kernel<<<cdiv(R, 8), 256, 0, st>>>(fin, B, N, Np, w.fh, w.fhT, w.inv); LAUNCHED();
  TRY(lin_fwd(w, fin, 128, R, 128, p + o.at("classification.0.weight"), 32, p + o.at("classification.0.bias"), w.c1, 32, nullptr, st));
  relu_inplace_kernel<<<nblk((long long)R * 32), 256, 0, st>>>(w.c1, (long long)R * 32); LAUNCHED();
  TRY(lin_fwd(w, w.c1, 32, R, 32, p + o.at("classification.2.weight"), 32, p + o.at("classification.2.bias"), w.c2, 32, nullptr, st));
  relu_inplace_kernel<<<nblk((long long)R * 32), 256, 0, st>>>(w.c2, (long long)R * 32); LAUNCHED();
  TRY(lin_fwd(w, w.c2, 32, R, 32, p + o.at("classification.4.weight"), 1, p + o.at("classification.4.bias"), w.logit, 1, nullptr, st));
  CU(cudaMemsetAsync(w.acc, 0, 4 * sizeof(double), st));
  if (gt_labels) {
    CU(cudaMemsetAsync(w.cnt, 0, (B + 1) * sizeof(double), st));
    label_count_kernel<<<B, 256, 0, st>>>(gt_labels, B, N, w.cnt); LAUNCHED();
    bce_kernel<<<nblk(R), 256, 0, st>>>(w.logit, gt_labels, R, balanced, w.cnt + B, w_class, w.acc + 2, w.dlogit); LAUNCHED();
    static std::atomic<unsigned long long> configured{0};
    if (cudaError_t e = ensure_dyn_smem(sm_loss_fused_kernel, kSmlSmem, configured)) return fail_cuda(e, "sm_loss_fused_kernel attribute");
    sm_loss_fused_kernel<<<dim3(Np / 64, B), 256, kSmlSmem, st>>>(w.fh, w.fhT, gt_labels, B, N, Np, p + o.at("sigma"), balanced, w.cnt, w_sm, w.acc, w.dfh);
    LAUNCHED();
    loss_finalize_kernel<<<1, 1, 0, st>>>(w.acc, w_class, w_sm, losses, nullptr); LAUNCHED();
  }
  if (M) {      // the materialised training-mode output, for callers that compute their own loss on it (the reference's trainer does)
    TRY(pgemm(w, w.fh, 128, (size_t)N * 128, N, 128, 0, w.fh, 128, (size_t)N * 128, N, 0, M, N, (size_t)N * N, 1.f, nullptr, nullptr, B, st));
    m_from_s_kernel<<<nblk((long long)B * N * N), 256, 0, st>>>(M, N, (long long)B * N * N, p + o.at("sigma")); LAUNCHED();
  }
  if (logits) CU(cudaMemcpyAsync(logits, w.logit, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
  if (features) CU(cudaMemcpyAsync(features, fin, (size_t)R * 128 * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Uses what gmf_pointdsc_train_forward left in `workspace` (same shapes, params and inputs).  grads (gmf_pointdsc_param_count floats) is
// overwritten; d_p_tokens / d_q_tokens [B, T, 128] (gradients flowing on into the image backbone) may be NULL.
int gmf_pointdsc_train_backward(int device, int num_layers, const float* params, const float* corr_pos, const float* p_tokens, const float* q_tokens, int B, int N,
                                int T, float w_class, float w_sm, int tf32x3, const float* d_logits, const float* d_M, const float* d_features, float* grads,
                                float* d_p_tokens, float* d_q_tokens, void* workspace, size_t workspace_bytes, void* stream) {
  if (!params || !corr_pos || !p_tokens || !q_tokens || !grads) return fail(GMF_ERR_INVALID, "NULL argument");
  TRY(pt_check(num_layers, B, N, T));
  CU(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)stream;
  PtWs w;
  TRY(pt_ws(w, workspace, workspace_bytes, num_layers, B, N, T, tf32x3));
  const PtOff o(num_layers);
  const float* p = params;
  float* G = grads;
  const int R = B * N, RT = B * T;
  CU(cudaMemsetAsync(G, 0, o.total * sizeof(float), st));
  CU(cudaMemsetAsync(w.dimg, 0, (size_t)RT * 128 * sizeof(float), st));
  const float* feat = w.blk[num_layers - 1].out;
  const bool external = d_logits || d_M || d_features;
  if (!external) {
    loss_finalize_kernel<<<1, 1, 0, st>>>(w.acc, w_class, w_sm, nullptr, G + o.at("sigma")); LAUNCHED();      // fused loss head of the forward
  } else {
    // autograd entry: d loss / d logits [B,N], d loss / d M [B,N,N], d loss / d features [B,N,128] come from the caller (any may be NULL)
    if (d_logits) CU(cudaMemcpyAsync(w.dlogit, d_logits, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
    else CU(cudaMemsetAsync(w.dlogit, 0, (size_t)R * 4, st));
    if (d_M) {
      const size_t sp = (size_t)N * N, sf = (size_t)N * 128;
      const long long nn = (long long)B * N * N;
      CU(cudaMemsetAsync(w.acc + 3, 0, sizeof(double), st));
      TRY(pgemm(w, w.fh, 128, sf, N, 128, 0, w.fh, 128, sf, N, 0, w.dP, N, sp, 1.f, nullptr, nullptr, B, st));                     // S = fh fh^T again
      ds_from_dm_kernel<<<nblk(nn), 256, 0, st>>>(w.dP, d_M, N, nn, p + o.at("sigma"), w.acc + 3); LAUNCHED();
      store_double_as_float_kernel<<<1, 1, 0, st>>>(w.acc + 3, G + o.at("sigma")); LAUNCHED();
      TRY(pgemm(w, w.dP, N, sp, N, N, 0, w.fh, 128, sf, 128, 1, w.dfh, 128, sf, 1.f, nullptr, nullptr, B, st));                    // G fh
      TRY(pgemm(w, w.dP, N, sp, N, N, 1, w.fh, 128, sf, 128, 1, w.dmsg, 128, sf, 1.f, nullptr, nullptr, B, st));                   // G^T fh
      add_inplace_kernel<<<nblk((long long)R * 128), 256, 0, st>>>(w.dfh, w.dmsg, (long long)R * 128); LAUNCHED();
    } else {
      CU(cudaMemsetAsync(w.dfh, 0, (size_t)R * 128 * 4, st));
    }
  }
  // classifier 128-32-32-1 (PointDSC.py:175-181)
  TRY(lin_bwd(w, w.dlogit, 1, R, 1, w.c2, 32, 32, p + o.at("classification.4.weight"), w.dc2, nullptr, G + o.at("classification.4.weight"), G + o.at("classification.4.bias"), st));
  relu_bwd_kernel<<<nblk((long long)R * 32), 256, 0, st>>>(w.dc2, w.c2, (long long)R * 32); LAUNCHED();
  TRY(lin_bwd(w, w.dc2, 32, R, 32, w.c1, 32, 32, p + o.at("classification.2.weight"), w.dc1, nullptr, G + o.at("classification.2.weight"), G + o.at("classification.2.bias"), st));
  relu_bwd_kernel<<<nblk((long long)R * 32), 256, 0, st>>>(w.dc1, w.c1, (long long)R * 32); LAUNCHED();
  TRY(lin_bwd(w, w.dc1, 32, R, 32, feat, 128, 128, p + o.at("classification.0.weight"), w.dA, nullptr, G + o.at("classification.0.weight"), G + o.at("classification.0.bias"), st));
  normalize_bwd_kernel<<<cdiv(R, 8), 256, 0, st>>>(w.dfh, w.fh, w.inv, R, w.dA, 1); LAUNCHED();
  if (d_features) { add_inplace_kernel<<<nblk((long long)R * 128), 256, 0, st>>>(w.dA, d_features, (long long)R * 128); LAUNCHED(); }
  float *dcur = w.dA, *dnext = w.dB;
  for (int i = num_layers - 1; i >= 0; --i) {
    const BlkP bo = blk_offsets(o, i);
    const BlkAct& k = w.blk[i];
    const float* fin = i ? w.blk[i - 1].out : w.f0;
    // fc_message (PointDSC.py:13-21): out = conv6(m2) + fused
    TRY(lin_bwd(w, dcur, 128, R, 128, k.m2, 64, 64, p + bo.m6w, w.d64a, nullptr, G + bo.m6w, G + bo.m6b, st));
    TRY(bn_relu_bwd(w, w.d64a, k.m2, k.z2, R, 64, p, bo.m4, k.st2, w.d64a, G, st));
    TRY(lin_bwd(w, w.d64a, 64, R, 64, k.m1, 64, 64, p + bo.m3w, w.d64b, nullptr, G + bo.m3w, G + bo.m3b, st));
    TRY(bn_relu_bwd(w, w.d64b, k.m1, k.z1, R, 64, p, bo.m1, k.st1, w.d64b, G, st));
    TRY(lin_bwd(w, w.d64b, 64, R, 64, k.msg, 128, 128, p + bo.m0w, w.dmsg, nullptr, G + bo.m0w, G + bo.m0b, st));
    // SC-guided attention (PointDSC.py:54-64)
    const size_t sq = (size_t)N * 384, sp = (size_t)N * N, sm = (size_t)N * 128;
    TRY(pgemm(w, w.dmsg, 128, sm, N, 128, 0, k.qkv + 256, 384, sq, N, 0, w.dP, N, sp, 1.f, nullptr, nullptr, B, st));                        // dP = dmsg V^T
    TRY(pgemm(w, k.P, N, sp, N, N, 1, w.dmsg, 128, sm, 128, 1, w.dqkv + 256, 384, sq, 1.f, nullptr, nullptr, B, st));                        // dV = P^T dmsg
    softmax_mul_bwd_kernel<<<R, 256, 0, st>>>(k.P, w.dP, w.compat, N, kScScale); LAUNCHED();
    TRY(pgemm(w, w.dP, N, sp, N, N, 0, k.qkv + 128, 384, sq, 128, 1, w.dqkv, 384, sq, 1.f, nullptr, nullptr, B, st));                        // dQ = dS K
    TRY(pgemm(w, w.dP, N, sp, N, N, 1, k.qkv, 384, sq, 128, 1, w.dqkv + 128, 384, sq, 1.f, nullptr, nullptr, B, st));                        // dK = dS^T Q
    for (int t = 0; t < 3; ++t)
      TRY(lin_bwd(w, w.dqkv + t * 128, 384, R, 128, k.a, 128, 128, p + bo.qw[t], w.dasc, t ? w.dasc : nullptr, G + bo.qw[t], G + bo.qb[t], st));
    // Fusion-2: queries = this layer's features, context = the Fusion-1 output shared by all layers (PointDSC.py:68-71)
    TRY(fusion_train_bwd(w, p, G, bo.f, k.f, k.a, w.f1.out, dcur, B, N, T, w.dxf, w.dimg, 1, st));
    add_inplace_kernel<<<nblk((long long)R * 128), 256, 0, st>>>(w.dasc, w.dxf, (long long)R * 128); LAUNCHED();
    // PointCN: conv - BatchNorm - ReLU (PointDSC.py:104-109)
    TRY(bn_relu_bwd(w, w.dasc, k.a, k.z, R, 128, p, bo.pbn, k.st, w.dasc, G, st));
    TRY(lin_bwd(w, w.dasc, 128, R, 128, fin, 128, 128, p + bo.pw, dnext, nullptr, G + bo.pw, G + bo.pb, st));
    std::swap(dcur, dnext);
  }
  TRY(lin_bwd(w, dcur, 128, R, 128, corr_pos, 6, 6, p + o.at("encoder.layer0.weight"), nullptr, nullptr, G + o.at("encoder.layer0.weight"),
              G + o.at("encoder.layer0.bias"), st));
  float* dq = d_q_tokens ? d_q_tokens : w.dh;            // Fusion-1 input gradients: scratch (dead by then) when the caller does not want them
  float* dp = d_p_tokens ? d_p_tokens : w.dcn;
  TRY(fusion_train_bwd(w, p, G, fus_offsets(o, "encoder.fusion_layer_1.", false), w.f1, q_tokens, p_tokens, w.dimg, B, T, T, dq, dp, 0, st));
  return 0;
}

// torch.optim.Adam step on flat buffers (train_3DMatch.py:52-58; amsgrad False).  `step` is 1 for the first call; `mask` (one byte per
// element, NULL = all) marks the trainable entries - running statistics and sigma_spat live in the same flat buffer and are left alone.
int gmf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const unsigned char* mask, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float grad_scale, int step, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return fail(GMF_ERR_INVALID, "gmf_adam_step: bad argument");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, mask, n, lr, beta1, beta2, eps, weight_decay,
                                                                                 grad_scale, bc1, sqrtf(bc2));
  LAUNCHED();
  return 0;
}

}  // extern "C"
