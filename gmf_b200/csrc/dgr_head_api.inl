// C-ABI of the DGR bottleneck fusion head (include/gmf_b200.h, "DGR head"); included at the end of gmf_api.cu (shares its
// error/launch-count plumbing and pack_linear).  Reference: GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py:140-221.

namespace {

constexpr int kDgrLatent = 256, kDgrCtx = 128, kDgrHead = 128, kDgrHidden = 1024;

std::vector<Spec> dgr_spec(bool pe) {
  std::vector<Spec> s;
  if (pe) {
    s.push_back({"cpe.proj_q.weight", kDgrLatent * 3}); s.push_back({"cpe.proj_q.bias", kDgrLatent});
    s.push_back({"cpe.proj_content.weight", kDgrCtx * 3}); s.push_back({"cpe.proj_content.bias", kDgrCtx});
  }
  const std::string a = "cross_attend_blocks.0.", f = "cross_attend_blocks.1.";
  s.push_back({a + "norm.weight", kDgrLatent}); s.push_back({a + "norm.bias", kDgrLatent});
  s.push_back({a + "norm_context.weight", kDgrCtx}); s.push_back({a + "norm_context.bias", kDgrCtx});
  s.push_back({a + "fn.to_q.weight", kDgrHead * kDgrLatent}); s.push_back({a + "fn.to_kv.weight", 2 * kDgrHead * kDgrCtx});
  s.push_back({a + "fn.to_out.weight", kDgrLatent * kDgrHead}); s.push_back({a + "fn.to_out.bias", kDgrLatent});
  s.push_back({f + "norm.weight", kDgrLatent}); s.push_back({f + "norm.bias", kDgrLatent});
  s.push_back({f + "fn.net.0.weight", 2 * kDgrHidden * kDgrLatent}); s.push_back({f + "fn.net.0.bias", 2 * kDgrHidden});
  s.push_back({f + "fn.net.2.weight", kDgrLatent * kDgrHidden}); s.push_back({f + "fn.net.2.bias", kDgrLatent});
  return s;
}

struct DgrWork {
  float *x0, *xq_img, *c_img, *o, *o_img, *x1, *ln_img, *g_img, *part_o, *part_l;
  __nv_bfloat16 *q_t, *k_t, *vt_t;
};

size_t dgr_carve(DgrWork& w, uint8_t* base, int M, int T) {
  const size_t mt = cdiv(M, 128), kt = cdiv(T, 128);
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off += (bytes + 1023) & ~(size_t)1023;
    return p;
  };
  w.x0 = (float*)take((size_t)M * kDgrLatent * 4);
  w.xq_img = (float*)take(mt * 128 * kDgrLatent * 4);
  w.c_img = (float*)take(kt * 128 * kDgrCtx * 4);
  w.o = (float*)take((size_t)M * kDgrHead * 4);
  w.o_img = (float*)take(mt * 128 * kDgrHead * 4);
  w.x1 = (float*)take((size_t)M * kDgrLatent * 4);
  w.ln_img = (float*)take(mt * 128 * kDgrLatent * 4);
  w.g_img = (float*)take(mt * 128 * kDgrHidden * 4);
  w.q_t = (__nv_bfloat16*)take(mt * 128 * 128 * 2);
  w.k_t = (__nv_bfloat16*)take(kt * 128 * 128 * 2);
  w.vt_t = (__nv_bfloat16*)take(kt * 128 * 128 * 2);
  w.part_o = (float*)take(kt * (size_t)M * kDgrHead * 4);        // up to one split per context tile
  w.part_l = (float*)take(kt * (size_t)M * 2 * 4);
  return off + 1024;
}

}  // namespace

struct gmf_dgr_head {
  int device = 0;
  bool pe = true, loaded = false;
  float* blob = nullptr;
  const float *cpe_q_w = nullptr, *cpe_q_b = nullptr, *cpe_c_w = nullptr, *cpe_c_b = nullptr;
  const float *lnq_g, *lnq_b, *lnc_g, *lnc_b, *lnf_g, *lnf_b, *wq, *wkv, *wo, *bo, *w1, *b1, *w2, *b2;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // neutral distance-feature tiles (compat == 1) for the attention kernel.  They live in their OWN allocation: inside the per-call
  // workspace their offsets would move with M (the active-voxel count changes every call) while a cache keyed on the tile count
  // would still call them valid.
  uint8_t* feat = nullptr;   // [kpts | aq | bd] for feat_tiles tiles
  int feat_tiles = 0;
};

extern "C" {

int gmf_dgr_head_create(gmf_dgr_head** out, int device, int latent_dim, int context_dim, int head_dim, int pe) {
  if (!out) return fail(GMF_ERR_INVALID, "out is NULL");
  if (latent_dim != kDgrLatent || context_dim != kDgrCtx || head_dim != kDgrHead)
    return fail(GMF_ERR_INVALID, "gmf_dgr_head: only the reference's bottleneck shape latent_dim=256, dim=128, cross_dim_head=128 is built "
                                 "(latent_dim=128 / cross_dim_head=64 is the image_fusion block: use gmf_fusion_layer)");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(GMF_ERR_INVALID, "no such CUDA device");
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(GMF_ERR_STATE, "gmf_b200 kernels are built for sm_100a only (no fallback path)");
  gmf_dgr_head* h = new gmf_dgr_head();
  h->device = device; h->pe = pe != 0;
  *out = h;
  return 0;
}

void gmf_dgr_head_destroy(gmf_dgr_head* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->blob) cudaFree(h->blob);
  if (h->ws) cudaFree(h->ws);
  if (h->feat) cudaFree(h->feat);
  delete h;
}

int gmf_dgr_head_weight_count(int pe) { return (int)dgr_spec(pe != 0).size(); }

int gmf_dgr_head_weight_spec(int pe, int index, char* name, int cap, int64_t* numel) {
  const std::vector<Spec> s = dgr_spec(pe != 0);
  if (index < 0 || index >= (int)s.size()) return fail(GMF_ERR_INVALID, "weight index out of range");
  if (name && cap > 0) { strncpy(name, s[index].name.c_str(), cap - 1); name[cap - 1] = 0; }
  if (numel) *numel = s[index].numel;
  return 0;
}

int gmf_dgr_head_load_weights(gmf_dgr_head* h, const float* host_flat, int64_t numel) {
  if (!h || !host_flat) return fail(GMF_ERR_INVALID, "NULL argument");
  const std::vector<Spec> spec = dgr_spec(h->pe);
  int64_t total = 0;
  for (const Spec& s : spec) total += s.numel;
  if (numel != total) return fail(GMF_ERR_INVALID, "gmf_dgr_head_load_weights: expected " + std::to_string(total) + " floats");
  size_t idx = 0;
  const float* cur = host_flat;
  auto next = [&](int64_t n) -> std::vector<float> {
    if (spec[idx].numel != n) abort();
    std::vector<float> v(cur, cur + n);
    cur += n; ++idx;
    return v;
  };
  Blob blob;
  size_t o_cqw = 0, o_cqb = 0, o_ccw = 0, o_ccb = 0;
  if (h->pe) {
    o_cqw = blob.push(next(kDgrLatent * 3)); o_cqb = blob.push(next(kDgrLatent));
    o_ccw = blob.push(next(kDgrCtx * 3)); o_ccb = blob.push(next(kDgrCtx));
  }
  const size_t o_lnqg = blob.push(next(kDgrLatent)), o_lnqb = blob.push(next(kDgrLatent));
  const size_t o_lncg = blob.push(next(kDgrCtx)), o_lncb = blob.push(next(kDgrCtx));
  std::vector<float> wq = next(kDgrHead * kDgrLatent);
  const float qs = kLog2e / std::sqrt((float)kDgrHead);             // softmax scale (perceiver_io.py:76) and log2(e) folded into to_q
  for (float& v : wq) v *= qs;
  const size_t o_wq = blob.push(pack_linear(wq, kDgrHead, kDgrLatent, 32, 128));
  const size_t o_wkv = blob.push(pack_linear(next(2 * kDgrHead * kDgrCtx), 2 * kDgrHead, kDgrCtx, 32, 128));   // rows 0-127 = K, 128-255 = V (:91)
  const size_t o_wo = blob.push(pack_linear(next(kDgrLatent * kDgrHead), kDgrLatent, kDgrHead, 32, 128));
  const size_t o_bo = blob.push(next(kDgrLatent));
  const size_t o_lnfg = blob.push(next(kDgrLatent)), o_lnfb = blob.push(next(kDgrLatent));
  // GEGLU (:53-56): output columns 0..1023 = value, 1024..2047 = gate; block b of the GEMM holds value cols [128b,128b+128) then their gates
  std::vector<int> rowmap(2 * kDgrHidden);
  for (int b = 0; b < kDgrHidden / 128; ++b)
    for (int n = 0; n < 256; ++n) rowmap[b * 256 + n] = n < 128 ? b * 128 + n : kDgrHidden + b * 128 + (n - 128);
  const size_t o_w1 = blob.push(pack_linear(next(2 * kDgrHidden * kDgrLatent), 2 * kDgrHidden, kDgrLatent, 32, 256, &rowmap));
  const size_t o_b1 = blob.push(next(2 * kDgrHidden));
  const size_t o_w2 = blob.push(pack_linear(next(kDgrLatent * kDgrHidden), kDgrLatent, kDgrHidden, 32, 128));
  const size_t o_b2 = blob.push(next(kDgrLatent));

  CU(cudaSetDevice(h->device));
  if (h->blob) { CU(cudaDeviceSynchronize()); cudaFree(h->blob); h->blob = nullptr; }
  CU(cudaMalloc(&h->blob, blob.h.size() * sizeof(float)));
  CU(cudaMemcpy(h->blob, blob.h.data(), blob.h.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* d = h->blob;
  if (h->pe) { h->cpe_q_w = d + o_cqw; h->cpe_q_b = d + o_cqb; h->cpe_c_w = d + o_ccw; h->cpe_c_b = d + o_ccb; }
  h->lnq_g = d + o_lnqg; h->lnq_b = d + o_lnqb; h->lnc_g = d + o_lncg; h->lnc_b = d + o_lncb; h->lnf_g = d + o_lnfg; h->lnf_b = d + o_lnfb;
  h->wq = d + o_wq; h->wkv = d + o_wkv; h->wo = d + o_wo; h->bo = d + o_bo; h->w1 = d + o_w1; h->b1 = d + o_b1; h->w2 = d + o_w2; h->b2 = d + o_b2;
  h->loaded = true;
  return 0;
}

// PerceiverIO.forward(data=image_feat, queries_encoder=latents) (perceiver_io.py:187-221): latents [M,256], image_feat [T,128] -> out [M,256]
int gmf_dgr_head_forward(gmf_dgr_head* h, const float* latents, const float* image_feat, int M, int T, float* out, void* stream) {
  if (!h || !latents || !image_feat || !out) return fail(GMF_ERR_INVALID, "NULL argument");
  if (!h->loaded) return fail(GMF_ERR_STATE, "gmf_dgr_head: weights not loaded");
  if (M < 1 || T < 1) return fail(GMF_ERR_INVALID, "gmf_dgr_head_forward: M and T must be >= 1");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int mt = cdiv(M, 128), kt = cdiv(T, 128), xt = std::max(mt, kt);
  DgrWork w;
  const size_t need = dgr_carve(w, nullptr, M, T);
  if (need > h->ws_bytes) {
    CU(cudaDeviceSynchronize());
    if (h->ws) cudaFree(h->ws);
    h->ws = nullptr; h->ws_bytes = 0;
    CU(cudaMalloc(&h->ws, need));
    h->ws_bytes = need;
  }
  dgr_carve(w, (uint8_t*)(((uintptr_t)h->ws + 1023) & ~(uintptr_t)1023), M, T);
  if (h->feat_tiles < xt) {
    // neutral spatial-consistency operand: all-zero coordinates give DA = 0, Y = -1 in the attention kernel, i.e. compat == 1.
    // Every row is the same, so a buffer generated for xt tiles serves any smaller (M, T) as well.
    CU(cudaDeviceSynchronize());
    if (h->feat) cudaFree(h->feat);
    h->feat = nullptr; h->feat_tiles = 0;
    const size_t kp_bytes = (size_t)xt * 128 * 8 * 4, f_bytes = (size_t)xt * 128 * 64 * 2;
    CU(cudaMalloc(&h->feat, kp_bytes + 2 * f_bytes));
    CU(cudaMemsetAsync(h->feat, 0, kp_bytes, st));
    dist_feature_scaled_kernel<<<dim3(xt, 1), 128, 0, st>>>((const float*)h->feat, xt * 128, 1.0f, (__nv_bfloat16*)(h->feat + kp_bytes),
                                                           (__nv_bfloat16*)(h->feat + kp_bytes + f_bytes));
    LAUNCHED();
    h->feat_tiles = xt;
  }
  const __nv_bfloat16* n_aq = (const __nv_bfloat16*)(h->feat + (size_t)h->feat_tiles * 128 * 8 * 4);
  const __nv_bfloat16* n_bd = n_aq + (size_t)h->feat_tiles * 128 * 64;
  const float* resid0 = latents;
  if (h->pe) {
    rows_to_img_kernel<kDgrLatent, true, true><<<mt * 16, 256, 0, st>>>(latents, M, mt, h->cpe_q_w, h->cpe_q_b, h->lnq_g, h->lnq_b, w.x0, w.xq_img);
    LAUNCHED();
    rows_to_img_kernel<kDgrCtx, true, true><<<kt * 16, 256, 0, st>>>(image_feat, T, kt, h->cpe_c_w, h->cpe_c_b, h->lnc_g, h->lnc_b, nullptr, w.c_img);
    LAUNCHED();
    resid0 = w.x0;
  } else {
    rows_to_img_kernel<kDgrLatent, false, true><<<mt * 16, 256, 0, st>>>(latents, M, mt, nullptr, nullptr, h->lnq_g, h->lnq_b, nullptr, w.xq_img);
    LAUNCHED();
    rows_to_img_kernel<kDgrCtx, false, true><<<kt * 16, 256, 0, st>>>(image_feat, T, kt, nullptr, nullptr, h->lnc_g, h->lnc_b, nullptr, w.c_img);
    LAUNCHED();
  }
  cudaError_t e;
  {  // to_q (scaled) -> bf16 Q tiles
    ImgGemmArgs a{};
    a.a_img = w.xq_img; a.w_packed = h->wq; a.K = kDgrLatent; a.L = M; a.tiles = mt; a.t0 = w.q_t;
    e = launch_img_gemm<128, DE_QIMG>(a, 1, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "dgr to_q launch");
  }
  {  // to_kv -> bf16 K tiles, V^T tiles
    ImgGemmArgs a{};
    a.a_img = w.c_img; a.w_packed = h->wkv; a.K = kDgrCtx; a.L = T; a.tiles = kt; a.t0 = w.k_t; a.t1 = w.vt_t;
    e = launch_img_gemm<128, DE_KVIMG>(a, 2, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "dgr to_kv launch");
  }
  {  // softmax(q k^T / sqrt(128)) v
    ScAttnArgs sa{};
    sa.q_t = w.q_t; sa.k_t = w.k_t; sa.vt_t = w.vt_t; sa.aq_t = n_aq; sa.bd_t = n_bd; sa.out = w.o;
    sa.N = T; sa.tiles = kt; sa.Nq = M; sa.q_tiles = mt;
    // few query tiles: split the keys over enough CTAs to cover the SMs (flash-decoding style); the combine is folded into the
    // kernel that builds the to_out GEMM's operand image
    int splits = std::max(1, std::min(kt, 148 / mt));
    const int per = cdiv(kt, splits);
    splits = cdiv(kt, per);
    if (splits > 1) {
      sa.tiles_per_split = per; sa.part_o = w.part_o; sa.part_l = w.part_l;
      e = launch_sc_attn_v9_split(sa, splits, st);
    } else {
      e = launch_sc_attn_v9<0, 2>(sa, 1, st);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "dgr attention launch");
    if (splits > 1) {
      attn_combine_img_kernel<<<mt * 16, 256, 0, st>>>(w.part_o, w.part_l, splits, M, mt, w.o_img);
    } else {
      rows_to_img_kernel<kDgrHead, false, false><<<mt * 16, 256, 0, st>>>(w.o, M, mt, nullptr, nullptr, nullptr, nullptr, nullptr, w.o_img);
    }
    LAUNCHED();
  }
  {  // to_out + bias + residual
    ImgGemmArgs a{};
    a.a_img = w.o_img; a.w_packed = h->wo; a.K = kDgrHead; a.L = M; a.tiles = mt; a.bias = h->bo; a.residual = resid0; a.out = w.x1; a.ld = kDgrLatent;
    e = launch_img_gemm<128, DE_RES>(a, kDgrLatent / 128, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "dgr to_out launch");
  }
  rows_to_img_kernel<kDgrLatent, false, true><<<mt * 16, 256, 0, st>>>(w.x1, M, mt, nullptr, nullptr, h->lnf_g, h->lnf_b, nullptr, w.ln_img);
  LAUNCHED();
  {  // Linear(256, 2048) + GEGLU -> A image of the second FFN GEMM
    ImgGemmArgs a{};
    a.a_img = w.ln_img; a.w_packed = h->w1; a.K = kDgrLatent; a.L = M; a.tiles = mt; a.bias = h->b1; a.out_img = w.g_img;
    a.out_chunks = kDgrHidden / 32; a.hidden = kDgrHidden;
    e = launch_img_gemm<256, DE_GEGLU>(a, kDgrHidden / 128, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "dgr ffn1 launch");
  }
  {  // Linear(1024, 256) + bias + residual
    ImgGemmArgs a{};
    a.a_img = w.g_img; a.w_packed = h->w2; a.K = kDgrHidden; a.L = M; a.tiles = mt; a.bias = h->b2; a.residual = w.x1; a.out = out; a.ld = kDgrLatent;
    e = launch_img_gemm<128, DE_RES>(a, kDgrLatent / 128, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "dgr ffn2 launch");
  }
  return 0;
}

}  // extern "C"
