// Training-mode output `M` of PointDSC.forward (models/PointDSC.py:231-234): the feature-space compatibility matrix
//   M = clamp(1 - (1 - Fn Fn^T) / sigma^2, 0, 1),  diag = 0,   Fn = F.normalize(corr_features)
// as one tensor-pipe GEMM at fp32 accuracy (error-compensated tf32, K = 384: knn_operand_kernel + img_gemm_kernel<128, DE_COMPAT>).
// Included at the end of gmf_api.cu.

extern "C" {

size_t gmf_feature_compat_workspace_bytes(int B, int N) {
  if (B < 1 || N < 1) return 0;
  const size_t nt = (size_t)cdiv(N, 128);
  return (size_t)B * N * 128 * 4 + (size_t)B * N * 4 + 2 * (size_t)B * nt * 12 * 4096 * 4 + 8192;
}

int gmf_feature_compat(gmf_ctx* ctx, const float* feat, int B, int N, float* M, void* workspace, size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (!feat || !M) return fail(GMF_ERR_INVALID, "gmf_feature_compat: NULL argument");
  if (B < 1 || N < 1) return fail(GMF_ERR_INVALID, "gmf_feature_compat: need B, N >= 1");
  if (!workspace || workspace_bytes < gmf_feature_compat_workspace_bytes(B, N)) return fail(GMF_ERR_STATE, "gmf_feature_compat: workspace too small");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nt = cdiv(N, 128);
  Bump b{(uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023)};
  float* normed = b.take<float>((size_t)B * N * 128);
  float* conf = b.take<float>((size_t)B * N);
  float* img_a = b.take<float>((size_t)B * nt * 12 * 4096);
  float* img_b = b.take<float>((size_t)B * nt * 12 * 4096);
  const long long rows = (long long)B * N;
  TRY(run_classify(ctx, feat, rows, normed, conf, st));          // F.normalize (:229); the logits are a by-product here
  knn_operand_kernel<<<dim3(nt * 16, B), 256, 0, st>>>(normed, nullptr, 1, N, N, nt, img_a);
  LAUNCHED();
  knn_operand_kernel<<<dim3(nt * 16, B), 256, 0, st>>>(normed, nullptr, 0, N, N, nt, img_b);
  LAUNCHED();
  ImgGemmArgs a{};
  a.a_img = img_a; a.w_packed = img_b; a.K = 384; a.L = N; a.tiles = nt; a.out = M; a.ld = N; a.ncols = N;
  a.a_pair_stride = (size_t)nt * 12 * 4096; a.w_pair_stride = a.a_pair_stride; a.out_pair_stride = (size_t)N * N;
  a.scale = 1.0f / (ctx->sigma * ctx->sigma);
  cudaError_t e = launch_img_gemm<128, DE_COMPAT>(a, nt, st, B);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "feature compat GEMM launch");
  return 0;
}

}  // extern "C"
