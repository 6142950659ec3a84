// C-ABI of the classical SM baseline (include/gmf_b200.h, "SM baseline"); included at the end of gmf_api.cu.
// Reference: GMF_PointDSC/baseline_scripts/baseline_3DMatch.py:19-53.

namespace {
struct SmWork { float4 *src4, *tgt4; float *y, *v, *w; int* top; };
size_t sm_carve(SmWork& w, uint8_t* base, int B, int N, int S) {
  Bump b{base};
  w.src4 = b.take<float4>((size_t)B * N); w.tgt4 = b.take<float4>((size_t)B * N);
  w.y = b.take<float>((size_t)B * N); w.v = b.take<float>((size_t)B * N); w.w = b.take<float>((size_t)B * N);
  w.top = b.take<int>((size_t)B * std::max(S, 1));
  return b.off + 1024;
}
__global__ void sm_pack_points_kernel(const float* __restrict__ src, const float* __restrict__ tgt, long long n, float4* __restrict__ s4, float4* __restrict__ t4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  s4[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
  t4[i] = make_float4(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2], 0.f);
}
}  // namespace

extern "C" {

size_t gmf_sm_workspace_bytes(int B, int N, double top_ratio) {
  if (B < 1 || N < 2) return 0;
  SmWork w;
  return sm_carve(w, nullptr, B, N, (int)((double)N * top_ratio)) + 1024;
}

int gmf_sm_baseline(gmf_ctx* ctx, const float* src, const float* tgt, int B, int N, float inlier_threshold, double top_ratio, int iters,
                    float* trans, float* labels, float* leading_eig, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (!src || !tgt || !trans || !labels) return fail(GMF_ERR_INVALID, "gmf_sm_baseline: NULL argument");
  if (B < 1 || N < 2 || iters < 1 || !(inlier_threshold > 0.f)) return fail(GMF_ERR_INVALID, "gmf_sm_baseline: need B >= 1, N >= 2, iters >= 1, threshold > 0");
  const int S = (int)((double)N * top_ratio);                 // int(leading_eig.shape[1] * top_ratio) (:45)
  if (S < 1 || S > N) return fail(GMF_ERR_INVALID, "gmf_sm_baseline: int(N * top_ratio) must be in [1, N]");
  int np2 = 1;
  while (np2 < N) np2 <<= 1;
  if (np2 > 16384) return fail(GMF_ERR_INVALID, "gmf_sm_baseline supports N <= 16384");
  if (!workspace) return fail(GMF_ERR_INVALID, "workspace is NULL");
  SmWork w;
  const size_t need = sm_carve(w, nullptr, B, N, S) + 1024;
  if (workspace_bytes < need) return fail(GMF_ERR_STATE, "workspace too small: need " + std::to_string(need) + " bytes");
  sm_carve(w, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), B, N, S);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * N;
  sm_pack_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, tgt, n, w.src4, w.tgt4);
  LAUNCHED();
  const float sigma = inlier_threshold / 3.0f;                 // :33
  const float inv_2s2 = 1.0f / (2.0f * sigma * sigma);
  for (int it = 0; it < iters; ++it) {
    sm_matvec_kernel<<<dim3(cdiv(N, 128), B), 512, 0, st>>>(w.src4, w.tgt4, it == 0 ? nullptr : w.v, N, inv_2s2, w.y);
    LAUNCHED();
    sm_normalize_kernel<<<B, 1024, 0, st>>>(w.y, N, w.v);
    LAUNCHED();
  }
  static std::atomic<unsigned long long> configured{0};
  CU(ensure_dyn_smem(topk_sort_kernel, 16384 * 8, configured));
  topk_sort_kernel<<<B, 1024, (size_t)np2 * 8, st>>>(w.v, 1, N, np2, S, w.top);
  LAUNCHED();
  sm_labels_kernel<<<dim3(cdiv(N, 256), B), 256, 0, st>>>(w.v, w.top, N, S, labels, w.w);
  LAUNCHED();
  sm_scatter_kernel<<<dim3(cdiv(S, 256), B), 256, 0, st>>>(w.v, w.top, N, S, labels, w.w);
  LAUNCHED();
  rigid_transform_kernel<<<cdiv(B, 4), 128, 0, st>>>(src, tgt, w.w, B, N, trans);
  LAUNCHED();
  if (leading_eig) CU(cudaMemcpyAsync(leading_eig, w.v, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
