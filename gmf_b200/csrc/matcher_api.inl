// C-ABI of the correspondence construction step (include/gmf_b200.h, "correspondence construction"); included at the end of gmf_api.cu.

extern "C" {

size_t gmf_match_workspace_bytes(int B, int Ns, int Nt) {
  if (B < 1 || Ns < 1 || Nt < 1) return 0;
  return ((size_t)B * Ns + (size_t)B * Nt) * sizeof(unsigned long long) + 2048;
}

int gmf_build_correspondences(gmf_ctx* ctx, const float* src_desc, const float* tgt_desc, const float* src_keypts, const float* tgt_keypts,
                              int B, int Ns, int Nt, int D, int use_mutual, int32_t* source_idx, int32_t* corr, int32_t* n_corr,
                              float* src_sel, float* tgt_sel, float* corr_pos, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (!src_desc || !tgt_desc || !src_keypts || !tgt_keypts || !source_idx || !corr || !n_corr || !src_sel || !tgt_sel || !corr_pos)
    return fail(GMF_ERR_INVALID, "gmf_build_correspondences: NULL argument");
  if (B < 1 || Ns < 1 || Nt < 1 || D < 1 || D > 1024) return fail(GMF_ERR_INVALID, "gmf_build_correspondences: need B, Ns, Nt >= 1 and 1 <= D <= 1024");
  if (!workspace || workspace_bytes < gmf_match_workspace_bytes(B, Ns, Nt)) return fail(GMF_ERR_STATE, "gmf_build_correspondences: workspace too small");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* best_src = (unsigned long long*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  unsigned long long* best_tgt = best_src + (size_t)B * Ns;
  CU(cudaMemsetAsync(best_src, 0xff, ((size_t)B * Ns + (size_t)(use_mutual ? B : 0) * Nt) * sizeof(unsigned long long), st));
  auto launch = [&](const float* A, const float* Bm, int Na, int Nb, unsigned long long* best) -> int {
    const int rt = cdiv(Na, kNnTile), ctl = cdiv(Nb, kNnTile);
    // enough column chunks to give every SM ~2 CTAs, never more chunks than column tiles
    int chunks = std::max(1, std::min(ctl, cdiv(2 * 148, rt * B)));
    const int per = cdiv(ctl, chunks);
    chunks = cdiv(ctl, per);
    if (D <= 32 && (D & 3) == 0 && (((uintptr_t)A | (uintptr_t)Bm) & 15) == 0) nn_argmin_kernel<true><<<dim3(rt, chunks, B), 256, 0, st>>>(A, Bm, Na, Nb, D, per, best);
    else nn_argmin_kernel<false><<<dim3(rt, chunks, B), 256, 0, st>>>(A, Bm, Na, Nb, D, per, best);
    LAUNCHED();
    return 0;
  };
  TRY(launch(src_desc, tgt_desc, Ns, Nt, best_src));                       // source_idx = argmin(distance, axis=1)
  if (use_mutual) TRY(launch(tgt_desc, src_desc, Nt, Ns, best_tgt));       // target_idx = argmin(distance, axis=0)
  corr_build_kernel<<<B, 1024, 0, st>>>(best_src, use_mutual ? best_tgt : nullptr, src_keypts, tgt_keypts, Ns, Nt, use_mutual ? 1 : 0, source_idx,
                                        corr, src_sel, tgt_sel, corr_pos, n_corr);
  LAUNCHED();
  return 0;
}

}  // extern "C"
