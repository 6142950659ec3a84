// C-ABI of the correspondence construction step (include/gmf_b200.h, "correspondence construction"); included at the end of gmf_api.cu.

extern "C" {

size_t gmf_match_workspace_bytes(int B, int Ns, int Nt, int D) {
  if (B < 1 || Ns < 1 || Nt < 1 || D < 1) return 0;
  const size_t kd = (size_t)cdiv(D, 32), img = (size_t)3 * kd * 4096 * sizeof(float);      // bytes of one 128-row operand tile
  // best keys (src, tgt) + split-tf32 operand images: src as A and as B rows, tgt as B and as A rows
  return ((size_t)B * Ns + (size_t)B * Nt) * sizeof(unsigned long long) + 2 * (size_t)B * (cdiv(Ns, 128) + cdiv(Nt, 128)) * img + 8192;
}

int gmf_build_correspondences(gmf_ctx* ctx, const float* src_desc, const float* tgt_desc, const float* src_keypts, const float* tgt_keypts,
                              int B, int Ns, int Nt, int D, int use_mutual, int32_t* source_idx, int32_t* corr, int32_t* n_corr,
                              float* src_sel, float* tgt_sel, float* corr_pos, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (!src_desc || !tgt_desc || !src_keypts || !tgt_keypts || !source_idx || !corr || !n_corr || !src_sel || !tgt_sel || !corr_pos)
    return fail(GMF_ERR_INVALID, "gmf_build_correspondences: NULL argument");
  if (B < 1 || Ns < 1 || Nt < 1 || D < 1 || D > 1024) return fail(GMF_ERR_INVALID, "gmf_build_correspondences: need B, Ns, Nt >= 1 and 1 <= D <= 1024");
  if (!workspace || workspace_bytes < gmf_match_workspace_bytes(B, Ns, Nt, D)) return fail(GMF_ERR_STATE, "gmf_build_correspondences: workspace too small");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  Bump b{(uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023)};
  unsigned long long* best_src = b.take<unsigned long long>((size_t)B * Ns + (size_t)B * Nt);
  unsigned long long* best_tgt = best_src + (size_t)B * Ns;
  CU(cudaMemsetAsync(best_src, 0xff, ((size_t)B * Ns + (size_t)(use_mutual ? B : 0) * Nt) * sizeof(unsigned long long), st));
  if (ctx->match_impl >= 1) {
    // tensor pipe: error-compensated tf32 (x = hi + lo, K = 96 ceil(D / 32)) through the image GEMM kernel with the argmin in its epilogue
    const int kd = cdiv(D, 32), rs = cdiv(Ns, 128), rt = cdiv(Nt, 128);
    const size_t tile_f = (size_t)3 * kd * 4096;
    float* src_a = b.take<float>((size_t)B * rs * tile_f);
    float* tgt_b = b.take<float>((size_t)B * rt * tile_f);
    desc_operand_kernel<<<dim3(rs * 16, B), 256, 0, st>>>(src_desc, D, kd, 1, Ns, rs, src_a);
    LAUNCHED();
    desc_operand_kernel<<<dim3(rt * 16, B), 256, 0, st>>>(tgt_desc, D, kd, 0, Nt, rt, tgt_b);
    LAUNCHED();
    auto gemm = [&](const float* A, const float* Bm, int Na, int Nb, unsigned long long* best) -> int {
      ImgGemmArgs a{};
      a.a_img = A; a.w_packed = Bm; a.K = 96 * kd; a.L = Na; a.tiles = cdiv(Na, 128); a.ncols = Nb; a.best = best;
      a.a_pair_stride = (size_t)cdiv(Na, 128) * tile_f; a.w_pair_stride = (size_t)cdiv(Nb, 128) * tile_f;
      cudaError_t e = launch_img_gemm<128, DE_ARGMIN>(a, cdiv(Nb, 128), st, B);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (e != cudaSuccess) return fail_cuda(e, "matcher GEMM launch");
      return 0;
    };
    TRY(gemm(src_a, tgt_b, Ns, Nt, best_src));                              // source_idx = argmin(distance, axis=1)
    if (use_mutual) {                                                       // target_idx = argmin(distance, axis=0)
      float* tgt_a = b.take<float>((size_t)B * rt * tile_f);
      float* src_b = b.take<float>((size_t)B * rs * tile_f);
      desc_operand_kernel<<<dim3(rt * 16, B), 256, 0, st>>>(tgt_desc, D, kd, 1, Nt, rt, tgt_a);
      LAUNCHED();
      desc_operand_kernel<<<dim3(rs * 16, B), 256, 0, st>>>(src_desc, D, kd, 0, Ns, rs, src_b);
      LAUNCHED();
      TRY(gemm(tgt_a, src_b, Nt, Ns, best_tgt));
    }
  } else {
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
  }
  corr_build_kernel<<<B, 1024, 0, st>>>(best_src, use_mutual ? best_tgt : nullptr, src_keypts, tgt_keypts, Ns, Nt, use_mutual ? 1 : 0, source_idx,
                                        corr, src_sel, tgt_sel, corr_pos, n_corr);
  LAUNCHED();
  return 0;
}

}  // extern "C"
