// Spatial-consistency guided non-local attention (PointDSC.py:56-64) with the pairwise squared distances ALSO on the tensor
// pipe:   softmax_j( c_ij * q_i.k_j / sqrt(128) ) v_j,   c_ij = max(0, 1 - (|s_i-s_j| - |t_i-t_j|)^2 / sigma_d^2)   (PointDSC.py:216-221)
//
// |s_i - s_j|^2 = |s_i|^2 + |s_j|^2 - 2 s_i.s_j is bilinear in per-point feature vectors, so it is produced by a K=32 bf16 MMA
// next to S = Q K^T.  fp32 coordinates are split into three bf16 terms (x = x0 + x1 + x2, 24 mantissa bits); the six
// significant cross products per coordinate plus the split squared norms fill 24 of the 32 K slots, the accumulator is fp32,
// so D2 carries the same ~1e-7 |s|^2 absolute error as the fp32 FMA expansion it replaces.  The softmax threads then read
// S, D2S and D2T from TMEM and spend ~10 instructions per (i,j) instead of ~24, and no longer pull key points through the
// shared-memory broadcast path (which was the co-limiter with the issue slots).  The N x N matrix never exists in HBM.
//
// CTA = one 128-query row tile; 64-key pipeline stages {K, V^T, Bd}; TMEM: S[2] | D2S[2] | D2T[2] | O (512 columns).
//   warps 0-7   softmax: two threads per score row (32 columns each), lazy-rescale online softmax, bf16 P -> smem (2 buffers)
//   warp 8      producer: bulk-async copies (TMA engine) into a 3-stage mbarrier ring
//   warp 9      MMA issuer (converged, one elected lane): per key tile 8 (S) + 2 (D2S) + 2 (D2T) + 4 (PV) tcgen05.mma
#pragma once
#include "common.cuh"
#ifndef GMF_SC_RSQ
#define GMF_SC_RSQ 0
#endif
#ifndef GMF_SC_DBG
#define GMF_SC_DBG 0   // timing experiments only: 1 = skip D2 TMEM loads, 2 = no MUFU, 3 = no P smem write
#endif

namespace gmf {

struct ScAttnArgs {
  const __nv_bfloat16* q_t;   // [pairs][tiles][128*128]   (scale * log2e folded into the projection)
  const __nv_bfloat16* k_t;   // [pairs][tiles][128*128]
  const __nv_bfloat16* vt_t;  // [pairs][tiles][128*128]   V^T tiles
  const __nv_bfloat16* aq_t;  // [pairs][tiles][128*64]    query-side distance features (s-part | t-part)
  const __nv_bfloat16* bd_t;  // [pairs][tiles][128*64]    key-side distance features
  float* out;                 // [pairs][N][128] fp32
  int N, tiles;
  float neg_inv_sigma2;
  // gen 9 only: fused head of fc_message (PointDSC.py:13-21,65).  fc1_w != NULL switches it on: instead of msg the kernel
  // writes m2 = ReLU(BN(conv64x64(ReLU(BN(conv128x64(msg))))))  [pairs][N][64]  (BN folded into the packed weights / biases)
  const float* fc1_w;         // pack_linear(W, 64, 128, 32, 64)
  const float* fc1_b;
  const float* fc2_w;         // pack_linear(W, 64, 64, 64, 64)
  const float* fc2_b;
  float* m2_out;
  long long* trace;           // optional timeline of CTA (0,0): [role 0..3][tile][4] clock64 stamps (GMF_SC_TRACE)
};

struct ScCfg {
  static constexpr int D = 128, BN = 64, NS = 3, NV = 3, PB = 2;   // NS: {K, Bd} ring depth, NV: V^T ring depth
  static constexpr int Q_BYTES = 128 * D * 2, AQ_BYTES = 128 * 64 * 2;
  static constexpr int K_BYTES = BN * D * 2, V_BYTES = D * BN * 2, BD_BYTES = BN * 64 * 2;
  static constexpr int KSTAGE_BYTES = K_BYTES + BD_BYTES;   // released as soon as S/D2 of that tile retired (early)
  static constexpr int P_TILE = 128 * BN * 2;
  static constexpr int XCH_BYTES = 3 * 4 * 128 * 4;
  static constexpr int SMEM = 1024 + Q_BYTES + AQ_BYTES + NS * KSTAGE_BYTES + NV * V_BYTES + PB * P_TILE + XCH_BYTES + 512;
  static constexpr int COL_S = 0, COL_DS = 128, COL_DT = 256, COL_O = 384;
};

// CL = CTAs per cluster (1 or 2).  With CL = 2 the two CTAs own adjacent row tiles of the same pair and SHARE the K / V^T / Bd
// stream: each producer fetches half of every stage and multicasts it into both CTAs' shared memory, halving L2 -> SM traffic
// (the limiter once the distances moved to the tensor pipe).  A stage is refilled only after BOTH CTAs retired its MMAs
// (tcgen05.commit multicast onto a 2-arrival mbarrier).
template <int CL, int TPR>   // TPR = softmax threads per score row (2 or 4): 4*TPR softmax warps + producer + MMA warp
__global__ void __launch_bounds__(128 * TPR + 64, 1) sc_attn_tc_kernel(const ScAttnArgs a) {
  using Cfg = ScCfg;
  constexpr int D = Cfg::D, BN = Cfg::BN, NS = Cfg::NS, NV = Cfg::NV, PB = Cfg::PB, HC = BN / TPR;
  constexpr int WP = 4 * TPR, WM = 4 * TPR + 1;           // producer / MMA warp index
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sAq = sQ + Cfg::Q_BYTES;
  uint8_t* sK = sAq + Cfg::AQ_BYTES;                     // [NS] x {K, Bd}
  uint8_t* sV = sK + NS * Cfg::KSTAGE_BYTES;             // [NV] x V^T
  uint8_t* sP = sV + NV * Cfg::V_BYTES;                  // [PB]
  float* sX = (float*)(sP + PB * Cfg::P_TILE);           // [3][TPR][128]
  uint64_t* bars = (uint64_t*)((uint8_t*)sX + Cfg::XCH_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;        // [NS]
  uint64_t* k_empty = k_full + NS;    // [NS]
  uint64_t* v_full = k_empty + NS;    // [NV]
  uint64_t* v_empty = v_full + NV;    // [NV]
  uint64_t* s_full = v_empty + NV;    // [2]
  uint64_t* s_free = s_full + 2;      // [2]
  uint64_t* p_ready = s_free + 2;     // [PB]
  uint64_t* pv_done = p_ready + PB;   // [PB]
  uint32_t* tmem_slot = (uint32_t*)(pv_done + PB + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.y;
  const int qt = min((int)blockIdx.x, a.tiles - 1);          // ghost CTA of an odd tile count recomputes the last tile (no store)
  const bool store = (int)blockIdx.x < a.tiles;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  const int nt = (a.N + BN - 1) / BN;

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], CL); }
    for (int i = 0; i < NV; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128 * TPR); }
    for (int i = 0; i < PB; ++i) { mbar_init(&p_ready[i], 128 * TPR); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
  }
  if (warp == WP) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();      // barrier inits visible cluster-wide before any remote arrive
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == WP) {
    // ------------------------------------ producer ------------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const size_t tq = (size_t)pair * a.tiles + qt;
    mbar_expect_tx_p(q_full, Cfg::Q_BYTES + Cfg::AQ_BYTES, leader);
    bulk_g2s_p(sQ, a.q_t + tq * (128 * D), Cfg::Q_BYTES, q_full, leader);
    bulk_g2s_p(sAq, a.aq_t + tq * (128 * 64), Cfg::AQ_BYTES, q_full, leader);
    // two independent rings: {K, Bd} of tile j is needed when S_j is issued (two tiles ahead of the softmax) and is free again
    // as soon as those MMAs retire; V^T of tile j is needed only at PV_j.  The producer always serves the older request first.
    int jk = 0, jv = 0;
    while (jk < nt || jv < nt) {
      if (jk < nt && (jv >= nt || jk <= jv + 2)) {
        const int st = jk % NS;
        if (jk >= NS) mbar_wait(&k_empty[st], ((jk / NS) - 1) & 1);
        uint8_t* dst = sK + st * Cfg::KSTAGE_BYTES;
        mbar_expect_tx_p(&k_full[st], Cfg::KSTAGE_BYTES, leader);
        const size_t tix = (size_t)pair * a.tiles + (jk >> 1);
        const int h = jk & 1;
        const uint8_t* ksrc = (const uint8_t*)(a.k_t + tix * (128 * D)) + h * 8192;
        const uint8_t* bsrc = (const uint8_t*)(a.bd_t + tix * (128 * 64)) + h * Cfg::BD_BYTES;
        if (CL == 1) {
          bulk_g2s_p(dst, ksrc, 8192, &k_full[st], leader);
          bulk_g2s_p(dst + 8192, ksrc + 16384, 8192, &k_full[st], leader);
          bulk_g2s_p(dst + Cfg::K_BYTES, bsrc, Cfg::BD_BYTES, &k_full[st], leader);
        } else {
          bulk_g2s_mc_p(dst + crank * 8192, ksrc + crank * 16384, 8192, &k_full[st], 0x3, leader);
          bulk_g2s_mc_p(dst + Cfg::K_BYTES + crank * 4096, bsrc + crank * 4096, 4096, &k_full[st], 0x3, leader);
        }
        ++jk;
      } else {
        const int sv_ = jv % NV;
        if (jv >= NV) mbar_wait(&v_empty[sv_], ((jv / NV) - 1) & 1);
        uint8_t* dst = sV + sv_ * Cfg::V_BYTES;
        mbar_expect_tx_p(&v_full[sv_], Cfg::V_BYTES, leader);
        const size_t tix = (size_t)pair * a.tiles + (jv >> 1);
        const uint8_t* vsrc = (const uint8_t*)(a.vt_t + tix * (128 * D)) + (jv & 1) * Cfg::V_BYTES;
        if (CL == 1) bulk_g2s_p(dst, vsrc, Cfg::V_BYTES, &v_full[sv_], leader);
        else bulk_g2s_mc_p(dst + crank * 8192, vsrc + crank * 8192, 8192, &v_full[sv_], 0x3, leader);
        ++jv;
      }
    }
  } else if (warp == WM) {
    // ------------------------------------ MMA issuer ------------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc_s = umma_idesc(128, BN, kFmtBF16);
    const uint32_t idesc_o = umma_idesc(128, D, kFmtBF16);
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t aq_desc = umma_desc_sw128(smem_u32(sAq));
    const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
    const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK));
    const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV));
    auto issue_sd = [&](int j) {
      const int st = j % NS, bb = j & 1;
      mbar_wait(&k_full[st], (j / NS) & 1);
      if (j >= 2) mbar_wait(&s_free[bb], ((j >> 1) - 1) & 1);
      tc_fence_after();
      const uint64_t kd = umma_desc_adv(k_desc0, st * Cfg::KSTAGE_BYTES);
      const uint64_t bd = umma_desc_adv(kd, Cfg::K_BYTES);
#pragma unroll
      for (int at = 0; at < 2; ++at)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_bf16_p(tmem + Cfg::COL_S + bb * BN, umma_desc_adv(q_desc, at * 16384 + ks * 32), umma_desc_adv(kd, at * 8192 + ks * 32),
                        idesc_s, (at | ks) ? 1u : 0u, leader);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)        // source-cloud squared distances: K slots 0..31
        tc_mma_bf16_p(tmem + Cfg::COL_DS + bb * BN, umma_desc_adv(aq_desc, ks * 32), umma_desc_adv(bd, ks * 32), idesc_s, ks ? 1u : 0u, leader);
#pragma unroll
      for (int ks = 2; ks < 4; ++ks)        // target-cloud squared distances: K slots 32..63
        tc_mma_bf16_p(tmem + Cfg::COL_DT + bb * BN, umma_desc_adv(aq_desc, ks * 32), umma_desc_adv(bd, ks * 32), idesc_s, ks > 2 ? 1u : 0u, leader);
      tc_commit_p(&s_full[bb], leader);
      if (CL == 1) tc_commit_p(&k_empty[st], leader);          // the {K, Bd} stage is free once these MMAs retire
      else tc_commit_mc_p(&k_empty[st], 0x3, leader);
    };
    mbar_wait(q_full, 0);
    for (int jj = 0; jj < 2 && jj < nt; ++jj) issue_sd(jj);
    for (int j = 0; j < nt; ++j) {
      const int sv_ = j % NV, pb = j % PB;
      // scores run two key tiles ahead: S/D2 of tile j+2 only need the S buffer drained by softmax(j) (early in its tile) and
      // a {K, Bd} stage that was prefetched long ago, so they are issued BEFORE blocking on P_j
      if (j + 2 < nt) issue_sd(j + 2);
      mbar_wait(&p_ready[pb], (j / PB) & 1);
      mbar_wait(&v_full[sv_], (j / NV) & 1);
      tc_fence_after();
      const uint64_t pd = umma_desc_adv(p_desc, pb * Cfg::P_TILE);
      const uint64_t vd = umma_desc_adv(v_desc0, sv_ * Cfg::V_BYTES);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        tc_mma_bf16_p(tmem + Cfg::COL_O, umma_desc_adv(pd, ks * 32), umma_desc_adv(vd, ks * 32), idesc_o, (j > 0 || ks > 0) ? 1u : 0u, leader);
      tc_commit_p(&pv_done[pb], leader);
      if (CL == 1) tc_commit_p(&v_empty[sv_], leader);
      else tc_commit_mc_p(&v_empty[sv_], 0x3, leader);        // both CTAs must retire a stage before either producer refills it
    }
  } else {
    // ------------------------------------ softmax (TPR threads per query row) ------------------------------------
    const int q = warp & 3, h = warp >> 2;               // TMEM lane quadrant, column slice of the 64-key tile
    const int r = q * 32 + lane;
    const int gq = qt * 128 + r;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t trow = tlane + h * HC;
    const int bar_id = 1 + q;
    // Single-pass online softmax with a DEFERRED reference maximum: tile j is exponentiated against the reference left by
    // tile j-1 (no pre-exp barrier between the threads sharing a row); the row maximum seen in tile j is exchanged after
    // P_j has been published and, if it exceeded the reference by more than 2^8, the O accumulator / row sum are rescaled
    // before P_{j+1} is published.  bf16/fp32 have 8 exponent bits, so one tile of un-normalised weights is harmless.
    // (Rows always contain zero logits here - c_ij = 0 for incompatible pairs - so the initial reference 0 cannot underflow.)
    float m_ref = 0.f, l_sum = 0.f, pend_shift = 0.f;     // pend_shift: rescale decided after the previous tile, not yet applied to O
    const float nis2 = a.neg_inv_sigma2;
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t pk[HC / 2];
      float tmax = -INFINITY;
      {
        uint32_t us[HC], ua[HC], ub[HC];
        if (HC == 32) {
          tmem_ld32(trow + Cfg::COL_S + b * BN, reinterpret_cast<uint32_t(&)[32]>(us));
          tmem_ld32(trow + Cfg::COL_DS + b * BN, reinterpret_cast<uint32_t(&)[32]>(ua));
          tmem_ld32(trow + Cfg::COL_DT + b * BN, reinterpret_cast<uint32_t(&)[32]>(ub));
        } else {
          tmem_ld16(trow + Cfg::COL_S + b * BN, reinterpret_cast<uint32_t(&)[16]>(us));
          tmem_ld16(trow + Cfg::COL_DS + b * BN, reinterpret_cast<uint32_t(&)[16]>(ua));
          tmem_ld16(trow + Cfg::COL_DT + b * BN, reinterpret_cast<uint32_t(&)[16]>(ub));
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[b]);
        const int nvalid = a.N - j * BN - h * HC;
        float ps0 = 0.f, ps1 = 0.f;
        // (|ds| - |dt|)^2 = ds^2 + dt^2 - 2 sqrt(ds^2 dt^2); |.| guards tiny negative d^2 from cancellation
#pragma unroll
        for (int c = 0; c < HC; c += 2) {
          float t0, t1;
          {
            const float d2s = __uint_as_float(ua[c]), d2t = __uint_as_float(ub[c]);
            const float x = fmaf(-2.f, sqrt_approx(fabsf(d2s * d2t)), d2s + d2t);
            t0 = fmaf(__uint_as_float(us[c]), __saturatef(fmaf(x, nis2, 1.f)), -m_ref);
          }
          {
            const float d2s = __uint_as_float(ua[c + 1]), d2t = __uint_as_float(ub[c + 1]);
            const float x = fmaf(-2.f, sqrt_approx(fabsf(d2s * d2t)), d2s + d2t);
            t1 = fmaf(__uint_as_float(us[c + 1]), __saturatef(fmaf(x, nis2, 1.f)), -m_ref);
          }
          if (nvalid < HC) {                               // ragged last tile (warp-uniform)
            if (c >= nvalid) t0 = -INFINITY;
            if (c + 1 >= nvalid) t1 = -INFINITY;
          }
          tmax = fmaxf(tmax, fmaxf(t0, t1));
          const float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
          ps0 += p0; ps1 += p1;
          pk[c >> 1] = pack_bf16(p0, p1);
        }
        l_sum += ps0 + ps1;
      }
      const int pb = j % PB;
      uint8_t* myP = sP + pb * Cfg::P_TILE;
      if (j >= PB) mbar_wait(&pv_done[pb], ((j / PB) - 1) & 1);        // P buffer free again (PV of tile j-PB retired)
      if (__any_sync(0xffffffffu, pend_shift != 0.f)) {
        // rare: the previous tile raised the reference; P_j above already used the new reference, O still holds the old one
        mbar_wait(&pv_done[(j - 1) % PB], ((j - 1) / PB) & 1);
        tc_fence_after();
        const float alpha = ex2_approx(-pend_shift);
#pragma unroll
        for (int c = 0; c < D / TPR / 32; ++c) {
          uint32_t u[32];
          tmem_ld32(tlane + Cfg::COL_O + h * (D / TPR) + c * 32, u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(__uint_as_float(u[i]) * alpha);
          tmem_st32(tlane + Cfg::COL_O + h * (D / TPR) + c * 32, u);
        }
        tmem_st_wait();
        pend_shift = 0.f;
      }
#if GMF_SC_DBG != 3
#pragma unroll
      for (int c8 = 0; c8 < HC / 8; ++c8)
        *reinterpret_cast<uint4*>(myP + swz_off(r, h * (HC / 8) + c8)) = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
      fence_proxy_async();
#else
      if (pk[0] == 0x12345678u) *reinterpret_cast<uint4*>(myP + swz_off(r, h * (HC / 8))) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
#endif
      tc_fence_before();
      mbar_arrive(&p_ready[pb]);
      // deferred row maximum: off the MMA critical path
#if GMF_SC_DBG != 4
      sX[(b * TPR + h) * 128 + r] = tmax;
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * TPR) : "memory");
#pragma unroll
      for (int o = 1; o < TPR; ++o) tmax = fmaxf(tmax, sX[(b * TPR + ((h + o) % TPR)) * 128 + r]);
#endif
      if (tmax > 8.f && j + 1 < nt) {                      // identical decision in all TPR threads of the row
        m_ref += tmax;
        l_sum *= ex2_approx(-tmax);
        pend_shift = tmax;
      }
    }
    sX[(2 * TPR + h) * 128 + r] = l_sum;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * TPR) : "memory");
#pragma unroll
    for (int o = 1; o < TPR; ++o) l_sum += sX[(2 * TPR + ((h + o) % TPR)) * 128 + r];
    mbar_wait(&pv_done[(nt - 1) % PB], ((nt - 1) / PB) & 1);
    tc_fence_after();
    const float inv = 1.f / l_sum;
    float* op = a.out + ((size_t)pair * a.N + gq) * D + h * (D / TPR);
#pragma unroll
    for (int c = 0; c < D / TPR / 32; ++c) {
      uint32_t u[32];
      tmem_ld32(tlane + Cfg::COL_O + h * (D / TPR) + c * 32, u);
      tmem_ld_wait();
      if (store && gq < a.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(op + c * 32 + 4 * i) =
              make_float4(__uint_as_float(u[4 * i]) * inv, __uint_as_float(u[4 * i + 1]) * inv,
                          __uint_as_float(u[4 * i + 2]) * inv, __uint_as_float(u[4 * i + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();      // the peer may still arrive on / multicast into this CTA's smem
  if (warp == WP) tmem_dealloc(tmem, 512);
}

template <int CL, int TPR>
inline cudaError_t launch_sc_attn(const ScAttnArgs& a, int pairs, cudaStream_t st) {
  static bool configured = false;
  auto kern = sc_attn_tc_kernel<CL, TPR>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ScCfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(((a.tiles + CL - 1) / CL) * CL, pairs);
  cfg.blockDim = dim3(128 * TPR + 64);
  cfg.dynamicSmemBytes = ScCfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a);
}

// --------------------------------------------------------------------------------------------------------------------
// distance-feature tiles: per point, 64 bf16 = [s-part (32) | t-part (32)], written in the 128B-swizzled tile image.
// query side (A):  per coord c: (-2u0,-2u0,-2u1,-2u1,-2u0,-2u2), then |u|^2 split (n0,n1,n2), then (1,1,1), zeros
// key side   (B):  per coord c: (  w0,  w1,  w0,  w1,  w2,  w0), then (1,1,1), then |w|^2 split (m0,m1,m2), zeros
// --------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split3(float x, __nv_bfloat16& h0, __nv_bfloat16& h1, __nv_bfloat16& h2) {
  h0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h0);
  h1 = __float2bfloat16_rn(r1);
  h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
}

__global__ void dist_feature_kernel(const float* __restrict__ kpts, int Np, __nv_bfloat16* __restrict__ aq_t, __nv_bfloat16* __restrict__ bd_t) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // padded point index
  if (i >= Np) return;
  const float4 s4 = *reinterpret_cast<const float4*>(kpts + ((size_t)pair * Np + i) * 8);
  const float4 t4 = *reinterpret_cast<const float4*>(kpts + ((size_t)pair * Np + i) * 8 + 4);
  __align__(16) __nv_bfloat16 A[64];
  __align__(16) __nv_bfloat16 B[64];
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f);
#pragma unroll
  for (int k = 0; k < 64; ++k) { A[k] = zero; B[k] = zero; }
  const float pts[2][4] = {{s4.x, s4.y, s4.z, s4.w}, {t4.x, t4.y, t4.z, t4.w}};
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const int o = part * 32;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      __nv_bfloat16 h0, h1, h2;
      split3(pts[part][c], h0, h1, h2);
      const __nv_bfloat16 m0 = __float2bfloat16_rn(-2.f * __bfloat162float(h0)), m1 = __float2bfloat16_rn(-2.f * __bfloat162float(h1)),
                          m2 = __float2bfloat16_rn(-2.f * __bfloat162float(h2));
      A[o + 6 * c + 0] = m0; B[o + 6 * c + 0] = h0;
      A[o + 6 * c + 1] = m0; B[o + 6 * c + 1] = h1;
      A[o + 6 * c + 2] = m1; B[o + 6 * c + 2] = h0;
      A[o + 6 * c + 3] = m1; B[o + 6 * c + 3] = h1;
      A[o + 6 * c + 4] = m0; B[o + 6 * c + 4] = h2;
      A[o + 6 * c + 5] = m2; B[o + 6 * c + 5] = h0;
    }
    __nv_bfloat16 n0, n1, n2;
    split3(pts[part][3], n0, n1, n2);
    A[o + 18] = n0; A[o + 19] = n1; A[o + 20] = n2; B[o + 18] = one; B[o + 19] = one; B[o + 20] = one;
    A[o + 21] = one; A[o + 22] = one; A[o + 23] = one; B[o + 21] = n0; B[o + 22] = n1; B[o + 23] = n2;
  }
  const int tile = i >> 7, r = i & 127;
  const size_t tbase = ((size_t)pair * (Np >> 7) + tile) * (128 * 64);
  uint8_t* ad = (uint8_t*)(aq_t + tbase);
  uint8_t* bd = (uint8_t*)(bd_t + tbase);
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    *reinterpret_cast<uint4*>(ad + swz_off(r, ch)) = *reinterpret_cast<const uint4*>(&A[ch * 8]);
    *reinterpret_cast<uint4*>(bd + swz_off(r, ch)) = *reinterpret_cast<const uint4*>(&B[ch * 8]);
  }
}

}  // namespace gmf
