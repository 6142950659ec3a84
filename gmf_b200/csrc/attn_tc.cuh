// Flash-style attention on tcgen05/TMEM for the two attention types of GMF-PointDSC:
//   * fusion cross-attention (Fusion-1 / Fusion-2), head dim 64:   softmax(q k^T / sqrt(64)) v      fusion_layer.py:82-94
//   * spatial-consistency guided non-local attention, head dim 128: softmax(c_ij * q k^T / sqrt(128)) v  PointDSC.py:56-64
//     with c_ij = max(0, 1 - (|s_i-s_j| - |t_i-t_j|)^2 / sigma_d^2)  (PointDSC.py:216-221) recomputed on the fly from the
//     point coordinates inside the softmax loop — the N x N matrix never exists in HBM.
//
// Operands arrive as bf16 tiles already in the 128B-swizzled K-major image (see common.cuh) written by the projection
// GEMM epilogues: Q tiles [128 x D], K tiles [128 keys x D], V^T tiles [D x 128 keys].  The 1/sqrt(D) * log2(e) factor is
// folded into the Q projection weights, so scores are in log2 units.
//
// CTA = TWO 128-query row tiles of one pair (256 queries), key tiles of 64, 18 warps:
//   warps 0-7 / 8-15  softmax group 0 / 1 (one per row tile).  Two threads share a score row (TMEM lane): warp (q, h) owns
//              lanes 32q..32q+31 and the 32-column half h of the 64-key tile: tcgen05.ld -> (x c_ij) -> online softmax with
//              lazy rescale (row max exchanged between the two halves through shared memory + a 64-thread named barrier)
//              -> bf16 P written to shared memory in the UMMA image.  4 softmax warps per scheduler hide each other's
//              MUFU / TMEM / barrier latencies.
//   warp 16    producer: one thread issues bulk-async copies (TMA engine) of K / V^T / key-point half-tiles into an
//              mbarrier ring shared by both row tiles
//   warp 17    MMA: one thread, event driven: S = Q K^T (per row tile double-buffered in TMEM) and O += P V
#pragma once
#include "common.cuh"

namespace gmf {

struct AttnArgs {
  const __nv_bfloat16* q_t;   // [pairs][q_tiles][128*D]
  const __nv_bfloat16* k_t;   // [pairs][k_tiles][128*D]
  const __nv_bfloat16* vt_t;  // [pairs][k_tiles][D*128]
  const float* kpts;          // SC only: [pairs][k_tiles*128][8] = (sx,sy,sz,|s|^2,tx,ty,tz,|t|^2), centred
  float* out;                 // [pairs][Lq][D] fp32
  int Lq, Lk, q_tiles, k_tiles;   // tiles of 128 rows
  float neg_inv_sigma2;       // SC only: -1/sigma_d^2
  // fusion attention gen 2 only: fused to_out (fusion_layer.py:94) + residual.  wo_packed != NULL switches it on:
  //   xout[B, Lq, 128] = softmax(..) v . Wo^T + bo + resid        (`out` is then unused)
  const float* wo_packed;     // [128 x 64] tf32, swizzled K-major image (pack_linear(wo, 128, 64, 64, 128))
  const float* bo;            // [128]
  const float* resid;         // [B, Lq, 128]
  float* xout;                // [B, Lq, 128]
};

template <int D, bool SC>
struct AttnCfg {
  static constexpr int BN = 64;                        // keys per pipeline stage
  static constexpr int NSTAGE = SC ? 3 : 6;            // K/V ring depth (TMA latency ~ one softmax period for D=64)
  static constexpr int PBUF = SC ? 1 : 2;              // P buffers per row tile
  static constexpr int Q_TILE = 128 * D * 2;           // one row tile of Q
  static constexpr int K_BYTES = BN * D * 2;           // D/64 atoms of 64 rows x 128 B
  static constexpr int V_BYTES = D * BN * 2;           // one atom: D rows x 128 B
  static constexpr int PTS_BYTES = SC ? BN * 32 : 0;
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES + PTS_BYTES;
  static constexpr int P_TILE = 128 * BN * 2;
  static constexpr int XCH_BYTES = 2 * 3 * 2 * 128 * 4;   // row max / row sum exchange [row tile][slot][half][row]
  static constexpr int SMEM = 1024 + 2 * Q_TILE + NSTAGE * STAGE_BYTES + 2 * PBUF * P_TILE + XCH_BYTES + 512;
  static constexpr int TMEM_COLS = 512;                // S[t][b] at (2t+b)*64, O[t] at 256 + 128 t
  static constexpr int O_COL = 256;
};

template <int D, bool SC>
__global__ void __launch_bounds__(576, 1) attn_tc_kernel(const AttnArgs a) {
  using Cfg = AttnCfg<D, SC>;
  constexpr int BN = Cfg::BN, NS = Cfg::NSTAGE, PB = Cfg::PBUF;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS)
  uint8_t* sQ = smem;                                   // [2] row tiles
  uint8_t* sStage = sQ + 2 * Cfg::Q_TILE;               // [NS] x {K, V^T, pts}
  uint8_t* sP = sStage + NS * Cfg::STAGE_BYTES;         // [2 row tiles][PBUF]
  float* sX = (float*)(sP + 2 * PB * Cfg::P_TILE);    // [2][3][2][128]
  uint64_t* bars = (uint64_t*)((uint8_t*)sX + Cfg::XCH_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [NS]
  uint64_t* kv_empty = kv_full + NS;  // [NS]
  uint64_t* s_full = kv_empty + NS;   // [2][2]
  uint64_t* s_free = s_full + 4;      // [2][2]
  uint64_t* p_ready = s_free + 4;     // [2][PB]
  uint64_t* pv_done = p_ready + 2 * PB;   // [2][PB]
  uint32_t* tmem_slot = (uint32_t*)(pv_done + 2 * PB + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.y;
  const int qt0 = blockIdx.x * 2;                                    // first row tile of this CTA
  const int ntile = (qt0 + 1 < a.q_tiles) ? 2 : 1;                  // active row tiles
  const int nt = (a.Lk + BN - 1) / BN;                               // key half-tiles

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 256); }
    for (int i = 0; i < 2 * PB; ++i) { mbar_init(&p_ready[i], 256); mbar_init(&pv_done[i], 1); }
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 16) {
    // ------------------------------------ producer (whole warp converged; one elected lane issues) -------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    mbar_expect_tx_p(q_full, ntile * Cfg::Q_TILE, leader);
    bulk_g2s_p(sQ, a.q_t + (size_t)(pair * a.q_tiles + qt0) * (128 * D), ntile * Cfg::Q_TILE, q_full, leader);
    for (int j = 0; j < nt; ++j) {
      const int st = j % NS;
      if (j >= NS) mbar_wait(&kv_empty[st], ((j / NS) - 1) & 1);
      uint8_t* dst = sStage + st * Cfg::STAGE_BYTES;
      mbar_expect_tx_p(&kv_full[st], Cfg::STAGE_BYTES, leader);
      const size_t tix = (size_t)pair * a.k_tiles + (j >> 1);      // 128-key tile, half h
      const int h = j & 1;
      const uint8_t* ksrc = (const uint8_t*)(a.k_t + tix * (128 * D)) + h * 8192;
#pragma unroll
      for (int at = 0; at < D / 64; ++at) bulk_g2s_p(dst + at * 8192, ksrc + at * 16384, 8192, &kv_full[st], leader);
      bulk_g2s_p(dst + Cfg::K_BYTES, (const uint8_t*)(a.vt_t + tix * (128 * D)) + h * Cfg::V_BYTES, Cfg::V_BYTES, &kv_full[st], leader);
      if (SC) bulk_g2s_p(dst + Cfg::K_BYTES + Cfg::V_BYTES, a.kpts + (tix * 128 + h * 64) * 8, Cfg::PTS_BYTES, &kv_full[st], leader);
    }
  } else if (warp == 17) {
    // ------------------------------------ MMA issuer (whole warp converged; one elected lane issues) -----------
    // In-order schedule per key tile j and row tile t: O_t += P_{t,j} V_j, then S_{t,j+2} = Q_t K_{j+2}^T, i.e. the scores
    // run two key tiles ahead of the softmax (two S buffers per row tile).  Descriptors are built once; per MMA only the
    // 14-bit start-address field advances.
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc_s = umma_idesc(128, BN, kFmtBF16);
    const uint32_t idesc_o = umma_idesc(128, D, kFmtBF16);
    const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
    const uint64_t p_desc = umma_desc_sw128(smem_u32(sP));
    const uint64_t st_desc = umma_desc_sw128(smem_u32(sStage));
    auto issue_s = [&](int t, int j) {
      const int st = j % NS, bb = j & 1;
      mbar_wait(&kv_full[st], (j / NS) & 1);
      if (j >= 2) mbar_wait(&s_free[t * 2 + bb], ((j >> 1) - 1) & 1);
      tc_fence_after();
      const uint64_t qd = umma_desc_adv(q_desc, t * Cfg::Q_TILE);
      const uint64_t kd = umma_desc_adv(st_desc, st * Cfg::STAGE_BYTES);
#pragma unroll
      for (int at = 0; at < D / 64; ++at)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_bf16_p(tmem + (t * 2 + bb) * BN, umma_desc_adv(qd, at * 16384 + ks * 32), umma_desc_adv(kd, at * 8192 + ks * 32),
                        idesc_s, (at | ks) ? 1u : 0u, leader);
      tc_commit_p(&s_full[t * 2 + bb], leader);
    };
    mbar_wait(q_full, 0);
    for (int jj = 0; jj < 2 && jj < nt; ++jj)
      for (int t = 0; t < ntile; ++t) issue_s(t, jj);
    for (int j = 0; j < nt; ++j) {
      const int st = j % NS, pb = j % PB;
      for (int t = 0; t < ntile; ++t) {
        if (j + 2 < nt) issue_s(t, j + 2);                    // only needs the S buffer drained by softmax(t, j): issue before blocking on P
        mbar_wait(&p_ready[t * PB + pb], (j / PB) & 1);
        tc_fence_after();
        const uint64_t pd = umma_desc_adv(p_desc, (t * PB + pb) * Cfg::P_TILE);
        const uint64_t vd = umma_desc_adv(st_desc, st * Cfg::STAGE_BYTES + Cfg::K_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_bf16_p(tmem + Cfg::O_COL + t * 128, umma_desc_adv(pd, ks * 32), umma_desc_adv(vd, ks * 32), idesc_o,
                        (j > 0 || ks > 0) ? 1u : 0u, leader);
        tc_commit_p(&pv_done[t * PB + pb], leader);
        if (t == ntile - 1) tc_commit_p(&kv_empty[st], leader);     // both row tiles are done with this K/V stage
      }
    }
  } else if ((warp >> 3) < ntile) {
    // ------------------------------------ softmax (two threads per query row) ------------------------------------
    constexpr int HC = BN / 2;                            // 32 score columns per thread
    const int t = warp >> 3;                              // row tile / softmax group
    const int q = warp & 3, h = (warp >> 2) & 1;          // TMEM lane quadrant, column half
    const int r = q * 32 + lane;                          // row in tile == TMEM lane
    const int gq = (qt0 + t) * 128 + r;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int bar_id = 1 + t * 4 + q;                     // named barrier shared by the two halves of these 32 rows
    float* xch = sX + t * 768;
    float qsx = 0.f, qsy = 0.f, qsz = 0.f, qsn = 0.f, qtx = 0.f, qty = 0.f, qtz = 0.f, qtn = 0.f;
    if (SC) {
      // queries == keys in self attention: read the query point from the same (zero-padded) key-point array
      const float4 s4 = *reinterpret_cast<const float4*>(a.kpts + ((size_t)pair * a.k_tiles * 128 + gq) * 8);
      const float4 t4 = *reinterpret_cast<const float4*>(a.kpts + ((size_t)pair * a.k_tiles * 128 + gq) * 8 + 4);
      qsx = -2.f * s4.x; qsy = -2.f * s4.y; qsz = -2.f * s4.z; qsn = s4.w;
      qtx = -2.f * t4.x; qty = -2.f * t4.y; qtz = -2.f * t4.z; qtn = t4.w;
    }
    float m_ref = 0.f, l_sum = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1, st = j % NS;
      mbar_wait(&s_full[t * 2 + b], (j >> 1) & 1);
      tc_fence_after();
      float sv[HC];
      {
        uint32_t u0[32];
        tmem_ld32(trow + (t * 2 + b) * BN + h * HC, u0);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HC; ++i) sv[i] = __uint_as_float(u0[i]);
      }
      tc_fence_before();
      mbar_arrive(&s_free[t * 2 + b]);
      // pass 1: t' = S * c_ij - m_ref (stale reference max), running max of this half
      float tmax = -INFINITY;
      if (SC) {
        mbar_wait(&kv_full[st], (j / NS) & 1);           // acquire the TMA-written key points
        const float4* kp = reinterpret_cast<const float4*>(sStage + st * Cfg::STAGE_BYTES + Cfg::K_BYTES + Cfg::V_BYTES) + 2 * h * HC;
#pragma unroll
        for (int c = 0; c < HC; ++c) {
          const float4 ks = kp[2 * c], kt = kp[2 * c + 1];
          const float d2s = fmaf(qsx, ks.x, fmaf(qsy, ks.y, fmaf(qsz, ks.z, qsn + ks.w)));
          const float d2t = fmaf(qtx, kt.x, fmaf(qty, kt.y, fmaf(qtz, kt.z, qtn + kt.w)));
          // (|ds| - |dt|)^2 = ds^2 + dt^2 - 2 sqrt(ds^2 dt^2): one MUFU instead of two; |.| guards tiny negative d^2
          const float x = fmaf(-2.f, sqrt_approx(fabsf(d2s * d2t)), d2s + d2t);
          const float cij = __saturatef(fmaf(x, a.neg_inv_sigma2, 1.f));
          sv[c] = fmaf(sv[c], cij, -m_ref);
          tmax = fmaxf(tmax, sv[c]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < HC; ++c) tmax = fmaxf(tmax, sv[c]);
      }
      const int nvalid = a.Lk - j * BN - h * HC;           // valid columns of this half
      if (nvalid < HC) {                                   // ragged last tile (warp-uniform branch)
        tmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < HC; ++c) {
          if (c >= nvalid) sv[c] = -INFINITY;
          tmax = fmaxf(tmax, sv[c]);
        }
      }
      // row max over both halves
      xch[(b * 2 + h) * 128 + r] = tmax;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      tmax = fmaxf(tmax, xch[(b * 2 + (h ^ 1)) * 128 + r]);
      // lazy rescale: move the reference max only when the tile max exceeds it by more than 8 (P <= 2^8); the first tile
      // also re-centres a very negative row so that nothing underflows
      const float rel = SC ? tmax : tmax - m_ref;          // SC scores already carry -m_ref
      const bool need = (rel > 8.f) || (j == 0 && rel < -8.f);
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(-rel);
        m_ref += rel;
        l_sum *= alpha;
      }
      float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
      const float sh = SC ? (need ? rel : 0.f) : m_ref;
#pragma unroll
      for (int c = 0; c < HC; c += 4) {
        sv[c] = ex2_approx(sv[c] - sh); ps0 += sv[c];
        sv[c + 1] = ex2_approx(sv[c + 1] - sh); ps1 += sv[c + 1];
        sv[c + 2] = ex2_approx(sv[c + 2] - sh); ps2 += sv[c + 2];
        sv[c + 3] = ex2_approx(sv[c + 3] - sh); ps3 += sv[c + 3];
      }
      l_sum += (ps0 + ps1) + (ps2 + ps3);
      const int pb = j % PB;
      uint8_t* myP = sP + (t * PB + pb) * Cfg::P_TILE;
      if (j >= PB) mbar_wait(&pv_done[t * PB + pb], ((j / PB) - 1) & 1);      // P buffer free again (PV of tile j-PB retired)
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rare: rescale this thread's half of the O accumulator row in TMEM; every earlier PV must have retired
        mbar_wait(&pv_done[t * PB + (j - 1) % PB], ((j - 1) / PB) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
          uint32_t u[32];
          tmem_ld32(trow + Cfg::O_COL + t * 128 + h * (D / 2) + c * 32, u);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(__uint_as_float(u[i]) * alpha);
          tmem_st32(trow + Cfg::O_COL + t * 128 + h * (D / 2) + c * 32, u);
        }
        tmem_st_wait();
      }
#pragma unroll
      for (int c8 = 0; c8 < HC / 8; ++c8) {
        uint4 pk;
        pk.x = pack_bf16(sv[8 * c8], sv[8 * c8 + 1]); pk.y = pack_bf16(sv[8 * c8 + 2], sv[8 * c8 + 3]);
        pk.z = pack_bf16(sv[8 * c8 + 4], sv[8 * c8 + 5]); pk.w = pack_bf16(sv[8 * c8 + 6], sv[8 * c8 + 7]);
        *reinterpret_cast<uint4*>(myP + swz_off(r, h * (HC / 8) + c8)) = pk;
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&p_ready[t * PB + pb]);
    }
    // combine the two halves' row sums, then each thread normalises and stores its half of the output row
    xch[(2 * 2 + h) * 128 + r] = l_sum;                  // dedicated third slot
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
    l_sum += xch[(2 * 2 + (h ^ 1)) * 128 + r];
    mbar_wait(&pv_done[t * PB + (nt - 1) % PB], ((nt - 1) / PB) & 1);
    tc_fence_after();
    const float inv = 1.f / l_sum;
    float* op = a.out + ((size_t)pair * a.Lq + gq) * D + h * (D / 2);
#pragma unroll
    for (int c = 0; c < D / 64; ++c) {
      uint32_t u[32];
      tmem_ld32(trow + Cfg::O_COL + t * 128 + h * (D / 2) + c * 32, u);
      tmem_ld_wait();
      if (gq < a.Lq) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(op + c * 32 + 4 * i) =
              make_float4(__uint_as_float(u[4 * i]) * inv, __uint_as_float(u[4 * i + 1]) * inv,
                          __uint_as_float(u[4 * i + 2]) * inv, __uint_as_float(u[4 * i + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

template <int D, bool SC>
inline cudaError_t launch_attn(const AttnArgs& a, int pairs, cudaStream_t st) {
  using Cfg = AttnCfg<D, SC>;
  static bool configured = false;
  auto kern = attn_tc_kernel<D, SC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<dim3((a.q_tiles + 1) / 2, pairs), 576, Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
