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
// CTA = one 128-query tile of one pair, 6 warps:
//   warps 0-3  softmax: thread r owns score row r (TMEM lane r): tcgen05.ld -> (x c_ij) -> online softmax with lazy
//              rescale -> bf16 P written to shared memory in the UMMA image
//   warp 4     producer: one thread issues bulk-async copies (TMA engine) of K / V^T / key-point tiles, 2-stage ring
//   warp 5     MMA: one thread issues S = Q K^T (double-buffered in TMEM, issued one tile ahead) and O += P V
#pragma once
#include "common.cuh"

namespace gmf {

struct AttnArgs {
  const __nv_bfloat16* q_t;   // [pairs][q_tiles][128*D]
  const __nv_bfloat16* k_t;   // [pairs][k_tiles][128*D]
  const __nv_bfloat16* vt_t;  // [pairs][k_tiles][D*128]
  const float* kpts;          // SC only: [pairs][k_tiles*128][8] = (sx,sy,sz,|s|^2,tx,ty,tz,|t|^2), centred
  float* out;                 // [pairs][Lq][D] fp32
  int Lq, Lk, q_tiles, k_tiles;
  float neg_inv_sigma2;       // SC only: -1/sigma_d^2
};

template <int D, bool SC>
struct AttnCfg {
  static constexpr int Q_BYTES = 128 * D * 2;
  static constexpr int K_BYTES = 128 * D * 2;
  static constexpr int V_BYTES = D * 128 * 2;
  static constexpr int P_BYTES = 128 * 128 * 2;
  static constexpr int PTS_BYTES = SC ? 128 * 32 : 0;
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES + PTS_BYTES;
  static constexpr int SMEM = 1024 + Q_BYTES + 2 * STAGE_BYTES + P_BYTES + 256;
  static constexpr int TMEM_COLS = 512;   // S: 2 x 128, O: D
  static constexpr int O_COL = 256;
};

template <int D, bool SC>
__global__ void __launch_bounds__(192, 1) attn_tc_kernel(const AttnArgs a) {
  using Cfg = AttnCfg<D, SC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sStage = sQ + Cfg::Q_BYTES;                 // [2] x {K, V^T, pts}
  uint8_t* sP = sStage + 2 * Cfg::STAGE_BYTES;
  uint64_t* bars = (uint64_t*)(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_free = bars + 7;     // [2]
  uint64_t* p_ready = bars + 9;
  uint64_t* pv_done = bars + 10;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, pair = blockIdx.y;
  const int nt = a.k_tiles;

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128);
    }
    mbar_init(p_ready, 128); mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------ producer ------------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
      bulk_g2s(sQ, a.q_t + (size_t)(pair * a.q_tiles + qt) * (128 * D), Cfg::Q_BYTES, q_full);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(&kv_empty[s], ((j >> 1) - 1) & 1);
        uint8_t* st = sStage + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&kv_full[s], Cfg::STAGE_BYTES);
        const size_t tix = (size_t)pair * nt + j;
        bulk_g2s(st, a.k_t + tix * (128 * D), Cfg::K_BYTES, &kv_full[s]);
        bulk_g2s(st + Cfg::K_BYTES, a.vt_t + tix * (128 * D), Cfg::V_BYTES, &kv_full[s]);
        if (SC) bulk_g2s(st + Cfg::K_BYTES + Cfg::V_BYTES, a.kpts + tix * (128 * 8), Cfg::PTS_BYTES, &kv_full[s]);
      }
    }
  } else if (warp == 5) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 128, kFmtBF16);
      const uint32_t idesc_o = umma_idesc(128, D, kFmtBF16);
      const uint32_t q_base = smem_u32(sQ), p_base = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&s_free[s], ((j >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sStage + s * Cfg::STAGE_BYTES);
#pragma unroll
        for (int at = 0; at < D / 64; ++at)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(tmem + s * 128, umma_desc_sw128(q_base + at * 16384 + ks * 32),
                        umma_desc_sw128(k_base + at * 16384 + ks * 32), idesc_s, (at | ks) ? 1u : 0u);
        tc_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_s(j + 1);
        const int s = j & 1;
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(sStage + s * Cfg::STAGE_BYTES + Cfg::K_BYTES);
#pragma unroll
        for (int at = 0; at < 2; ++at)          // 128 keys = 2 atoms of 64
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(tmem + Cfg::O_COL, umma_desc_sw128(p_base + at * 16384 + ks * 32),
                        umma_desc_sw128(v_base + at * (D * 128) + ks * 32), idesc_o, (j > 0 || at > 0 || ks > 0) ? 1u : 0u);
        tc_commit(&kv_empty[s]);
        tc_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------ softmax (one thread per query row) ------------------------------------
    const int r = tid;                                   // row in tile == TMEM lane
    const int gq = qt * 128 + r;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    float qsx = 0.f, qsy = 0.f, qsz = 0.f, qsn = 0.f, qtx = 0.f, qty = 0.f, qtz = 0.f, qtn = 0.f;
    if (SC) {
      // query coordinates come from the same key-point array (queries == keys in self attention); rows beyond Lq are
      // zero-padded there
      const float4 s4 = *reinterpret_cast<const float4*>(a.kpts + ((size_t)pair * nt * 128 + gq) * 8);
      const float4 t4 = *reinterpret_cast<const float4*>(a.kpts + ((size_t)pair * nt * 128 + gq) * 8 + 4);
      qsx = -2.f * s4.x; qsy = -2.f * s4.y; qsz = -2.f * s4.z; qsn = s4.w;
      qtx = -2.f * t4.x; qty = -2.f * t4.y; qtz = -2.f * t4.z; qtn = t4.w;
    }
    float m_ref = 0.f, l_sum = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      float sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t t[32];
        tmem_ld32(trow + s * 128 + c * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sv[c * 32 + i] = __uint_as_float(t[i]);
      }
      tc_fence_before();
      mbar_arrive(&s_free[s]);
      const int nvalid = min(128, a.Lk - j * 128);
      if (SC) {
        mbar_wait(&kv_full[s], (j >> 1) & 1);            // acquire the TMA-written key points
        const float4* kp = reinterpret_cast<const float4*>(sStage + s * Cfg::STAGE_BYTES + Cfg::K_BYTES + Cfg::V_BYTES);
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const float4 ks = kp[2 * c], kt = kp[2 * c + 1];
          float d2s = fmaf(qsx, ks.x, fmaf(qsy, ks.y, fmaf(qsz, ks.z, qsn + ks.w)));
          float d2t = fmaf(qtx, kt.x, fmaf(qty, kt.y, fmaf(qtz, kt.z, qtn + kt.w)));
          d2s = fmaxf(d2s, 0.f); d2t = fmaxf(d2t, 0.f);
          // (|ds| - |dt|)^2 = ds^2 + dt^2 - 2 sqrt(ds^2 dt^2): one MUFU instead of two
          const float x = fmaf(-2.f, sqrt_approx(d2s * d2t), d2s + d2t);
          const float cij = __saturatef(fmaf(x, a.neg_inv_sigma2, 1.f));
          sv[c] *= cij;
        }
      }
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        if (c >= nvalid) sv[c] = -INFINITY;
        tmax = fmaxf(tmax, sv[c]);
      }
      // lazy rescale (scores are log2-scaled): keep the reference max until it is exceeded by > 8 (P <= 256)
      float alpha = 1.f;
      bool need = false;
      if (j == 0) {
        m_ref = tmax;
      } else if (tmax > m_ref + 8.f) {
        alpha = ex2_approx(m_ref - tmax);
        m_ref = tmax;
        l_sum *= alpha;
        need = true;
      }
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        sv[c] = ex2_approx(sv[c] - m_ref);
        psum += sv[c];
      }
      l_sum += psum;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);                 // P buffer and O accumulator are quiescent
        if (__any_sync(0xffffffffu, need)) {
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t t[32];
            tmem_ld32(trow + Cfg::O_COL + c * 32, t);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st32(trow + Cfg::O_COL + c * 32, t);
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int c8 = 0; c8 < 16; ++c8) {
        uint4 pk;
        pk.x = pack_bf16(sv[8 * c8], sv[8 * c8 + 1]); pk.y = pack_bf16(sv[8 * c8 + 2], sv[8 * c8 + 3]);
        pk.z = pack_bf16(sv[8 * c8 + 4], sv[8 * c8 + 5]); pk.w = pack_bf16(sv[8 * c8 + 6], sv[8 * c8 + 7]);
        *reinterpret_cast<uint4*>(sP + (c8 >> 3) * 16384 + swz_off(r, c8 & 7)) = pk;
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    mbar_wait(pv_done, (nt - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_sum;
    float* op = a.out + ((size_t)pair * a.Lq + gq) * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t t[32];
      tmem_ld32(trow + Cfg::O_COL + c * 32, t);
      tmem_ld_wait();
      if (gq < a.Lq) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(op + c * 32 + 4 * i) =
              make_float4(__uint_as_float(t[4 * i]) * inv, __uint_as_float(t[4 * i + 1]) * inv,
                          __uint_as_float(t[4 * i + 2]) * inv, __uint_as_float(t[4 * i + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, Cfg::TMEM_COLS);
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
  kern<<<dim3(a.q_tiles, pairs), 192, Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
