// Fusion cross-attention (Fusion-1 / Fusion-2, fusion_layer.py:82-94), head dim 64:  out = softmax(q k^T / 8) v, generation 2.
// Same recipe as the gen-9 SC kernel (sc_attn_v9.cuh; measurements in profiles/r01_sc_attention.md):
//   * every MMA takes its A operand from TENSOR MEMORY: Q (written once per CTA) for the scores, bf16 P for P.V — shared memory
//     only streams K and V^T, so an M128 N64 K16 step costs 32 clk instead of the shared-memory-bound 48;
//   * fixed softmax reference per pass (0 first) + block-wide "repeat once with the exact row maxima" vote instead of a per-tile
//     maximum exchange / accumulator rescale;
//   * a score buffer is handed back to the issuer as soon as its tile is in registers (scores run two key tiles ahead);
//   * MMA issue loops unrolled over the ring period with warp-uniform operands, issued from an elected lane.
// CTA = TWO 128-query row tiles of one pair sharing the K / V^T stream (halves the L2 -> SM traffic).  The row tiles are
// independent pipelines: 8 softmax warps (two threads per score row, 32 columns each), one P.V issuer warp and one score issuer
// warp per row tile; one producer warp.  21 warps.
// TMEM (512 columns): S[t][b] at 64 (2t + b) | O[t] at 256 + 64 t | P[t] at 384 + 32 t | Q[t] at 448 + 32 t.
#pragma once
#include "attn_args.cuh"

namespace gmf {

struct Fa2Cfg {
  static constexpr int D = 64, BN = 64, NR = 4;                 // NR: K ring depth = V^T ring depth = unroll period
  static constexpr int K_BYTES = BN * D * 2, V_BYTES = D * BN * 2;
  static constexpr int XCH_BYTES = 2 * 2 * 2 * 128 * 4;         // [max | sum][row tile][half][row]
  static constexpr int WO_BYTES = 128 * 64 * 4;                 // to_out weight (tf32) for the fused output projection
  static constexpr int STG_BYTES = 16 * 32 * 32 * 4;            // per-warp 32 x 32 transpose tiles for coalesced epilogue I/O
  static constexpr int WQ_BYTES = 64 * 128 * 2;                 // to_q weight (fp16) for the fused query projection
  static constexpr int SMEM = 1024 + NR * K_BYTES + NR * V_BYTES + WO_BYTES + WQ_BYTES + STG_BYTES + XCH_BYTES + 512;
  static constexpr int COL_O = 256, COL_P = 384, COL_Q = 448;
  static constexpr float WINDOW = 80.f;
};

struct Fa2Mma {
  uint32_t tmem, idesc, idesc_s, leader;
  uint64_t k_desc0, v_desc0;
  BarArr k_full, k_empty, v_full, v_empty, s_full, s_free, p_ready, pv_done, o_full;   // mbarriers by 32-bit shared address (common.cuh)
  int nt, t;                                                   // key tiles, this issuer's row tile
};

// S[t][T&1] = Q_t K_j^T for the key tile at ring position T
template <int T>
__device__ __forceinline__ void fa2_issue_s(const Fa2Mma& m, uint32_t ring_parity) {
  using Cfg = Fa2Cfg;
  mbar_wait(&m.k_full[T], ring_parity);
  tc_fence_after();
  if (m.leader) {
    const uint32_t col = m.tmem + (uint32_t)(2 * m.t + (T & 1)) * 64u;
    const uint32_t qcol = m.tmem + Cfg::COL_Q + (uint32_t)m.t * 32u;
    const uint64_t kd = umma_desc_adv(m.k_desc0, T * Cfg::K_BYTES);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) tc_mma_bf16_ts(col, qcol + ks * 8, umma_desc_adv(kd, ks * 32), m.idesc_s, ks ? 1u : 0u);
    tc_commit(&m.s_full[2 * m.t + (T & 1)]);
    tc_commit(&m.k_empty[T]);
  }
  __syncwarp();
}

template <int T>
__device__ __forceinline__ void fa2_s_step(const Fa2Mma& m, int j0, uint32_t ph) {
  const int j = j0 + T;
  if (j + 2 >= m.nt) return;
  mbar_wait(&m.s_free[2 * m.t + (T & 1)], (T >> 1) & 1);       // softmax pulled tile j out of buffer T&1
  fa2_issue_s<(T + 2) % 4>(m, T + 2 < 4 ? ph : ph ^ 1u);
}

template <int T>
__device__ __forceinline__ void fa2_pv_step(const Fa2Mma& m, int j0, uint32_t ph) {
  using Cfg = Fa2Cfg;
  const int j = j0 + T;
  if (j >= m.nt) return;
  mbar_wait2(&m.p_ready[m.t], T & 1, &m.v_full[T], ph);
  tc_fence_after();
  if (m.leader) {
    const uint32_t pcol = m.tmem + Cfg::COL_P + (uint32_t)m.t * 32u;
    const uint64_t vd = umma_desc_adv(m.v_desc0, T * Cfg::V_BYTES);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      tc_mma_bf16_ts(m.tmem + Cfg::COL_O + (uint32_t)m.t * 64u, pcol + ks * 8, umma_desc_adv(vd, ks * 32), m.idesc, (j > 0 || ks > 0) ? 1u : 0u);
    tc_commit(&m.pv_done[m.t]);
    tc_commit(&m.v_empty[T]);
    if (j == m.nt - 1) tc_commit(&m.o_full[m.t]);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(672, 1) fus_attn_v2_kernel(const AttnArgs a) {
  using Cfg = Fa2Cfg;
  constexpr int D = Cfg::D, BN = Cfg::BN, NR = Cfg::NR;
  constexpr int WP = 16;                                        // producer; 17,18: P.V issuers; 19,20: score issuers
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                                    // [NR] x K (64 keys x 64)
  uint8_t* sV = sK + NR * Cfg::K_BYTES;                  // [NR] x V^T (64 x 64 keys)
  uint8_t* sWo = sV + NR * Cfg::V_BYTES;                 // to_out weight image (fused output projection)
  uint8_t* sWq = sWo + Cfg::WO_BYTES;                    // to_q weight image (fused query projection)
  float* sStg = (float*)(sWq + Cfg::WQ_BYTES);           // [16 warps][32][32]; before the main loop: LN(x0) operand images of the two row tiles
  float* sX = (float*)((uint8_t*)sStg + Cfg::STG_BYTES); // [2][2][2][128]
  uint64_t* bars = (uint64_t*)((uint8_t*)sX + Cfg::XCH_BYTES);
  const BarArr q_full{smem_u32(bars)};   // [2]  every barrier below is "this address + constant" (no shared-window re-derivation per operation)
  const BarArr k_full = q_full + 2;      // [NR]
  const BarArr k_empty = k_full + NR;    // [NR]
  const BarArr v_full = k_empty + NR;    // [NR]
  const BarArr v_empty = v_full + NR;    // [NR]
  const BarArr s_full = v_empty + NR;    // [2][2]
  const BarArr s_free = s_full + 4;      // [2][2]
  const BarArr p_ready = s_free + 4;     // [2]
  const BarArr pv_done = p_ready + 2;    // [2]
  const BarArr o_full = pv_done + 2;     // [2]
  const BarArr on_ready = o_full + 2;    // [2] normalised, tf32-rounded O written back to TMEM
  const BarArr x_full = on_ready + 2;    // [2] O . Wo^T complete
  const BarArr wo_full = x_full + 2;     // 1
  const BarArr wq_full = wo_full + 1;    // 1
  const BarArr qa_ready = wq_full + 1;   // [2] LN(x0) operand image of a row tile written (256 threads)
  const BarArr qacc_full = qa_ready + 2; // [2] Q = LN(x0) . Wq^T in the (still idle) first score buffer of the row tile
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 + 4 * NR + 4 + 4 + 2 + 2 + 2 + 2 + 2 + 1 + 1 + 2 + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.y;
  const int qt0 = blockIdx.x * 2;
#ifdef GMF_FFN_TRACE
#define FTR(i) do { if (a.trace && blockIdx.x == 5 && blockIdx.y == 33 && tid == 0) a.trace[i] = clock64(); } while (0)
#else
#define FTR(i) do { } while (0)
#endif
  FTR(0);
  const int ntile = (qt0 + 1 < a.q_tiles) ? 2 : 1;                  // active row tiles
  const int nt = (a.Lk + BN - 1) / BN;

  auto init_bars = [&]() {
    for (int i = 0; i < NR; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], ntile); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], ntile); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 256); }
    for (int i = 0; i < 2; ++i) { mbar_init(&p_ready[i], 256); mbar_init(&pv_done[i], 1); mbar_init(&o_full[i], 1); }
  };
  if (tid == 0) {
    mbar_init(&q_full[0], 256); mbar_init(&q_full[1], 256);
    mbar_init(&on_ready[0], 256); mbar_init(&on_ready[1], 256); mbar_init(&x_full[0], 1); mbar_init(&x_full[1], 1); mbar_init(wo_full, 1);
    mbar_init(wq_full, 1); mbar_init(&qa_ready[0], 256); mbar_init(&qa_ready[1], 256); mbar_init(&qacc_full[0], 1); mbar_init(&qacc_full[1], 1);
    init_bars();
    fence_mbar_init();
  }
  if (warp == WP) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int t = (warp >> 3) & 1;                            // softmax warps: row tile
  const int h = (warp >> 2) & 1;                            // column half of a key tile
  const int r = (warp & 3) * 32 + lane;                     // row in tile == TMEM lane
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const bool softmax_role = warp < 16 && t < ntile;
  float ref = 0.f, l_sum = 0.f, rmax = -INFINITY;
  int pass = 0;

  FTR(1);
  if (softmax_role && a.xq) {
    // ---------------- fused query side: x0 = x + dwconv(x), LN_q(x0) -> fp16 operand image, Q = LN(x0) . Wq^T -> fp16 -> tensor memory ----------------
    // The 8 warps of a row tile own 16 rows each (two batches of 8, lanes across the 128 channels).
    const int c4 = lane * 4;
    const int row_lo = (qt0 + t) * 128 + (warp & 7) * 16;        // first row of this warp
    const float* xp = a.xq + (size_t)pair * a.Lq * 128;
    for (int i = lane >> 2; i < 18; i += 8) {                    // pull the 18 rows (16 + halo) into L2 ahead of the two batches
      const int gr = row_lo - 1 + i;
      if (gr >= 0 && gr < a.Lq) asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + (size_t)gr * 128 + (lane & 3) * 32));
    }
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.lnq_g + c4)), b4 = __ldg(reinterpret_cast<const float4*>(a.lnq_b + c4));
    const uint64_t g01 = pack2(g4.x, g4.y), g23 = pack2(g4.z, g4.w), b01 = pack2(b4.x, b4.y), b23 = pack2(b4.z, b4.w);
    uint64_t w0a = pack2(0.f, 0.f), w0b = w0a, w2a = w0a, w2b = w0a, cba = w0a, cbb = w0a, w1a = pack2(1.f, 1.f), w1b = w1a;
    if (a.cpe_w) {                                               // taps of channel c at cpe_w[3 c ..]: previous, current, next row
      const float* cw = a.cpe_w + c4 * 3;
      const float4 cw0 = __ldg(reinterpret_cast<const float4*>(cw)), cw1 = __ldg(reinterpret_cast<const float4*>(cw + 4)), cw2 = __ldg(reinterpret_cast<const float4*>(cw + 8));
      const float4 cb = __ldg(reinterpret_cast<const float4*>(a.cpe_b + c4));
      w0a = pack2(cw0.x, cw0.w); w0b = pack2(cw1.z, cw2.y);
      w1a = pack2(1.0f + cw0.y, 1.0f + cw1.x); w1b = pack2(1.0f + cw1.w, 1.0f + cw2.z);
      w2a = pack2(cw0.z, cw1.y); w2b = pack2(cw2.x, cw2.w);
      cba = pack2(cb.x, cb.y); cbb = pack2(cb.z, cb.w);
    }
    uint8_t* img = (uint8_t*)sStg + t * 32768 + (lane >> 4) * 16384 + (lane & 1) * 8;
    const uint32_t piece = (lane & 15) >> 1;
#pragma unroll 1
    for (int bt = 0; bt < 2; ++bt) {
      const int rb = (warp & 7) * 16 + bt * 8;                   // tile row of rv[1]
      const int g0 = (qt0 + t) * 128 + rb - 1;                   // global row of rv[0]
      float4 rv[10];
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g0 + i >= 0 && g0 + i < a.Lq) rv[i] = *reinterpret_cast<const float4*>(xp + (size_t)(g0 + i) * 128 + c4);
      }
      uint64_t d01[8], d23[8];
      float mean[8], rs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        d01[i] = ffma2(w0a, pack2(rv[i].x, rv[i].y), ffma2(w1a, pack2(rv[i + 1].x, rv[i + 1].y), ffma2(w2a, pack2(rv[i + 2].x, rv[i + 2].y), cba)));
        d23[i] = ffma2(w0b, pack2(rv[i].z, rv[i].w), ffma2(w1b, pack2(rv[i + 1].z, rv[i + 1].w), ffma2(w2b, pack2(rv[i + 2].z, rv[i + 2].w), cbb)));
        float s0, s1;
        unpack2(fadd2(d01[i], d23[i]), s0, s1);
        mean[i] = s0 + s1;
        if (a.x0 && g0 + 1 + i < a.Lq) {                           // residual stream of the output epilogue: stays in L2 until this CTA reads it back
          float4 o;
          unpack2(d01[i], o.x, o.y);
          unpack2(d23[i], o.z, o.w);
          *reinterpret_cast<float4*>(a.x0 + ((size_t)pair * a.Lq + g0 + 1 + i) * 128 + c4) = o;
        }
      }
      warp_sum8(mean, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float nm = mean[i] * (-1.0f / 128.0f);
        const uint64_t nm2 = pack2(nm, nm);
        d01[i] = fadd2(d01[i], nm2);
        d23[i] = fadd2(d23[i], nm2);
        float s0, s1;
        unpack2(ffma2(d23[i], d23[i], fmul2(d01[i], d01[i])), s0, s1);
        rs[i] = s0 + s1;
      }
      warp_sum8(rs, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float r_ = rsqrtf(rs[i] * (1.0f / 128.0f) + 1e-5f);
        const uint64_t r2 = pack2(r_, r_);
        float y0, y1, y2, y3;
        unpack2(ffma2(fmul2(d01[i], r2), g01, b01), y0, y1);
        unpack2(ffma2(fmul2(d23[i], r2), g23, b23), y2, y3);
        uint2 pk;
        pk.x = pack_f16(y0, y1);
        pk.y = pack_f16(y2, y3);
        *reinterpret_cast<uint2*>(img + (rb + i) * 128 + ((piece ^ (uint32_t)i) << 4)) = pk;      // (row & 7) == i
      }
    }
    FTR(2);
    fence_proxy_async();
    mbar_arrive(&qa_ready[t]);
    // this thread's half of Q row r: accumulator (score buffer 0 of the row tile) -> fp16 pairs -> Q columns
    mbar_wait(&qacc_full[t], 0);
    tc_fence_after();
    uint32_t u[32], w[16];
    tmem_ld32(tlane + (uint32_t)(2 * t) * 64u + h * 32, u);
    tmem_ld_wait();
    const bool live = (qt0 + t) * 128 + r < a.Lq;
#pragma unroll
    for (int c = 0; c < 16; ++c) w[c] = live ? pack_f16(__uint_as_float(u[2 * c]), __uint_as_float(u[2 * c + 1])) : 0u;
    tmem_st16(tlane + Cfg::COL_Q + t * 32 + h * 16, w);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&q_full[t]);
    FTR(3);
  } else if (softmax_role) {
    // this thread's half of Q row r -> tensor memory (two bf16 per 32-bit column)
    const uint8_t* qsrc = (const uint8_t*)(a.q_t + (size_t)(pair * a.q_tiles + qt0 + t) * (128 * D));
    uint32_t w[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(qsrc + swz_off(r, h * 4 + c)));
      w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
    }
    tmem_st16(tlane + Cfg::COL_Q + t * 32 + h * 16, w);
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&q_full[t]);
  }

  for (;; ++pass) {
    int bad = 0;
    if (warp == WP) {
      // ------------------------------------ producer ------------------------------------
      const uint32_t leader = elect_one() ? 1u : 0u;
      if (pass == 0 && a.xq) {                                 // needed first: the query projection precedes every score MMA
        mbar_expect_tx_p(wq_full, Cfg::WQ_BYTES, leader);
        bulk_g2s_p(sWq, a.wq16, Cfg::WQ_BYTES, wq_full, leader);
      }
      if (pass == 0 && a.wo_packed) {
        mbar_expect_tx_p(wo_full, Cfg::WO_BYTES, leader);
        bulk_g2s_p(sWo, a.wo_packed, Cfg::WO_BYTES, wo_full, leader);
      }
      for (int j = 0; j < nt; ++j) {
        const int st = j % NR;
        const size_t tix = (size_t)pair * a.k_tiles + (j >> 1);
        const int hh = j & 1;
        if (j >= NR) mbar_wait(&k_empty[st], ((j / NR) - 1) & 1);
        mbar_expect_tx_p(&k_full[st], Cfg::K_BYTES, leader);
        bulk_g2s_p(sK + st * Cfg::K_BYTES, (const uint8_t*)(a.k_t + tix * (128 * D)) + hh * Cfg::K_BYTES, Cfg::K_BYTES, &k_full[st], leader);
        // V^T of tile j - 1 (needed one softmax period after its K): the producer never blocks on V before the K that is due first
        if (j >= 1) {
          const int jv = j - 1, sv = jv % NR;
          const size_t tv = (size_t)pair * a.k_tiles + (jv >> 1);
          if (jv >= NR) mbar_wait(&v_empty[sv], ((jv / NR) - 1) & 1);
          mbar_expect_tx_p(&v_full[sv], Cfg::V_BYTES, leader);
          bulk_g2s_p(sV + sv * Cfg::V_BYTES, (const uint8_t*)(a.vt_t + tv * (128 * D)) + (jv & 1) * Cfg::V_BYTES, Cfg::V_BYTES, &v_full[sv], leader);
        }
      }
      {
        const int jv = nt - 1, sv = jv % NR;
        const size_t tv = (size_t)pair * a.k_tiles + (jv >> 1);
        if (jv >= NR) mbar_wait(&v_empty[sv], ((jv / NR) - 1) & 1);
        mbar_expect_tx_p(&v_full[sv], Cfg::V_BYTES, leader);
        bulk_g2s_p(sV + sv * Cfg::V_BYTES, (const uint8_t*)(a.vt_t + tv * (128 * D)) + (jv & 1) * Cfg::V_BYTES, Cfg::V_BYTES, &v_full[sv], leader);
      }
    } else if (warp > WP) {
      // ------------------------------------ MMA issuers ------------------------------------
      Fa2Mma m;
      m.t = (warp - 17) & 1;
      const bool is_pv = warp < 19;
      m.tmem = __shfl_sync(0xffffffffu, tmem, 0); m.idesc = umma_idesc(128, BN, kFmtBF16); m.idesc_s = umma_idesc(128, BN, kFmtF16);   // S = Q K^T on fp16 operands, P V on bf16
      m.leader = elect_one() ? 1u : 0u;
      m.k_desc0 = umma_desc_sw128(smem_u32(sK)); m.v_desc0 = umma_desc_sw128(smem_u32(sV));
      m.k_full = k_full; m.k_empty = k_empty; m.v_full = v_full; m.v_empty = v_empty; m.s_full = s_full; m.s_free = s_free;
      m.p_ready = p_ready; m.pv_done = pv_done; m.o_full = o_full; m.nt = nt;
      if (m.t < ntile) {
        if (!is_pv) {
          if (pass == 0 && a.xq) {
            // Q_t = LN(x0_t) . Wq^T: A = the row tile's fp16 operand image (staging area), B = Wq, D = score buffer 0 of the row tile
            mbar_wait2(&qa_ready[m.t], 0, wq_full, 0);
            tc_fence_after();
            if (m.leader) {
              const uint64_t ad = umma_desc_sw128(smem_u32(sStg) + (uint32_t)m.t * 32768u), wd = umma_desc_sw128(smem_u32(sWq));
#pragma unroll
              for (int i = 0; i < 8; ++i)
                tc_mma_bf16(m.tmem + (uint32_t)(2 * m.t) * 64u, umma_desc_adv(ad, (i >> 2) * 16384 + (i & 3) * 32), umma_desc_adv(wd, (i >> 2) * 8192 + (i & 3) * 32),
                            m.idesc_s, i ? 1u : 0u);
              tc_commit(&qacc_full[m.t]);
            }
            __syncwarp();
          }
          if (pass == 0) { mbar_wait(&q_full[m.t], 0); tc_fence_after(); }
          fa2_issue_s<0>(m, 0u);
          if (nt > 1) fa2_issue_s<1>(m, 0u);
#pragma unroll 1
          for (int j0 = 0; j0 + 2 < nt; j0 += 4) {
            const uint32_t ph = (uint32_t)(j0 / 4) & 1u;
            fa2_s_step<0>(m, j0, ph); fa2_s_step<1>(m, j0, ph); fa2_s_step<2>(m, j0, ph); fa2_s_step<3>(m, j0, ph);
          }
        } else {
#pragma unroll 1
          for (int j0 = 0; j0 < nt; j0 += 4) {
            const uint32_t ph = (uint32_t)(j0 / 4) & 1u;
            fa2_pv_step<0>(m, j0, ph); fa2_pv_step<1>(m, j0, ph); fa2_pv_step<2>(m, j0, ph); fa2_pv_step<3>(m, j0, ph);
          }
          mbar_wait(&o_full[m.t], 0);                           // every MMA and commit of this row tile has retired before the vote
        }
      }
    } else if (softmax_role) {
      // ------------------------------------ softmax: row tile t, column half h ------------------------------------
      float ps0 = 0.f, ps1 = 0.f;
      uint64_t psA = pack2(0.f, 0.f), psB = psA;
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1;
        mbar_wait(&s_full[2 * t + b], (j >> 1) & 1);
        tc_fence_after();
        uint32_t us[32], pk[16];
        tmem_ld32(tlane + (uint32_t)(2 * t + b) * 64u + h * 32, us);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[2 * t + b]);                     // the score issuer may refill this buffer with tile j + 2
        const int nvalid = a.Lk - j * BN - h * 32;
        if (nvalid >= 32) {
          // packed fp32x2 adds (FADD2) for the reference shift and the row sums: 3 issue slots per element besides the MUFU
          const uint64_t nref2 = pack2(-ref, -ref);
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            float t0, t1, t2, t3;
            unpack2(fadd2(pack2(__uint_as_float(us[c]), __uint_as_float(us[c + 1])), nref2), t0, t1);
            unpack2(fadd2(pack2(__uint_as_float(us[c + 2]), __uint_as_float(us[c + 3])), nref2), t2, t3);
            rmax = fmaxf(rmax, fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)));
            const float p0 = ex2_approx(t0), p1 = ex2_approx(t1), p2 = ex2_approx(t2), p3 = ex2_approx(t3);
            psA = fadd2(psA, pack2(p0, p1)); psB = fadd2(psB, pack2(p2, p3));
            pk[c >> 1] = pack_bf16(p0, p1); pk[(c >> 1) + 1] = pack_bf16(p2, p3);
          }
        } else {                                             // ragged last key tile (warp-uniform)
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float t0 = c < nvalid ? __uint_as_float(us[c]) - ref : -INFINITY;
            const float t1 = c + 1 < nvalid ? __uint_as_float(us[c + 1]) - ref : -INFINITY;
            rmax = fmaxf(rmax, fmaxf(t0, t1));
            const float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
            ps0 += p0; ps1 += p1;
            pk[c >> 1] = pack_bf16(p0, p1);
          }
        }
        if (j >= 1) { mbar_wait(&pv_done[t], (j - 1) & 1); tc_fence_after(); }     // P.V of the previous tile has read the P columns
        tmem_st16(tlane + Cfg::COL_P + t * 32 + h * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[t]);
      }
      { float a0, a1, b0, b1; unpack2(psA, a0, a1); unpack2(psB, b0, b1); l_sum += (ps0 + ps1) + ((a0 + a1) + (b0 + b1)); }
      sX[(t * 2 + h) * 128 + r] = rmax;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
      const float comb = fmaxf(rmax, sX[(t * 2 + (h ^ 1)) * 128 + r]);
      if (pass == 0 && !(comb >= -Cfg::WINDOW && comb <= Cfg::WINDOW)) bad = 1;
      rmax = comb;
    }
    const int redo = __syncthreads_or(bad);
    if (!redo || pass == 1) break;
    ref = rmax;                                              // exact row maximum (softmax threads; unused elsewhere)
    l_sum = 0.f; rmax = -INFINITY;
    if (tid == 0) { init_bars(); fence_mbar_init(); }        // every async arrival of the pass has landed (issuers waited on o_full)
    __syncthreads();
  }

  FTR(4);
  const bool fused_out = a.wo_packed != nullptr;
  if (fused_out && (warp == 17 || warp == 18) && (warp - 17) < ntile) {
    // ------------------------------------ P.V issuer of row tile tt: X = (O / l) . Wo^T  (tf32, A operand = O in tensor memory) -----------
    const int tt = warp - 17;
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc_x = umma_idesc(128, 128, kFmtTF32);
    const uint64_t wo_desc = umma_desc_sw128(smem_u32(sWo));
    mbar_wait2(&on_ready[tt], 0, wo_full, 0);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        tc_mma_tf32_ts(tm + (uint32_t)tt * 128u, tm + Cfg::COL_O + (uint32_t)tt * 64u + i * 8, umma_desc_adv(wo_desc, (i >> 2) * 16384 + (i & 3) * 32), idesc_x,
                       i ? 1u : 0u);
      tc_commit(&x_full[tt]);
    }
    __syncwarp();
  }
  if (softmax_role) {
    sX[512 + (t * 2 + h) * 128 + r] = l_sum;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
    const float inv = 1.f / (l_sum + sX[512 + (t * 2 + (h ^ 1)) * 128 + r]);
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const int gq = (qt0 + t) * 128 + r;
    uint32_t u[32];
    tmem_ld32(tlane + Cfg::COL_O + t * 64 + h * 32, u);
    tmem_ld_wait();
    if (!fused_out) {
      float* op = a.out + ((size_t)pair * a.Lq + gq) * D + h * 32;
      if (gq < a.Lq) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(op + 4 * i) =
              make_float4(__uint_as_float(u[4 * i]) * inv, __uint_as_float(u[4 * i + 1]) * inv,
                          __uint_as_float(u[4 * i + 2]) * inv, __uint_as_float(u[4 * i + 3]) * inv);
      }
    } else {
      // normalised attention output, rounded to tf32, back into its TMEM columns: it is the A operand of the output projection
#pragma unroll
      for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(to_tf32(__uint_as_float(u[i]) * inv));
      tmem_st32(tlane + Cfg::COL_O + t * 64 + h * 32, u);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&on_ready[t]);
      // coalesced I/O through a per-warp XOR-swizzled 32 x 32 staging tile (global accesses touch 4 rows x 128 B per instruction)
      float* stg = sStg + warp * 1024;
      const int srow = lane >> 3, sj = lane & 7;
      const int grow0 = (qt0 + t) * 128 + (warp & 3) * 32;             // first row of this warp's lane quadrant
      // the residual rows of both 32-column chunks are fetched while the output projection runs (u[] is dead until X is read back)
      float4 rr[2][8];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          rr[c][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow0 + rw < a.Lq)
            rr[c][i] = *reinterpret_cast<const float4*>(a.resid + ((size_t)pair * a.Lq + grow0 + rw) * 128 + h * 64 + c * 32 + sj * 4);
        }
      mbar_wait(&x_full[t], 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(tlane + (uint32_t)t * 128u + h * 64 + c * 32, u);      // X lives in the (now idle) score buffers of row tile t
        tmem_ld_wait();
        const int col0 = h * 64 + c * 32;
        const size_t gbase = ((size_t)pair * a.Lq + grow0) * 128 + col0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          *reinterpret_cast<float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2)) = rr[c][i];
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bo + col0) + j);
          float4* slot = reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2));
          const float4 res = *slot;
          *slot = make_float4(__uint_as_float(u[4 * j]) + bb.x + res.x, __uint_as_float(u[4 * j + 1]) + bb.y + res.y,
                              __uint_as_float(u[4 * j + 2]) + bb.z + res.z, __uint_as_float(u[4 * j + 3]) + bb.w + res.w);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (grow0 + rw < a.Lq)
            *reinterpret_cast<float4*>(a.xout + gbase + (size_t)rw * 128 + sj * 4) = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
        }
        __syncwarp();
      }
    }
  }
  FTR(5);
  tc_fence_before();
  __syncthreads();
  if (warp == WP) tmem_dealloc(tmem, 512);
}

inline cudaError_t launch_fus_attn_v2(const AttnArgs& a, int pairs, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = ensure_dyn_smem(fus_attn_v2_kernel, Fa2Cfg::SMEM, configured)) return e;
  fus_attn_v2_kernel<<<dim3((a.q_tiles + 1) / 2, pairs), 672, Fa2Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
