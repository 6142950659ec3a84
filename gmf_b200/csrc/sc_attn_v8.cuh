// Spatial-consistency guided non-local attention, generation 8 (PointDSC.py:56-64, 216-221):
//     msg_i = softmax_j( c_ij * q_i.k_j / sqrt(128) ) v_j,    c_ij = max(0, 1 - (|s_i-s_j| - |t_i-t_j|)^2 / sigma_d^2)
// The N x N matrices never exist in HBM: per 128-query x 64-key tile the tensor pipe produces THREE fp32 accumulators in TMEM,
//     S   = Q K^T                       (log2 units: log2(e)/sqrt(128) folded into the Q projection)
//     DA  = |s_i - s_j|^2 / sigma^2     (K = 32 bf16 MMA over 3-term split coordinates, see dist_feature_kernel<1>)
//     DB  = 1 - |t_i - t_j|^2 / sigma^2 (the constant and the sign ride in the feature columns)
// and the softmax threads evaluate  c = sat(2 sqrt(DA (1 - DB)) + DB - DA),  p = exp2(S c - ref)  in 8 issue slots per element.
//
// Measured design rules on B200 (tools/ubench/sm_rates.cu): MUFU 16/clk/SM, SS-mode MMA is shared-memory-read bound
// (128 B/clk: M128 N64 K16 = 48 clk, not 32), tcgen05.ld ~110 clk latency but ~900 B/clk pipelined.  Hence:
//   * two softmax groups (4 warps each, ONE thread per score row) alternate key tiles, so each SM sub-partition always has one
//     warp in its MUFU/FMA phase while the other waits on TMEM / mbarriers (the gen-7 kernel ran all softmax warps in lock step
//     on the same tile: XU 44 % busy, issue 36 %);
//   * P is written back to TMEM (tcgen05.st, over the S columns it was computed from) and is the A operand of the PV MMA:
//     no P round trip through shared memory, 32 KB less shared-memory traffic per tile;
//   * the softmax reference is FIXED per pass (0 in the first pass), there is no per-tile maximum exchange or accumulator
//     rescale: floating point is scale invariant, so any reference within 2^+-80 of the true row maximum gives the same
//     result.  Each thread tracks its row maximum; if any row of the CTA leaves the window the CTA repeats the key loop once
//     with the exact row maxima as reference (block-wide vote, all roles take part);
//   * POLY of every 4 exponentials are evaluated on the FMA pipe (Cody-Waite + degree-3 minimax, rel. err 7.5e-5, well below
//     the bf16 rounding of P) to balance the MUFU and issue limits.
//
// CTA = one 128-query row tile of one pair, 10 warps: 0-3 softmax group 0 (even virtual tiles), 4-7 group 1 (odd), 8 producer
// (bulk-async copies into a {K, Bd} ring and a V^T ring), 9 MMA issuer.  TMEM: buffer g at 192 g: S | DA | DB (64 columns each),
// P_g aliases S_g columns 0..31, O at 384..511.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "sc_attn_tc.cuh"

namespace gmf {

struct Sc8Cfg {
  static constexpr int D = 128, BN = 64, NS = 4, NV = 4;
  static constexpr int Q_BYTES = 128 * D * 2, AQ_BYTES = 128 * 64 * 2;
  static constexpr int K_BYTES = BN * D * 2, V_BYTES = D * BN * 2, BD_BYTES = BN * 64 * 2;
  static constexpr int KSTAGE_BYTES = K_BYTES + BD_BYTES;
  static constexpr int XCH_BYTES = 2 * 2 * 128 * 4;       // [max | sum][group][row]
  static constexpr int SMEM = 1024 + Q_BYTES + AQ_BYTES + NS * KSTAGE_BYTES + NV * V_BYTES + XCH_BYTES + 256;
  static constexpr int COL_O = 384;
  static constexpr float WINDOW = 80.f;                   // |row max - reference| allowed before the CTA repeats the key loop
};

// 2^x on the FMA/ALU pipes for x <= ~100: round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax of 2^f, exponent add.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float fl = x + 12582912.f;                         // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (fl - 12582912.f);
  float p = fmaf(0.0551716648f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(fl) << 23));
}

template <int POLY>
__global__ void __launch_bounds__(320, 1) sc_attn_v8_kernel(const ScAttnArgs a) {
  using Cfg = Sc8Cfg;
  constexpr int D = Cfg::D, BN = Cfg::BN, NS = Cfg::NS, NV = Cfg::NV;
  constexpr int WP = 8, WM = 9;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sAq = sQ + Cfg::Q_BYTES;
  uint8_t* sK = sAq + Cfg::AQ_BYTES;                     // [NS] x {K, Bd}
  uint8_t* sV = sK + NS * Cfg::KSTAGE_BYTES;             // [NV] x V^T
  float* sX = (float*)(sV + NV * Cfg::V_BYTES);          // [2][2][128]
  uint64_t* bars = (uint64_t*)((uint8_t*)sX + Cfg::XCH_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;        // [NS]
  uint64_t* k_empty = k_full + NS;    // [NS]
  uint64_t* v_full = k_empty + NS;    // [NV]
  uint64_t* v_empty = v_full + NV;    // [NV]
  uint64_t* s_full = v_empty + NV;    // [2]
  uint64_t* p_ready = s_full + 2;     // [2]
  uint64_t* o_full = p_ready + 2;     // 1
  uint32_t* tmem_slot = (uint32_t*)(o_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pair = blockIdx.y, qt = blockIdx.x;
  const int nt = (a.N + BN - 1) / BN;

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < NV; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); }
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == WP) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // softmax-thread state (other roles carry it along unused)
  const int g = warp >> 2;                                  // softmax group
  const int r = (warp & 3) * 32 + lane;                     // score row == TMEM lane
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float ref = 0.f, l_sum = 0.f, rmax = -INFINITY;
  int vbase = 0, pass = 0;

  for (;; ++pass) {
    int bad = 0;
    if (warp == WP) {
      // ------------------------------------ producer ------------------------------------
      const uint32_t leader = elect_one() ? 1u : 0u;
      if (pass == 0) {
        const size_t tq = (size_t)pair * a.tiles + qt;
        mbar_expect_tx_p(q_full, Cfg::Q_BYTES + Cfg::AQ_BYTES, leader);
        bulk_g2s_p(sQ, a.q_t + tq * (128 * D), Cfg::Q_BYTES, q_full, leader);
        bulk_g2s_p(sAq, a.aq_t + tq * (128 * 64), Cfg::AQ_BYTES, q_full, leader);
      }
      // {K, Bd} of a tile is needed two tiles before its V^T: serve the older request first
      int vk = vbase, vv = vbase;
      const int vend = vbase + nt;
      while (vk < vend || vv < vend) {
        if (vk < vend && (vv >= vend || vk <= vv + 2)) {
          const int st = vk % NS, j = vk - vbase;
          if (vk >= NS) mbar_wait(&k_empty[st], ((vk / NS) - 1) & 1);
          uint8_t* dst = sK + st * Cfg::KSTAGE_BYTES;
          mbar_expect_tx_p(&k_full[st], Cfg::KSTAGE_BYTES, leader);
          const size_t tix = (size_t)pair * a.tiles + (j >> 1);
          const int h = j & 1;
          const uint8_t* ksrc = (const uint8_t*)(a.k_t + tix * (128 * D)) + h * 8192;
          bulk_g2s_p(dst, ksrc, 8192, &k_full[st], leader);
          bulk_g2s_p(dst + 8192, ksrc + 16384, 8192, &k_full[st], leader);
          bulk_g2s_p(dst + Cfg::K_BYTES, (const uint8_t*)(a.bd_t + tix * (128 * 64)) + h * Cfg::BD_BYTES, Cfg::BD_BYTES, &k_full[st], leader);
          ++vk;
        } else {
          const int sv_ = vv % NV, j = vv - vbase;
          if (vv >= NV) mbar_wait(&v_empty[sv_], ((vv / NV) - 1) & 1);
          mbar_expect_tx_p(&v_full[sv_], Cfg::V_BYTES, leader);
          const size_t tix = (size_t)pair * a.tiles + (j >> 1);
          bulk_g2s_p(sV + sv_ * Cfg::V_BYTES, (const uint8_t*)(a.vt_t + tix * (128 * D)) + (j & 1) * Cfg::V_BYTES, Cfg::V_BYTES, &v_full[sv_], leader);
          ++vv;
        }
      }
    } else if (warp == WM) {
      // ------------------------------------ MMA issuer ------------------------------------
      const uint32_t leader = elect_one() ? 1u : 0u;
      const uint32_t idesc_s = umma_idesc(128, BN, kFmtBF16);
      const uint32_t idesc_o = umma_idesc(128, D, kFmtBF16);
      const uint64_t q_desc = umma_desc_sw128(smem_u32(sQ));
      const uint64_t aq_desc = umma_desc_sw128(smem_u32(sAq));
      const uint64_t k_desc0 = umma_desc_sw128(smem_u32(sK));
      const uint64_t v_desc0 = umma_desc_sw128(smem_u32(sV));
      auto issue_sd = [&](int v) {
        const int st = v % NS;
        const uint32_t col = tmem + (uint32_t)(v & 1) * 192u;
        mbar_wait(&k_full[st], (v / NS) & 1);
        tc_fence_after();
        const uint64_t kd = umma_desc_adv(k_desc0, st * Cfg::KSTAGE_BYTES);
        const uint64_t bd = umma_desc_adv(kd, Cfg::K_BYTES);
#pragma unroll
        for (int at = 0; at < 2; ++at)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16_p(col, umma_desc_adv(q_desc, at * 16384 + ks * 32), umma_desc_adv(kd, at * 8192 + ks * 32), idesc_s, (at | ks) ? 1u : 0u, leader);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          tc_mma_bf16_p(col + 64, umma_desc_adv(aq_desc, ks * 32), umma_desc_adv(bd, ks * 32), idesc_s, ks ? 1u : 0u, leader);
#pragma unroll
        for (int ks = 2; ks < 4; ++ks)
          tc_mma_bf16_p(col + 128, umma_desc_adv(aq_desc, ks * 32), umma_desc_adv(bd, ks * 32), idesc_s, ks > 2 ? 1u : 0u, leader);
        tc_commit_p(&s_full[v & 1], leader);
        tc_commit_p(&k_empty[st], leader);
      };
      if (pass == 0) mbar_wait(q_full, 0);
      for (int jj = 0; jj < 2 && jj < nt; ++jj) issue_sd(vbase + jj);
      for (int j = 0; j < nt; ++j) {
        const int v = vbase + j, sv_ = v % NV;
        mbar_wait(&p_ready[v & 1], (v >> 1) & 1);
        mbar_wait(&v_full[sv_], (v / NV) & 1);
        tc_fence_after();
        const uint32_t pcol = tmem + (uint32_t)(v & 1) * 192u;
        const uint64_t vd = umma_desc_adv(v_desc0, sv_ * Cfg::V_BYTES);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          tc_mma_bf16_ts_p(tmem + Cfg::COL_O, pcol + ks * 8, umma_desc_adv(vd, ks * 32), idesc_o, (j > 0 || ks > 0) ? 1u : 0u, leader);
        tc_commit_p(&v_empty[sv_], leader);
        if (j == nt - 1) tc_commit_p(o_full, leader);
        // the tensor pipe executes in issue order: S/DA/DB of tile v+2 overwrite buffer v&1 (and P_v in it) only after PV_v has read it
        if (j + 2 < nt) issue_sd(v + 2);
      }
    } else {
      // ------------------------------------ softmax group g: virtual tiles v with (v & 1) == g ------------------------------------
      const uint32_t tbuf = tlane + (uint32_t)g * 192u;
      float ps0 = 0.f, ps1 = 0.f;
      for (int v = vbase + ((vbase ^ g) & 1); v < vbase + nt; v += 2) {
        const int j = v - vbase;
        mbar_wait(&s_full[g], (v >> 1) & 1);
        tc_fence_after();
        const int nvalid = a.N - j * BN;
        auto tile_body = [&](auto ragged_tag) {
          constexpr bool RAGGED = decltype(ragged_tag)::value;
          uint32_t us[2][16], ua[2][16], ub[2][16];
          tmem_ld16(tbuf, us[0]); tmem_ld16(tbuf + 64, ua[0]); tmem_ld16(tbuf + 128, ub[0]);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int cur = ch & 1;
            tmem_ld_wait();
            if (ch < 3) {
              tmem_ld16(tbuf + (ch + 1) * 16, us[cur ^ 1]); tmem_ld16(tbuf + 64 + (ch + 1) * 16, ua[cur ^ 1]);
              tmem_ld16(tbuf + 128 + (ch + 1) * 16, ub[cur ^ 1]);
            }
            uint32_t pk[8];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
              float t0, t1;
              {
                const float da = __uint_as_float(ua[cur][c]), db = __uint_as_float(ub[cur][c]);
                const float rt = sqrt_approx(fabsf(fmaf(-da, db, da)));
                t0 = fmaf(__uint_as_float(us[cur][c]), __saturatef(fmaf(rt, 2.f, db - da)), -ref);
              }
              {
                const float da = __uint_as_float(ua[cur][c + 1]), db = __uint_as_float(ub[cur][c + 1]);
                const float rt = sqrt_approx(fabsf(fmaf(-da, db, da)));
                t1 = fmaf(__uint_as_float(us[cur][c + 1]), __saturatef(fmaf(rt, 2.f, db - da)), -ref);
              }
              if (RAGGED) {
                if (ch * 16 + c >= nvalid) t0 = -INFINITY;
                if (ch * 16 + c + 1 >= nvalid) t1 = -INFINITY;
              }
              rmax = fmaxf(rmax, fmaxf(t0, t1));
              const float p0 = (!RAGGED && (c & 3) < POLY) ? ex2_poly(t0) : ex2_approx(t0);
              const float p1 = (!RAGGED && ((c + 1) & 3) < POLY) ? ex2_poly(t1) : ex2_approx(t1);
              ps0 += p0; ps1 += p1;
              pk[c >> 1] = pack_bf16(p0, p1);
            }
            tmem_st8(tbuf + ch * 8, pk);                     // P over S columns 0..31 (already consumed)
          }
        };
        if (nvalid >= BN) tile_body(std::false_type{});
        else tile_body(std::true_type{});                    // ragged last tile (CTA-uniform)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[g]);
      }
      l_sum += ps0 + ps1;
      // combined row maximum of the two groups decides whether the fixed reference was good enough
      sX[g * 128 + r] = rmax;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float comb = fmaxf(rmax, sX[(g ^ 1) * 128 + r]);
      if (pass == 0 && !(comb >= -Cfg::WINDOW && comb <= Cfg::WINDOW)) { bad = 1; ref = comb; }
    }
    const int redo = __syncthreads_or(bad);
    if (!redo || pass == 1) break;
    if (bad == 0 && warp < 8) ref = fmaxf(rmax, sX[((warp >> 2) ^ 1) * 128 + r]);   // rows that were fine also move to their exact maximum
    vbase += nt; l_sum = 0.f; rmax = -INFINITY;
    __syncthreads();                                         // sX is rewritten by the next pass
  }

  if (warp < 8) {
    sX[256 + g * 128 + r] = l_sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.f / (l_sum + sX[256 + (g ^ 1) * 128 + r]);
    mbar_wait(o_full, pass & 1);
    tc_fence_after();
    const int gq = qt * 128 + r;
    float* op = a.out + ((size_t)pair * a.N + gq) * D + g * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t u[32];
      tmem_ld32(tlane + Cfg::COL_O + g * 64 + c * 32, u);
      tmem_ld_wait();
      if (gq < a.N) {
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
  if (warp == WP) tmem_dealloc(tmem, 512);
}

template <int POLY>
inline cudaError_t launch_sc_attn_v8(const ScAttnArgs& a, int pairs, cudaStream_t st) {
  static bool configured = false;
  auto kern = sc_attn_v8_kernel<POLY>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Sc8Cfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  kern<<<dim3(a.tiles, pairs), 320, Sc8Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

// distance-feature tiles for the gen-8 kernel: coordinates are pre-divided by sigma_d, the t-part carries the flipped sign and the
// constant so that the accumulators are  DA = |ds|^2 / sigma^2  and  DB = 1 - |dt|^2 / sigma^2.
//   s-part A: per coord (-2u0,-2u0,-2u1,-2u1,-2u0,-2u2), |u|^2 split (n0,n1,n2), (1,1,1)       B: (w0,w1,w0,w1,w2,w0), (1,1,1), |w|^2 split
//   t-part A: per coord (+2u0,...),                       -|u|^2 split,          (-1,-1,-1), 1  B: same as s-part,                         , 1
__global__ void dist_feature_scaled_kernel(const float* __restrict__ kpts, int Np, float inv_sigma, __nv_bfloat16* __restrict__ aq_t,
                                           __nv_bfloat16* __restrict__ bd_t) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // padded point index
  if (i >= Np) return;
  const float4 s4 = *reinterpret_cast<const float4*>(kpts + ((size_t)pair * Np + i) * 8);
  const float4 t4 = *reinterpret_cast<const float4*>(kpts + ((size_t)pair * Np + i) * 8 + 4);
  __align__(16) __nv_bfloat16 A[64];
  __align__(16) __nv_bfloat16 B[64];
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f), one = __float2bfloat16_rn(1.f), mone = __float2bfloat16_rn(-1.f);
#pragma unroll
  for (int k = 0; k < 64; ++k) { A[k] = zero; B[k] = zero; }
  float pts[2][4] = {{s4.x * inv_sigma, s4.y * inv_sigma, s4.z * inv_sigma, 0.f}, {t4.x * inv_sigma, t4.y * inv_sigma, t4.z * inv_sigma, 0.f}};
#pragma unroll
  for (int part = 0; part < 2; ++part) pts[part][3] = fmaf(pts[part][0], pts[part][0], fmaf(pts[part][1], pts[part][1], pts[part][2] * pts[part][2]));
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const int o = part * 32;
    const float sg = part ? 2.f : -2.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      __nv_bfloat16 h0, h1, h2;
      split3(pts[part][c], h0, h1, h2);
      const __nv_bfloat16 m0 = __float2bfloat16_rn(sg * __bfloat162float(h0)), m1 = __float2bfloat16_rn(sg * __bfloat162float(h1)),
                          m2 = __float2bfloat16_rn(sg * __bfloat162float(h2));
      A[o + 6 * c + 0] = m0; B[o + 6 * c + 0] = h0;
      A[o + 6 * c + 1] = m0; B[o + 6 * c + 1] = h1;
      A[o + 6 * c + 2] = m1; B[o + 6 * c + 2] = h0;
      A[o + 6 * c + 3] = m1; B[o + 6 * c + 3] = h1;
      A[o + 6 * c + 4] = m0; B[o + 6 * c + 4] = h2;
      A[o + 6 * c + 5] = m2; B[o + 6 * c + 5] = h0;
    }
    __nv_bfloat16 n0, n1, n2;
    split3(part ? -pts[part][3] : pts[part][3], n0, n1, n2);           // A side: +-|u|^2
    __nv_bfloat16 w0, w1, w2;
    split3(pts[part][3], w0, w1, w2);                                   // B side: |w|^2, multiplied by +-1 from the A side
    const __nv_bfloat16 sgn1 = part ? mone : one;
    A[o + 18] = n0; A[o + 19] = n1; A[o + 20] = n2; B[o + 18] = one; B[o + 19] = one; B[o + 20] = one;
    A[o + 21] = sgn1; A[o + 22] = sgn1; A[o + 23] = sgn1; B[o + 21] = w0; B[o + 22] = w1; B[o + 23] = w2;
    if (part) { A[o + 24] = one; B[o + 24] = one; }
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
