// Fused GEGLU feed-forward block of the fusion layers (fusion_layer.py:54-69 behind PreNorm :32-52, residual :191):
//     out = x + W2 . ( (W1v . LN(x) + b1v) * gelu(W1g . LN(x) + b1g) ) + b2          x: [L, 128], hidden 512, TF32 tensor pipe
// One kernel per 128-token tile; the 128 x 512 hidden activation never leaves the SM (the unfused pair of kernels wrote it to HBM
// as fp32 and read it back: 1.3 GB per layer at cfg#2, 8 of the 11 activation-sized transfers of the block).
//
// 8 passes of 64 hidden columns.  Pass p:  ACC1[p&1] = LN(x) . [W1v_p | W1g_p]^T   (16 SS-mode tf32 MMAs, N = 128, K = 128)
//                                          H[p&1]    = GEGLU(ACC1[p&1])            (epilogue warps, TMEM -> registers -> TMEM)
//                                          OUT      += H[p&1] . W2_p^T             (8 TS-mode tf32 MMAs: A operand = H in TMEM)
// TMEM (512 columns): ACC1 2 x 128 | H 2 x 64 | OUT 128.  Shared memory: LN(x) tile image 64 KB (reused as epilogue staging),
// W1 ring 3 x 32 KB, W2 ring 2 x 32 KB.  Warps: 0-15 workers (LayerNorm prologue, GEGLU, output epilogue), 16 MMA1 issuer, 17 MMA2
// issuer, 18 W1 producer, 19 W2 producer.  Issue loops are fully unrolled (all descriptors = uniform base + constant).
#pragma once
#include <cuda_fp16.h>
#include "linear_tc.cuh"

// GMF_FFN_F16 = 1 (default): the first GEMM runs in kind::f16 with fp16 operands (LN(x) and W1 rounded to half precision: the same 11-bit
// significand as TF32, |LN(x)| and |W1| are far inside the fp16 range) - W1 is the bulk of the weight stream (64 of 96 KB per pass), and the
// pass cadence was set by the 3 x 32 KB W1 ring holding only 1.5 passes; in fp16 a pass is ONE 32 KB chunk and the ring holds three.
#ifndef GMF_FFN_F16
#define GMF_FFN_F16 1
#endif

namespace gmf {

struct FfnCfg {
  static constexpr int A_BYTES = 128 * 128 * 4;            // LN(x) as 4 swizzle atoms of 128 rows x 32 floats
  static constexpr int W_BYTES = 128 * 64 * 4;             // one weight stage: 128 rows x 64 k
  static constexpr int N1 = 3, N2 = 2;                     // W1 / W2 ring depths (bulk copies have ~1500 clk latency: a single W2
                                                           // buffer serialised load -> MMA2 -> load and set the pass cadence)
  static constexpr int SMEM = 1024 + A_BYTES + (N1 + N2) * W_BYTES + 512;
  static constexpr int COL_H = 256, COL_OUT = 384;
  static constexpr int PASSES = 8;
};

struct FfnArgs {
  const float* x;          // [B, L, 128] block input (also the residual)
  int L, tiles;
  const float* ln_g;
  const float* ln_b;
  const float* w1_packed;  // tf32 build: 8 passes x 2 k-chunks x [128 rows (64 value | 64 gate) x 64 k] swizzled tf32;
                           // fp16 build: 8 passes x [128 rows x 128 k] swizzled fp16 (2 atoms of 64 k)
  const float* b1;         // [1024]: value 0..511, gate 512..1023
  const float* w2_packed;  // 8 chunks x [128 out rows x 64 hidden] swizzled tf32
  const float* b2;         // [128]
  float* out;              // [B, L, 128]
#ifdef GMF_FFN_TRACE
  long long* trace;
#endif
  // optional fused tail of NonLocalBlock.forward (PointDSC.py:65,73): out += fc_message.6(m2) = m2 . W3^T + b3
  const float* m2;         // [B, L, 64] (ReLU(BN(conv(...))) output of fc_message.4) or NULL
  const float* w3_packed;  // [128 out rows x 64] swizzled tf32
  const float* b3;         // [128]
  // out_img != NULL: instead of the row-major `out`, write the SPLIT fp16 K-major tile image [B][tiles][hi 2 x 16 KB | lo 2 x 16 KB] (x = hi + lo,
  // common.cuh split_f16; 64 KB like the fp32 tile) that the next layer's
  // chained PointCN/QKV kernel loads with one bulk copy (the per-warp staging tiles already are 32-row blocks of that image)
  float* out_img;
};

__global__ void __launch_bounds__(640, 1) ffn_fused_kernel(const FfnArgs a) {
  using Cfg = FfnCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW1 = sA + Cfg::A_BYTES;                       // [N1] stages
  uint8_t* sW2 = sW1 + Cfg::N1 * Cfg::W_BYTES;
  uint64_t* bars = (uint64_t*)(sW2 + Cfg::N2 * Cfg::W_BYTES);
  uint64_t* a_ready = bars;            // 256
  uint64_t* full1 = bars + 1;          // [3]
  uint64_t* empty1 = bars + 4;         // [3]
  uint64_t* full2 = bars + 7;          // [2]
  uint64_t* empty2 = bars + 19;        // [2]
  uint64_t* acc1_full = bars + 9;      // [2]
  uint64_t* acc1_free = bars + 11;     // [2] 256
  uint64_t* h_ready = bars + 13;       // [2] 256
  uint64_t* h_free = bars + 15;        // [2]
  uint64_t* out_full = bars + 17;      // 1
  uint64_t* m2_ready = bars + 18;      // 512
  uint32_t* tmem_slot = (uint32_t*)(bars + 21);
  float* sStg = (float*)sA;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, pair = blockIdx.y;
  const int row0 = tile * 128;
#ifdef GMF_FFN_TRACE
  const bool trc = a.trace && blockIdx.x == 3 && blockIdx.y == 1;
#define TR(role, idx) do { if (trc && lane == 0) a.trace[(role) * 64 + (idx)] = clock64(); } while (0)
#else
#define TR(role, idx) do {} while (0)
#endif
  TR(0, 0);

  if (tid == 0) {
    mbar_init(a_ready, 512);
    for (int i = 0; i < 3; ++i) { mbar_init(&full1[i], 1); mbar_init(&empty1[i], 1); }
    mbar_init(&full2[0], 1); mbar_init(&full2[1], 1); mbar_init(&empty2[0], 1); mbar_init(&empty2[1], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&acc1_free[i], 512); mbar_init(&h_ready[i], 512); mbar_init(&h_free[i], 1); }
    mbar_init(out_full, 1); mbar_init(m2_ready, 512);
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 18) {
    // ------------------------------- W1 producer: 16 stages (pass, k-half) through a 3-deep ring -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint8_t* src = (const uint8_t*)a.w1_packed;
#pragma unroll 1
    for (int s = 0; s < (GMF_FFN_F16 ? 1 : 2) * Cfg::PASSES; ++s) {
      const int slot = s % Cfg::N1;
      if (s >= Cfg::N1) mbar_wait(&empty1[slot], ((s / Cfg::N1) - 1) & 1);
      mbar_expect_tx_p(&full1[slot], Cfg::W_BYTES, leader);
      bulk_g2s_p(sW1 + slot * Cfg::W_BYTES, src + (size_t)s * Cfg::W_BYTES, Cfg::W_BYTES, &full1[slot], leader);
    }
  } else if (warp == 19) {
    // ------------------------------- W2 producer: one chunk per pass -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint8_t* src = (const uint8_t*)a.w2_packed;
#pragma unroll 1
    for (int p = 0; p < Cfg::PASSES; ++p) {
      const int slot = p & 1;
      if (p >= 2) mbar_wait(&empty2[slot], ((p >> 1) - 1) & 1);
      mbar_expect_tx_p(&full2[slot], Cfg::W_BYTES, leader);
      bulk_g2s_p(sW2 + slot * Cfg::W_BYTES, src + (size_t)p * Cfg::W_BYTES, Cfg::W_BYTES, &full2[slot], leader);
    }
    if (a.m2) {                                              // ninth chunk: fc_message.6 weight for the fused block tail
      mbar_wait(&empty2[0], 1);                              // MMA2 of pass 6 (4th use of slot 0) retired
      mbar_expect_tx_p(&full2[0], Cfg::W_BYTES, leader);
      bulk_g2s_p(sW2, a.w3_packed, Cfg::W_BYTES, &full2[0], leader);
    }
  } else if (warp == 16) {
    // ------------------------------- MMA1 issuer: ACC1[p&1] = LN(x) . W1_p^T -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, GMF_FFN_F16 ? kFmtF16 : kFmtTF32);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW1));
    mbar_wait(a_ready, 0);
#pragma unroll
    for (int p = 0; p < Cfg::PASSES; ++p) {
      TR(1, 4 * p);
      if (p >= 2) mbar_wait(&acc1_free[p & 1], ((p >> 1) - 1) & 1);
      TR(1, 4 * p + 1);
#if GMF_FFN_F16
      {
        const int slot = p % Cfg::N1;
        mbar_wait(&full1[slot], (p / Cfg::N1) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int at = 0; at < 2; ++at)                       // 64 k per swizzle atom, 16 k per MMA
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_bf16(tm + (p & 1) * 128, umma_desc_adv(a_desc0, at * 16384 + ks * 32), umma_desc_adv(w_desc0, slot * Cfg::W_BYTES + at * 16384 + ks * 32),
                          idesc, (at | ks) ? 1u : 0u);
          tc_commit(&empty1[slot]);
          tc_commit(&acc1_full[p & 1]);
        }
        __syncwarp();
        TR(1, 4 * p + 2);
      }
#else
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        const int s = 2 * p + kc, slot = s % Cfg::N1;
        mbar_wait(&full1[slot], (s / Cfg::N1) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int at = 0; at < 2; ++at)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_tf32(tm + (p & 1) * 128, umma_desc_adv(a_desc0, (2 * kc + at) * 16384 + ks * 32),
                          umma_desc_adv(w_desc0, slot * Cfg::W_BYTES + at * 16384 + ks * 32), idesc, (kc | at | ks) ? 1u : 0u);
          tc_commit(&empty1[slot]);
          if (kc == 1) tc_commit(&acc1_full[p & 1]);
        }
        __syncwarp();
        TR(1, 4 * p + 2 + kc);
      }
#endif
    }
  } else if (warp == 17) {
    // ------------------------------- MMA2 issuer: OUT += H[p&1] . W2_p^T (A operand from tensor memory) -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, kFmtTF32);
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW2));
#pragma unroll
    for (int p = 0; p < Cfg::PASSES; ++p) {
      TR(2, 2 * p);
      mbar_wait2(&h_ready[p & 1], (p >> 1) & 1, &full2[p & 1], (p >> 1) & 1);
      tc_fence_after();
      TR(2, 2 * p + 1);
      if (leader) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          tc_mma_tf32_ts(tm + Cfg::COL_OUT, tm + Cfg::COL_H + (p & 1) * 64 + i * 8, umma_desc_adv(w_desc0, (p & 1) * Cfg::W_BYTES + (i >> 2) * 16384 + (i & 3) * 32), idesc,
                         (p | i) ? 1u : 0u);
        tc_commit(&empty2[p & 1]);
        tc_commit(&h_free[p & 1]);
        if (p == Cfg::PASSES - 1 && !a.m2) tc_commit(out_full);
      }
      __syncwarp();
    }
    if (a.m2) {                                              // OUT += m2 . W3^T : A operand = m2 (tf32) parked in the idle H[0] columns
      mbar_wait2(m2_ready, 0, &full2[0], 0);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          tc_mma_tf32_ts(tm + Cfg::COL_OUT, tm + Cfg::COL_H + i * 8, umma_desc_adv(w_desc0, (i >> 2) * 16384 + (i & 3) * 32), idesc, 1u);
        tc_commit(out_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- workers: LayerNorm prologue -> A operand (tf32, swizzled) -------------------------------
    {
      const int c4 = lane * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(a.ln_g + c4);
      const float4 b4 = *reinterpret_cast<const float4*>(a.ln_b + c4);
      const float* xp = a.x + (size_t)pair * a.L * 128;
      constexpr int RPW = 8;
      const int rbase = warp * RPW;
      float4 rv[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int gr = row0 + rbase + i;
        rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < a.L) rv[i] = *reinterpret_cast<const float4*>(xp + (size_t)gr * 128 + c4);
      }
      // branch-free sweeps so that the shuffle-reduction chains of the 8 rows interleave (rows past L are zeros, never stored)
      float mean[RPW], rs[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) mean[i] = rv[i].x + rv[i].y + rv[i].z + rv[i].w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < RPW; ++i) mean[i] += __shfl_xor_sync(0xffffffffu, mean[i], o);
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        mean[i] *= (1.0f / 128.0f);
        const float dx = rv[i].x - mean[i], dy = rv[i].y - mean[i], dz = rv[i].z - mean[i], dw = rv[i].w - mean[i];
        rv[i] = make_float4(dx, dy, dz, dw);
        rs[i] = dx * dx + dy * dy + dz * dz + dw * dw;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < RPW; ++i) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const float r_ = rsqrtf(rs[i] * (1.0f / 128.0f) + 1e-5f);
        const float4 v = make_float4(fmaf(rv[i].x * r_, g4.x, b4.x), fmaf(rv[i].y * r_, g4.y, b4.y), fmaf(rv[i].z * r_, g4.z, b4.z), fmaf(rv[i].w * r_, g4.w, b4.w));
#if GMF_FFN_F16
        {
          const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&h01); pk.y = *reinterpret_cast<const uint32_t*>(&h23);
          // channels 4 lane .. 4 lane + 3 of row r: atom lane >> 4 (64 channels = 128 B per row), 16-byte piece (lane & 15) >> 1, half of it lane & 1
          *reinterpret_cast<uint2*>(sA + (lane >> 4) * 16384 + swz_off(rbase + i, (lane & 15) >> 1) + (lane & 1) * 8) = pk;
        }
#else
        *reinterpret_cast<float4*>(sA + (lane >> 3) * 16384 + swz_off(rbase + i, lane & 7)) = to_tf32(v);
#endif
      }
      fence_proxy_async();
      mbar_arrive(a_ready);
      if (warp == 0) TR(0, 1);
    }
    // ------------------------------- workers: GEGLU between the two GEMMs (16 warps: 4 lane quadrants x 4 column quarters) -----------
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int p = 0; p < Cfg::PASSES; ++p) {
      const int b = p & 1;
      const int hc0 = p * 64 + cq * 16;                        // hidden column of v[0]
      float4 b1v[4], b1g[4];                                   // bias loads in flight while waiting for the accumulator
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        b1v[i] = __ldg(reinterpret_cast<const float4*>(a.b1 + hc0) + i);
        b1g[i] = __ldg(reinterpret_cast<const float4*>(a.b1 + 512 + hc0) + i);
      }
      if (warp == 0) TR(3, 4 * p);
      mbar_wait(&acc1_full[b], (p >> 1) & 1);
      tc_fence_after();
      if (warp == 0) TR(3, 4 * p + 1);
      uint32_t v[16], g[16];
      tmem_ld16(trow + b * 128 + cq * 16, v);
      tmem_ld16(trow + b * 128 + 64 + cq * 16, g);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc1_free[b]);                              // MMA1 of pass p + 2 may overwrite the accumulator
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b1 = b1v[i], b2 = b1g[i];
        float o0, o1, o2, o3;
        unpack2(geglu2(pack2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), pack2(b1.x, b1.y),
                       pack2(__uint_as_float(g[4 * i]), __uint_as_float(g[4 * i + 1])), pack2(b2.x, b2.y)), o0, o1);
        unpack2(geglu2(pack2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), pack2(b1.z, b1.w),
                       pack2(__uint_as_float(g[4 * i + 2]), __uint_as_float(g[4 * i + 3])), pack2(b2.z, b2.w)), o2, o3);
        v[4 * i] = __float_as_uint(to_tf32(o0)); v[4 * i + 1] = __float_as_uint(to_tf32(o1));
        v[4 * i + 2] = __float_as_uint(to_tf32(o2)); v[4 * i + 3] = __float_as_uint(to_tf32(o3));
      }
      if (warp == 0) TR(3, 4 * p + 2);
      if (p >= 2) { mbar_wait(&h_free[b], ((p >> 1) - 1) & 1); tc_fence_after(); }   // MMA2 of pass p - 2 has read H[b]
      if (warp == 0) TR(3, 4 * p + 3);
      tmem_st16(trow + Cfg::COL_H + b * 64 + cq * 16, v);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&h_ready[b]);
    }
    if (a.m2) {
      // this thread's 16 columns of m2 row r -> H[0] (free once MMA2 of pass 6 has retired: 4th completion of h_free[0])
      uint32_t mv[16];
      const int gr = row0 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < a.L) x4 = __ldg(reinterpret_cast<const float4*>(a.m2 + ((size_t)pair * a.L + gr) * 64 + cq * 16) + i);
        x4 = to_tf32(x4);
        mv[4 * i] = __float_as_uint(x4.x); mv[4 * i + 1] = __float_as_uint(x4.y); mv[4 * i + 2] = __float_as_uint(x4.z); mv[4 * i + 3] = __float_as_uint(x4.w);
      }
      mbar_wait(&h_free[0], 1);
      tc_fence_after();
      tmem_st16(trow + Cfg::COL_H + cq * 16, mv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(m2_ready);
    }
    // ------------------------------- workers: out = OUT + b2 + x (coalesced through a per-warp staging tile) -------------------------------
    if (warp == 0) TR(0, 2);
    float* stg = sStg + warp * 1024;                           // the A image is dead: every MMA1 retired before the last GEGLU pass
    const int srow = lane >> 3, sj = lane & 7;
    {
      const int c = cq;                                        // one 32-column chunk per warp
      const int col0 = c * 32;
      const size_t gbase = ((size_t)pair * a.L + row0 + q * 32) * 128 + col0;
      // the residual rows are fetched into the staging tile while the last MMA2s are still running
      float4 rr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rw = i * 4 + srow;
        rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + q * 32 + rw < a.L) rr[i] = *reinterpret_cast<const float4*>(a.x + gbase + (size_t)rw * 128 + sj * 4);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rw = i * 4 + srow;
        *reinterpret_cast<float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2)) = rr[i];
      }
      __syncwarp();
      mbar_wait(out_full, 0);
      tc_fence_after();
      if (warp == 0) TR(0, 3);
      uint32_t v[32];
      tmem_ld32(trow + Cfg::COL_OUT + c * 32, v);
      tmem_ld_wait();
      const bool live = row0 + r < a.L;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 bb = *reinterpret_cast<const float4*>(a.b2 + col0 + 4 * j);
        if (a.m2) { const float4 b3 = *reinterpret_cast<const float4*>(a.b3 + col0 + 4 * j); bb.x += b3.x; bb.y += b3.y; bb.z += b3.z; bb.w += b3.w; }
        float4* slot = reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2));
        const float4 res = *slot;
        const float4 o = make_float4(__uint_as_float(v[4 * j]) + bb.x + res.x, __uint_as_float(v[4 * j + 1]) + bb.y + res.y,
                                     __uint_as_float(v[4 * j + 2]) + bb.z + res.z, __uint_as_float(v[4 * j + 3]) + bb.w + res.w);
        if (a.out_img) {                                       // kept in registers: the image is assembled below
          v[4 * j] = __float_as_uint(live ? o.x : 0.f); v[4 * j + 1] = __float_as_uint(live ? o.y : 0.f);
          v[4 * j + 2] = __float_as_uint(live ? o.z : 0.f); v[4 * j + 3] = __float_as_uint(live ? o.w : 0.f);
        } else {
          *slot = o;
        }
      }
      __syncwarp();
      if (a.out_img) {
        // Split fp16 image (x = hi + lo): fp16 swizzle atoms are 64 columns wide, so warps (q, 2m) and (q, 2m + 1) share the 32-row blocks of
        // atom m.  Their two 4 KB staging tiles become that block of the hi image (even warp's tile) and of the lo image (odd warp's tile);
        // each warp then ships the block that sits in its own tile.
        const int m = cq >> 1;
        const int pair_bar = 1 + m * 4 + q;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // both warps have read their residual rows out of the tiles
        uint8_t* blk_hi = (uint8_t*)(sStg + ((2 * m) * 4 + q) * 1024);
        uint8_t* blk_lo = (uint8_t*)(sStg + ((2 * m + 1) * 4 + q) * 1024);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 H, Lw;
          split_f16x2(__uint_as_float(v[8 * jj]), __uint_as_float(v[8 * jj + 1]), H.x, Lw.x);
          split_f16x2(__uint_as_float(v[8 * jj + 2]), __uint_as_float(v[8 * jj + 3]), H.y, Lw.y);
          split_f16x2(__uint_as_float(v[8 * jj + 4]), __uint_as_float(v[8 * jj + 5]), H.z, Lw.z);
          split_f16x2(__uint_as_float(v[8 * jj + 6]), __uint_as_float(v[8 * jj + 7]), H.w, Lw.w);
          const uint32_t off = swz_off(lane, (cq & 1) * 4 + jj);
          *reinterpret_cast<uint4*>(blk_hi + off) = H;
          *reinterpret_cast<uint4*>(blk_lo + off) = Lw;
        }
        fence_proxy_async();
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      }
      if (!a.out_img) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (row0 + q * 32 + rw < a.L)
            *reinterpret_cast<float4*>(a.out + gbase + (size_t)rw * 128 + sj * 4) = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
        }
        __syncwarp();
      }
    }
    if (a.out_img) {
      // 16 x 4 KB bulk stores in flight, no CTA-wide barrier: tile layout [hi atom 0 | hi atom 1 | lo atom 0 | lo atom 1], 16 KB each
      if (lane == 0) {
        uint8_t* img = (uint8_t*)(a.out_img + (size_t)(pair * a.tiles + tile) * (128 * 128));
        bulk_s2g(img + (cq & 1) * 32768 + (cq >> 1) * 16384 + q * 4096, stg, 4096);
        bulk_commit_wait_read();
      }
      __syncwarp();
    }
  }
  if (warp == 0) TR(0, 4);
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

inline cudaError_t launch_ffn_fused(const FfnArgs& a, int pairs, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = ensure_dyn_smem(ffn_fused_kernel, FfnCfg::SMEM, configured)) return e;
  ffn_fused_kernel<<<dim3(a.tiles, pairs), 640, FfnCfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
