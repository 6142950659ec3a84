// PointCN + Q/K/V projection of one encoder layer as ONE chained-GEMM kernel (PointDSC.py:104-111 then :56-58):
//     feat1 = ReLU(BN(conv128x128(feat)))          -> fp32 [L, 128] to HBM (Fusion-2 needs it) AND, rounded to tf32, kept in TMEM
//     Q | K | V = conv128x384(feat1)               -> bf16 tile images for the SC attention kernel (Q pre-scaled, V transposed)
// The second GEMM takes its A operand straight from tensor memory (TS-mode tf32 MMA on the accumulator columns of the first),
// so feat1 is neither re-read from HBM nor staged in shared memory; one launch and one exposed prologue instead of two.
// TMEM (512 columns): feat1 0..127 | Q 128..255 | K 256..383 | V 384..511.  Shared memory: feat tile image 64 KB, weight ring
// 4 x 32 KB (8 chunks: 2 PointCN + 6 QKV, the three bf16 tile images reuse the ring once the MMAs have retired), staging 32 KB.
#pragma once
#include "linear_tc.cuh"

namespace gmf {

struct PcnQkvCfg {
  static constexpr int A_BYTES = 128 * 128 * 4, W_BYTES = 128 * 64 * 4, NBUF = 4, NSTAGE = 8;
  static constexpr int STG_BYTES = 8 * 32 * 32 * 4;
  static constexpr int SMEM = 1024 + A_BYTES + NBUF * W_BYTES + 256 + STG_BYTES;
};

struct PcnQkvArgs {
  const float* x;          // [B, L, 128] layer input
  const float* x_img;      // or (preferred) its tf32 tile image [B][tiles][128 x 128] written by the previous layer's FFN kernel
  int L, tiles;
  const float* w_packed;   // 8 chunks of [128 rows x 64 k] tf32: PointCN (BN folded) k-halves, then q, k, v blocks x k-halves
  const float* pcn_bias;   // [128] (BN folded)
  const float* qkv_bias;   // [384] (q part pre-scaled)
  float* feat1;            // [B, L, 128]
  __nv_bfloat16 *tq, *tk, *tv;   // [B][tiles][128*128] tile images
};

__global__ void __launch_bounds__(288, 1) pcn_qkv_kernel(const PcnQkvArgs a) {
  using Cfg = PcnQkvCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::A_BYTES;
  uint64_t* bars = (uint64_t*)(sB + Cfg::NBUF * Cfg::W_BYTES);
  uint64_t* full = bars;           // [4]
  uint64_t* mma_done = bars + 4;   // [4]
  uint64_t* a_ready = bars + 8;    // 256
  uint64_t* acc0_full = bars + 9;
  uint64_t* f1_ready = bars + 10;  // 256
  uint64_t* acc_full = bars + 11;
  uint64_t* img_full = bars + 13;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);   // (bars + 13 is img_full)
  float* sStg = (float*)(sB + Cfg::NBUF * Cfg::W_BYTES + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, pair = blockIdx.y;
  const int row0 = tile * 128;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&mma_done[i], 1); }
    mbar_init(a_ready, 256); mbar_init(acc0_full, 1); mbar_init(f1_ready, 256); mbar_init(acc_full, 1); mbar_init(img_full, 1);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------- control warp: weight stream + both GEMMs (fully unrolled, uniform operands) -------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, kFmtTF32);
    const uint8_t* wsrc = (const uint8_t*)a.w_packed;
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB));
    auto issue_load = [&](int it) {
      const int buf = it % Cfg::NBUF;
      mbar_expect_tx_p(&full[buf], Cfg::W_BYTES, leader);
      bulk_g2s_p(sB + buf * Cfg::W_BYTES, wsrc + (size_t)it * Cfg::W_BYTES, Cfg::W_BYTES, &full[buf], leader);
    };
    if (a.x_img) {
      mbar_expect_tx_p(img_full, Cfg::A_BYTES, leader);
      bulk_g2s_p(sA, a.x_img + (size_t)(pair * a.tiles + tile) * (128 * 128), Cfg::A_BYTES, img_full, leader);
    }
#pragma unroll
    for (int it = 0; it < Cfg::NBUF - 1; ++it) issue_load(it);
#pragma unroll
    for (int it = 0; it < Cfg::NSTAGE; ++it) {
      const int buf = it % Cfg::NBUF, nxt = it + Cfg::NBUF - 1;
      if (nxt < Cfg::NSTAGE) {
        if (nxt >= Cfg::NBUF) mbar_wait(&mma_done[nxt % Cfg::NBUF], ((nxt / Cfg::NBUF) - 1) & 1);
        issue_load(nxt);
      }
      if (it == 0) { if (a.x_img) mbar_wait(img_full, 0); else mbar_wait(a_ready, 0); }
      if (it == 2) mbar_wait(f1_ready, 0);                     // feat1 (tf32) is back in TMEM columns 0..127
      mbar_wait(&full[buf], (it / Cfg::NBUF) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t bd = umma_desc_adv(b_desc0, buf * Cfg::W_BYTES);
        const int kc = it & 1;
        if (it < 2) {                                          // PointCN: A = feat tile image in shared memory
#pragma unroll
          for (int at = 0; at < 2; ++at)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_tf32(tm, umma_desc_adv(a_desc0, (2 * kc + at) * 16384 + ks * 32), umma_desc_adv(bd, at * 16384 + ks * 32), idesc, (kc | at | ks) ? 1u : 0u);
        } else {                                               // Q / K / V: A = feat1 in tensor memory
          const int nb = (it - 2) >> 1;
#pragma unroll
          for (int at = 0; at < 2; ++at)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_tf32_ts(tm + 128 + nb * 128, tm + kc * 64 + at * 32 + ks * 8, umma_desc_adv(bd, at * 16384 + ks * 32), idesc, (kc | at | ks) ? 1u : 0u);
        }
        tc_commit(&mma_done[buf]);
        if (it == 1) tc_commit(acc0_full);
        if (it == Cfg::NSTAGE - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- workers: A operand (tf32, swizzled) unless the producer kernel already wrote the image ------
    if (!a.x_img) {
      const int c4 = lane * 4;
      const float* xp = a.x + (size_t)pair * a.L * 128;
      constexpr int RPW = 16;
      const int rbase = warp * RPW;
      float4 rv[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int gr = row0 + rbase + i;
        rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < a.L) rv[i] = *reinterpret_cast<const float4*>(xp + (size_t)gr * 128 + c4);
      }
#pragma unroll
      for (int i = 0; i < RPW; ++i) *reinterpret_cast<float4*>(sA + (lane >> 3) * 16384 + swz_off(rbase + i, lane & 7)) = to_tf32(rv[i]);
      fence_proxy_async();
      mbar_arrive(a_ready);
    }
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const bool valid = row0 + r < a.L;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    // ------------------------------- epilogue 0: feat1 = ReLU(acc + b) -> HBM (coalesced) and back to TMEM as tf32 -------------------
    mbar_wait(acc0_full, 0);
    tc_fence_after();
    {
      float* stg = sStg + warp * 1024;
      const int srow = lane >> 3, sj = lane & 7;
#pragma unroll 1
      for (int c = half; c < 4; c += 2) {
        uint32_t v[32];
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        const int col0 = c * 32;
        const size_t gbase = ((size_t)pair * a.L + row0 + q * 32) * 128 + col0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(a.pcn_bias + col0 + 4 * j);
          const float4 o = make_float4(fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.f), fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f),
                                       fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f));
          *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = o;
          const float4 t4 = to_tf32(o);
          v[4 * j] = __float_as_uint(t4.x); v[4 * j + 1] = __float_as_uint(t4.y); v[4 * j + 2] = __float_as_uint(t4.z); v[4 * j + 3] = __float_as_uint(t4.w);
        }
        tmem_st32(trow + c * 32, v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (row0 + q * 32 + rw < a.L)
            *reinterpret_cast<float4*>(a.feat1 + gbase + (size_t)rw * 128 + sj * 4) = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
        }
        __syncwarp();
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(f1_ready);
    }
    // ------------------------------- epilogue 1: bf16 Q / K / V^T tile images (assembled in the dead weight ring) ------------------
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = half; c < 12; c += 2) {
      uint32_t v[32];
      tmem_ld32(trow + 128 + c * 32, v);
      tmem_ld_wait();
      const int col0 = c * 32, which = col0 >> 7, dcol0 = col0 & 127;
      float o[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.qkv_bias + col0) + i);
        o[4 * i] = valid ? __uint_as_float(v[4 * i]) + b4.x : 0.f;
        o[4 * i + 1] = valid ? __uint_as_float(v[4 * i + 1]) + b4.y : 0.f;
        o[4 * i + 2] = valid ? __uint_as_float(v[4 * i + 2]) + b4.z : 0.f;
        o[4 * i + 3] = valid ? __uint_as_float(v[4 * i + 3]) + b4.w : 0.f;
      }
      uint8_t* img = sB + which * 32768;
      if (which < 2) {
        uint8_t* dst = img + (dcol0 >> 6) * 16384;
        const int cc0 = (dcol0 & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          pk.x = pack_bf16(o[8 * j], o[8 * j + 1]); pk.y = pack_bf16(o[8 * j + 2], o[8 * j + 3]);
          pk.z = pack_bf16(o[8 * j + 4], o[8 * j + 5]); pk.w = pack_bf16(o[8 * j + 6], o[8 * j + 7]);
          *reinterpret_cast<uint4*>(dst + swz_off(r, cc0 + j)) = pk;
        }
      } else {
        uint8_t* dst = img + (r >> 6) * (128 * 128) + (r & 7) * 2;
        const int kchunk = (r & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 32; ++i) *reinterpret_cast<__nv_bfloat16*>(dst + swz_off(dcol0 + i, kchunk)) = __float2bfloat16_rn(o[i]);
      }
    }
    fence_proxy_async();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) {
      const size_t tix = (size_t)(pair * a.tiles + tile) * (128 * 128);
      bulk_s2g(a.tq + tix, sB, 32768);
      bulk_s2g(a.tk + tix, sB + 32768, 32768);
      bulk_s2g(a.tv + tix, sB + 65536, 32768);
      bulk_commit_wait_read();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

inline cudaError_t launch_pcn_qkv(const PcnQkvArgs& a, int pairs, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pcn_qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PcnQkvCfg::SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  pcn_qkv_kernel<<<dim3(a.tiles, pairs), 288, PcnQkvCfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
