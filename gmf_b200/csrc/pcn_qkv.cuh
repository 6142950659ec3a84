// PointCN + Q/K/V projection of one encoder layer as ONE chained-GEMM kernel (PointDSC.py:104-111 then :56-58):
//     feat1 = ReLU(BN(conv128x128(feat)))          -> fp32 [L, 128] to HBM (Fusion-2 needs it) AND, split into fp16 hi + lo, kept in TMEM
//     Q | K | V = conv128x384(feat1)               -> fp16 Q / K and bf16 V^T tile images for the SC attention kernel (Q pre-scaled)
// The second GEMM takes its A operand straight from tensor memory (TS-mode MMA on the accumulator columns of the first),
// so feat1 is neither re-read from HBM nor staged in shared memory; one launch and one exposed prologue instead of two.
// Arithmetic: error-compensated fp16 (kind::f16): x = x_hi + x_lo, w = w_hi + w_lo, x w ~ x_hi w_hi + x_lo w_hi + x_hi w_lo, fp32
// accumulate - 22 significand bits per operand for 1.5x the tensor time of single TF32.  Plain TF32 (11 bits) on these two layers alone
// costs 1.3e-2 on the final logits at the KITTI shape (60 m coordinates through random-init weights; tools/probe_precision.py), above
// the 1e-2 the path is held to; the kernel is HBM / latency bound, so the extra MMAs are hidden.
// TMEM (512 columns): feat1 0..127 | Q 128..255 | K 256..383 | V 384..511.  Shared memory: feat tile image 64 KB, weight ring
// 4 x 32 KB (8 chunks: 2 PointCN + 6 QKV).  The feat tile area doubles as fp32 staging for feat1 and then holds the Q image; the K and V
// images are assembled in ring buffers whose MMAs have retired, so each image leaves as a bulk store while the next block is computed.
#pragma once
#include "linear_tc.cuh"

namespace gmf {

struct PcnQkvCfg {
  static constexpr int A_BYTES = 128 * 128 * 4, W_BYTES = 128 * 64 * 4, NBUF = 4, NSTAGE = 8;
  static constexpr int BIAS_BYTES = 512 * 4;
  static constexpr int SMEM = 1024 + A_BYTES + NBUF * W_BYTES + 256 + BIAS_BYTES;
};

struct PcnQkvArgs {
  const float* x;          // [B, L, 128] layer input
  const float* x_img;      // or (preferred) its split fp16 tile image [B][tiles][hi 32 KB | lo 32 KB] written by the previous layer's FFN kernel
  int L, tiles;
  const float* w_packed;   // 8 chunks of [128 rows x 64 k] as {fp16 hi atom 16 KB | fp16 lo atom 16 KB}: PointCN (BN folded) k-halves, then q, k, v blocks x k-halves
  const float* pcn_bias;   // [128] (BN folded)
  const float* qkv_bias;   // [384] (q part pre-scaled)
  float* feat1;            // [B, L, 128]
  __nv_bfloat16 *tq, *tk, *tv;   // [B][tiles][128*128] tile images
#ifdef GMF_FFN_TRACE
  long long* trace;        // development build: clock64 stamps of CTA 3's third tile (tools/pcn_trace.py)
#endif
};

// 17 warps: 16 workers (TMEM lane quadrant = warp & 3, 32-column chunk = warp >> 2) + one control warp
__global__ void __launch_bounds__(544, 1) pcn_qkv_kernel(const PcnQkvArgs a) {
  using Cfg = PcnQkvCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::A_BYTES;
  uint64_t* bars = (uint64_t*)(sB + Cfg::NBUF * Cfg::W_BYTES);
  uint64_t* full = bars;           // [4]
  uint64_t* mma_done = bars + 4;   // [4]
  uint64_t* a_ready = bars + 8;    // 512
  uint64_t* acc0_full = bars + 9;
  uint64_t* f1_ready = bars + 10;  // 512
  uint64_t* acc_full = bars + 11;
  uint64_t* img_full = bars + 13;
  uint64_t* accq_full = bars + 14;  // Q block of the second GEMM complete (stage 3), K block (stage 5): their epilogues overlap the rest
  uint64_t* acck_full = bars + 15;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);   // (bars + 13 is img_full)
  float* sBias = (float*)(sB + Cfg::NBUF * Cfg::W_BYTES + 256);   // PointCN bias [128] | Q,K,V biases [384]: read once, not per chunk from L2

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, pair = blockIdx.y;
  const int row0 = tile * 128;
  if (tid < 512) sBias[tid] = tid < 128 ? a.pcn_bias[tid] : a.qkv_bias[tid - 128];
#ifdef GMF_PCN_TRACE
  __shared__ long long tr_[16];
  const bool tr_on = blockIdx.x == 7 && (blockIdx.y == 3 || blockIdx.y == 40);
#define PTR(i) do { if (tr_on) tr_[i] = clock64(); } while (0)
  if (tid == 0) PTR(0);
#else
#define PTR(i) do { } while (0)
#endif

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&mma_done[i], 1); }
    mbar_init(a_ready, 512); mbar_init(acc0_full, 1); mbar_init(f1_ready, 512); mbar_init(acc_full, 1); mbar_init(img_full, 1); mbar_init(accq_full, 1); mbar_init(acck_full, 1);
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 16) {
    // ------------------------------- control warp: weight stream + both GEMMs (fully unrolled, uniform operands) -------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, kFmtF16);
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
      const int buf = it % Cfg::NBUF;
      // refill schedule: chunk 3 right away; chunks 4 and 5 as soon as the PointCN MMAs have retired (i.e. BEFORE waiting for feat1 to
      // come back into tensor memory - a bulk copy takes ~1.7k cycles to land and the Q/K/V MMAs of a chunk only ~0.65k), then 6 and 7
      // behind the MMAs of chunks 2 and 3
      auto refill = [&](int nxt) {
        if (nxt >= Cfg::NBUF) mbar_wait(&mma_done[nxt % Cfg::NBUF], ((nxt / Cfg::NBUF) - 1) & 1);
        issue_load(nxt);
      };
      if (it == 0) refill(3);
      if (it == 2) { refill(4); refill(5); }
      if (it == 3) refill(6);
      if (it == 4) refill(7);
      if (it == 0) { if (leader) PTR(1); if (a.x_img) mbar_wait(img_full, 0); else mbar_wait(a_ready, 0); if (leader) PTR(2); }
      if (it == 2) { mbar_wait(f1_ready, 0); if (leader) PTR(5); }                     // feat1 (tf32) is back in TMEM columns 0..127
      mbar_wait(&full[buf], (it / Cfg::NBUF) & 1);
      if (leader && it == 0) PTR(3);
      if (leader && it == 7) PTR(6);
      tc_fence_after();
      if (leader) {
        const uint64_t bd = umma_desc_adv(b_desc0, buf * Cfg::W_BYTES);
        const int kc = it & 1;
        // chunk = k-half kc: {w_hi atom | w_lo atom}; three products per K16 step: hi hi, lo hi, hi lo
        if (it < 2) {                                          // PointCN: A = split feat tile image in shared memory (hi atoms 0,1 | lo atoms 0,1)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ah = umma_desc_adv(a_desc0, kc * 16384 + ks * 32), al = umma_desc_adv(a_desc0, 32768 + kc * 16384 + ks * 32);
            const uint64_t bh = umma_desc_adv(bd, ks * 32), bl = umma_desc_adv(bd, 16384 + ks * 32);
            tc_mma_bf16(tm, ah, bh, idesc, (kc | ks) ? 1u : 0u);
            tc_mma_bf16(tm, al, bh, idesc, 1u);
            tc_mma_bf16(tm, ah, bl, idesc, 1u);
          }
        } else {                                               // Q / K / V: A = feat1 in tensor memory (hi: columns 0..63, lo: 64..127, two halfs per column)
          const int nb = (it - 2) >> 1;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ah = tm + kc * 32 + ks * 8, al = tm + 64 + kc * 32 + ks * 8;
            const uint64_t bh = umma_desc_adv(bd, ks * 32), bl = umma_desc_adv(bd, 16384 + ks * 32);
            tc_mma_bf16_ts(tm + 128 + nb * 128, ah, bh, idesc, (kc | ks) ? 1u : 0u);
            tc_mma_bf16_ts(tm + 128 + nb * 128, al, bh, idesc, 1u);
            tc_mma_bf16_ts(tm + 128 + nb * 128, ah, bl, idesc, 1u);
          }
        }
        tc_commit(&mma_done[buf]);
        if (it == 1) tc_commit(acc0_full);
        if (it == 3) tc_commit(accq_full);
        if (it == 5) tc_commit(acck_full);
        if (it == Cfg::NSTAGE - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- workers: A operand (tf32, swizzled) unless the producer kernel already wrote the image ------
    if (!a.x_img) {
      const int c4 = lane * 4;
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
      // columns 4 lane .. 4 lane + 3 -> fp16 atom lane / 16, 16-byte chunk (lane & 15) / 2, 8-byte half (lane & 1)
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        uint2 H, Lw;
        split_f16x2(rv[i].x, rv[i].y, H.x, Lw.x);
        split_f16x2(rv[i].z, rv[i].w, H.y, Lw.y);
        const uint32_t off = (lane >> 4) * 16384 + swz_off(rbase + i, (lane & 15) >> 1) + (lane & 1) * 8;
        *reinterpret_cast<uint2*>(sA + off) = H;
        *reinterpret_cast<uint2*>(sA + 32768 + off) = Lw;
      }
      fence_proxy_async();
      mbar_arrive(a_ready);
    }
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;
    const bool valid = row0 + r < a.L;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    // ------------------------------- epilogue 0: feat1 = ReLU(acc + b) -> HBM (coalesced) and back to TMEM as fp16 hi | lo --------------
    mbar_wait(acc0_full, 0);
    if (tid == 0) PTR(4);
    tc_fence_after();
    {
      // the feat tile image is dead once acc0_full fired: its 64 KB hold both chunks of every warp, so the TMEM round trip (which the
      // second GEMM waits for) finishes before any global store is issued; the stores then overlap the Q MMAs
      const int srow = lane >> 3, sj = lane & 7;
      {
        const int c = part;
        float* stg = (float*)sA + warp * 1024;
        uint32_t v[32], hw[16], lw[16];
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        const int col0 = c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(sBias + col0 + 4 * j);
          const float4 o = make_float4(fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.f), fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f),
                                       fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f));
          *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = o;
          // columns 32 c + 4 j .. + 3 -> packed pairs 2 j, 2 j + 1 of this warp's 16 hi words and 16 lo words
          split_f16x2(o.x, o.y, hw[2 * j], lw[2 * j]);
          split_f16x2(o.z, o.w, hw[2 * j + 1], lw[2 * j + 1]);
        }
        // the packed operand (hi: columns 0..63, lo: 64..127) overwrites accumulator columns other warps may still be reading
        asm volatile("bar.sync 1, 512;" ::: "memory");
        tmem_st16(trow + c * 16, hw);
        tmem_st16(trow + 64 + c * 16, lw);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(f1_ready);
      __syncwarp();
      {
        const int col0 = part * 32;
        const float* stg = (const float*)sA + warp * 1024;
        const size_t gbase = ((size_t)pair * a.L + row0 + q * 32) * 128 + col0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (row0 + q * 32 + rw < a.L)
            *reinterpret_cast<float4*>(a.feat1 + gbase + (size_t)rw * 128 + sj * 4) = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");            // every warp has read its staging tile: the area is reused for the Q image
    }
    // ------------------------------- epilogue 1: bf16 Q / K / V^T tile images (assembled in the dead weight ring) ------------------
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {                  // 0 = Q, 1 = K (row-major tiles), 2 = V (transposed tile)
      mbar_wait(which == 0 ? accq_full : which == 1 ? acck_full : acc_full, 0);
      if (tid == 0 && which == 2) PTR(7);
      tc_fence_after();
      // Q is assembled in the (dead) feat tile / staging area; K and V in ring buffers 0 and 1, whose last MMAs (stages 4, 5) have retired
      uint8_t* img = which == 0 ? sA : sB + (which - 1) * 32768;
      {
        const int cw = part;                                     // 32-column chunk within the block
        uint32_t v[32];
        tmem_ld32(trow + 128 + which * 128 + cw * 32, v);
        tmem_ld_wait();
        const int col0 = which * 128 + cw * 32, dcol0 = cw * 32;
        float o[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + 128 + col0 + 4 * i);
          o[4 * i] = valid ? __uint_as_float(v[4 * i]) + b4.x : 0.f;
          o[4 * i + 1] = valid ? __uint_as_float(v[4 * i + 1]) + b4.y : 0.f;
          o[4 * i + 2] = valid ? __uint_as_float(v[4 * i + 2]) + b4.z : 0.f;
          o[4 * i + 3] = valid ? __uint_as_float(v[4 * i + 3]) + b4.w : 0.f;
        }
        if (which < 2) {
          uint8_t* dst = img + (dcol0 >> 6) * 16384;
          const int cc0 = (dcol0 & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_f16(o[8 * j], o[8 * j + 1]); pk.y = pack_f16(o[8 * j + 2], o[8 * j + 3]);
            pk.z = pack_f16(o[8 * j + 4], o[8 * j + 5]); pk.w = pack_f16(o[8 * j + 6], o[8 * j + 7]);
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
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (tid == 0) {
        if (which == 2) PTR(8);
        const size_t tix = (size_t)(pair * a.tiles + tile) * (128 * 128);
        bulk_s2g((which == 0 ? a.tq : which == 1 ? a.tk : a.tv) + tix, img, 32768);
        bulk_commit();                                           // leaves while the next block's epilogue runs
      }
    }
    if (tid == 0) { bulk_wait_read(); PTR(9); }
  }
  tc_fence_before();
  __syncthreads();
#ifdef GMF_PCN_TRACE
  if (tr_on && tid == 0) {
    const long long e = clock64();
    printf("pcn_qkv trace blk(%d,%d): setup %lld | img wait start %lld | img landed %lld | W0 landed %lld | acc0 %lld | f1 back %lld | last W landed %lld | acc_full %lld | images built %lld | stores read %lld | end %lld\n",
           blockIdx.x, blockIdx.y, tr_[1] - tr_[0], tr_[1] - tr_[0], tr_[2] - tr_[0], tr_[3] - tr_[0], tr_[4] - tr_[0], tr_[5] - tr_[0], tr_[6] - tr_[0], tr_[7] - tr_[0],
           tr_[8] - tr_[0], tr_[9] - tr_[0], e - tr_[0]);
  }
#endif
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// PERSISTENT variant for layers whose input arrives as the split fp16 tile image (all but the first): one CTA per SM walks the tiles.
// The one-tile-per-CTA kernel above spends 2.3-4 k cycles waiting for its tile image, ~3 k in ring-slot reuse stalls of the second GEMM,
// ~2.8 k draining its image stores and ~2.5 k in launch / setup per 15-18 k cycle tile (GMF_PCN_TRACE).  Here the image of tile t+1 lands in
// the second 64 KB buffer and its PointCN GEMM runs while the workers still assemble the Q / K / V images of tile t; weights stream
// continuously through a 3-deep ring from their own warp; bulk stores drain under the next tile.
//   warps 0-15 workers (epilogues), 16 MMA issuer, 17 weight producer, 18 tile-image loader.
//   Shared memory: A[2] 64 KB each (image -> feat1 fp32 staging -> Q | K images -> V image), ring 3 x 32 KB, biases, barriers = 227 KB
//   (the dynamic window starts 1 KB into the SM's shared memory, i.e. 1024-aligned: the kernel traps if more than 768 B of padding were needed).
// ------------------------------------------------------------------------------------------------
struct PcnQkvPCfg {
  static constexpr int A_BYTES = 128 * 128 * 4, W_BYTES = 128 * 64 * 4, NBUF = 3, NCHUNK = 8;
  static constexpr int BIAS_BYTES = 512 * 4;
  static constexpr int SMEM = 2 * A_BYTES + NBUF * W_BYTES + BIAS_BYTES + 256 + 768;
  static constexpr int THREADS = 608;
};

// The 8 weight chunks of one tile, fully unrolled with COMPILE-TIME ring slots ((S0 + c) % 3; S0 = first slot of the tile, period 3 tiles): with
// run-time slots every MMA paid ~80 cycles of vector -> uniform register traffic for its descriptors (1 k cycles per 12-MMA chunk).
struct PcnIssue {
  uint32_t tm, idesc, leader, tpar, ppar;      // ppar: parity of the previous tile (blk_free)
  uint64_t a_desc, b_desc0;                    // A = this tile's image buffer
  BarArr w_full, w_empty, acc0_full, f1_ready, blk_full, blk_free;
  int s, it;                                   // ring stage of chunk 0
};
template <int S0>
__device__ __forceinline__ void pcn_issue_tile(const PcnIssue& m) {
  using Cfg = PcnQkvPCfg;
#pragma unroll
  for (int c = 0; c < Cfg::NCHUNK; ++c) {
    constexpr int dummy = 0; (void)dummy;
    const int slot = (S0 + c) % Cfg::NBUF;                       // compile-time after unrolling
    const int kc = c & 1;
    if (c == 2) mbar_wait(m.f1_ready, m.tpar);                   // feat1 (fp16 hi | lo) is back in TMEM columns 0..127
    if (c >= 2 && kc == 0 && m.it > 0) mbar_wait(&m.blk_free[(c - 2) >> 1], m.ppar);   // the previous tile's block has been read out
    mbar_wait(&m.w_full[slot], (uint32_t)((m.s + c) / Cfg::NBUF) & 1u);
    tc_fence_after();
    if (m.leader) {
      const uint64_t bd = umma_desc_adv(m.b_desc0, slot * Cfg::W_BYTES);
      // chunk = k-half kc: {w_hi atom | w_lo atom}; three products per K16 step: hi hi, lo hi, hi lo
      if (c < 2) {                                               // PointCN: A = split tile image in shared memory (hi atoms 0,1 | lo atoms 0,1)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = umma_desc_adv(m.a_desc, kc * 16384 + ks * 32), al = umma_desc_adv(m.a_desc, 32768 + kc * 16384 + ks * 32);
          const uint64_t bh = umma_desc_adv(bd, ks * 32), bl = umma_desc_adv(bd, 16384 + ks * 32);
          tc_mma_bf16(m.tm, ah, bh, m.idesc, (kc | ks) ? 1u : 0u);
          tc_mma_bf16(m.tm, al, bh, m.idesc, 1u);
          tc_mma_bf16(m.tm, ah, bl, m.idesc, 1u);
        }
      } else {                                                   // Q / K / V: A = feat1 in tensor memory (hi: columns 0..63, lo: 64..127)
        const int nb = (c - 2) >> 1;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ah = m.tm + kc * 32 + ks * 8, al = m.tm + 64 + kc * 32 + ks * 8;
          const uint64_t bh = umma_desc_adv(bd, ks * 32), bl = umma_desc_adv(bd, 16384 + ks * 32);
          tc_mma_bf16_ts(m.tm + 128 + nb * 128, ah, bh, m.idesc, (kc | ks) ? 1u : 0u);
          tc_mma_bf16_ts(m.tm + 128 + nb * 128, al, bh, m.idesc, 1u);
          tc_mma_bf16_ts(m.tm + 128 + nb * 128, ah, bl, m.idesc, 1u);
        }
      }
      tc_commit(&m.w_empty[slot]);
      if (c == 1) tc_commit(m.acc0_full);
      if (c >= 3 && kc == 1) tc_commit(&m.blk_full[(c - 2) >> 1]);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(PcnQkvPCfg::THREADS, 1) pcn_qkv_persist_kernel(const PcnQkvArgs a, const int pairs) {
  using Cfg = PcnQkvPCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (pad > 768u) __trap();
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;                                     // [2]
  uint8_t* sB = sA + 2 * Cfg::A_BYTES;                    // [3]
  float* sBias = (float*)(sB + Cfg::NBUF * Cfg::W_BYTES); // PointCN bias [128] | Q,K,V biases [384]
  uint64_t* bars = (uint64_t*)((uint8_t*)sBias + Cfg::BIAS_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  const BarArr img_full{bar0};                      // [2]      tile image landed in A[b]
  const BarArr st_ready = BarArr{bar0} + 2;        // [3] 512  Q / K / V image of the tile assembled in A[b] -> the loader thread bulk-stores it
  const BarArr q_drained = BarArr{bar0} + 19;      //          Q's bulk store has read its 32 KB: the V image may be written there
  const BarArr w_full = BarArr{bar0} + 5;          // [3]
  const BarArr w_empty = BarArr{bar0} + 8;         // [3]
  const BarArr acc0_full = BarArr{bar0} + 11;      //          PointCN accumulator complete
  const BarArr f1_ready = BarArr{bar0} + 12;       // 512      feat1 (fp16 hi | lo) back in tensor memory
  const BarArr blk_full = BarArr{bar0} + 13;       // [3]      Q / K / V accumulator complete
  const BarArr blk_free = BarArr{bar0} + 16;       // [3] 512  workers have read the Q / K / V accumulator of the previous tile
  uint32_t* tmem_slot = (uint32_t*)(bars + 20);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = a.tiles * pairs;
#ifdef GMF_FFN_TRACE
#define PTRC(role, idx) do { if (a.trace && blockIdx.x == 3 && it == 2 && lane == 0) a.trace[(role) * 32 + (idx)] = clock64(); } while (0)
#else
#define PTRC(role, idx) do {} while (0)
#endif
  if (tid < 512) sBias[tid] = tid < 128 ? a.pcn_bias[tid] : a.qkv_bias[tid - 128];
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&img_full[i], 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); mbar_init(&blk_full[i], 1); mbar_init(&blk_free[i], 512); mbar_init(&st_ready[i], 512); }
    mbar_init(acc0_full, 1); mbar_init(f1_ready, 512); mbar_init(q_drained, 1);
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 18) {
    // ------------------------------- tile-image loader AND image storer: one thread owns every bulk copy that touches A[] -------------------------------
    // so it knows without asking when a buffer's stores have drained (bulk async-groups are per thread) and no worker ever waits for a store.
    if (elect_one()) {
      const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      mbar_expect_tx(&img_full[0], Cfg::A_BYTES);
      bulk_g2s(sA, a.x_img + (size_t)blockIdx.x * (128 * 128), Cfg::A_BYTES, &img_full[0]);
#pragma unroll 1
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        const size_t g = blockIdx.x + (size_t)it * gridDim.x;
        if (it + 1 < my_tiles) {                                 // image of the next tile -> the other buffer, once the stores of tile it - 1 have read it
          if (it >= 1) bulk_wait_read();
          mbar_expect_tx(&img_full[b ^ 1], Cfg::A_BYTES);
          bulk_g2s(sA + (b ^ 1) * Cfg::A_BYTES, a.x_img + (g + gridDim.x) * (128 * 128), Cfg::A_BYTES, &img_full[b ^ 1]);
        }
#pragma unroll 1
        for (int which = 0; which < 3; ++which) {
          mbar_wait(&st_ready[which], it & 1);
          bulk_s2g((which == 0 ? a.tq : which == 1 ? a.tk : a.tv) + g * (128 * 128), sA + b * Cfg::A_BYTES + (which == 1 ? 32768 : 0), 32768);
          bulk_commit();
          if (which == 1) {                                      // Q's store (all but the most recent group) has read its half of A[b]
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            mbar_arrive(q_drained);
          }
        }
      }
      bulk_wait_read();
    }
    __syncwarp();
  } else if (warp == 17) {
    // ------------------------------- weight producer: 8 chunks per tile through a 3-deep ring, running across tiles -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint8_t* wsrc = (const uint8_t*)a.w_packed;
    int s = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x)
#pragma unroll 1
      for (int c = 0; c < Cfg::NCHUNK; ++c, ++s) {
        const int slot = s % Cfg::NBUF;
        if (s >= Cfg::NBUF) mbar_wait(&w_empty[slot], ((s / Cfg::NBUF) - 1) & 1);
        mbar_expect_tx_p(&w_full[slot], Cfg::W_BYTES, leader);
        bulk_g2s_p(sB + slot * Cfg::W_BYTES, wsrc + (size_t)c * Cfg::W_BYTES, Cfg::W_BYTES, &w_full[slot], leader);
      }
  } else if (warp == 16) {
    // ------------------------------- MMA issuer -------------------------------
    PcnIssue m;
    m.leader = elect_one() ? 1u : 0u;
    m.tm = __shfl_sync(0xffffffffu, tmem, 0);
    m.idesc = umma_idesc(128, 128, kFmtF16);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    m.b_desc0 = umma_desc_sw128(smem_u32(sB));
    m.w_full = w_full; m.w_empty = w_empty; m.acc0_full = acc0_full; m.f1_ready = f1_ready; m.blk_full = blk_full; m.blk_free = blk_free;
    int it = 0, ph = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
      const int b = it & 1;
      m.tpar = it & 1; m.ppar = (it - 1) & 1; m.it = it; m.s = it * Cfg::NCHUNK;
      m.a_desc = umma_desc_adv(a_desc0, b * Cfg::A_BYTES);
      mbar_wait(&img_full[b], (it >> 1) & 1);
      if (ph == 0) pcn_issue_tile<0>(m);                         // first ring slot of tile it = (8 it) % 3 = 0, 2, 1, 0, ...
      else if (ph == 1) pcn_issue_tile<2>(m);
      else pcn_issue_tile<1>(m);
      ph = ph == 2 ? 0 : ph + 1;
    }
  } else if (warp < 16) {
    // ------------------------------- workers -------------------------------
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int srow = lane >> 3, sj = lane & 7;
    int it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t tpar = it & 1;
      const int pair = g / a.tiles, tile = g - pair * a.tiles;
      const int row0 = tile * 128;
      const bool valid = row0 + r < a.L;
      uint8_t* sAb = sA + b * Cfg::A_BYTES;
      // ---------------- epilogue 0: feat1 = ReLU(acc + b) -> HBM (coalesced) and back to TMEM as fp16 hi | lo ----------------
      if (warp == 0) PTRC(0, 0);
      mbar_wait(acc0_full, tpar);
      tc_fence_after();
      if (warp == 0) PTRC(0, 1);
      {
        float* stg = (float*)sAb + warp * 1024;                  // the tile image is dead once acc0_full fired
        uint32_t v[32], hw[16], lw[16];
        tmem_ld32(trow + part * 32, v);
        tmem_ld_wait();
        const int col0 = part * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(sBias + col0 + 4 * j);
          const float4 o = make_float4(fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.f), fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.f),
                                       fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.f));
          *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = o;
          split_f16x2(o.x, o.y, hw[2 * j], lw[2 * j]);
          split_f16x2(o.z, o.w, hw[2 * j + 1], lw[2 * j + 1]);
        }
        // the packed operand (hi: columns 0..63, lo: 64..127) overwrites accumulator columns the other three warps of this lane quadrant may
        // still be reading
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
        tmem_st16(trow + part * 16, hw);
        tmem_st16(trow + 64 + part * 16, lw);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(f1_ready);
        if (warp == 0) PTRC(0, 2);
        __syncwarp();
        const size_t gbase = ((size_t)pair * a.L + row0 + q * 32) * 128 + col0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (row0 + q * 32 + rw < a.L)
            *reinterpret_cast<float4*>(a.feat1 + gbase + (size_t)rw * 128 + sj * 4) = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
        }
      }
      // ---------------- epilogue 1: Q / K / V^T tile images: Q -> A[b] lower half, K -> upper half, V -> lower half once Q's store has drained ------
#pragma unroll 1
      for (int which = 0; which < 3; ++which) {
        if (warp == 0) PTRC(0, 3 + 3 * which);
        mbar_wait(&blk_full[which], tpar);
        tc_fence_after();
        if (warp == 0) PTRC(0, 4 + 3 * which);
        uint32_t v[32];
        tmem_ld32(trow + 128 + which * 128 + part * 32, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&blk_free[which]);
        uint8_t* img = sAb + (which == 1 ? 32768 : 0);
        const int col0 = which * 128 + part * 32, dcol0 = part * 32;
        float o[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(sBias + 128 + col0 + 4 * i);
          o[4 * i] = valid ? __uint_as_float(v[4 * i]) + b4.x : 0.f;
          o[4 * i + 1] = valid ? __uint_as_float(v[4 * i + 1]) + b4.y : 0.f;
          o[4 * i + 2] = valid ? __uint_as_float(v[4 * i + 2]) + b4.z : 0.f;
          o[4 * i + 3] = valid ? __uint_as_float(v[4 * i + 3]) + b4.w : 0.f;
        }
        // Q (and K, in the upper half) overwrite the feat1 staging tiles: every warp must have stored its feat1 rows; V overwrites the Q image:
        // its bulk store must have read it
        if (which == 0) asm volatile("bar.sync 1, 512;" ::: "memory");
        if (which == 2) mbar_wait(q_drained, tpar);
        if (which < 2) {
          uint8_t* dst = img + (dcol0 >> 6) * 16384;
          const int cc0 = (dcol0 & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_f16(o[8 * j], o[8 * j + 1]); pk.y = pack_f16(o[8 * j + 2], o[8 * j + 3]);
            pk.z = pack_f16(o[8 * j + 4], o[8 * j + 5]); pk.w = pack_f16(o[8 * j + 6], o[8 * j + 7]);
            *reinterpret_cast<uint4*>(dst + swz_off(r, cc0 + j)) = pk;
          }
        } else {
          uint8_t* dst = img + (r >> 6) * (128 * 128) + (r & 7) * 2;
          const int kchunk = (r & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 32; ++i) *reinterpret_cast<__nv_bfloat16*>(dst + swz_off(dcol0 + i, kchunk)) = __float2bfloat16_rn(o[i]);
        }
        fence_proxy_async();
        mbar_arrive(&st_ready[which]);                           // the loader thread ships the image
        if (warp == 0) PTRC(0, 5 + 3 * which);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

inline cudaError_t launch_pcn_qkv(const PcnQkvArgs& a, int pairs, cudaStream_t st) {
  if (a.x_img) {                                                 // tile-image input: persistent kernel
    static std::atomic<unsigned long long> configured_p{0};
    if (cudaError_t e = ensure_dyn_smem(pcn_qkv_persist_kernel, PcnQkvPCfg::SMEM, configured_p)) return e;
    const int total = a.tiles * pairs;
    if (total <= 0) return cudaSuccess;
    pcn_qkv_persist_kernel<<<min(total, device_sm_count()), PcnQkvPCfg::THREADS, PcnQkvPCfg::SMEM, st>>>(a, pairs);
    return cudaGetLastError();
  }
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = ensure_dyn_smem(pcn_qkv_kernel, PcnQkvCfg::SMEM, configured)) return e;
  pcn_qkv_kernel<<<dim3(a.tiles, pairs), 544, PcnQkvCfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
