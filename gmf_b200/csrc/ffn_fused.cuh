// Fused GEGLU feed-forward block of the fusion layers (fusion_layer.py:54-69 behind PreNorm :32-52, residual :191):
//     out = x + W2 . ( (W1v . LN(x) + b1v) * gelu(W1g . LN(x) + b1g) ) + b2          x: [L, 128], hidden 512
// PERSISTENT kernel: one CTA per SM walks over the 128-token tiles of all pairs; the 128 x 512 hidden activation never leaves the SM, and
// the LayerNorm of the NEXT tile, the weight stream and the first GEMM passes of the next tile run underneath the GEGLU arithmetic and the
// output epilogue of the current one (round-2 timeline of the one-tile-per-CTA version: 7.7 k cycles LayerNorm prologue + 12.5 k passes +
// 2.5 k fc_message tail + 6 k epilogue + launch gap per tile; only the passes are arithmetic - profiles/r02_summary.md).
//
// 8 passes of 64 hidden columns per tile.  Pass p:  ACC1[p&1] = LN(x) . [W1v_p | W1g_p]^T   (8 SS-mode fp16 MMAs, N = 128, K = 128)
//                                                   H[.]      = GEGLU(ACC1[p&1])            (worker warps, TMEM -> registers -> TMEM, fp16)
//                                                   OUT      += H[.] . W2_p^T               (4 TS-mode fp16 MMAs: A operand = H in TMEM)
// fp16 operands carry the same 11-bit significand as the TF32 path they replace (values saturate at the fp16 maximum); accumulation is fp32.
// With the fused NonLocalBlock tail (m2 != NULL) the LayerNorm warps also park m2 (fp16) in tensor memory and a ninth MMA2 step adds m2 . W3^T.
// TMEM (512 columns): ACC1 2 x 128 | H 2 x 32 | m2 2 x 32 | OUT 128.  Shared memory: LN(x) tile image 2 x 32 KB (double buffered across tiles), W1 ring
// 2 x 32 KB, W2 ring 2 x 16 KB, epilogue staging 16 x 4 KB.  Warps: 0-15 workers (GEGLU, output epilogue), 16 MMA1 issuer, 17 MMA2 issuer,
// 18 / 19 W1 / W2 producers, 20-23 LayerNorm of the tile after the current one (one per SM sub-partition: the kernel is bound by instruction issue, and
// a single LayerNorm warp competing with four GEGLU warps on its sub-partition needed 40 k cycles per tile).  768 threads x 80 registers.
#pragma once
#include <cuda_fp16.h>
#include "linear_tc.cuh"

namespace gmf {

struct FfnCfg {
  static constexpr int A_BYTES = 128 * 128 * 2;            // LN(x) as 2 swizzle atoms of 128 rows x 64 halfs
  static constexpr int W1_BYTES = 128 * 128 * 2;           // one pass of W1: 128 rows (64 value | 64 gate) x 128 k
  static constexpr int W2_BYTES = 128 * 64 * 2;            // one pass of W2: 128 out rows x 64 hidden
  static constexpr int STG_BYTES = 16 * 4096;              // one 32 x 32 fp32 tile per worker warp
  static constexpr int SMEM = 1024 + 2 * A_BYTES + 2 * W1_BYTES + 2 * W2_BYTES + STG_BYTES + 512 + 512;   // barriers; b2 (+ b3)
  static constexpr int COL_H = 256, COL_M2 = 320, COL_OUT = 384;
  static constexpr int PASSES = 8;
  static constexpr int THREADS = 768;
};

struct FfnArgs {
  const float* x;          // [B, L, 128] block input (also the residual)
  int L, tiles, pairs;
  const float* ln_g;
  const float* ln_b;
  const float* w1_packed;  // 8 passes x [128 rows (64 value | 64 gate) x 128 k] swizzled fp16 (2 atoms of 64 k)
  const float* b1;         // [1024]: value 0..511, gate 512..1023
  const float* w2_packed;  // 8 chunks x [128 out rows x 64 hidden] swizzled fp16 (1 atom)
  const float* b2;         // [128]
  float* out;              // [B, L, 128]
#ifdef GMF_FFN_TRACE
  long long* trace;
#endif
  // optional fused tail of NonLocalBlock.forward (PointDSC.py:65,73): out += fc_message.6(m2) = m2 . W3^T + b3
  const float* m2;         // [B, L, 64] (ReLU(BN(conv(...))) output of fc_message.4) or NULL
  const float* w3_packed;  // [128 out rows x 64] swizzled fp16
  const float* b3;         // [128]
  // out_img != NULL: instead of the row-major `out`, write the SPLIT fp16 K-major tile image [B][tiles][hi 2 x 16 KB | lo 2 x 16 KB] (x = hi + lo,
  // common.cuh split_f16; 64 KB like the fp32 tile) that the next layer's
  // chained PointCN/QKV kernel loads with one bulk copy (the per-warp staging tiles already are 32-row blocks of that image)
  float* out_img;
};

// residual rows [32 x 32 floats] of one worker warp -> its swizzled staging tile, 8 x 16-byte asynchronous copies per lane
__device__ __noinline__ void ffn_stage_residual(float* stg, const float* src, const float* valid_ptr, int rows_left, int lane) {
  const int srow = lane >> 3, sj = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rw = i * 4 + srow;
    const bool ok = rw < rows_left;                              // rows past L: zero fill, the (unread) source address stays inside the tensor
    cp_async16(stg + rw * 32 + ((sj ^ (rw & 7)) << 2), ok ? src + (size_t)rw * 128 + sj * 4 : valid_ptr, ok);
  }
  cp_async_commit();
}

__global__ void __launch_bounds__(FfnCfg::THREADS, 1) ffn_fused_kernel(const FfnArgs a) {
  using Cfg = FfnCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                     // [2] LN(x) images
  uint8_t* sW1 = sA + 2 * Cfg::A_BYTES;                   // [2] stages
  uint8_t* sW2 = sW1 + 2 * Cfg::W1_BYTES;                 // [2] stages
  float* sStg = (float*)(sW2 + 2 * Cfg::W2_BYTES);
  uint64_t* bars = (uint64_t*)((uint8_t*)sStg + Cfg::STG_BYTES);
  const uint32_t bar0 = smem_u32(bars);                    // barriers are addressed as bar0 + 8 i (common.cuh BarArr): no per-op address re-derivation
  const BarArr a_ready{bar0};                     // [2] 128  LayerNorm warps -> MMA1
  const BarArr a_free = BarArr{bar0} + 2;  // [2]      MMA1 of a tile retired -> LayerNorm warp (tile + 2)
  const BarArr full1 = BarArr{bar0} + 4;  // [2]
  const BarArr empty1 = BarArr{bar0} + 6;  // [2]
  const BarArr full2 = BarArr{bar0} + 8;  // [2]
  const BarArr empty2 = BarArr{bar0} + 10;  // [2]
  const BarArr acc1_full = BarArr{bar0} + 12;  // [2]
  const BarArr acc1_free = BarArr{bar0} + 14;  // [2] 512
  const BarArr h_ready = BarArr{bar0} + 16;  // [2] 512
  const BarArr h_free = BarArr{bar0} + 18;  // [2]
  const BarArr out_full = BarArr{bar0} + 20;  //          last MMA2 of a tile retired -> workers
  const BarArr out_free = BarArr{bar0} + 21;  // 512      workers have read OUT -> MMA2 of the next tile
  const BarArr m2_ready = BarArr{bar0} + 22;  // [2] 128  LayerNorm warps parked m2 (fp16) in tensor memory -> MMA2
  const BarArr m2_free = BarArr{bar0} + 24;  // [2]      tail MMA of a tile retired -> LayerNorm warps (tile + 2)
  uint32_t* tmem_slot = (uint32_t*)(bars + 26);
  float* sBias = (float*)(bars + 64);                     // [128] b2 (+ b3): output bias of the block

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = a.tiles * a.pairs;
  const int NP = a.m2 ? Cfg::PASSES + 1 : Cfg::PASSES;     // H / W2 pipeline steps per tile
#ifdef GMF_FFN_TRACE
#define TR(role, idx) do { if (a.trace && blockIdx.x == 3 && it == 1 && lane == 0) a.trace[(role) * 64 + (idx)] = clock64(); } while (0)
#else
#define TR(role, idx) do {} while (0)
#endif

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_ready[i], 128); mbar_init(&a_free[i], 1);
      mbar_init(&full1[i], 1); mbar_init(&empty1[i], 1); mbar_init(&full2[i], 1); mbar_init(&empty2[i], 1);
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_free[i], 512); mbar_init(&h_ready[i], 512); mbar_init(&h_free[i], 1);
      mbar_init(&m2_ready[i], 128); mbar_init(&m2_free[i], 1);
    }
    mbar_init(out_full, 1); mbar_init(out_free, 512);
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (tid < 128) sBias[tid] = a.b2[tid] + (a.m2 ? a.b3[tid] : 0.f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 18) {
    // ------------------------------- W1 producer: one 32 KB stage per pass through a 2-deep ring, running across tiles -------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint8_t* w1 = (const uint8_t*)a.w1_packed;
    int s1 = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
#pragma unroll 1
      for (int p = 0; p < Cfg::PASSES; ++p, ++s1) {
        const int slot = s1 & 1;
        if (s1 >= 2) mbar_wait(&empty1[slot], ((s1 >> 1) - 1) & 1);
        mbar_expect_tx_p(&full1[slot], Cfg::W1_BYTES, leader);
        bulk_g2s_p(sW1 + slot * Cfg::W1_BYTES, w1 + (size_t)p * Cfg::W1_BYTES, Cfg::W1_BYTES, &full1[slot], leader);
      }
    }
  } else if (warp == 19) {
    // ------------------------------- W2 producer: one 16 KB chunk per MMA2 step (chunk 8 = fc_message.6 weight of the fused block tail) -----
    // A warp of its own: behind the W1 loads in one instruction stream, a wait for a W2 slot (MMA2 two steps back) delayed the W1 stage of the
    // next pass and with it MMA1.
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint8_t* w2 = (const uint8_t*)a.w2_packed;
    int s2 = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x) {
#pragma unroll 1
      for (int c = 0; c < NP; ++c, ++s2) {
        const int slot = s2 & 1;
        if (s2 >= 2) mbar_wait(&empty2[slot], ((s2 >> 1) - 1) & 1);
        mbar_expect_tx_p(&full2[slot], Cfg::W2_BYTES, leader);
        bulk_g2s_p(sW2 + slot * Cfg::W2_BYTES, c < Cfg::PASSES ? w2 + (size_t)c * Cfg::W2_BYTES : (const uint8_t*)a.w3_packed, Cfg::W2_BYTES,
                   &full2[slot], leader);
      }
    }
  } else if (warp == 16) {
    // ------------------------------- MMA1 issuer: ACC1[p&1] = LN(x) . W1_p^T -------------------------------
    // 8 passes per tile and 2-deep rings: buffer indices and barrier parities of pass p are the same in every tile.
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, kFmtF16);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW1));
    int it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
      const int ab = it & 1;
      mbar_wait(&a_ready[ab], (it >> 1) & 1);
      const uint64_t a_desc = umma_desc_adv(a_desc0, ab * Cfg::A_BYTES);
#pragma unroll
      for (int p = 0; p < Cfg::PASSES; ++p) {
        const int b = p & 1;
        TR(1, 4 * p);
        if (it > 0 || p >= 2) mbar_wait(&acc1_free[b], ((p >> 1) - 1) & 1);
        TR(1, 4 * p + 1);
        mbar_wait(&full1[b], (p >> 1) & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int at = 0; at < 2; ++at)                       // 64 k per swizzle atom, 16 k per MMA
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_bf16(tm + b * 128, umma_desc_adv(a_desc, at * 16384 + ks * 32), umma_desc_adv(w_desc0, b * Cfg::W1_BYTES + at * 16384 + ks * 32),
                          idesc, (at | ks) ? 1u : 0u);
          tc_commit(&empty1[b]);
          tc_commit(&acc1_full[b]);
          if (p == Cfg::PASSES - 1) tc_commit(&a_free[ab]);     // the LN(x) image of this tile is dead
        }
        __syncwarp();
        TR(1, 4 * p + 2);
      }
    }
  } else if (warp == 17) {
    // ------------------------------- MMA2 issuer: OUT += H . W2_p^T (A operand from tensor memory) -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, kFmtF16);
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW2));
    int it = 0, s2 = 0;                                        // s2: running W2 ring stage (8 or 9 per tile)
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
#pragma unroll 1
      for (int vp = 0; vp < NP; ++vp, ++s2) {
        const int ws = s2 & 1;
        const uint32_t wpar = (s2 >> 1) & 1;
        const bool tail = vp == Cfg::PASSES;                     // fused block tail: A operand = m2 parked by the LayerNorm warps
        const int hb = vp & 1;                                   // 8 H steps per tile: buffer and parity of step vp are the same in every tile
        TR(2, 2 * vp);
        if (tail) mbar_wait2(&m2_ready[it & 1], (it >> 1) & 1, &full2[ws], wpar);
        else mbar_wait2(&h_ready[hb], (vp >> 1) & 1, &full2[ws], wpar);
        if (vp == 0 && it > 0) mbar_wait(out_free, (it - 1) & 1);      // the epilogue of the previous tile has read OUT
        tc_fence_after();
        TR(2, 2 * vp + 1);
        if (leader) {
          const uint32_t acol = tail ? tm + Cfg::COL_M2 + (it & 1) * 32 : tm + Cfg::COL_H + hb * 32;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            tc_mma_bf16_ts(tm + Cfg::COL_OUT, acol + i * 8, umma_desc_adv(w_desc0, ws * Cfg::W2_BYTES + i * 32), idesc, (vp | i) ? 1u : 0u);
          tc_commit(&empty2[ws]);
          if (tail) tc_commit(&m2_free[it & 1]); else tc_commit(&h_free[hb]);
          if (vp == NP - 1) tc_commit(out_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 20) {
    // ------------------------------- LayerNorm warps: A operand (fp16, swizzled) of the tile AFTER the one being multiplied ---------------
    // Four warps x 32 rows per tile in batches of 8 rows (lanes across the 128 channels); the lines of the next tile are pulled into L2 while
    // this one is reduced.
    const int lw = warp - 20;
    const int c4 = lane * 4;
    const float4 g4 = *reinterpret_cast<const float4*>(a.ln_g + c4);
    const float4 b4 = *reinterpret_cast<const float4*>(a.ln_b + c4);
    const uint64_t g01 = pack2(g4.x, g4.y), g23 = pack2(g4.z, g4.w), b01 = pack2(b4.x, b4.y), b23 = pack2(b4.z, b4.w);
    constexpr int RPW = 8;
    // byte offset of this lane's 8-byte piece in row (8 k + i) of the image: atom lane >> 4 (64 channels = 128 B per row), 16-byte piece
    // (lane & 15) >> 1 swizzled with the row, half of it lane & 1
    const uint32_t lane_off = (lane >> 4) * 16384 + (lane & 1) * 8;
    const uint32_t piece = (lane & 15) >> 1;
    int it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
      const int ab = it & 1;
      const int pair = g / a.tiles, tile = g - pair * a.tiles;
      const float* xp = a.x + ((size_t)pair * a.L + tile * 128 + lw * 32) * 128;
      const int rows_left = a.L - (tile * 128 + lw * 32);      // rows of this warp's block that exist
      {                                                         // next tile of this CTA: 32 rows x 4 lines of 128 B per warp
        const int gn = g + gridDim.x;
        if (gn < total) {
          const int pn = gn / a.tiles, tn = gn - pn * a.tiles;
          const float* xn = a.x + ((size_t)pn * a.L + tn * 128 + lw * 32) * 128;
          const int left = a.L - (tn * 128 + lw * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = k * 8 + (lane >> 2);
            if (row < left) asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + (size_t)row * 128 + (lane & 3) * 32));
          }
        }
      }
      TR(4, 0);
      if (it >= 2) mbar_wait(&a_free[ab], ((it >> 1) - 1) & 1);
      TR(4, 1);
      uint8_t* img = sA + ab * Cfg::A_BYTES + lane_off;
#pragma unroll 1
      for (int bt = 0; bt < 4; ++bt) {
        const int rbase = bt * 8;
        float4 rv[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rbase + i < rows_left) rv[i] = *reinterpret_cast<const float4*>(xp + (size_t)(rbase + i) * 128 + c4);
        }
        // the 8 rows of a batch are reduced together (common.cuh warp_sum8; rows past L are zeros)
        float mean[RPW], rs[RPW];
        uint64_t d01[RPW], d23[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i) mean[i] = (rv[i].x + rv[i].y) + (rv[i].z + rv[i].w);
        warp_sum8(mean, lane);
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const float nm = mean[i] * (-1.0f / 128.0f);
          const uint64_t nm2 = pack2(nm, nm);
          d01[i] = fadd2(pack2(rv[i].x, rv[i].y), nm2);
          d23[i] = fadd2(pack2(rv[i].z, rv[i].w), nm2);
          float s0, s1;
          unpack2(ffma2(d23[i], d23[i], fmul2(d01[i], d01[i])), s0, s1);
          rs[i] = s0 + s1;
        }
        warp_sum8(rs, lane);
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const float r_ = rsqrtf(rs[i] * (1.0f / 128.0f) + 1e-5f);
          const uint64_t r2 = pack2(r_, r_);
          float y0, y1, y2, y3;
          unpack2(ffma2(fmul2(d01[i], r2), g01, b01), y0, y1);
          unpack2(ffma2(fmul2(d23[i], r2), g23, b23), y2, y3);
          uint2 pk;
          pk.x = pack_f16(y0, y1);
          pk.y = pack_f16(y2, y3);
          // row lw * 32 + rbase + i: (row & 7) == i
          *reinterpret_cast<uint2*>(img + (lw * 32 + rbase + i) * 128 + ((piece ^ (uint32_t)i) << 4)) = pk;
        }
      }
      fence_proxy_async();
      mbar_arrive(&a_ready[ab]);
      TR(4, 2);
      if (a.m2) {
        // fused block tail: row lw * 32 + lane of m2 (64 floats) -> fp16 -> tensor memory (lane = row: this warp's lane quadrant is lw)
        if (it >= 2) { mbar_wait(&m2_free[ab], ((it >> 1) - 1) & 1); tc_fence_after(); }
        const int row = tile * 128 + lw * 32 + lane;
        const float4* mp = reinterpret_cast<const float4*>(a.m2 + ((size_t)pair * a.L + row) * 64);
        const uint32_t tdst = tmem + ((uint32_t)(lw * 32) << 16) + Cfg::COL_M2 + ab * 32;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t hw[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < a.L) m4 = __ldg(mp + h * 8 + i);
            hw[2 * i] = pack_f16(m4.x, m4.y);
            hw[2 * i + 1] = pack_f16(m4.z, m4.w);
          }
          tmem_st16(tdst + h * 16, hw);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&m2_ready[ab]);
      }
    }
  } else if (warp < 16) {
    // ------------------------------- workers (16 warps: 4 lane quadrants x 4 column quarters) -------------------------------
    const int q = warp & 3, cq = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    float* stg = sStg + warp * 1024;
    const int srow = lane >> 3, sj = lane & 7;
    // Software pipeline over the tiles of this CTA: the output epilogue of tile i - 1 runs AFTER the first GEGLU pass of tile i, so that the
    // last MMA2s of tile i - 1 (and the fused block tail) retire underneath that pass instead of being waited for.
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
#pragma unroll 1
    for (int it = 0; it <= my_tiles; ++it) {
      const bool has = it < my_tiles;                            // it == my_tiles: only the epilogue of the last tile is left
      const int g = blockIdx.x + it * gridDim.x;
      const int pair = g / a.tiles, tile = g - pair * a.tiles;
      const int row0 = tile * 128;
      if (warp == 0) TR(0, 0);
      // ---------------- GEGLU between the two GEMMs ----------------
#pragma unroll 1
      for (int p = 0; p < (has ? Cfg::PASSES : 1); ++p) {
        if (has) {
          const int b = p & 1;                                     // 8 passes per tile: buffers and parities of pass p are the same in every tile
          const int hc0 = p * 64 + cq * 16;                        // hidden column of v[0]
          float4 b1v[4], b1g[4];                                   // bias loads in flight while waiting for the accumulator
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            b1v[i] = __ldg(reinterpret_cast<const float4*>(a.b1 + hc0) + i);
            b1g[i] = __ldg(reinterpret_cast<const float4*>(a.b1 + 512 + hc0) + i);
          }
          if (p == Cfg::PASSES - 3) {
            // residual rows of this warp's 32 x 32 output block -> staging tile (asynchronous, zero-filled past L); the image blocks of the
            // previous tile have long left the tile (their bulk stores were issued more than five passes ago).  Out of line: inlined, its
            // address arithmetic was hoisted into every pass.
            if (a.out_img) {
              if (lane == 0) bulk_wait_read();
              __syncwarp();
            }
            ffn_stage_residual(stg, a.x + ((size_t)pair * a.L + row0 + q * 32) * 128 + cq * 32, a.x, a.L - (row0 + q * 32), lane);
          }
          if (warp == 0) TR(3, 4 * p);
          mbar_wait(&acc1_full[b], (p >> 1) & 1);
          tc_fence_after();
          if (warp == 0) TR(3, 4 * p + 1);
          uint32_t v[16], gt[16];
          tmem_ld16(trow + b * 128 + cq * 16, v);
          tmem_ld16(trow + b * 128 + 64 + cq * 16, gt);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&acc1_free[b]);                              // MMA1 two passes on may overwrite the accumulator
          uint32_t hw[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b1 = b1v[i], b2 = b1g[i];
            float o0, o1, o2, o3;
            unpack2(geglu2(pack2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), pack2(b1.x, b1.y),
                           pack2(__uint_as_float(gt[4 * i]), __uint_as_float(gt[4 * i + 1])), pack2(b2.x, b2.y)), o0, o1);
            unpack2(geglu2(pack2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), pack2(b1.z, b1.w),
                           pack2(__uint_as_float(gt[4 * i + 2]), __uint_as_float(gt[4 * i + 3])), pack2(b2.z, b2.w)), o2, o3);
            hw[2 * i] = pack_f16(o0, o1);
            hw[2 * i + 1] = pack_f16(o2, o3);
          }
          if (warp == 0) TR(3, 4 * p + 2);
          if (it > 0 || p >= 2) { mbar_wait(&h_free[b], ((p >> 1) - 1) & 1); tc_fence_after(); }   // MMA2 two passes back has read H[b]
          if (warp == 0) TR(3, 4 * p + 3);
          tmem_st8(trow + Cfg::COL_H + b * 32 + cq * 8, hw);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&h_ready[b]);
      
        }
        if (p == 0 && it > 0) {
          const int eg = g - gridDim.x, eit = it - 1;
          const int epair = eg / a.tiles, erow0 = (eg - epair * a.tiles) * 128;
          // ---------------- out = OUT + b2 (+ b3) + x, assembled through the per-warp staging tile (one 32-column chunk per warp) ----------------
          if (warp == 0) TR(0, 2);
          const int col0 = cq * 32;
          const size_t gbase = ((size_t)epair * a.L + erow0 + q * 32) * 128 + col0;
          cp_async_wait_all();                                       // residual rows have landed in the staging tile
          __syncwarp();
          if (warp == 0) TR(0, 7);
          mbar_wait(out_full, eit & 1);
          tc_fence_after();
          if (warp == 0) TR(0, 3);
          uint32_t v[32];
          tmem_ld32(trow + Cfg::COL_OUT + cq * 32, v);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(out_free);                                     // MMA2 of the next tile may start accumulating
          if (warp == 0) TR(0, 8);
          const bool live = erow0 + r < a.L;
  #pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = *reinterpret_cast<const float4*>(sBias + col0 + 4 * j);
            float4* slot = reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2));
            const float4 res = *slot;
            const float4 o = make_float4(__uint_as_float(v[4 * j]) + bb.x + res.x, __uint_as_float(v[4 * j + 1]) + bb.y + res.y,
                                         __uint_as_float(v[4 * j + 2]) + bb.z + res.z, __uint_as_float(v[4 * j + 3]) + bb.w + res.w);
            if (a.out_img) {                                         // kept in registers: the image is assembled below
              v[4 * j] = __float_as_uint(live ? o.x : 0.f); v[4 * j + 1] = __float_as_uint(live ? o.y : 0.f);
              v[4 * j + 2] = __float_as_uint(live ? o.z : 0.f); v[4 * j + 3] = __float_as_uint(live ? o.w : 0.f);
            } else {
              *slot = o;
            }
          }
          __syncwarp();
          if (warp == 0) TR(0, 9);
          if (a.out_img) {
            // Split fp16 image (x = hi + lo): fp16 swizzle atoms are 64 columns wide, so warps (q, 2m) and (q, 2m + 1) share the 32-row blocks of
            // atom m.  Their two 4 KB staging tiles become that block of the hi image (even warp's tile) and of the lo image (odd warp's tile);
            // each warp then ships the block that sits in its own tile.
            const int m = cq >> 1;
            const int pair_bar = 1 + m * 4 + q;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // both warps have read their residual rows out of the tiles
            if (warp == 0) TR(0, 10);
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
            if (warp == 0) TR(0, 11);
            // 16 x 4 KB bulk stores in flight, no CTA-wide barrier: tile layout [hi atom 0 | hi atom 1 | lo atom 0 | lo atom 1], 16 KB each
            if (lane == 0) {
              uint8_t* img = (uint8_t*)(a.out_img + (size_t)eg * (128 * 128));
              bulk_s2g(img + (cq & 1) * 32768 + (cq >> 1) * 16384 + q * 4096, stg, 4096);
              bulk_commit();
            }
            __syncwarp();
          } else {
  #pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rw = i * 4 + srow;
              if (erow0 + q * 32 + rw < a.L)
                *reinterpret_cast<float4*>(a.out + gbase + (size_t)rw * 128 + sj * 4) = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
            }
            __syncwarp();
          }
          if (warp == 0) TR(0, 4);
        }
      }
    }
    if (a.out_img && lane == 0) bulk_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

inline cudaError_t launch_ffn_fused(FfnArgs a, int pairs, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = ensure_dyn_smem(ffn_fused_kernel, FfnCfg::SMEM, configured)) return e;
  a.pairs = pairs;
  const int total = a.tiles * pairs;
  if (total <= 0) return cudaSuccess;
  ffn_fused_kernel<<<min(total, device_sm_count()), FfnCfg::THREADS, FfnCfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
