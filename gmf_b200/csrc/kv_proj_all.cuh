// Context K / V projections of ALL encoder layers in one persistent kernel (fusion_layer.py:172-188 for the context side of every
// fusion_layer_2: x_l = ctx + dwconv_l(ctx) (conditional position encoding, k = 3 along the token axis), LN_l(x_l) . Wkv_l^T -> K_l | V_l).
// Every layer's context is the same Fusion-1 output (PointDSC.py:137), so a 128-token tile of it is read from HBM ONCE, kept in shared
// memory as fp32 (with its two halo rows) and pushed through the 12 layers' position encoding / LayerNorm / GEMM / tile-image epilogue,
// instead of 12 launches that each re-read the 157 MB context and expose their LayerNorm prologue (round 1-2: 12 x 0.115 ms at 0.31 of the
// HBM roofline; an in-CTA layer loop without role overlap did not help).  Roles, one CTA per SM, 640 threads:
//   warps 0-7   position encoding + LayerNorm of layer l -> A[n&1] (fp16, swizzled; 16 rows per warp), one step ahead of the MMA
//   warps 8-15  epilogue of step n: ACC[n&1] -> K tile image (fp16, K-major) | V^T tile image (bf16) in a staging buffer -> two 16 KB bulk stores
//               (history: four LayerNorm warps + 16 epilogue warps: 6.4 k cycles per layer, everything waiting for the LayerNorm; 16 warps
//               alternating both jobs: 5.2 k, the two phases serialised in every warp)
//   warp  16    MMA issuer: ACC[n&1] = A[n&1] . W_l^T   (8 fp16 MMAs, N = 128, K = 128; fp16 operands = the 11-bit significand of TF32)
//   warp  17    weight producer (32 KB per layer through a 2-deep ring)
//   warp  18    context loader (one bulk copy per tile: 130 rows x 512 B, rows outside the sequence zero-filled)
// Step counter n = tile iteration x layers + layer; all double buffers are indexed by n & 1 with parity (n >> 1) & 1.
#pragma once
#include <cuda_fp16.h>
#include "linear_tc.cuh"

namespace gmf {

struct KvAllCfg {
  static constexpr int MAX_LAYERS = 16;
  static constexpr int A_BYTES = 128 * 128 * 2;            // LN output as 2 swizzle atoms of 128 rows x 64 halfs
  static constexpr int W_BYTES = 128 * 128 * 2;            // to_kv weight of one layer: 128 rows (64 K | 64 V) x 128 k
  static constexpr int OUT_BYTES = 2 * 128 * 64 * 2;       // K image 16 KB | V^T image 16 KB
  static constexpr int X_ROWS = 130;                       // tile rows -1 .. 128 (depthwise conv halo)
  static constexpr int X_BYTES = X_ROWS * 512;
  static constexpr int SMEM = 1024 + 2 * A_BYTES + 2 * W_BYTES + OUT_BYTES + X_BYTES + 256;
  static constexpr int THREADS = 640;
};

struct KvLayer {
  const float* cpe_w;      // [128][3] depthwise conv taps (previous, current, next token)
  const float* cpe_b;      // [128]
  const float* ln_g;
  const float* ln_b;
  const float* w_packed;   // fp16 image of to_kv.weight (gmf_api.cu pack_linear_f16(W, 128, 128, 128))
  __nv_bfloat16* k_out;    // [B][tiles] K tile images (fp16 payload, 128 tokens x 64)
  __nv_bfloat16* v_out;    // [B][tiles] V^T tile images (bf16, 64 x 128 tokens)
};

struct KvAllArgs {
  const float* x;          // [B, L, 128] context tokens
  int L, tiles, pairs, layers;
  KvLayer layer[KvAllCfg::MAX_LAYERS];
#ifdef GMF_FFN_TRACE
  long long* trace;        // development build: clock64 stamps of CTA 3's second tile (tools/kv_trace.py)
#endif
};

__global__ void __launch_bounds__(KvAllCfg::THREADS, 1) kv_proj_all_kernel(const __grid_constant__ KvAllArgs a) {
  using Cfg = KvAllCfg;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                                     // [2]
  uint8_t* sW = sA + 2 * Cfg::A_BYTES;                    // [2]
  uint8_t* sOut = sW + 2 * Cfg::W_BYTES;                  // K image | V^T image
  float* sX = (float*)(sOut + Cfg::OUT_BYTES);            // [130][128] fp32, row s = token row0 - 1 + s
  uint64_t* bars = (uint64_t*)((uint8_t*)sX + Cfg::X_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  const BarArr a_ready{bar0};                      // [2] 256  LayerNorm warps -> MMA
  const BarArr a_free = BarArr{bar0} + 2;          // [2]      MMA retired -> LayerNorm warps (step + 2)
  const BarArr w_full = BarArr{bar0} + 4;          // [2]
  const BarArr w_empty = BarArr{bar0} + 6;         // [2]
  const BarArr acc_full = BarArr{bar0} + 8;        // [2]
  const BarArr acc_free = BarArr{bar0} + 10;       // [2] 256  epilogue warps
  const BarArr x_full = BarArr{bar0} + 12;         //          context tile landed
  const BarArr x_free = BarArr{bar0} + 13;         // 256      LayerNorm warps have read the tile for the last layer
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = a.tiles * a.pairs;
  const int NL = a.layers;
#ifdef GMF_FFN_TRACE
#define KTR(role, idx) do { if (a.trace && blockIdx.x == 3 && it == 1 && lane == 0) a.trace[(role) * 64 + (idx)] = clock64(); } while (0)
#else
#define KTR(role, idx) do {} while (0)
#endif

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_ready[i], 256); mbar_init(&a_free[i], 1); mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], 256);
    }
    mbar_init(x_full, 1); mbar_init(x_free, 256);
    fence_mbar_init();
  }
  if (warp == 16) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 17) {
    // ------------------------------- weight producer -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    int n = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x)
#pragma unroll 1
      for (int l = 0; l < NL; ++l, ++n) {
        const int slot = n & 1;
        if (n >= 2) mbar_wait(&w_empty[slot], ((n >> 1) - 1) & 1);
        mbar_expect_tx_p(&w_full[slot], Cfg::W_BYTES, leader);
        bulk_g2s_p(sW + slot * Cfg::W_BYTES, a.layer[l].w_packed, Cfg::W_BYTES, &w_full[slot], leader);
      }
  } else if (warp == 18) {
    // ------------------------------- context loader: rows row0 - 1 .. row0 + 128 of the pair's token sequence -------------------------------
    int it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
      const int pair = g / a.tiles, tile = g - pair * a.tiles;
      const int row0 = tile * 128;
      if (it > 0) mbar_wait(x_free, (it - 1) & 1);
      const int lo = max(row0 - 1, 0), hi = min(row0 + 129, a.L);        // valid token rows [lo, hi)
      const int s_lo = lo - (row0 - 1), s_hi = hi - (row0 - 1);          // their rows in the shared tile
      for (int s = 0; s < Cfg::X_ROWS; ++s)                              // rows outside the sequence are zeros (conv padding, ragged tail)
        if (s < s_lo || s >= s_hi) *reinterpret_cast<float4*>(sX + s * 128 + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(s_hi - s_lo) * 512u;
        mbar_expect_tx(x_full, bytes);
        bulk_g2s(sX + s_lo * 128, a.x + ((size_t)pair * a.L + lo) * 128, bytes, x_full);
      }
      __syncwarp();
    }
  } else if (warp == 16) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, 128, kFmtF16);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW));
    int n = 0, it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it)
#pragma unroll 1
      for (int l = 0; l < NL; ++l, ++n) {
        const int b = n & 1;
        const uint32_t par = (n >> 1) & 1;
        KTR(1, 3 * l);
        if (n >= 2) mbar_wait(&acc_free[b], par ^ 1);
        KTR(1, 3 * l + 1);
        mbar_wait2(&a_ready[b], par, &w_full[b], par);
        KTR(1, 3 * l + 2);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int at = 0; at < 2; ++at)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_bf16(tm + b * 128, umma_desc_adv(a_desc0, b * Cfg::A_BYTES + at * 16384 + ks * 32),
                          umma_desc_adv(w_desc0, b * Cfg::W_BYTES + at * 16384 + ks * 32), idesc, (at | ks) ? 1u : 0u);
          tc_commit(&w_empty[b]);
          tc_commit(&a_free[b]);
          tc_commit(&acc_full[b]);
        }
        __syncwarp();
      }
  } else if (warp < 8) {
    // ------------------------------- LayerNorm warps: position encoding + LayerNorm of layer l -> A[n & 1] (fp16, swizzled) -------------------------------
    // Warp w owns tile rows 16 w .. 16 w + 15: two batches of 8 rows, lanes across the 128 channels.  They run one step ahead of the MMA.
    const int c4 = lane * 4;
    constexpr int RPW = 8;
    const uint32_t lane_off = (lane >> 4) * 16384 + (lane & 1) * 8;
    const uint32_t piece = (lane & 15) >> 1;
    // per-layer parameters of this lane's 4 channels: 3 KB per layer and 12 layers do not stay in the 28 KB L1, so the next layer's set is
    // fetched while this one is normalised (fetched at the point of use it cost 1.5 k cycles per step)
    float4 pg4, pb4, pcw0, pcw1, pcw2, pcb;
    auto fetch_params = [&](int layer) {
      const KvLayer& ly = a.layer[layer];
      const float* cw = ly.cpe_w + c4 * 3;
      pg4 = __ldg(reinterpret_cast<const float4*>(ly.ln_g + c4));
      pb4 = __ldg(reinterpret_cast<const float4*>(ly.ln_b + c4));
      pcw0 = __ldg(reinterpret_cast<const float4*>(cw)); pcw1 = __ldg(reinterpret_cast<const float4*>(cw + 4)); pcw2 = __ldg(reinterpret_cast<const float4*>(cw + 8));
      pcb = __ldg(reinterpret_cast<const float4*>(ly.cpe_b + c4));
    };
    fetch_params(0);
    int n = 0, it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
#pragma unroll 1
      for (int l = 0; l < NL; ++l, ++n) {
        const int ab = n & 1;
        const float4 g4 = pg4, b4 = pb4, cw0 = pcw0, cw1 = pcw1, cw2 = pcw2, cb = pcb;
        fetch_params(l + 1 == NL ? 0 : l + 1);
        // x + dwconv(x) = w0 x[r-1] + (1 + w1) x[r] + w2 x[r+1] + b, two channels per packed operand; taps of channel c at cw[3 c ..]
        const uint64_t w0a = pack2(cw0.x, cw0.w), w0b = pack2(cw1.z, cw2.y);
        const uint64_t w1a = pack2(1.0f + cw0.y, 1.0f + cw1.x), w1b = pack2(1.0f + cw1.w, 1.0f + cw2.z);
        const uint64_t w2a = pack2(cw0.z, cw1.y), w2b = pack2(cw2.x, cw2.w);
        const uint64_t cba = pack2(cb.x, cb.y), cbb = pack2(cb.z, cb.w);
        const uint64_t g01 = pack2(g4.x, g4.y), g23 = pack2(g4.z, g4.w), b01 = pack2(b4.x, b4.y), b23 = pack2(b4.z, b4.w);
        if (warp == 0) KTR(0, 3 * l);
        if (l == 0) mbar_wait(x_full, it & 1);
        if (n >= 2) mbar_wait(&a_free[ab], ((n >> 1) - 1) & 1);
        if (warp == 0) KTR(0, 3 * l + 1);
        uint8_t* img = sA + ab * Cfg::A_BYTES + lane_off;
#pragma unroll 1
        for (int bt = 0; bt < 2; ++bt) {
          const int rbase = warp * 16 + bt * RPW;                // tile row of rv[1]; shared row = tile row + 1
          float4 rv[RPW + 2];
#pragma unroll
          for (int i = 0; i < RPW + 2; ++i) rv[i] = *reinterpret_cast<const float4*>(sX + (rbase + i) * 128 + c4);
          uint64_t d01[RPW], d23[RPW];
          float mean[RPW], rs[RPW];
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            d01[i] = ffma2(w0a, pack2(rv[i].x, rv[i].y), ffma2(w1a, pack2(rv[i + 1].x, rv[i + 1].y), ffma2(w2a, pack2(rv[i + 2].x, rv[i + 2].y), cba)));
            d23[i] = ffma2(w0b, pack2(rv[i].z, rv[i].w), ffma2(w1b, pack2(rv[i + 1].z, rv[i + 1].w), ffma2(w2b, pack2(rv[i + 2].z, rv[i + 2].w), cbb)));
            float s0, s1;
            unpack2(fadd2(d01[i], d23[i]), s0, s1);
            mean[i] = s0 + s1;
          }
          warp_sum8(mean, lane);
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
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
          for (int i = 0; i < RPW; ++i) {
            const float r_ = rsqrtf(rs[i] * (1.0f / 128.0f) + 1e-5f);
            const uint64_t r2 = pack2(r_, r_);
            float y0, y1, y2, y3;
            unpack2(ffma2(fmul2(d01[i], r2), g01, b01), y0, y1);
            unpack2(ffma2(fmul2(d23[i], r2), g23, b23), y2, y3);
            uint2 pk;
            pk.x = pack_f16(y0, y1);
            pk.y = pack_f16(y2, y3);
            *reinterpret_cast<uint2*>(img + (rbase + i) * 128 + ((piece ^ (uint32_t)i) << 4)) = pk;   // (row & 7) == i
          }
        }
        fence_proxy_async();
        mbar_arrive(&a_ready[ab]);
        if (warp == 0) KTR(0, 3 * l + 2);
        if (l == NL - 1) mbar_arrive(x_free);                    // the loader may overwrite the context tile
      }
    }
  } else if (warp < 16) {
    // ------------------------------- epilogue warps: ACC[n & 1] -> K | V^T tile images -> two bulk stores -------------------------------
    // 4 lane quadrants x 2 halves: half 0 = the 64 K columns (fp16, K-major rows), half 1 = the 64 V columns (bf16, transposed)
    const int e = warp - 8;
    const int q = e & 3, half = e >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    int n = 0, it = 0;
#pragma unroll 1
    for (int g = blockIdx.x; g < total; g += gridDim.x, ++it) {
      const int pair = g / a.tiles, tile = g - pair * a.tiles;
      const bool valid = tile * 128 + r < a.L;
#pragma unroll 1
      for (int l = 0; l < NL; ++l, ++n) {
        const int b = n & 1;
        if (warp == 8) KTR(2, 4 * l);
        mbar_wait(&acc_full[b], (n >> 1) & 1);
        tc_fence_after();
        if (warp == 8) KTR(2, 4 * l + 1);
        uint32_t v[64];
        tmem_ld32(trow + b * 128 + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(trow + b * 128 + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_free[b]);
        if (tid == 256) bulk_wait_read();                        // the previous step's images have left the staging buffer
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 8) KTR(2, 4 * l + 2);
        if (half == 0) {                                         // K: token r, dims 0..63
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 pk;
            pk.x = valid ? pack_f16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])) : 0u;
            pk.y = valid ? pack_f16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])) : 0u;
            pk.z = valid ? pack_f16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])) : 0u;
            pk.w = valid ? pack_f16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])) : 0u;
            *reinterpret_cast<uint4*>(sOut + swz_off(r, j)) = pk;
          }
        } else {                                                 // V^T: rows = dims, 64 tokens per 128-byte row, two token halves
          uint8_t* dst = sOut + 16384 + (r >> 6) * 8192 + (r & 7) * 2;
          const int kchunk = (r & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 64; ++i)
            *reinterpret_cast<__nv_bfloat16*>(dst + swz_off(i, kchunk)) = __float2bfloat16_rn(valid ? __uint_as_float(v[i]) : 0.f);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 256) {
          const size_t tix = (size_t)g * (128 * 64);
          bulk_s2g(a.layer[l].k_out + tix, sOut, 16384);
          bulk_s2g(a.layer[l].v_out + tix, sOut + 16384, 16384);
          bulk_commit();
        }
        if (warp == 8) KTR(2, 4 * l + 3);
      }
    }
    if (tid == 256) bulk_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 256);
}

inline cudaError_t launch_kv_proj_all(const KvAllArgs& a, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = ensure_dyn_smem(kv_proj_all_kernel, KvAllCfg::SMEM, configured)) return e;
  const int total = a.tiles * a.pairs;
  if (total <= 0 || a.layers <= 0) return cudaSuccess;
  kv_proj_all_kernel<<<min(total, device_sm_count()), KvAllCfg::THREADS, KvAllCfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
