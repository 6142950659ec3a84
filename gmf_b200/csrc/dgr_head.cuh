// DGR bottleneck fusion head: PerceiverIO(depth=0, dim=128, latent_dim=256, cross_heads=1, cross_dim_head=128,
// pe=True).forward  (GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py:187-221; built at model/resunet_new.py:516-525, called
// from ResUNet2.transformer :694-705 with ALL active stride-8 voxels of the batch as one latent sequence [1, M, 256] and the image
// tokens [1, T, 128] as context).
//
// The head is a few hundred MFLOP..GFLOP on M = 512..2048 rows, i.e. launch/latency bound: the design spreads every GEMM over
// (row tile x column block) CTAs instead of maximising per-CTA efficiency, and every operand crosses HBM (L2, really) as the
// ready-to-multiply image of its consumer:
//   rows_to_img_kernel   ConvPosEnc (:105-136) and/or LayerNorm (PreNorm :31-51) of token rows -> tf32 A-operand tile image
//   img_gemm_kernel      one 128 x NB output block per CTA: bulk-async (TMA engine) stage ring -> tcgen05.mma kind::tf32 -> TMEM ->
//                        fused epilogue: bf16 Q / K / V^T tile images for the attention kernel, bias + residual rows, or
//                        bias + GEGLU (:53-56) emitted as the next GEMM's A image
//   attention            sc_attn_v9_kernel with neutral distance features (compat == 1) and Nq != Nk: softmax(q k^T / sqrt(128)) v
#pragma once
#include "common.cuh"
#include "linear_tc.cuh"

namespace gmf {

// One warp per token row.  x [L][C] row-major -> img [tiles][C/32][128 rows x 128 B] (tf32, 128B-swizzled K-major chunks); rows >= L
// of the last tile are written as zeros.  CPE: v = x + dwconv3(x) along the token axis (zero padding); x0_out keeps that residual stream.
template <int C, bool CPE, bool LN>
__global__ void __launch_bounds__(256) rows_to_img_kernel(const float* __restrict__ x, int L, int tiles, const float* __restrict__ cpe_w,
                                                          const float* __restrict__ cpe_b, const float* __restrict__ ln_g,
                                                          const float* __restrict__ ln_b, float* __restrict__ x0_out, float* __restrict__ img) {
  constexpr int G = C / 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= tiles * 128) return;
  const bool valid = r < L;
  float4 v[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int col = g * 128 + lane * 4;
    v[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      v[g] = *reinterpret_cast<const float4*>(x + (size_t)r * C + col);
      if (CPE) {
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f), n = p;
        if (r > 0) p = *reinterpret_cast<const float4*>(x + (size_t)(r - 1) * C + col);
        if (r + 1 < L) n = *reinterpret_cast<const float4*>(x + (size_t)(r + 1) * C + col);
        const float* w = cpe_w + col * 3;                 // depthwise taps [C][1][3]: index [t-1, t, t+1]
        const float4 cb = *reinterpret_cast<const float4*>(cpe_b + col);
        v[g].x += fmaf(w[0], p.x, fmaf(w[1], v[g].x, fmaf(w[2], n.x, cb.x)));
        v[g].y += fmaf(w[3], p.y, fmaf(w[4], v[g].y, fmaf(w[5], n.y, cb.y)));
        v[g].z += fmaf(w[6], p.z, fmaf(w[7], v[g].z, fmaf(w[8], n.z, cb.z)));
        v[g].w += fmaf(w[9], p.w, fmaf(w[10], v[g].w, fmaf(w[11], n.w, cb.w)));
        if (x0_out) *reinterpret_cast<float4*>(x0_out + (size_t)r * C + col) = v[g];
      }
    }
  }
  if (LN) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) s += v[g].x + v[g].y + v[g].z + v[g].w;
    const float mean = warp_sum(s) * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      v[g] = make_float4(v[g].x - mean, v[g].y - mean, v[g].z - mean, v[g].w - mean);
      sq += v[g].x * v[g].x + v[g].y * v[g].y + v[g].z * v[g].z + v[g].w * v[g].w;
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int col = g * 128 + lane * 4;
      const float4 g4 = *reinterpret_cast<const float4*>(ln_g + col), b4 = *reinterpret_cast<const float4*>(ln_b + col);
      v[g] = make_float4(fmaf(v[g].x * rstd, g4.x, b4.x), fmaf(v[g].y * rstd, g4.y, b4.y), fmaf(v[g].z * rstd, g4.z, b4.z),
                         fmaf(v[g].w * rstd, g4.w, b4.w));
    }
  }
  const int tile = r >> 7, rr = r & 127;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    uint8_t* chunk = (uint8_t*)(img + ((size_t)tile * (C / 32) + g * 4 + (lane >> 3)) * 4096);
    *reinterpret_cast<float4*>(chunk + swz_off(rr, lane & 7)) = valid ? to_tf32(v[g]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Combine the split-key partial attention results (sc_attn_v9_kernel<.., SPLIT>) and emit the A-operand image of the to_out GEMM:
//   out = sum_s O_s 2^(ref_s - ref*) / sum_s l_s 2^(ref_s - ref*),  ref* = max_s ref_s.   One warp per query row (128 columns).
__global__ void __launch_bounds__(256) attn_combine_img_kernel(const float* __restrict__ part_o, const float* __restrict__ part_l, int splits, int M,
                                                               int tiles, float* __restrict__ img) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= tiles * 128) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < M) {
    float rmax = -INFINITY;
    for (int s = 0; s < splits; ++s) rmax = fmaxf(rmax, part_l[((size_t)s * M + r) * 2 + 1]);
    float L = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float sc = exp2f(part_l[((size_t)s * M + r) * 2 + 1] - rmax);
      L = fmaf(part_l[((size_t)s * M + r) * 2], sc, L);
      const float4 o = *reinterpret_cast<const float4*>(part_o + ((size_t)s * M + r) * 128 + lane * 4);
      acc.x = fmaf(o.x, sc, acc.x); acc.y = fmaf(o.y, sc, acc.y); acc.z = fmaf(o.z, sc, acc.z); acc.w = fmaf(o.w, sc, acc.w);
    }
    const float inv = 1.0f / L;
    acc = to_tf32(make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
  }
  const int tile = r >> 7, rr = r & 127;
  uint8_t* chunk = (uint8_t*)(img + ((size_t)tile * 4 + (lane >> 3)) * 4096);
  *reinterpret_cast<float4*>(chunk + swz_off(rr, lane & 7)) = acc;
}

// Operands of the seed kNN distance GEMM (models/common.py:53-75 restricted to the seed rows) for the tensor pipe at fp32 accuracy:
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  <a, b> ~ a_hi b_hi + a_hi b_lo + a_lo b_hi  (the dropped lo*lo term is 2^-22
// relative), written as ONE K = 384 product:  A row = [a_hi | a_hi | a_lo],  B row = [b_hi | b_lo | b_hi].
// rows: idx == NULL -> row r of x, else row idx[r] (the seeds); a_side selects the A-row / B-row chunk order.  One warp per row; image
// layout as above, 12 chunks.
__global__ void __launch_bounds__(256) knn_operand_kernel(const float* __restrict__ x, const int* __restrict__ idx, int a_side, int L, int nrows,
                                                          int tiles, float* __restrict__ img) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp, pair = blockIdx.y;
  if (r >= tiles * 128) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < nrows) {
    const int src = idx ? idx[(size_t)pair * nrows + r] : r;
    v = *reinterpret_cast<const float4*>(x + ((size_t)pair * L + src) * 128 + lane * 4);
  }
  const float4 hi = to_tf32(v);
  const float4 lo = to_tf32(make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
  const int tile = r >> 7, rr = r & 127;
  float* base = img + ((size_t)pair * tiles + tile) * (12 * 4096);
  const uint32_t off = swz_off(rr, lane & 7);
  const int ch = lane >> 3;
  *reinterpret_cast<float4*>((uint8_t*)(base + (size_t)ch * 4096) + off) = hi;
  *reinterpret_cast<float4*>((uint8_t*)(base + (size_t)(4 + ch) * 4096) + off) = a_side ? hi : lo;
  *reinterpret_cast<float4*>((uint8_t*)(base + (size_t)(8 + ch) * 4096) + off) = a_side ? lo : hi;
}

// Same split for short descriptor rows of any width D (FCGF 32, FPFH 33): kd = ceil(D / 32) chunks per segment, K = 96 kd.
__global__ void __launch_bounds__(256) desc_operand_kernel(const float* __restrict__ x, int D, int kd, int a_side, int nrows, int tiles,
                                                           float* __restrict__ img) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp, pair = blockIdx.y;
  if (r >= tiles * 128) return;
  const int tile = r >> 7, rr = r & 127;
  float* base = img + ((size_t)pair * tiles + tile) * ((size_t)3 * kd * 4096);
  for (int j = 0; j < kd; ++j) {
    const int c = j * 32 + lane;
    float v = 0.f;
    if (r < nrows && c < D) v = x[((size_t)pair * nrows + r) * D + c];
    const float hi = to_tf32(v), lo = to_tf32(v - hi);
    const uint32_t off = swz_off(rr, lane >> 2) + (lane & 3) * 4;
    *reinterpret_cast<float*>((uint8_t*)(base + (size_t)j * 4096) + off) = hi;
    *reinterpret_cast<float*>((uint8_t*)(base + (size_t)(kd + j) * 4096) + off) = a_side ? hi : lo;
    *reinterpret_cast<float*>((uint8_t*)(base + (size_t)(2 * kd + j) * 4096) + off) = a_side ? lo : hi;
  }
}

enum { DE_RES = 0, DE_GEGLU = 1, DE_QIMG = 2, DE_KVIMG = 3, DE_DIST = 4, DE_COMPAT = 5, DE_ARGMIN = 6, DE_STORE = 7, DE_STORE_X3 = 8 };

struct ImgGemmArgs {
  const float* a_img;      // [tiles][K/32][128 x 32] tf32 chunks (rows_to_img_kernel / DE_GEGLU epilogue)
  const float* w_packed;   // [column blocks][K/32][NB x 32] (pack_linear(W, nout, K, 32, NB))
  int K, L, tiles;
  const float* bias;       // DE_RES: [ld]; DE_GEGLU: [2 * hidden] (value | gate)
  const float* residual;   // DE_RES: [L][ld]
  float* out;              // DE_RES: [L][ld]
  int ld;
  float* out_img;          // DE_GEGLU: [tiles][out_chunks][128 x 32]
  int out_chunks, hidden;
  __nv_bfloat16* t0;       // DE_QIMG: Q tiles; DE_KVIMG: K tiles        [tiles][128 x 128] bf16
  __nv_bfloat16* t1;       // DE_KVIMG: V^T tiles
  // batched use (blockIdx.z = pair): element strides of a_img / w_packed / out between pairs; DE_DIST: valid output columns
  size_t a_pair_stride, w_pair_stride, out_pair_stride;
  int ncols;
  float scale;             // DE_COMPAT: 1 / sigma^2; DE_STORE: out = scale * acc (+ bias[col]) (+ residual[row][col]), any L / ncols / ld (training path)
  unsigned long long* best;   // DE_ARGMIN: [pairs][L] running (distance bits << 32 | column) minima, preset to all-ones
};

template <int NB, int EPI>
struct IgCfg {
  // many small output blocks (seed kNN distances, feature compat, matcher): two stages only, so that two CTAs share an SM and one's
  // epilogue overlaps the other's loads and MMAs; the GEMMs of the DGR head (few CTAs) keep a 4-deep ring
  static constexpr bool SMALL = (EPI == DE_DIST || EPI == DE_COMPAT || EPI == DE_ARGMIN);
  // DE_STORE_X3 (training path, error-compensated "3xTF32" products): every K chunk of an operand image is a (hi, lo) pair of tf32 chunks,
  // hi = tf32(v), lo = tf32(v - hi); a stage holds both pairs and the issuer accumulates hi hi + hi lo + lo hi from ONE load of the stage
  static constexpr bool X3 = (EPI == DE_STORE_X3);
  static constexpr bool STORE = (EPI == DE_STORE || EPI == DE_STORE_X3);
  static constexpr int NSTG = SMALL ? 2 : (X3 ? 3 : 4);
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = NB * 128;
  static constexpr int STAGE = (A_BYTES + B_BYTES) * (X3 ? 2 : 1);
  static constexpr int STG_BYTES = (EPI == DE_RES || EPI == DE_DIST || EPI == DE_COMPAT || STORE) ? 4 * 4096 : 0;
  static constexpr int SMEM = 1024 + NSTG * STAGE + 256 + STG_BYTES;
};

// grid (row tiles, column blocks); 6 warps: 0 = bulk-copy producer, 1 = MMA issuer (+ TMEM owner), 2..5 = epilogue (one TMEM lane
// quadrant each, one thread per accumulator row)
template <int NB, int EPI>
__global__ void __launch_bounds__(192, IgCfg<NB, EPI>::SMALL ? 2 : 1) img_gemm_kernel(const ImgGemmArgs a) {
  using Cfg = IgCfg<NB, EPI>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + Cfg::NSTG * Cfg::STAGE);
  uint64_t* full = bars;            // [4]
  uint64_t* empty = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  float* sStg = (float*)(smem + Cfg::NSTG * Cfg::STAGE + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, cb = blockIdx.y;
  const int nkc = a.K >> 5;

  if (tid == 0) {
    for (int i = 0; i < Cfg::NSTG; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, NB); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    constexpr int PER = Cfg::X3 ? 2 : 1;                 // tf32 chunks per K chunk of an operand image
    const float* asrc = a.a_img + blockIdx.z * a.a_pair_stride + (size_t)tile * nkc * (4096 * PER);
    const float* wsrc = a.w_packed + blockIdx.z * a.w_pair_stride + (size_t)cb * nkc * (NB * 32 * PER);
#pragma unroll 1
    for (int kc = 0; kc < nkc; ++kc) {
      const int st = kc % Cfg::NSTG;
      if (kc >= Cfg::NSTG) mbar_wait(&empty[st], ((kc / Cfg::NSTG) - 1) & 1);
      uint8_t* dst = smem + st * Cfg::STAGE;
      mbar_expect_tx_p(&full[st], Cfg::STAGE, leader);
      bulk_g2s_p(dst, asrc + (size_t)kc * (4096 * PER), Cfg::A_BYTES * PER, &full[st], leader);
      bulk_g2s_p(dst + Cfg::A_BYTES * PER, wsrc + (size_t)kc * (NB * 32 * PER), Cfg::B_BYTES * PER, &full[st], leader);
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, NB, kFmtTF32);
#pragma unroll 1
    for (int kc = 0; kc < nkc; ++kc) {
      const int st = kc % Cfg::NSTG;
      mbar_wait(&full[st], (kc / Cfg::NSTG) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t ad = umma_desc_sw128(smem_u32(smem + st * Cfg::STAGE));
        const uint64_t bd = umma_desc_sw128(smem_u32(smem + st * Cfg::STAGE + Cfg::A_BYTES * (Cfg::X3 ? 2 : 1)));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc_mma_tf32(tm, umma_desc_adv(ad, ks * 32), umma_desc_adv(bd, ks * 32), idesc, (kc > 0 || ks > 0) ? 1u : 0u);
        if (Cfg::X3) {                                    // + hi lo + lo hi (stage = A_hi | A_lo | B_hi | B_lo)
          const uint64_t al = umma_desc_adv(ad, Cfg::A_BYTES), bl = umma_desc_adv(bd, Cfg::B_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc_mma_tf32(tm, umma_desc_adv(ad, ks * 32), umma_desc_adv(bl, ks * 32), idesc, 1u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc_mma_tf32(tm, umma_desc_adv(al, ks * 32), umma_desc_adv(bd, ks * 32), idesc, 1u);
        }
        tc_commit(&empty[st]);
        if (kc == nkc - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int row0 = tile * 128;
    const bool valid = row0 + r < a.L;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    unsigned long long amin = ~0ull;                     // DE_ARGMIN: this row's best key / largest dot product within this column block
    float atop = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      if (EPI == DE_ARGMIN) {
        tmem_ld_wait();
        // descriptor-space nearest neighbour (datasets/ThreeDMatch.py:384-385): distance = sqrt(2 - 2 dot + 1e-6) is monotone in the dot
        // product and the columns are visited in increasing order, so the exact fp32 expression is evaluated only for a new maximum
        const int col0 = cb * 128 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float dp = __uint_as_float(v[i]);
          if (col0 + i < a.ncols && dp > atop) {
            atop = dp;
            const float d = __fsqrt_rn(__fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, dp)), 1e-6f));
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(col0 + i);
            amin = key < amin ? key : amin;
          }
        }
      } else if (EPI == DE_GEGLU) {
        uint32_t gt[32];
        tmem_ld32(trow + 128 + c * 32, gt);
        tmem_ld_wait();
        const int hc = cb * 128 + c * 32;                     // hidden column of this chunk
        const float4* bv = reinterpret_cast<const float4*>(a.bias + hc);
        const float4* bg = reinterpret_cast<const float4*>(a.bias + a.hidden + hc);
        uint8_t* dst = (uint8_t*)(a.out_img + ((size_t)tile * a.out_chunks + cb * 4 + c) * 4096);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b1 = __ldg(bv + j), b2 = __ldg(bg + j);
          float4 o = make_float4((__uint_as_float(v[4 * j]) + b1.x) * gelu_erf(__uint_as_float(gt[4 * j]) + b2.x),
                                 (__uint_as_float(v[4 * j + 1]) + b1.y) * gelu_erf(__uint_as_float(gt[4 * j + 1]) + b2.y),
                                 (__uint_as_float(v[4 * j + 2]) + b1.z) * gelu_erf(__uint_as_float(gt[4 * j + 2]) + b2.z),
                                 (__uint_as_float(v[4 * j + 3]) + b1.w) * gelu_erf(__uint_as_float(gt[4 * j + 3]) + b2.w));
          *reinterpret_cast<float4*>(dst + swz_off(r, j)) = valid ? to_tf32(o) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else if (EPI == DE_DIST || EPI == DE_COMPAT || Cfg::STORE) {
        tmem_ld_wait();
        // DE_DIST: squared feature distance of unit vectors, 2 - 2 <a, b> (models/common.py:64-66), rows = seeds, columns = points
        // DE_COMPAT: feature compatibility clamp(1 - (1 - <a, b>) / sigma^2, 0, 1) with a zero diagonal (models/PointDSC.py:231-234)
        float* stg = sStg + q * 1024;
        const int srow = lane >> 3, sj = lane & 7;
        const int col0 = cb * 128 + c * 32;
        float* obase = a.out + blockIdx.z * a.out_pair_stride + (size_t)(row0 + q * 32) * a.ld + col0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          if (Cfg::STORE) {
            o = make_float4(__uint_as_float(v[4 * j]) * a.scale, __uint_as_float(v[4 * j + 1]) * a.scale, __uint_as_float(v[4 * j + 2]) * a.scale,
                            __uint_as_float(v[4 * j + 3]) * a.scale);
          } else if (EPI == DE_DIST) {
            o = make_float4(fmaf(-2.0f, __uint_as_float(v[4 * j]), 2.0f), fmaf(-2.0f, __uint_as_float(v[4 * j + 1]), 2.0f),
                            fmaf(-2.0f, __uint_as_float(v[4 * j + 2]), 2.0f), fmaf(-2.0f, __uint_as_float(v[4 * j + 3]), 2.0f));
          } else {
            const int grow = row0 + r, gc = col0 + 4 * j;      // this thread's row, first of its four columns
            float e[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) e[t] = (grow == gc + t) ? 0.f : __saturatef(1.0f - (1.0f - __uint_as_float(v[4 * j + t])) * a.scale);
            o = make_float4(e[0], e[1], e[2], e[3]);
          }
          *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (row0 + q * 32 + rw < a.L) {
            float4 o = *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
            float* dst = obase + (size_t)rw * a.ld + sj * 4;
            const int cc = col0 + sj * 4;
            if (Cfg::STORE) {
              float e[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (cc + t < a.ncols) {
                  if (a.bias) e[t] += a.bias[cc + t];
                  if (a.residual) e[t] += a.residual[(size_t)(row0 + q * 32 + rw) * a.ld + cc + t];
                }
              o = make_float4(e[0], e[1], e[2], e[3]);
            }
            if (cc + 3 < a.ncols && (a.ld & 3) == 0 && ((uintptr_t)dst & 15) == 0) *reinterpret_cast<float4*>(dst) = o;   // flat gradient buffers: any offset
            else {
              if (cc < a.ncols) dst[0] = o.x;
              if (cc + 1 < a.ncols) dst[1] = o.y;
              if (cc + 2 < a.ncols) dst[2] = o.z;
              if (cc + 3 < a.ncols) dst[3] = o.w;
            }
          }
        }
        __syncwarp();
      } else if (EPI == DE_RES) {
        tmem_ld_wait();
        // coalesced row-major I/O through a per-warp XOR-swizzled 32 x 32 staging tile
        float* stg = sStg + q * 1024;
        const int srow = lane >> 3, sj = lane & 7;
        const int col0 = cb * 128 + c * 32;
        const size_t gbase = (size_t)(row0 + q * 32) * a.ld + col0;
        float4 rr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row0 + q * 32 + rw < a.L) rr[i] = *reinterpret_cast<const float4*>(a.residual + gbase + (size_t)rw * a.ld + sj * 4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          *reinterpret_cast<float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2)) = rr[i];
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + col0) + j);
          float4* slot = reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2));
          const float4 x = *slot;
          *slot = make_float4(__uint_as_float(v[4 * j]) + bb.x + x.x, __uint_as_float(v[4 * j + 1]) + bb.y + x.y,
                              __uint_as_float(v[4 * j + 2]) + bb.z + x.z, __uint_as_float(v[4 * j + 3]) + bb.w + x.w);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rw = i * 4 + srow;
          if (row0 + q * 32 + rw < a.L)
            *reinterpret_cast<float4*>(a.out + gbase + (size_t)rw * a.ld + sj * 4) =
                *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
        }
        __syncwarp();
      } else {
        tmem_ld_wait();
        // bf16 tile images in the attention kernel's layout (rows past L are zero: their A rows are zero and there is no bias)
        if (EPI == DE_QIMG || cb == 0) {
          uint8_t* img = (uint8_t*)(a.t0 + (size_t)tile * (128 * 128)) + (c >> 1) * 16384;
          const int cc0 = (c & 1) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_f16(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
            pk.y = pack_f16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            pk.z = pack_f16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            pk.w = pack_f16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            *reinterpret_cast<uint4*>(img + swz_off(r, cc0 + j)) = pk;
          }
        } else {
          // V^T: [2 halves of 64 keys][128 dims x 128 B]; the 8 lanes that share a 16-byte piece write it together
          uint8_t* img = (uint8_t*)(a.t1 + (size_t)tile * (128 * 128)) + (r >> 6) * 16384 + (r & 7) * 2;
          const int kchunk = (r & 63) >> 3;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int d = c * 32 + i;
            *reinterpret_cast<__nv_bfloat16*>(img + swz_off(d, kchunk)) = __float2bfloat16_rn(__uint_as_float(v[i]));
          }
        }
      }
    }
    if (EPI == DE_ARGMIN && valid) atomicMin(&a.best[(size_t)blockIdx.z * a.L + row0 + r], amin);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, NB);
}

template <int NB, int EPI>
inline cudaError_t launch_img_gemm(const ImgGemmArgs& a, int col_blocks, cudaStream_t st, int pairs = 1) {
  using Cfg = IgCfg<NB, EPI>;
  static std::atomic<unsigned long long> configured{0};
  auto kern = img_gemm_kernel<NB, EPI>;
  if (cudaError_t e = ensure_dyn_smem(kern, Cfg::SMEM, configured)) return e;
  kern<<<dim3(a.tiles, col_blocks, pairs), 192, Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
