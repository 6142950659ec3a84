// Token-wise linear layers of the GMF-PointDSC encoder as ONE tcgen05 (kind::tf32) kernel family with fused
// prologues (ConvPosEnc 3-tap stencil, LayerNorm) and epilogues (bias / folded-BN / ReLU / residual / GEGLU /
// bf16 Q,K,V^T tile emission for the flash-attention kernels).
//
//   out[128-row tile, NOUT] = epilogue( prologue(x)[128, K] . W[NOUT, K]^T )
//
// Reference ops replaced (GMF_PointDSC/models): PointCN conv+BN+ReLU PointDSC.py:104-111; projection_q/k/v
// :56-58; fc_message :13-21,65; PreNorm/LayerNorm fusion_layer.py:32-52; to_q/to_kv/to_out :84-94;
// FeedForward+GEGLU :54-69; ConvPosEnc :118-128.
//
// Structure per CTA (one 128-token tile of one pair): 8 worker warps build the A operand (prologue math, written to
// shared memory in the 128B-swizzled K-major image), one control thread streams pre-swizzled weight chunks with
// bulk-async copies (TMA engine) through a 2-deep mbarrier ring and issues tcgen05.mma into TMEM; the workers then
// drain TMEM (tcgen05.ld, one thread per accumulator row) through the fused epilogue.
#pragma once
#include "common.cuh"

namespace gmf {

enum { PRO_NONE = 0, PRO_LN = 1, PRO_CPE_LN = 2, PRO_TILED = 3 };
enum { EPI_BIAS_RELU = 0, EPI_BIAS_RES = 1, EPI_GEGLU_TILED = 2, EPI_QKV_SC = 3, EPI_Q_FUS = 4, EPI_KV_FUS = 5, EPI_BIAS = 6 };

struct LinArgs {
  const float* x;         // PRO_TILED: tiled activation image, else [B, L, K] row-major
  int L;                  // tokens per pair
  int tiles;              // ceil(L / 128)
  const float* ln_g;      // LayerNorm gamma/beta [K]
  const float* ln_b;
  const float* cpe_w;     // depthwise taps [K][3]
  const float* cpe_b;     // [K]
  float* x0_out;          // PRO_CPE_LN: x + dwconv(x) written here (residual stream), may be null
  const float* w_packed;  // weight chunks, swizzled, in issue order
  const float* bias;      // [NOUT] (already BN-folded / pre-scaled by the packer)
  const float* residual;  // EPI_BIAS_RES: [B, L, NOUT]
  float* out;             // fp32 output (row-major [B,L,NOUT] or tiled image for EPI_GEGLU_TILED)
  __nv_bfloat16* t0;      // tiled bf16 outputs (Q / K / V^T)
  __nv_bfloat16* t1;
  __nv_bfloat16* t2;
};

template <int K, int NOUT, int PRO>
struct LinCfg {
  // "LIGHT" kernels (one accumulator block of <= 128 columns) are HBM/latency bound: they are sized for TWO CTAs per SM
  // (<= 113 KB shared memory, <= 256 TMEM columns, <= 113 registers) so that one tile's global loads / epilogue stores overlap
  // the other's MMAs; weights stream through a small ring in 32-float K chunks.
  static constexpr bool LIGHT = NOUT <= 128;
  static constexpr bool STREAM = true;                                   // weights arrive as a stream of chunks
  static constexpr int KCH = LIGHT ? (K >= 128 ? 32 : K) : 64;          // K extent of one weight chunk
  static constexpr int NKC = K / KCH;
  static constexpr int NB = NOUT >= 128 ? 128 : NOUT;                   // MMA N = rows of one weight chunk
  static constexpr int PASSES = NOUT > 512 ? NOUT / 256 : 1;       // GEGLU: 4 passes of (128 value | 128 gate) columns
  static constexpr int NNB = NOUT / NB / PASSES;
  static constexpr int PASS_COLS = NNB * NB;
  static constexpr int TBUF = (PASSES > 1 && 2 * PASS_COLS <= 512) ? 2 : 1;   // double-buffered accumulators: MMAs of pass p+1
  static constexpr int TCOLS_RAW = TBUF * PASS_COLS;                          // overlap the epilogue of pass p
  static constexpr int TMEM_COLS = TCOLS_RAW <= 32 ? 32 : TCOLS_RAW <= 64 ? 64 : TCOLS_RAW <= 128 ? 128 : TCOLS_RAW <= 256 ? 256 : 512;
  static constexpr int NBUF = LIGHT ? (PRO == PRO_TILED ? 3 : (NKC > 1 ? 2 : 1)) : 4;   // weight ring depth
  static constexpr int A_BYTES = 128 * KCH * 4;            // per k-chunk
  static constexpr int A_TOTAL = (PRO == PRO_TILED) ? NBUF * A_BYTES : 128 * K * 4;
  static constexpr int B_BYTES = NB * KCH * 4;
  static constexpr int NSTAGE = PASSES * NNB * NKC;
  // per-warp 32x32 fp32 transpose staging for coalesced row-major epilogue I/O: aliases the idle second weight buffer
  // when the kernel streams a single weight chunk, else lives behind the barriers
  // LIGHT kernels: the A operand is dead once the accumulator is complete -> epilogue staging / tile images alias it
  static constexpr bool STG_ALIAS = LIGHT;
  static constexpr int STG_BYTES = 8 * 32 * 32 * 4;
  static_assert(!LIGHT || A_TOTAL >= STG_BYTES, "staging must fit in the A region");
  static constexpr int SMEM = 1024 + A_TOTAL + NBUF * B_BYTES + 256 + (STG_ALIAS ? 0 : STG_BYTES);
  static constexpr int MIN_CTAS = (LIGHT && SMEM <= 113 * 1024) ? 2 : 1;
};

// exact (erf) GELU, F.gelu default in the reference (fusion_layer.py:57):  gelu(x) = x/2 + |x|/2 (1 - erfc(|x| / sqrt2)) with
//     erfc(|x| / sqrt2) = 2^q(|x|),   q = degree-5 polynomial without constant term (weighted minimax fit of log2 erfc, tools/fit_gelu.py):
// |gelu error| <= 1.3e-6 over all x (q -> -inf for large |x|: 2^q -> 0, no clamp needed), 20 x below the Abramowitz-Stegun 7.1.25 form used
// before, with ONE special-function op (EX2) instead of two (RCP + EX2) and 5 instead of 8 FMA-pipe instructions: the GEGLU arithmetic of the
// fused FFN kernel is shared between the XU pipe and the issue port (profiles/r02_summary.md).
#define GMF_GELU_C1 (-1.1510913f)
#define GMF_GELU_C2 (-4.5925468e-01f)
#define GMF_GELU_C3 (-5.2561242e-02f)
#define GMF_GELU_C4 (7.3975129e-03f)
#define GMF_GELU_C5 (-5.2045897e-04f)
// Evaluated in n = -|x| (odd coefficients change sign) with the factor 1/2 folded into the exponent:  gelu(x) = max(x, 0) + n 2^(q(n) - 1).
__device__ __forceinline__ float gelu_erf(float x) {
  const float n = -fabsf(x);
  float q = fmaf(-GMF_GELU_C5, n, GMF_GELU_C4);
  q = fmaf(q, n, -GMF_GELU_C3);
  q = fmaf(q, n, GMF_GELU_C2);
  q = fmaf(q, n, -GMF_GELU_C1);
  return fmaf(n, ex2_approx(fmaf(q, n, -1.0f)), fmaxf(x, 0.0f));                     // ex2 = erfc(|x| / sqrt2) / 2
}

// GEGLU on two hidden units at once with packed fp32x2 arithmetic: (v + bv) * gelu_erf(g + bg) for both lanes of each 64-bit operand.
// Same formula as gelu_erf; the polynomial, the products and the bias adds issue as one FFMA2 / FMUL2 / FADD2 per pair, -|x|, max(x, 0) and the
// exponentials stay scalar: 9 packed + 6 scalar instructions per pair (the fused FFN kernel is bound by the issue rate of exactly this
// arithmetic: 65 536 hidden activations per 128-token tile).
__device__ __forceinline__ uint64_t geglu2(uint64_t v2, uint64_t bv2, uint64_t g2, uint64_t bg2) {
  const uint64_t x2 = fadd2(g2, bg2);
  float x0, x1;
  unpack2(x2, x0, x1);
  const uint64_t n2 = pack2(-fabsf(x0), -fabsf(x1));
  uint64_t q2 = ffma2(pack2(-GMF_GELU_C5, -GMF_GELU_C5), n2, pack2(GMF_GELU_C4, GMF_GELU_C4));
  q2 = ffma2(q2, n2, pack2(-GMF_GELU_C3, -GMF_GELU_C3));
  q2 = ffma2(q2, n2, pack2(GMF_GELU_C2, GMF_GELU_C2));
  q2 = ffma2(q2, n2, pack2(-GMF_GELU_C1, -GMF_GELU_C1));
  float a0, a1;
  unpack2(ffma2(q2, n2, pack2(-1.0f, -1.0f)), a0, a1);
  const uint64_t gelu = ffma2(n2, pack2(ex2_approx(a0), ex2_approx(a1)), pack2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
  return fmul2(fadd2(v2, bv2), gelu);
}

template <int K, int NOUT, int PRO, int EPI>
__global__ void __launch_bounds__(288, LinCfg<K, NOUT, PRO>::MIN_CTAS) linear_tc_kernel(const LinArgs a) {
  using Cfg = LinCfg<K, NOUT, PRO>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS)
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::A_TOTAL;
  uint64_t* bars = (uint64_t*)(sB + Cfg::NBUF * Cfg::B_BYTES);
  uint64_t* full = bars;          // [4] weights (+A chunk) landed
  uint64_t* mma_done = bars + 4;  // [4] MMAs that read stage buffers retired
  uint64_t* a_ready = bars + 8;   // workers finished the A operand
  uint64_t* acc_full = bars + 9;  // [2] accumulators of a pass complete (per TMEM buffer)
  uint64_t* tmem_free = bars + 11; // [2] workers drained the TMEM buffer of a pass
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);
  float* sStg = Cfg::STG_ALIAS ? (float*)sA : (float*)(sB + Cfg::NBUF * Cfg::B_BYTES + 256);
  uint8_t* sImg = Cfg::LIGHT ? sA : sB;                  // bf16 tile images are assembled here once the MMAs have retired

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x, pair = blockIdx.y;
  const int row0 = tile * 128;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&mma_done[i], 1); }
    mbar_init(a_ready, 256);
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1); mbar_init(&tmem_free[0], 256); mbar_init(&tmem_free[1], 256);
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------- control warp: one elected lane issues TMA + MMA ---------------
    // The stage loop is fully unrolled for short schedules so that every descriptor is "uniform base + constant" (run-time ring
    // indices cost ~8 extra SASS instructions per MMA in vector->uniform register moves).
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = umma_idesc(128, Cfg::NB, kFmtTF32);
    const uint8_t* wsrc = (const uint8_t*)a.w_packed;
    const uint8_t* asrc = (const uint8_t*)a.x + (size_t)(pair * a.tiles + tile) * (size_t)(Cfg::NKC * Cfg::A_BYTES);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t b_desc0 = umma_desc_sw128(smem_u32(sB));
    auto issue_load = [&](int it) {
      const int buf = it % Cfg::NBUF;
      if (PRO == PRO_TILED) {
        mbar_expect_tx_p(&full[buf], Cfg::B_BYTES + Cfg::A_BYTES, leader);
        bulk_g2s_p(sA + buf * Cfg::A_BYTES, asrc + (size_t)(it % Cfg::NKC) * Cfg::A_BYTES, Cfg::A_BYTES, &full[buf], leader);
      } else {
        mbar_expect_tx_p(&full[buf], Cfg::B_BYTES, leader);
      }
      bulk_g2s_p(sB + buf * Cfg::B_BYTES, wsrc + (size_t)it * Cfg::B_BYTES, Cfg::B_BYTES, &full[buf], leader);
    };
    constexpr int PRE = Cfg::NBUF > 1 ? Cfg::NBUF - 1 : 1;      // chunks in flight ahead of the MMAs
#pragma unroll
    for (int it = 0; it < PRE && it < Cfg::NSTAGE; ++it) issue_load(it);
    auto stage = [&](int it) {
      const int buf = it % Cfg::NBUF;
      const int nxt = it + PRE;
      if (nxt < Cfg::NSTAGE) {
        if (nxt >= Cfg::NBUF) mbar_wait(&mma_done[nxt % Cfg::NBUF], ((nxt / Cfg::NBUF) - 1) & 1);   // MMAs of the previous occupant retired
        issue_load(nxt);
      }
      const int pass = it / (Cfg::NNB * Cfg::NKC);
      const int nbi = (it / Cfg::NKC) % Cfg::NNB;
      const int kc = it % Cfg::NKC;
      if (it == 0 && PRO != PRO_TILED) mbar_wait(a_ready, 0);
      const int tb = pass % Cfg::TBUF;
      if (pass >= Cfg::TBUF && nbi == 0 && kc == 0) mbar_wait(&tmem_free[tb], ((pass / Cfg::TBUF) - 1) & 1);
      mbar_wait(&full[buf], (it / Cfg::NBUF) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t ad = umma_desc_adv(a_desc0, PRO == PRO_TILED ? buf * Cfg::A_BYTES : kc * (Cfg::KCH / 32) * 16384);
        const uint64_t bd = umma_desc_adv(b_desc0, buf * Cfg::B_BYTES);
#pragma unroll
        for (int at = 0; at < Cfg::KCH / 32; ++at) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_tf32(tm + tb * Cfg::PASS_COLS + nbi * Cfg::NB, umma_desc_adv(ad, at * 16384 + ks * 32), umma_desc_adv(bd, at * (Cfg::NB * 128) + ks * 32),
                        idesc, (kc > 0 || at > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&mma_done[buf]);
        if (nbi == Cfg::NNB - 1 && kc == Cfg::NKC - 1) tc_commit(&acc_full[tb]);
      }
      __syncwarp();
    };
    if constexpr (Cfg::NSTAGE <= 16) {
#pragma unroll
      for (int it = 0; it < Cfg::NSTAGE; ++it) stage(it);
    } else {
#pragma unroll 1
      for (int it = 0; it < Cfg::NSTAGE; ++it) stage(it);
    }
  } else {
    // ------------------------------- workers: A operand prologue -------------------------------
    if (PRO != PRO_TILED) {
      if (K == 128) {
        const int c4 = lane * 4;  // this lane's 4 channels
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 w0 = b4, w1 = b4, w2 = b4, cb = b4;
        if (PRO == PRO_LN || PRO == PRO_CPE_LN) {
          g4 = *reinterpret_cast<const float4*>(a.ln_g + c4);
          b4 = *reinterpret_cast<const float4*>(a.ln_b + c4);
        }
        if (PRO == PRO_CPE_LN) {
          const float* w = a.cpe_w + c4 * 3;
          w0 = make_float4(w[0], w[3], w[6], w[9]);
          w1 = make_float4(w[1], w[4], w[7], w[10]);
          w2 = make_float4(w[2], w[5], w[8], w[11]);
          cb = *reinterpret_cast<const float4*>(a.cpe_b + c4);
        }
        const float* xp = a.x + (size_t)pair * a.L * K;
        // warp w owns rows [16w, 16w+16): all global loads (plus the two CPE halo rows) are issued before any use so
        // that 16-18 independent 512-byte row reads are in flight per warp
        constexpr int RPW = 16;
        const int rbase = warp * RPW;
        float4 rv[RPW + 2];
#pragma unroll
        for (int i = 0; i < RPW + 2; ++i) {
          const int gr = row0 + rbase + i - 1;
          const bool want = (PRO == PRO_CPE_LN) ? true : (i >= 1 && i <= RPW);
          rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (want && gr >= 0 && gr < a.L) rv[i] = *reinterpret_cast<const float4*>(xp + (size_t)gr * K + c4);
        }
        // Rows are processed without per-row branches (rows past L are zeros and are never stored), in three sweeps, so that the
        // 16 shuffle-reduction chains of a warp interleave instead of running back to back (~150 cycles each).
        float4 xv[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          float4 v = rv[i + 1];
          if (PRO == PRO_CPE_LN) {
            const float4 pv = rv[i], nv = rv[i + 2];
            v.x += fmaf(w0.x, pv.x, fmaf(w1.x, v.x, fmaf(w2.x, nv.x, cb.x)));
            v.y += fmaf(w0.y, pv.y, fmaf(w1.y, v.y, fmaf(w2.y, nv.y, cb.y)));
            v.z += fmaf(w0.z, pv.z, fmaf(w1.z, v.z, fmaf(w2.z, nv.z, cb.z)));
            v.w += fmaf(w0.w, pv.w, fmaf(w1.w, v.w, fmaf(w2.w, nv.w, cb.w)));
            const int gr = row0 + rbase + i;
            if (a.x0_out && gr < a.L) *reinterpret_cast<float4*>(a.x0_out + ((size_t)pair * a.L + gr) * K + c4) = v;
          }
          xv[i] = v;
        }
        if (PRO == PRO_LN || PRO == PRO_CPE_LN) {
          float mean[RPW], rs[RPW];
#pragma unroll
          for (int i = 0; i < RPW; ++i) mean[i] = xv[i].x + xv[i].y + xv[i].z + xv[i].w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < RPW; ++i) mean[i] += __shfl_xor_sync(0xffffffffu, mean[i], o);
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            mean[i] *= (1.0f / 128.0f);
            const float dx = xv[i].x - mean[i], dy = xv[i].y - mean[i], dz = xv[i].z - mean[i], dw = xv[i].w - mean[i];
            xv[i] = make_float4(dx, dy, dz, dw);
            rs[i] = dx * dx + dy * dy + dz * dz + dw * dw;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < RPW; ++i) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            const float r_ = rsqrtf(rs[i] * (1.0f / 128.0f) + 1e-5f);
            xv[i] = make_float4(fmaf(xv[i].x * r_, g4.x, b4.x), fmaf(xv[i].y * r_, g4.y, b4.y), fmaf(xv[i].z * r_, g4.z, b4.z), fmaf(xv[i].w * r_, g4.w, b4.w));
          }
        }
#pragma unroll
        for (int i = 0; i < RPW; ++i)
          *reinterpret_cast<float4*>(sA + (lane >> 3) * 16384 + swz_off(rbase + i, lane & 7)) = to_tf32(xv[i]);
      } else {  // K == 64: half a warp per row
        const float* xp = a.x + (size_t)pair * a.L * K;
        const int ch = lane & 15;
        float4 rv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int gr = row0 + warp * 16 + i * 2 + (lane >> 4);
          rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gr < a.L) rv[i] = *reinterpret_cast<const float4*>(xp + (size_t)gr * K + ch * 4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = warp * 16 + i * 2 + (lane >> 4);
          *reinterpret_cast<float4*>(sA + (ch >> 3) * 16384 + swz_off(r, ch & 7)) = to_tf32(rv[i]);
        }
      }
      fence_proxy_async();
      mbar_arrive(a_ready);
    }

    // ------------------------------- workers: epilogue -------------------------------
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const int gr = row0 + r;
    const bool valid = gr < a.L;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const size_t grow = (size_t)pair * a.L + gr;
#pragma unroll 1
    for (int pass = 0; pass < Cfg::PASSES; ++pass) {
      const int tb = pass % Cfg::TBUF;
      const uint32_t tbase = trow + tb * Cfg::PASS_COLS;
      mbar_wait(&acc_full[tb], (pass / Cfg::TBUF) & 1);
      tc_fence_after();
      constexpr int NCHUNK = (EPI == EPI_GEGLU_TILED) ? 4 : Cfg::PASS_COLS / 32;
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        uint32_t v[32];
        tmem_ld32(tbase + c * 32, v);
        if (EPI == EPI_GEGLU_TILED) {
          uint32_t gt[32];
          tmem_ld32(tbase + 128 + c * 32, gt);
          tmem_ld_wait();
          const int oc0 = pass * 128 + c * 32;       // output (value) column
          const float4* bv = reinterpret_cast<const float4*>(a.bias + oc0);
          const float4* bg = reinterpret_cast<const float4*>(a.bias + 512 + oc0);
          float o[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b1 = __ldg(bv + i), b2 = __ldg(bg + i);
            o[4 * i] = (__uint_as_float(v[4 * i]) + b1.x) * gelu_erf(__uint_as_float(gt[4 * i]) + b2.x);
            o[4 * i + 1] = (__uint_as_float(v[4 * i + 1]) + b1.y) * gelu_erf(__uint_as_float(gt[4 * i + 1]) + b2.y);
            o[4 * i + 2] = (__uint_as_float(v[4 * i + 2]) + b1.z) * gelu_erf(__uint_as_float(gt[4 * i + 2]) + b2.z);
            o[4 * i + 3] = (__uint_as_float(v[4 * i + 3]) + b1.w) * gelu_erf(__uint_as_float(gt[4 * i + 3]) + b2.w);
          }
          // tiled image for the next GEMM (K=512 in 8 chunks of 64 floats = 2 swizzle atoms): the 32 KB chunk image is
          // assembled in shared memory by all 8 warps (round 0: chunks c=0,1 -> k-chunk 2*pass, round 1: c=2,3 -> 2*pass+1)
          // and leaves as ONE bulk-async store; the wait for the previous store's smem read overlaps the GELU math above.
          if (tid == 0) bulk_wait_read();                            // the previous chunk image has left shared memory
          asm volatile("bar.sync 1, 256;" ::: "memory");
          {
            uint8_t* dst = (uint8_t*)sStg + (c & 1) * 16384;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
              if (valid) ov = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);   // tensor core truncates to tf32
              *reinterpret_cast<float4*>(dst + swz_off(r, j)) = ov;
            }
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, 256;" ::: "memory");             // image complete
          if (tid == 0) {
            bulk_s2g((uint8_t*)a.out + ((size_t)(pair * a.tiles + tile) * 8 + (oc0 >> 6)) * 32768, sStg, 32768);
            bulk_commit();                                           // drained lazily: overlaps the next chunk's GELU math
          }
        } else {
          tmem_ld_wait();
          const int col0 = c * 32;
          if (EPI == EPI_BIAS_RELU || EPI == EPI_BIAS_RES || EPI == EPI_BIAS) {
            // coalesced I/O through a per-warp XOR-swizzled 32x32 staging tile: global accesses touch 4 rows x 128 B per
            // instruction instead of 32 rows x 16 B
            float* stg = sStg + warp * 1024;
            const int srow = lane >> 3, sj = lane & 7;
            const size_t gbase = ((size_t)pair * a.L + row0 + q * 32) * NOUT + col0;
            if (EPI == EPI_BIAS_RES) {
              float4 rr[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rw = i * 4 + srow;
                rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + q * 32 + rw < a.L) rr[i] = *reinterpret_cast<const float4*>(a.residual + gbase + (size_t)rw * NOUT + sj * 4);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rw = i * 4 + srow;
                *reinterpret_cast<float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2)) = rr[i];
              }
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = *reinterpret_cast<const float4*>(a.bias + col0 + 4 * j);
              float4 o = make_float4(__uint_as_float(v[4 * j]) + bb.x, __uint_as_float(v[4 * j + 1]) + bb.y,
                                     __uint_as_float(v[4 * j + 2]) + bb.z, __uint_as_float(v[4 * j + 3]) + bb.w);
              if (EPI == EPI_BIAS_RELU) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
              }
              float4* slot = reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2));
              if (EPI == EPI_BIAS_RES) {
                const float4 rr = *slot;
                o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
              }
              *slot = o;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rw = i * 4 + srow;
              if (row0 + q * 32 + rw < a.L)
                *reinterpret_cast<float4*>(a.out + gbase + (size_t)rw * NOUT + sj * 4) =
                    *reinterpret_cast<const float4*>(stg + rw * 32 + ((sj ^ (rw & 7)) << 2));
            }
            __syncwarp();
          } else {
            // bf16 tile emission for the attention kernels
            int which, dcol0;   // which: 0 = Q-like (row-major tile), 1 = K-like, 2 = V (transposed tile)
            int drows;          // head dim
            if (EPI == EPI_QKV_SC) { which = col0 >> 7; dcol0 = col0 & 127; drows = 128; }
            else if (EPI == EPI_Q_FUS) { which = 0; dcol0 = col0; drows = 64; }
            else { which = 1 + (col0 >> 6); dcol0 = col0 & 63; drows = 64; }
            float o[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (a.bias) b4 = __ldg(reinterpret_cast<const float4*>(a.bias + col0) + i);
              o[4 * i] = valid ? __uint_as_float(v[4 * i]) + b4.x : 0.f;
              o[4 * i + 1] = valid ? __uint_as_float(v[4 * i + 1]) + b4.y : 0.f;
              o[4 * i + 2] = valid ? __uint_as_float(v[4 * i + 2]) + b4.z : 0.f;
              o[4 * i + 3] = valid ? __uint_as_float(v[4 * i + 3]) + b4.w : 0.f;
            }
            // the bf16 tile images are assembled in shared memory (the weight ring is dead once acc_full fired) and leave
            // the SM as one bulk-async store per image instead of scattered 2..16-byte global stores
            const uint32_t img_bytes = 128u * drows * 2u;
            uint8_t* img = sImg + which * img_bytes;
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
              uint8_t* dst = img + (r >> 6) * (drows * 128) + (r & 7) * 2;
              const int kchunk = (r & 63) >> 3;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int d = dcol0 + i;
                *reinterpret_cast<__nv_bfloat16*>(dst + swz_off(d, kchunk)) = __float2bfloat16_rn(o[i]);
              }
            }
          }
        }
      }
      if (EPI == EPI_QKV_SC || EPI == EPI_Q_FUS || EPI == EPI_KV_FUS) {
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");                 // all 8 worker warps finished their image parts
        if (tid == 0) {
          const uint32_t drows = (EPI == EPI_QKV_SC) ? 128u : 64u;
          const uint32_t img_bytes = 128u * drows * 2u;
          const size_t tix = (size_t)(pair * a.tiles + tile) * (128 * drows);
          if (EPI == EPI_QKV_SC || EPI == EPI_Q_FUS) bulk_s2g(a.t0 + tix, sImg, img_bytes);
          if (EPI == EPI_QKV_SC || EPI == EPI_KV_FUS) {
            bulk_s2g(a.t1 + tix, sImg + img_bytes, img_bytes);
            bulk_s2g(a.t2 + tix, sImg + 2 * img_bytes, img_bytes);
          }
          bulk_commit_wait_read();
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_free[tb]);
    }
  }
  if (EPI == EPI_GEGLU_TILED && tid == 0) bulk_wait_read();         // last chunk image must leave smem before exit
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

template <int K, int NOUT, int PRO, int EPI>
inline cudaError_t launch_linear(const LinArgs& a, int pairs, cudaStream_t st) {
  using Cfg = LinCfg<K, NOUT, PRO>;
  static std::atomic<unsigned long long> configured{0};
  auto kern = linear_tc_kernel<K, NOUT, PRO, EPI>;
  if (cudaError_t e = ensure_dyn_smem(kern, Cfg::SMEM, configured)) return e;
  kern<<<dim3(a.tiles, pairs), 288, Cfg::SMEM, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gmf
