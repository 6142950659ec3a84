// Micro-benchmarks of the per-SM rates the SC-attention kernel design depends on (B200, sm_100a):
//   T1  tcgen05.ld 32x32b.x32 throughput vs. number of reading warps
//   T2  MUFU.EX2 / MUFU.SQRT throughput vs. number of warps
//   T3  the per-key-tile MMA sequence of sc_attn_tc_kernel (8 S + 2 + 2 D2 + 4 PV) issued back to back, cycles per tile
//   T4  T3 running concurrently with T1-style TMEM readers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gmf_b200/csrc tools/ubench/sm_rates.cu -o gpurun_out/sm_rates
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace gmf;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// mode 0: TMEM loads only; mode 1: MMA only; mode 2: both.  nld = tcgen05.ld.x32 per iteration per reader warp.
template <int SHAPE>   // 0: all MMAs of a tile, 1: only S (8 x N=64), 2: only PV (4 x N=128), 3: S as 4 x N=128 (128-key tile)
__global__ void __launch_bounds__(640, 1) k_tmem_mma(int mode, int reader_warps, int iters, int nld, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_end[32];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 32) t_end[tid] = 0;
  if (tid == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // bf16 ~0.0078
  if (warp == 16) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (warp == 17 && (mode == 1 || mode == 2)) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc_s = umma_idesc(128, 64, kFmtBF16), idesc_o = umma_idesc(128, 128, kFmtBF16);
    const uint64_t q_desc = umma_desc_sw128(smem_u32(smem));                 // 32 KB Q
    const uint64_t aq_desc = umma_desc_sw128(smem_u32(smem + 32768));        // 16 KB
    const uint64_t k_desc = umma_desc_sw128(smem_u32(smem + 49152));         // 16 KB K (+ 16 KB for the 128-key variant)
    const uint64_t bd_desc = umma_desc_sw128(smem_u32(smem + 81920));        // 8 KB
    const uint64_t p_desc = umma_desc_sw128(smem_u32(smem + 90112));         // 16 KB
    const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + 106496));        // 16 KB
    for (int it = 0; it < iters; ++it) {
      const int bb = it & 1;
      if (SHAPE == 0 || SHAPE == 1) {
#pragma unroll
        for (int at = 0; at < 2; ++at)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16_p(tmem + bb * 64, umma_desc_adv(q_desc, at * 16384 + ks * 32), umma_desc_adv(k_desc, at * 8192 + ks * 32), idesc_s, (at | ks) ? 1u : 0u, leader);
      }
      if (SHAPE == 3) {
#pragma unroll
        for (int at = 0; at < 2; ++at)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16_p(tmem + bb * 128, umma_desc_adv(q_desc, at * 16384 + ks * 32), umma_desc_adv(k_desc, at * 16384 + ks * 32), idesc_o, (at | ks) ? 1u : 0u, leader);
      }
      if (SHAPE == 0) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) tc_mma_bf16_p(tmem + 128 + bb * 64, umma_desc_adv(aq_desc, ks * 32), umma_desc_adv(bd_desc, ks * 32), idesc_s, ks ? 1u : 0u, leader);
#pragma unroll
        for (int ks = 2; ks < 4; ++ks) tc_mma_bf16_p(tmem + 256 + bb * 64, umma_desc_adv(aq_desc, ks * 32), umma_desc_adv(bd_desc, ks * 32), idesc_s, ks > 2 ? 1u : 0u, leader);
      }
      if (SHAPE == 0 || SHAPE == 2) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc_mma_bf16_p(tmem + 384, umma_desc_adv(p_desc, ks * 32), umma_desc_adv(v_desc, ks * 32), idesc_o, 1u, leader);
      }
    }
    tc_commit_p(&bar[0], leader);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    if (leader) t_end[warp] = clock64();
  } else if (warp < reader_warps && (mode == 0 || mode == 2)) {
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
      for (int l = 0; l < nld; ++l) {
        uint32_t u[32];
        tmem_ld32(tlane + ((it + l) & 7) * 32, u);
        tmem_ld_wait();
        acc += __uint_as_float(u[0]) + __uint_as_float(u[13]) + __uint_as_float(u[31]);
      }
    }
    if (acc == 123.456f) sink[tid] = acc;
    if ((tid & 31) == 0) t_end[warp] = clock64();
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    long long t1 = 0;
    for (int i = 0; i < 32; ++i) t1 = t_end[i] > t1 ? t_end[i] : t1;
    out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

// readers issue U loads before one wait (deeper pipelining)
__global__ void __launch_bounds__(512, 1) k_tmem_deep(int reader_warps, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_end[32];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 32) t_end[tid] = 0;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (warp < reader_warps) {
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
      uint32_t u0[32], u1[32], u2[32];
      tmem_ld32(tlane + 0, u0);
      tmem_ld32(tlane + 128, u1);
      tmem_ld32(tlane + 256, u2);
      tmem_ld_wait();
      acc += __uint_as_float(u0[3]) + __uint_as_float(u1[17]) + __uint_as_float(u2[31]);
    }
    if (acc == 123.456f) sink[tid] = acc;
    if ((tid & 31) == 0) t_end[warp] = clock64();
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    long long t1 = 0;
    for (int i = 0; i < 32; ++i) t1 = t_end[i] > t1 ? t_end[i] : t1;
    out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// T7: packed half-precision exponential: one MUFU op per TWO values?
__global__ void k_ex2_h2(int iters, long long* out, float* sink) {
  unsigned x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0x38003400u + threadIdx.x + i;   // two small halves
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
  }
  __syncthreads();
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= x[i];
  if (s == 0x12345678u) sink[threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>   // 0 ex2, 1 sqrt, 2 ffma
__global__ void k_mufu(int iters, long long* out, float* sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else if (OP == 1) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      else asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) sink[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// T9: does a MUFU warp-instruction get cheaper when most lanes are predicated off?  (The SC softmax could skip sqrt + ex2 for the ~87 % of
// score elements whose compatibility is exactly 0 if it did.)  MODE 0: all lanes; 1: lanes 0-3 only; 2: one lane in 8; 3: per-element
// pseudo-random 1/8 of the lanes; 4: whole warps skip 7 of 8 instructions through a uniform branch (the best any skipping could do).
template <int MODE>
__global__ void k_mufu_pred(int iters, long long* out, float* sink) {
  float x[16];
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      unsigned on;
      if (MODE == 0) on = 1u;
      else if (MODE == 1) on = lane < 4;
      else if (MODE == 2) on = (lane & 7) == 0;
      else if (MODE == 3) on = (((lane * 2654435761u) >> 7) + i * 5 + it) % 8 == 0;
      else on = ((i + it) & 7) == 0;
      if (MODE == 4) { if (on) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); }
      else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ex2.approx.ftz.f32 %0, %0;\n\t}" : "+f"(x[i]) : "r"(on));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456f) sink[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// T10: the arithmetic of one SC softmax tile share (16 score elements per thread: compat from DA / Y, logit, exp2, row max, row sum, bf16
// pack) in a synchronisation-free loop - the XU-pipe ceiling of the softmax body by itself.  POLY of every 4 exponentials are evaluated on
// the FMA pipe with packed fp32x2 arithmetic (Cody-Waite split + degree-3 polynomial).
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t t2) {
  // t in [-126, 126]; 2^t = 2^n * p(f), n = round(t), f = t - n in [-0.5, 0.5]
  const uint64_t magic = pack2(12582912.f, 12582912.f);
  const uint64_t fl = fadd2(t2, magic);
  const uint64_t r = fadd2(fl, pack2(-12582912.f, -12582912.f));
  const uint64_t f = ffma2(r, pack2(-1.f, -1.f), t2);
  uint64_t p = ffma2(pack2(0.0551716648f, 0.0551716648f), f, pack2(0.2426111251f, 0.2426111251f));
  p = ffma2(p, f, pack2(0.6932609677f, 0.6932609677f));
  p = ffma2(p, f, pack2(0.9999280572f, 0.9999280572f));
  float p0, p1, f0, f1;
  unpack2(p, p0, p1); unpack2(fl, f0, f1);
  p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(f0) << 23));
  p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(f1) << 23));
  return pack2(p0, p1);
}
template <int POLY>
__global__ void k_softmax_body(int iters, long long* out, float* sink) {
  float us[16], ua[16], ub[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { us[i] = 0.01f * (threadIdx.x + i); ua[i] = 1.0f + 0.001f * i + 0.01f * threadIdx.x; ub[i] = 0.5f + 0.002f * i; }
  float rmax = -1e30f;
  uint64_t psum2 = pack2(0.f, 0.f);
  uint32_t acc = 0;
  const uint64_t nref2 = pack2(-1.f, -1.f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[8];
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      const uint64_t A2 = pack2(ua[c], ua[c + 1]), Y2 = pack2(ub[c], ub[c + 1]);
      float q0, q1, u0, u1, t0_, t1_;
      unpack2(ffma2(A2, Y2, A2), q0, q1);
      unpack2(fadd2(A2, Y2), u0, u1);
      const float c0 = __saturatef(fmaf(sqrt_approx(fabsf(q0)), 2.f, -u0));
      const float c1 = __saturatef(fmaf(sqrt_approx(fabsf(q1)), 2.f, -u1));
      const uint64_t t2 = ffma2(pack2(us[c], us[c + 1]), pack2(c0, c1), nref2);
      unpack2(t2, t0_, t1_);
      rmax = fmaxf(rmax, fmaxf(t0_, t1_));
      uint64_t p2;
      if ((c & 7) < 2 * POLY) p2 = ex2_poly2(pack2(fmaxf(t0_, -126.f), fmaxf(t1_, -126.f)));
      else p2 = pack2(ex2_approx(t0_), ex2_approx(t1_));
      psum2 = fadd2(psum2, p2);
      float p0, p1;
      unpack2(p2, p0, p1);
      pk[c >> 1] = pack_bf16(p0, p1);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= pk[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) { us[i] += 1e-6f; ua[i] += 1e-7f; }    // keep the loop body from being hoisted
  }
  __syncthreads();
  long long t1 = clock64();
  float ps0, ps1;
  unpack2(psum2, ps0, ps1);
  if (rmax + ps0 + ps1 == 123.456f || acc == 0x12345u) sink[threadIdx.x] = rmax;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// T8: does packing two fp32 into bf16x2 (cvt.rn.bf16x2.f32, SASS F2FP) share the MUFU (XU) pipe?  MODE 0: 16 cvt per iteration;
// MODE 1: 16 ex2 + 8 cvt per iteration (the softmax mix: one pack per two exponentials); MODE 2: 16 ex2 + 8 PRMT-style truncating packs
template <int MODE>
__global__ void k_cvt(int iters, long long* out, float* sink) {
  float x[16];
  unsigned pk[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = 0.001f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        unsigned r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[(i + 1) & 15]));
        pk[i & 7] ^= r;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        unsigned r;
        if (MODE == 1) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
        else asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(x[2 * i])), "r"(__float_as_uint(x[2 * i + 1])));
        pk[i] ^= r;
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
  unsigned u = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) u ^= pk[i];
  if (s == 123.456f || u == 0x12345678u) sink[threadIdx.x] = s + u;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// back-to-back MMAs of one shape: M=128, N, K=16 (bf16), A from shared memory (TS=0) or tensor memory (TS=1)
template <int N, int TS>
__global__ void __launch_bounds__(128, 1) k_mma_rate(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_end;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t idesc = umma_idesc(128, N, kFmtBF16);
    const uint64_t a_desc = umma_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = umma_desc_sw128(smem_u32(smem + 32768));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (TS) tc_mma_bf16_ts_p(tmem + (it & 1) * 256, tmem + 448 + ks * 8, umma_desc_adv(b_desc, (ks >> 2) * (N * 128) + (ks & 3) * 32), idesc, ks ? 1u : 0u, leader);
        else tc_mma_bf16_p(tmem + (it & 1) * 256, umma_desc_adv(a_desc, (ks >> 2) * 16384 + (ks & 3) * 32), umma_desc_adv(b_desc, (ks >> 2) * (N * 128) + (ks & 3) * 32), idesc, ks ? 1u : 0u, leader);
      }
    }
    tc_commit_p(&bar, leader);
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (leader) t_end = clock64();
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) out[0] = t_end - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int TS>
void run_mma_rate(long long* d_out) {
  const int SM = 128 * 1024, iters = 2000;
  long long h;
  CK(cudaFuncSetAttribute(k_mma_rate<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  for (int rep = 0; rep < 2; ++rep) { k_mma_rate<N, TS><<<148, 128, SM>>>(iters, d_out); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
  printf("  M128 N%-3d K16 %s: %.1f clk per MMA (tensor floor %d)\n", N, TS ? "A in TMEM" : "A in smem", (double)h / (iters * 8), N / 2);
}

// T6: bulk-async (TMA engine) L2 -> shared streaming rate: every CTA streams the same `span` bytes (L2 resident) in `chunk`-byte
// copies through a DEPTH-deep ring, like the K/V stream of the attention kernels.
template <int DEPTH>
__global__ void __launch_bounds__(64, 1) k_tma_stream(const uint8_t* src, size_t span, int chunk, int iters, int ctas_per_region, long long* out, int skew = 0) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[DEPTH];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < DEPTH; ++i) mbar_init(&full[i], 1); fence_mbar_init(); }
  __syncthreads();
  long long t0 = clock64();
  if (tid == 0) {
    const uint8_t* base = src + (size_t)(blockIdx.x / ctas_per_region) * span;
    const int per = (int)(span / chunk);
    for (int it = 0; it < iters + DEPTH; ++it) {
      const int s = it % DEPTH;
      if (it >= DEPTH) mbar_wait(&full[s], ((it / DEPTH) - 1) & 1);
      if (it < iters) {
        mbar_expect_tx(&full[s], chunk);
        bulk_g2s(smem + (size_t)s * chunk, base + (size_t)((it + skew * (int)blockIdx.x * 7) % per) * chunk, chunk, &full[s]);
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  long long* d_out; float* d_sink;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_sink, 4096 * 4));
  long long h;
  const int SM = 200 * 1024;
  CK(cudaFuncSetAttribute(k_tmem_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  CK(cudaFuncSetAttribute(k_tmem_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  CK(cudaFuncSetAttribute(k_tmem_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  CK(cudaFuncSetAttribute(k_tmem_mma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
  const int iters = 2000;
  printf("== T1 tcgen05.ld.32x32b.x32 (4 KB per warp-instruction), ld;wait per load\n");
  for (int w : {1, 4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { k_tmem_mma<0><<<148, 640, SM>>>(0, w, iters, 3, d_out, d_sink); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  readers=%2d  %8lld clk  %.1f B/clk/SM  %.1f clk per ld.x32 per warp\n", w, h, (double)w * iters * 3 * 4096 / h, (double)h / (iters * 3));
  }
  printf("== T1b three loads in flight before the wait\n");
  for (int w : {4, 8, 16}) {
    for (int rep = 0; rep < 2; ++rep) { k_tmem_deep<<<148, 512>>>(w, iters, d_out, d_sink); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  readers=%2d  %8lld clk  %.1f B/clk/SM\n", w, h, (double)w * iters * 3 * 4096 / h);
  }
  printf("== T2 MUFU / FFMA throughput (16 independent chains per thread)\n");
  for (int w : {4, 8, 16, 32}) {
    k_mufu<0><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    const double ex2 = (double)w * 32 * 16 * iters / h;
    k_mufu<1><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    const double sq = (double)w * 32 * 16 * iters / h;
    k_mufu<2><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  warps=%2d  ex2 %.1f /clk/SM   sqrt %.1f /clk/SM   ffma %.1f /clk/SM\n", w, ex2, sq, (double)w * 32 * 16 * iters / h);
  }
  for (int w : {4, 8, 16}) {
    k_ex2_h2<<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  warps=%2d  ex2.approx.f16x2: %.1f instr-lanes/clk/SM = %.1f exponentials/clk/SM\n", w, (double)w * 32 * 16 * iters / h, 2.0 * w * 32 * 16 * iters / h);
  }
  printf("== T9 predicated MUFU.EX2: warp-instructions per clk per SM (16 warps; all lanes on = 0.5)\n");
  {
    auto run9 = [&](auto kern, const char* name, double frac) {
      kern<<<148, 512>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
      printf("  %-44s %8lld clk  %.3f issued warp-instr/clk/SM\n", name, h, 16.0 * 16 * iters * frac / h);
    };
    run9(k_mufu_pred<0>, "all 32 lanes on", 1.0);
    run9(k_mufu_pred<1>, "lanes 0-3 on (one quarter-warp pass)", 1.0);
    run9(k_mufu_pred<2>, "one lane in 8 on", 1.0);
    run9(k_mufu_pred<3>, "pseudo-random 1/8 of the lanes on", 1.0);
    run9(k_mufu_pred<4>, "uniform branch skips 7 of 8 instructions", 0.125);
  }
  printf("== T10 SC softmax body without synchronisation (16 warps x 16 elements per thread-iteration): cycles per 128 x 32 tile\n");
  {
    auto run10 = [&](auto kern, const char* name) {
      kern<<<148, 512>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
      // 16 warps x 32 lanes x 16 elements = 8192 elements per iteration = two 128 x 32 tiles
      printf("  %-44s %.1f clk per tile (XU floor 512 with 2 MUFU per element)\n", name, (double)h / iters / 2.0);
    };
    run10(k_softmax_body<0>, "all exponentials on MUFU");
    run10(k_softmax_body<1>, "1 of 4 exponentials on the FMA pipe (fp32x2)");
    run10(k_softmax_body<2>, "2 of 4 exponentials on the FMA pipe (fp32x2)");
    run10(k_softmax_body<3>, "3 of 4 exponentials on the FMA pipe (fp32x2)");
    run10(k_softmax_body<4>, "all exponentials on the FMA pipe (fp32x2)");
  }
  printf("== T8 fp32x2 -> bf16x2 packing vs the MUFU pipe (16 warps)\n");
  {
    const int w = 16;
    k_cvt<0><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  cvt.rn.bf16x2.f32 alone: %.1f /clk/SM\n", (double)w * 32 * 16 * iters / h);
    k_mufu<0><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    const double base = (double)h / iters;
    k_cvt<1><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  16 ex2 per thread-iteration: %.1f clk;  16 ex2 + 8 cvt: %.1f clk", base, (double)h / iters);
    k_cvt<2><<<148, 32 * w>>>(iters, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf(";  16 ex2 + 8 prmt: %.1f clk\n", (double)h / iters);
  }
  printf("== T3 MMA sequence, cycles per key tile (tensor floor: all 640, S 256, PV 256, S128 512 @ 8192 flop/clk)\n");
  {
    k_tmem_mma<0><<<148, 640, SM>>>(1, 0, iters, 0, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost)); printf("  all (8 S + 4 D2 + 4 PV): %.1f clk/tile\n", (double)h / iters);
    k_tmem_mma<1><<<148, 640, SM>>>(1, 0, iters, 0, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost)); printf("  S only (8 x M128 N64 K16): %.1f clk/tile\n", (double)h / iters);
    k_tmem_mma<2><<<148, 640, SM>>>(1, 0, iters, 0, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost)); printf("  PV only (4 x M128 N128 K16): %.1f clk/tile\n", (double)h / iters);
    k_tmem_mma<3><<<148, 640, SM>>>(1, 0, iters, 0, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost)); printf("  S128 only (8 x M128 N128 K16): %.1f clk/tile\n", (double)h / iters);
  }
  printf("== T4 MMA sequence + concurrent TMEM readers (3 x ld.x32 per warp per tile-iteration)\n");
  for (int w : {8, 16}) {
    k_tmem_mma<0><<<148, 640, SM>>>(2, w, iters, 3, d_out, d_sink); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    printf("  readers=%2d: %.1f clk/iteration (max of both sides)\n", w, (double)h / iters);
  }
  {
    printf("== T6 TMA (cp.async.bulk) L2 -> smem streaming, 148 CTAs, 4-deep ring\n");
    const size_t span = 3276800;   // one pair's K + Bd + V^T
    uint8_t* d_src; CK(cudaMalloc(&d_src, span * 8)); CK(cudaMemset(d_src, 1, span * 8));
    long long* d_o; CK(cudaMalloc(&d_o, 148 * 8));
    long long ho[148];
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    CK(cudaFuncSetAttribute(k_tma_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int chunk : {8192, 16384, 32768}) for (int cpr : {148, 37}) {
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) { k_tma_stream<4><<<148, 64, 4 * chunk + 2048>>>(d_src, span, chunk, iters, cpr, d_o); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(ho, d_o, sizeof(ho), cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = ho[i] > mx ? ho[i] : mx;
      const double bpc = 148.0 * iters * chunk / mx;
      printf("  chunk %5d B, %3d CTAs per 3.2 MB region: %.0f B/clk chip-wide, %.1f B/clk/SM  (%.2f TB/s at %d MHz nominal)\n", chunk, cpr, bpc, bpc / 148, bpc * clk_khz * 1e3 / 1e12, clk_khz / 1000);
    }
  }
  {
    printf("== T6b same, but every CTA starts at a different offset of its region (desynchronised streams, no L2 request merging)\n");
    const size_t span = 3276800;
    uint8_t* d_src; CK(cudaMalloc(&d_src, span * 8)); CK(cudaMemset(d_src, 1, span * 8));
    long long* d_o; CK(cudaMalloc(&d_o, 148 * 8));
    long long ho[148];
    for (int chunk : {16384, 32768}) for (int cpr : {148, 37}) {
      const int iters = 4000;
      for (int rep = 0; rep < 2; ++rep) { k_tma_stream<4><<<148, 64, 4 * chunk + 2048>>>(d_src, span, chunk, iters, cpr, d_o, 1); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(ho, d_o, sizeof(ho), cudaMemcpyDeviceToHost));
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = ho[i] > mx ? ho[i] : mx;
      const double bpc = 148.0 * iters * chunk / mx;
      printf("  chunk %5d B, %3d CTAs per 3.2 MB region: %.0f B/clk chip-wide, %.1f B/clk/SM (%.2f TB/s at 1965 MHz)\n", chunk, cpr, bpc, bpc / 148, bpc * 1.965e9 / 1e12);
    }
  }
  printf("== T5 MMA rate by shape and A-operand source\n");
  run_mma_rate<32, 0>(d_out); run_mma_rate<64, 0>(d_out); run_mma_rate<128, 0>(d_out); run_mma_rate<256, 0>(d_out);
  run_mma_rate<32, 1>(d_out); run_mma_rate<48, 1>(d_out); run_mma_rate<64, 1>(d_out); run_mma_rate<128, 1>(d_out); run_mma_rate<256, 1>(d_out);
  return 0;
}
