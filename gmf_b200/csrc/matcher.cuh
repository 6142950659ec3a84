// Correspondence construction: the step right before the PointDSC forward, NumPy on DataLoader workers in the
// reference (GMF_PointDSC/datasets/ThreeDMatch.py:384-391 mutual / one-way nearest neighbour in descriptor space, :401-402 keypoint
// gather, :411-414 corr_pos = [src | tgt] - mean; same code in datasets/KITTI.py:94-102).
//
//   nn_argmin_kernel     distance(i,j) = sqrt(2 - 2 <a_i, b_j> + 1e-6) evaluated exactly as the reference's fp32 expression, fused with
//                        the row argmin: the Ns x Nt matrix is never written.  128 x 128 register-blocked FP32 tiles (8 x 8 per thread);
//                        the running minimum travels as a 64-bit key (distance bits << 32 | column) so that an unsigned atomicMin
//                        implements np.argmin's first-index tie rule across threads, column chunks and CTAs.
//   corr_build_kernel    mutual test, order-preserving compaction (np.where order), keypoint gather, centring.  One CTA per pair.
#pragma once
#include "common.cuh"

namespace gmf {

constexpr int kNnTile = 128, kNnKb = 32, kNnLd = kNnTile + 4;

// grid (row tiles, column chunks, pairs).  A [pairs][Na][D], B [pairs][Nb][D]; best [pairs][Na] must be preset to all-ones.
// FAST: D <= 32 and D % 4 == 0 (FCGF descriptors): the A tile is loaded once, the next B tile travels through registers (four float4
// per thread, conflict-free transposing stores) while the current one is multiplied.
template <bool FAST>
__global__ void __launch_bounds__(256) nn_argmin_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int Na, int Nb, int D,
                                                        int tiles_per_chunk, unsigned long long* __restrict__ best) {
  __shared__ __align__(16) float As[kNnKb][kNnLd];
  __shared__ __align__(16) float Bs[kNnKb][kNnLd];
  __shared__ unsigned long long sbest[kNnTile];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int pair = blockIdx.z, row0 = blockIdx.x * kNnTile;
  const float* Ap = A + (size_t)pair * Na * D;
  const float* Bp = Bm + (size_t)pair * Nb * D;
  const int nkb = (D + kNnKb - 1) / kNnKb;
  const int ctiles = (Nb + kNnTile - 1) / kNnTile;
  const int ct0 = blockIdx.y * tiles_per_chunk, ct1 = min(ct0 + tiles_per_chunk, ctiles);
  if (tid < kNnTile) sbest[tid] = ~0ull;
  unsigned long long mine[8];
  float top[8];                                          // largest dot product this thread has seen per row
#pragma unroll
  for (int i = 0; i < 8; ++i) { mine[i] = ~0ull; top[i] = -INFINITY; }

  auto load_tile = [&](float (*S)[kNnLd], const float* P, int n, int r0, int k0) {
    // 128 rows x 32 k values, transposed into [k][row]; rows / k past the end read as zero
    for (int e = tid; e < kNnTile * kNnKb; e += 256) {
      const int r = e >> 5, k = e & 31;
      float v = 0.f;
      if (r0 + r < n && k0 + k < D) v = P[(size_t)(r0 + r) * D + k0 + k];
      S[k][r] = v;
    }
  };

  // FAST path staging: element e = tid + 256 i -> row e & 127, k-quad e >> 7
  auto fetch = [&](const float* P, int n, int r0, float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i, r = e & 127, k4 = e >> 7;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < n && k4 * 4 < D) v[i] = *reinterpret_cast<const float4*>(P + (size_t)(r0 + r) * D + k4 * 4);
    }
  };
  auto stash = [&](float (*S)[kNnLd], const float4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i, r = e & 127, k4 = e >> 7;
      S[k4 * 4][r] = v[i].x; S[k4 * 4 + 1][r] = v[i].y; S[k4 * 4 + 2][r] = v[i].z; S[k4 * 4 + 3][r] = v[i].w;
    }
  };
  float4 nxt[4];
  if (FAST) {
    float4 av4[4];
    fetch(Ap, Na, row0, av4);
    stash(As, av4);
    if (ct0 < ct1) fetch(Bp, Nb, ct0 * kNnTile, nxt);
  }

  for (int ct = ct0; ct < ct1; ++ct) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int kb = 0; kb < (FAST ? 1 : nkb); ++kb) {
      __syncthreads();
      if (FAST) {
        stash(Bs, nxt);
      } else {
        if (nkb > 1 || ct == ct0) load_tile(As, Ap, Na, row0, kb * kNnKb);
        load_tile(Bs, Bp, Nb, ct * kNnTile, kb * kNnKb);
      }
      __syncthreads();
      if (FAST && ct + 1 < ct1) fetch(Bp, Nb, (ct + 1) * kNnTile, nxt);      // in flight during the FMAs below
      const int kcount = FAST ? ((D + 3) & ~3) : min(kNnKb, D - kb * kNnKb);   // D = 33 (FPFH): the second block is one step, not 32
#pragma unroll 8
      for (int k = 0; k < kcount; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = ct * kNnTile + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col < Nb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // The distance is a monotone (non-increasing) function of the dot product, and a thread visits its columns in increasing
          // order, so only a dot product above the thread's running maximum can lower the distance or win a tie: the exact
          // expression is evaluated on that rare path only.
          if (acc[i][j] > top[i]) {
            top[i] = acc[i][j];
            // np.sqrt(2 - 2 * dot + 1e-6) in fp32, one rounding per operation (ThreeDMatch.py:384)
            const float d = __fsqrt_rn(__fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, acc[i][j])), 1e-6f));
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)col;
            mine[i] = key < mine[i] ? key : mine[i];
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    atomicMin(&sbest[r], mine[i]);
  }
  __syncthreads();
  if (tid < kNnTile && row0 + tid < Na) atomicMin(&best[(size_t)pair * Na + row0 + tid], sbest[tid]);
}

// One CTA (1024 threads) per pair.  best_src [pairs][Ns], best_tgt [pairs][Nt] (ignored unless mutual).
// Outputs (row capacity Ns per pair): source_idx [Ns], corr [Ns][2], src_sel / tgt_sel [Ns][3], corr_pos [Ns][6], n_corr [pairs].
__global__ void __launch_bounds__(1024) corr_build_kernel(const unsigned long long* __restrict__ best_src, const unsigned long long* __restrict__ best_tgt,
                                                          const float* __restrict__ src_kp, const float* __restrict__ tgt_kp, int Ns, int Nt, int mutual,
                                                          int* __restrict__ source_idx, int* __restrict__ corr, float* __restrict__ src_sel,
                                                          float* __restrict__ tgt_sel, float* __restrict__ corr_pos, int* __restrict__ n_corr) {
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  __shared__ double red[32][6];
  __shared__ float mean_s[6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = blockIdx.x;
  const unsigned long long* bs = best_src + (size_t)pair * Ns;
  const unsigned long long* bt = best_tgt ? best_tgt + (size_t)pair * Nt : nullptr;
  const float* sk = src_kp + (size_t)pair * Ns * 3;
  const float* tk = tgt_kp + (size_t)pair * Nt * 3;
  int* sidx_out = source_idx + (size_t)pair * Ns;
  int* co = corr + (size_t)pair * Ns * 2;
  float* ss = src_sel + (size_t)pair * Ns * 3;
  float* ts = tgt_sel + (size_t)pair * Ns * 3;
  float* cp = corr_pos + (size_t)pair * Ns * 6;
  if (tid == 0) base_s = 0;
  double sum[6] = {0, 0, 0, 0, 0, 0};
  __syncthreads();
  for (int i0 = 0; i0 < Ns; i0 += 1024) {
    const int i = i0 + tid;
    int j = 0;
    bool keep = false;
    if (i < Ns) {
      j = (int)(unsigned)(bs[i] & 0xffffffffull);
      sidx_out[i] = j;
      keep = !mutual || (int)(unsigned)(bt[j] & 0xffffffffull) == i;      // target_idx[source_idx] == arange (ThreeDMatch.py:388)
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int c = warp_cnt[w]; before += w < warp ? c : 0; total += c; }
    const int base = base_s;
    if (keep) {
      const int pos = base + before + __popc(m & ((1u << lane) - 1));
      co[2 * pos] = i; co[2 * pos + 1] = j;
      const float s0 = sk[3 * i], s1 = sk[3 * i + 1], s2 = sk[3 * i + 2], t0 = tk[3 * j], t1 = tk[3 * j + 1], t2 = tk[3 * j + 2];
      ss[3 * pos] = s0; ss[3 * pos + 1] = s1; ss[3 * pos + 2] = s2;
      ts[3 * pos] = t0; ts[3 * pos + 1] = t1; ts[3 * pos + 2] = t2;
      sum[0] += s0; sum[1] += s1; sum[2] += s2; sum[3] += t0; sum[4] += t1; sum[5] += t2;
    }
    __syncthreads();
    if (tid == 0) base_s = base + total;
    __syncthreads();
  }
  const int n = base_s;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double v = sum[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][c] = v;
  }
  __syncthreads();
  if (tid < 6) {
    double v = 0;
    for (int w = 0; w < 32; ++w) v += red[w][tid];
    mean_s[tid] = n > 0 ? (float)(v / n) : 0.f;                            // corr_pos.mean(0) (:413)
  }
  if (tid == 0) n_corr[pair] = n;
  __syncthreads();
  for (int e = tid; e < Ns * 6; e += 1024) {
    const int pos = e / 6, c = e - pos * 6;
    float v = 0.f;
    if (pos < n) v = (c < 3 ? ss[3 * pos + c] : ts[3 * pos + c - 3]) - mean_s[c];
    cp[e] = v;
    if (pos >= n) {                                                        // rows past n_corr are zero-filled
      if (c < 3) { ss[3 * pos + c] = 0.f; ts[3 * pos + c] = 0.f; }
      if (c < 2) co[2 * pos + c] = -1;
    }
  }
}

}  // namespace gmf
