// C-ABI entry points (include/gmf_b200.h), weight packing and the launch schedule of the GMF-PointDSC forward.
#include "../../include/gmf_b200.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "linear_tc.cuh"
#include "sc_attn_v9.cuh"
#include "fus_attn_v2.cuh"
#include "ffn_fused.cuh"
#include "kv_proj_all.cuh"
#include "pcn_qkv.cuh"
#include "tail.cuh"
#include "dgr_head.cuh"
#include "dgr_train.cuh"
#include "pdsc_train.cuh"
#include "matcher.cuh"
#include "sm_baseline.cuh"
#include "se3_refine.cuh"

using namespace gmf;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

// optional per-launch CUDA-event profiler (bench.py's roofline leg); categories are listed in include/gmf_b200.h
struct Prof {
  bool on = false;
  std::vector<cudaEvent_t> ev;     // pairs (start, stop)
  std::vector<int> cat;
  size_t used = 0;                 // events handed out
};
thread_local Prof* g_prof = nullptr;
struct ProfScope {
  cudaStream_t st; cudaEvent_t stop = nullptr;
  ProfScope(int category, cudaStream_t s) : st(s) {
    Prof* p = g_prof;
    if (!p || !p->on) return;
    if (p->used + 2 > p->ev.size()) {
      const size_t old = p->ev.size();
      p->ev.resize(old + 512);
      for (size_t i = old; i < p->ev.size(); ++i) cudaEventCreate(&p->ev[i]);
    }
    cudaEventRecord(p->ev[p->used], st);
    stop = p->ev[p->used + 1];
    p->cat.push_back(category);
    p->used += 2;
  }
  ~ProfScope() { if (stop) cudaEventRecord(stop, st); }
};
enum { CAT_QKV = 0, CAT_QFUS, CAT_KVFUS, CAT_FFN, CAT_ATTN_FUS, CAT_ATTN_SC, CAT_PREP, CAT_CLASSIFY, CAT_SEEDS, CAT_KNN, CAT_SPECTRAL,
       CAT_SCORE, CAT_OTHER, CAT_COUNT };
static_assert(CAT_COUNT == GMF_PROFILE_CATEGORIES, "profile categories out of sync with include/gmf_b200.h");

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return GMF_ERR_CUDA;
}
#define CU(x)                                     \
  do {                                            \
    cudaError_t e_ = (x);                         \
    if (e_ != cudaSuccess) return fail_cuda(e_, #x); \
  } while (0)
#define LAUNCHED()                                          \
  do {                                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);     \
    cudaError_t e_ = cudaGetLastError();                    \
    if (e_ != cudaSuccess) return fail_cuda(e_, "kernel launch"); \
  } while (0)
#define TRY(x)            \
  do {                    \
    int r_ = (x);         \
    if (r_ != 0) return r_; \
  } while (0)

constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------------
// state_dict table (mirror of gmf_b200/weights.py::hot_path_spec)
// ------------------------------------------------------------------------------------------------
struct Spec { std::string name; int64_t numel; };

void fusion_spec(std::vector<Spec>& s, const std::string& p, bool pe) {
  if (pe) {
    s.push_back({p + "cpe.proj_q.weight", 128 * 3}); s.push_back({p + "cpe.proj_q.bias", 128});
    s.push_back({p + "cpe.proj_content.weight", 128 * 3}); s.push_back({p + "cpe.proj_content.bias", 128});
  }
  const std::string a = p + "cross_attend_blocks.0.", f = p + "cross_attend_blocks.1.";
  s.push_back({a + "norm.weight", 128}); s.push_back({a + "norm.bias", 128});
  s.push_back({a + "norm_context.weight", 128}); s.push_back({a + "norm_context.bias", 128});
  s.push_back({a + "fn.to_q.weight", 64 * 128}); s.push_back({a + "fn.to_kv.weight", 128 * 128});
  s.push_back({a + "fn.to_out.weight", 128 * 64}); s.push_back({a + "fn.to_out.bias", 128});
  s.push_back({f + "norm.weight", 128}); s.push_back({f + "norm.bias", 128});
  s.push_back({f + "fn.net.0.weight", 1024 * 128}); s.push_back({f + "fn.net.0.bias", 1024});
  s.push_back({f + "fn.net.2.weight", 128 * 512}); s.push_back({f + "fn.net.2.bias", 128});
}
void bn_spec(std::vector<Spec>& s, const std::string& p, int ch) {
  s.push_back({p + "weight", ch}); s.push_back({p + "bias", ch});
  s.push_back({p + "running_mean", ch}); s.push_back({p + "running_var", ch});
}
std::vector<Spec> build_spec(int L) {
  std::vector<Spec> s;
  s.push_back({"sigma", 1}); s.push_back({"sigma_spat", 1});
  s.push_back({"encoder.layer0.weight", 128 * 6}); s.push_back({"encoder.layer0.bias", 128});
  fusion_spec(s, "encoder.fusion_layer_1.", false);
  for (int i = 0; i < L; ++i) {
    const std::string p = "encoder.blocks.PointCN_layer_" + std::to_string(i) + ".";
    s.push_back({p + "0.weight", 128 * 128}); s.push_back({p + "0.bias", 128});
    bn_spec(s, p + "1.", 128);
    const std::string n = "encoder.blocks.NonLocal_layer_" + std::to_string(i) + ".";
    s.push_back({n + "fc_message.0.weight", 64 * 128}); s.push_back({n + "fc_message.0.bias", 64});
    bn_spec(s, n + "fc_message.1.", 64);
    s.push_back({n + "fc_message.3.weight", 64 * 64}); s.push_back({n + "fc_message.3.bias", 64});
    bn_spec(s, n + "fc_message.4.", 64);
    s.push_back({n + "fc_message.6.weight", 128 * 64}); s.push_back({n + "fc_message.6.bias", 128});
    for (const char* q : {"q", "k", "v"}) {
      s.push_back({n + "projection_" + q + ".weight", 128 * 128}); s.push_back({n + "projection_" + q + ".bias", 128});
    }
    fusion_spec(s, n + "fusion_layer_2.", true);
  }
  s.push_back({"classification.0.weight", 32 * 128}); s.push_back({"classification.0.bias", 32});
  s.push_back({"classification.2.weight", 32 * 32}); s.push_back({"classification.2.bias", 32});
  s.push_back({"classification.4.weight", 32}); s.push_back({"classification.4.bias", 1});
  return s;
}

// ------------------------------------------------------------------------------------------------
// host-side weight packing
// ------------------------------------------------------------------------------------------------
float tf32_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

struct Blob {
  std::vector<float> h;
  size_t push(const float* p, size_t n) {   // 256-byte aligned offsets (bulk copies need 16 B)
    size_t off = (h.size() + 63) & ~(size_t)63;
    h.resize(off + n);
    memcpy(h.data() + off, p, n * sizeof(float));
    return off;
  }
  size_t push(const std::vector<float>& v) { return push(v.data(), v.size()); }
};

// W: [nout][k] row-major.  Emits the weight chunk images in the linear kernel's issue order: for each block of `nb`
// TMEM columns (rowmap gives the source row of every column) and each k-chunk of `kch`: [nb rows x kch floats], as
// kch/32 swizzle atoms of nb x 128 B.
std::vector<float> pack_linear(const std::vector<float>& W, int nout, int k, int kch, int nb, const std::vector<int>* rowmap = nullptr) {
  std::vector<float> out((size_t)nout * k, 0.f);
  size_t chunk = 0;
  for (int blk = 0; blk < nout / nb; ++blk)
    for (int kc = 0; kc < k / kch; ++kc, ++chunk) {
      uint8_t* img = (uint8_t*)(out.data() + chunk * (size_t)nb * kch);
      for (int n = 0; n < nb; ++n) {
        const int srow = rowmap ? (*rowmap)[blk * nb + n] : blk * nb + n;
        for (int kk = 0; kk < kch; ++kk) {
          const int atom = kk >> 5, c16 = (kk & 31) >> 2, within = kk & 3;
          float* dst = (float*)(img + (size_t)atom * nb * 128 + swz_off(n, c16)) + within;
          *dst = tf32_round(W[(size_t)srow * k + kc * kch + kk]);
        }
      }
    }
  return out;
}

// fp16 variant for kind::f16 B operands: per block of `nb` rows the whole K extent as k/64 swizzle atoms of nb x 128 B (64 halfs per row).
// Returned as a float vector (two halfs per element) so that it travels in the same weight blob.
std::vector<float> pack_linear_f16(const std::vector<float>& W, int nout, int k, int nb, const std::vector<int>* rowmap = nullptr) {
  std::vector<float> out((size_t)nout * k / 2, 0.f);
  for (int blk = 0; blk < nout / nb; ++blk) {
    uint8_t* img = (uint8_t*)out.data() + (size_t)blk * nb * k * 2;
    for (int n = 0; n < nb; ++n) {
      const int srow = rowmap ? (*rowmap)[blk * nb + n] : blk * nb + n;
      for (int kk = 0; kk < k; ++kk) {
        const int atom = kk >> 6, piece = (kk & 63) >> 3, within = kk & 7;
        __half h = __float2half_rn(W[(size_t)srow * k + kk]);
        memcpy(img + (size_t)atom * nb * 128 + swz_off(n, piece) + within * 2, &h, 2);
      }
    }
  }
  return out;
}

// Error-compensated fp16 operands (pcn_qkv.cuh): per block of `nb` rows and k-chunk of 64: {hi atom | lo atom}, each nb x 128 B (64 halfs per row),
// w = hi + lo.  Returned as floats (two halfs per element) so that it travels in the same weight blob; same byte count as fp32.
std::vector<float> pack_linear_split16(const std::vector<float>& W, int nout, int k, int nb) {
  std::vector<float> out((size_t)nout * k, 0.f);
  size_t chunk = 0;
  for (int blk = 0; blk < nout / nb; ++blk)
    for (int kc = 0; kc < k / 64; ++kc, ++chunk) {
      uint8_t* img = (uint8_t*)(out.data() + chunk * (size_t)nb * 64);
      for (int n = 0; n < nb; ++n)
        for (int kk = 0; kk < 64; ++kk) {
          const float w = W[(size_t)(blk * nb + n) * k + kc * 64 + kk];
          const __half hi = __float2half_rn(std::min(std::max(w, -65504.f), 65504.f));
          const __half lo = __float2half_rn(w - __half2float(hi));
          const size_t off = swz_off(n, kk >> 3) + (kk & 7) * 2;
          memcpy(img + off, &hi, 2);
          memcpy(img + (size_t)nb * 128 + off, &lo, 2);
        }
    }
  return out;
}

struct FusionW {
  bool pe = false;
  const float *cpe_q_w = nullptr, *cpe_q_b = nullptr, *cpe_c_w = nullptr, *cpe_c_b = nullptr;
  const float *lnq_g, *lnq_b, *lnc_g, *lnc_b, *lnf_g, *lnf_b;
  const float *wq, *wkv, *wo, *bo, *b1, *b2;
  const float *w1f, *w2f;   // fused-FFN packing (8 passes of 64 hidden columns)
  const float* wkv16 = nullptr;   // to_kv.weight as one fp16 image (kv_proj_all.cuh)
  const float* wq16 = nullptr;    // scaled to_q.weight as one fp16 image (query projection fused into fus_attn_v2.cuh)
};
struct LayerW {
  const float *pcn_b, *qkv_w, *qkv_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b, *fc3_w, *fc3_b;
  const float* pq_w;        // PointCN + QKV weight chunks for the chained kernel (pcn_qkv.cuh)
  FusionW f2;
};

}  // namespace

struct gmf_ctx {
  int device = 0;
  gmf_config cfg{};
  bool loaded = false;
  float* blob = nullptr;
  float sigma = 1.f, sigma_spat = 0.1f;
  const float *l0_w = nullptr, *l0_b = nullptr;
  FusionW f1;
  std::vector<LayerW> layers;
  ClsWeights cls{};
  int chunk_pairs = 64;
  int match_impl = 0;       // 0 = FP32 register-blocked matcher; 1 = tensor pipe (error-compensated tf32, argmin in the GEMM epilogue): correct but
                            // slower at K = 96 (6.3 vs 3.4 ms for 64 pairs x 5000^2): one 128 x 128 block per CTA is all fixed latency
  // staging for the host-buffer entry point
  uint8_t* stage = nullptr;
  size_t stage_bytes = 0;
  // side stream: the fusion attention of an encoder layer does not depend on its SC attention (both only need the PointCN / QKV outputs), so
  // the two run concurrently and the second kernel's CTAs fill the SMs the first one's last, partial wave leaves idle
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t copy_stream = nullptr;   // uploads of the next chunk overlap the current chunk's kernels
  cudaEvent_t copy_ev[64] = {};
  cudaEvent_t start_ev = nullptr;
  cudaEvent_t in_done[2] = {};          // forward that read input staging set 0 / 1 has finished (double-buffered across host calls)
  int in_set = 0;
  int in_shape[3] = {0, 0, 0};          // (B, N, T) of the previous host call: a different shape moves the staging layout
  Prof prof;
};

namespace {

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct Work {
  float *kpts; float4 *src4, *tgt4;
  float *feat_img;   // tf32 tile image of the layer output (handed from the FFN kernel to the next layer's PointCN/QKV kernel)
  float *imgfeat, *featA, *feat1, *x0, *x1, *m2;
  __nv_bfloat16 *qf, *kf, *vtf, *qs, *ks, *vts, *aq, *bd;
  __nv_bfloat16 *kf_all, *vtf_all;   // [layers] context K / V^T tiles of every encoder layer (projected up front on a side stream)
  size_t kv_stride;
  float *knn_a, *knn_b;   // split-tf32 operand images of the seed kNN distance GEMM (seed rows / all points)
  int *perm;          // points of each pair in descending-x order (windowed NMS)
  float *normed, *conf, *key, *seed_w, *seed_trans, *pre_refine, *dist, *seedM;
  int *seeds, *knn, *counts, *best;
  unsigned* pair_mask;
};
struct Bump {
  uint8_t* base; size_t off = 0;
  template <class T> T* take(size_t n) {
    off = (off + 1023) & ~(size_t)1023;
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

size_t carve(Work& w, uint8_t* base, int B, int N, int T, int S, int k, int layers) {
  Bump b{base};
  const int nt = cdiv(N, 128), tt = cdiv(std::max(T, 1), 128);
  const size_t Lm = (size_t)std::max(N, T), tm = (size_t)std::max(nt, tt);
  w.kpts = b.take<float>((size_t)B * nt * 128 * 8);
  w.src4 = b.take<float4>((size_t)B * N); w.tgt4 = b.take<float4>((size_t)B * N);
  w.imgfeat = b.take<float>((size_t)B * std::max(T, 1) * 128);
  w.featA = b.take<float>((size_t)B * N * 128); w.feat1 = b.take<float>((size_t)B * N * 128);
  w.feat_img = b.take<float>((size_t)B * nt * 128 * 128);
  w.x0 = b.take<float>(B * Lm * 128); w.x1 = b.take<float>(B * Lm * 128);
  w.m2 = b.take<float>((size_t)B * N * 64);
  w.qf = b.take<__nv_bfloat16>(B * tm * 128 * 64);
  w.kf = b.take<__nv_bfloat16>(B * tm * 128 * 64); w.vtf = b.take<__nv_bfloat16>(B * tm * 128 * 64);
  w.kv_stride = (size_t)B * tt * 128 * 64;
  w.kf_all = b.take<__nv_bfloat16>(w.kv_stride * layers); w.vtf_all = b.take<__nv_bfloat16>(w.kv_stride * layers);
  w.qs = b.take<__nv_bfloat16>((size_t)B * nt * 128 * 128);
  w.ks = b.take<__nv_bfloat16>((size_t)B * nt * 128 * 128); w.vts = b.take<__nv_bfloat16>((size_t)B * nt * 128 * 128);
  w.aq = b.take<__nv_bfloat16>((size_t)B * nt * 128 * 64); w.bd = b.take<__nv_bfloat16>((size_t)B * nt * 128 * 64);
  w.normed = b.take<float>((size_t)B * N * 128); w.conf = b.take<float>((size_t)B * N); w.key = b.take<float>((size_t)B * N); w.perm = b.take<int>((size_t)B * N);
  w.seed_w = b.take<float>((size_t)B * S * k); w.seed_trans = b.take<float>((size_t)B * S * 16);
  w.pre_refine = b.take<float>((size_t)B * 16);
  w.dist = b.take<float>((size_t)B * S * N);
  w.knn_a = b.take<float>((size_t)B * cdiv(S, 128) * 12 * 4096); w.knn_b = b.take<float>((size_t)B * nt * 12 * 4096);
  w.seedM = b.take<float>((size_t)B * S * 1600);
  w.seeds = b.take<int>((size_t)B * S); w.knn = b.take<int>((size_t)B * S * k); w.counts = b.take<int>((size_t)B * S);
  w.best = b.take<int>(B); w.pair_mask = b.take<unsigned>(B);
  return b.off + 1024;
}

inline int num_seeds(const gmf_ctx* c, int N) { return (int)((double)N * c->cfg.ratio); }   // int(num_corr * self.ratio) in double
inline int eff_k(const gmf_ctx* c, int N) { return std::min(c->cfg.k, N - 1); }

int check_ws(const gmf_ctx* ctx, Work& w, void* ws, size_t bytes, int B, int N, int T) {
  if (!ws) return fail(GMF_ERR_INVALID, "workspace is NULL");
  const int S = std::max(num_seeds(ctx, N), 1), k = std::max(eff_k(ctx, N), 1);
  const size_t need = carve(w, nullptr, B, N, T, S, k, ctx->cfg.num_layers);
  if (bytes < need) return fail(GMF_ERR_STATE, "workspace too small: need " + std::to_string(need) + " bytes");
  uint8_t* base = (uint8_t*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  carve(w, base, B, N, T, S, k, ctx->cfg.num_layers);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// launch schedule
// ------------------------------------------------------------------------------------------------
template <int K, int NOUT, int PRO, int EPI>
int run_linear(const LinArgs& a, int pairs, cudaStream_t st, int category = CAT_OTHER) {
  ProfScope ps(category, st);
  cudaError_t e = launch_linear<K, NOUT, PRO, EPI>(a, pairs, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "linear_tc launch");
  return 0;
}

LinArgs lin(const float* x, int L, const float* w, const float* bias) {
  LinArgs a{};
  a.x = x; a.L = L; a.tiles = cdiv(L, 128); a.w_packed = w; a.bias = bias;
  return a;
}

// context: (CPE) -> LN_ctx -> to_kv => bf16 K tiles, V^T tiles
int run_fusion_kv(const FusionW& f, const float* ctxk, int B, int Lk, __nv_bfloat16* kf, __nv_bfloat16* vtf, cudaStream_t st) {
  LinArgs a = lin(ctxk, Lk, f.wkv, nullptr);
  a.ln_g = f.lnc_g; a.ln_b = f.lnc_b; a.t1 = kf; a.t2 = vtf;
  if (f.pe) {
    a.cpe_w = f.cpe_c_w; a.cpe_b = f.cpe_c_b; a.x0_out = nullptr;
    TRY((run_linear<128, 128, PRO_CPE_LN, EPI_KV_FUS>(a, B, st, CAT_KVFUS)));
  } else {
    TRY((run_linear<128, 128, PRO_LN, EPI_KV_FUS>(a, B, st, CAT_KVFUS)));
  }
  return 0;
}
int run_fusion_core(const gmf_ctx* ctx, const FusionW& f, Work& w, const float* xq, const __nv_bfloat16* kf, const __nv_bfloat16* vtf, int B, int Lq, int Lk,
                    float* out, cudaStream_t st, const float* tail_m2, const float* tail_w3, const float* tail_b3, float* out_img = nullptr);

// context K / V^T of every encoder layer from the Fusion-1 output, one persistent kernel (kv_proj_all.cuh)
int run_fusion_kv_all(const gmf_ctx* ctx, Work& w, const float* ctxk, int B, int Lk, cudaStream_t st);

// FusionLayer.forward (fusion_layer.py:172-201), everything on one stream
int run_fusion(const gmf_ctx* ctx, const FusionW& f, Work& w, const float* xq, const float* ctxk, int B, int Lq, int Lk, float* out, cudaStream_t st,
               const float* tail_m2 = nullptr, const float* tail_w3 = nullptr, const float* tail_b3 = nullptr) {
  TRY(run_fusion_kv(f, ctxk, B, Lk, w.kf, w.vtf, st));
  return run_fusion_core(ctx, f, w, xq, w.kf, w.vtf, B, Lq, Lk, out, st, tail_m2, tail_w3, tail_b3);
}

// query side (position encoding, LayerNorm, to_q) + attention (+ to_out + residual) in one kernel: writes w.x1
int run_fusion_attn(const FusionW& f, Work& w, const float* xq, const __nv_bfloat16* kf, const __nv_bfloat16* vtf, int B, int Lq, int Lk, cudaStream_t st) {
  {
    AttnArgs a{};
    a.k_t = kf; a.vt_t = vtf; a.out = nullptr;
    a.Lq = Lq; a.Lk = Lk; a.q_tiles = cdiv(Lq, 128); a.k_tiles = cdiv(Lk, 128);
    a.wo_packed = f.wo; a.bo = f.bo; a.xout = w.x1;                                                   // to_out + bias + residual fused
    a.xq = xq; a.cpe_w = f.pe ? f.cpe_q_w : nullptr; a.cpe_b = f.pe ? f.cpe_q_b : nullptr; a.lnq_g = f.lnq_g; a.lnq_b = f.lnq_b; a.wq16 = f.wq16;
    a.x0 = f.pe ? w.x0 : nullptr; a.resid = f.pe ? w.x0 : xq;                                         // residual stream: x + dwconv(x) when there is a position encoding
#ifdef GMF_FFN_TRACE
    static int n_fa = 0;
    const bool fa_trace = (++n_fa == 20);
    if (fa_trace) { cudaMalloc(&a.trace, 64 * 8); cudaMemsetAsync(a.trace, 0, 64 * 8, st); }
#endif
    ProfScope ps(CAT_ATTN_FUS, st);
    cudaError_t e = launch_fus_attn_v2(a, B, st);
#ifdef GMF_FFN_TRACE
    if (fa_trace) {
      cudaStreamSynchronize(st);
      long long h[64];
      cudaMemcpy(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "fus_attn trace: setup %lld | LN done %lld | Q in TMEM %lld | main loop done %lld | epilogue done %lld\n", h[1] - h[0], h[2] - h[0], h[3] - h[0],
              h[4] - h[0], h[5] - h[0]);
    }
#endif
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "fusion attention launch");
  }
  return 0;
}

// the GEGLU feed-forward block of the fusion layer on w.x1 (+ the fused NonLocalBlock tail)
int run_fusion_ffn(const FusionW& f, Work& w, int B, int Lq, float* out, cudaStream_t st, const float* tail_m2, const float* tail_w3, const float* tail_b3,
                   float* out_img) {
  {   // LN -> Linear(128,1024) -> GEGLU -> Linear(512,128) + bias + residual, hidden activation on chip
    FfnArgs a{};
    a.x = w.x1; a.L = Lq; a.tiles = cdiv(Lq, 128); a.ln_g = f.lnf_g; a.ln_b = f.lnf_b;
    a.w1_packed = f.w1f; a.b1 = f.b1; a.w2_packed = f.w2f; a.b2 = f.b2; a.out = out;
    a.m2 = tail_m2; a.w3_packed = tail_w3; a.b3 = tail_b3; a.out_img = out_img;
#ifdef GMF_FFN_TRACE
    static int n_ffn = 0;
    const bool do_trace = (++n_ffn == 20);
    if (do_trace) { cudaMalloc(&a.trace, 512 * 8); cudaMemsetAsync(a.trace, 0, 512 * 8, st); }
#endif
    ProfScope ps(CAT_FFN, st);
    cudaError_t e = launch_ffn_fused(a, B, st);
#ifdef GMF_FFN_TRACE
    if (do_trace) {
      cudaStreamSynchronize(st);
      long long h[512];
      cudaMemcpy(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost);
      if (FILE* f = fopen("gpurun_out/ffn_trace.bin", "wb")) { fwrite(h, 1, sizeof(h), f); fclose(f); }
    }
#endif
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "ffn_fused launch");
  }
  return 0;
}

int run_fusion_core(const gmf_ctx* ctx, const FusionW& f, Work& w, const float* xq, const __nv_bfloat16* kf, const __nv_bfloat16* vtf, int B, int Lq, int Lk,
                    float* out, cudaStream_t st, const float* tail_m2, const float* tail_w3, const float* tail_b3, float* out_img) {
  (void)ctx;
  TRY(run_fusion_attn(f, w, xq, kf, vtf, B, Lq, Lk, st));
  return run_fusion_ffn(f, w, B, Lq, out, st, tail_m2, tail_w3, tail_b3, out_img);
}

int run_prep(const gmf_ctx* ctx, Work& w, const float* src, const float* tgt, int B, int N, cudaStream_t st, float sigma_d = 0.f) {
  ProfScope ps(CAT_PREP, st);
  const int Np = cdiv(N, 128) * 128;
  prep_points_kernel<<<B, 256, 0, st>>>(src, tgt, N, Np, w.kpts, w.src4, w.tgt4);
  LAUNCHED();
  dist_feature_scaled_kernel<<<dim3(Np / 128, B), 128, 0, st>>>(w.kpts, Np, 1.0f / (sigma_d > 0.f ? sigma_d : ctx->sigma_spat), w.aq, w.bd);
  LAUNCHED();
  return 0;
}

cudaError_t launch_sc_any(const gmf_ctx* ctx, const ScAttnArgs& sa, int B, cudaStream_t st) {
  (void)ctx;
  return launch_sc_attn_v9<0, 2>(sa, B, st);
}

// Q/K/V projections + SC-guided attention (PointDSC.py:56-64); feat1 = PointCN output
int run_sc_attention(const gmf_ctx* ctx, const LayerW& lw, Work& w, const float* feat1, int B, int N, float* msg, cudaStream_t st, float* fused_m2 = nullptr,
                     bool qkv_done = false) {
  if (!qkv_done) {
    LinArgs a = lin(feat1, N, lw.qkv_w, lw.qkv_b);
    a.t0 = w.qs; a.t1 = w.ks; a.t2 = w.vts;
    TRY((run_linear<128, 384, PRO_NONE, EPI_QKV_SC>(a, B, st, CAT_QKV)));
  }
  ScAttnArgs sa{};
  sa.q_t = w.qs; sa.k_t = w.ks; sa.vt_t = w.vts; sa.aq_t = w.aq; sa.bd_t = w.bd; sa.out = msg;
  sa.N = N; sa.tiles = cdiv(N, 128);
  if (fused_m2) { sa.fc1_w = lw.fc1_w; sa.fc1_b = lw.fc1_b; sa.fc2_w = lw.fc2_w; sa.fc2_b = lw.fc2_b; sa.m2_out = fused_m2; }
#ifdef GMF_SC_TRACE
  static int n_sc = 0;
  const bool do_trace = (++n_sc == 30) && B >= 2;
  if (do_trace) { cudaMalloc(&sa.trace, 7 * 64 * 8 * 8); cudaMemsetAsync(sa.trace, 0, 7 * 64 * 8 * 8, st); }
#endif
  ProfScope ps(CAT_ATTN_SC, st);
  cudaError_t e = launch_sc_any(ctx, sa, B, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "sc_attn_v9 launch");
#ifdef GMF_SC_TRACE
  if (do_trace) {
    cudaStreamSynchronize(st);
    static long long h[7 * 64 * 8];
    cudaMemcpy(h, sa.trace, sizeof(h), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen("gpurun_out/sc_trace.bin", "wb")) { fwrite(h, 1, sizeof(h), f); fclose(f); }
  }
#endif
  return 0;
}

// PointCN_layer_i + NonLocal_layer_i (PointDSC.py:140-142, 40-74)
int run_fusion_kv_all(const gmf_ctx* ctx, Work& w, const float* ctxk, int B, int Lk, cudaStream_t st) {
  const int L = ctx->cfg.num_layers;
  if (L > KvAllCfg::MAX_LAYERS) return fail(GMF_ERR_INVALID, "run_fusion_kv_all: too many layers");
  KvAllArgs a{};
  a.x = ctxk; a.L = Lk; a.tiles = cdiv(Lk, 128); a.pairs = B; a.layers = L;
  for (int li = 0; li < L; ++li) {
    const FusionW& f = ctx->layers[li].f2;
    a.layer[li] = KvLayer{f.cpe_c_w, f.cpe_c_b, f.lnc_g, f.lnc_b, f.wkv16, w.kf_all + (size_t)li * w.kv_stride, w.vtf_all + (size_t)li * w.kv_stride};
  }
#ifdef GMF_FFN_TRACE
  static int n_kv = 0;
  const bool do_trace = (++n_kv == 3);
  if (do_trace) { cudaMalloc(&a.trace, 512 * 8); cudaMemsetAsync(a.trace, 0, 512 * 8, st); }
#endif
  ProfScope ps(CAT_KVFUS, st);
  cudaError_t e = launch_kv_proj_all(a, st);
#ifdef GMF_FFN_TRACE
  if (do_trace) {
    cudaStreamSynchronize(st);
    long long h[512];
    cudaMemcpy(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen("gpurun_out/kv_trace.bin", "wb")) { fwrite(h, 1, sizeof(h), f); fclose(f); }
  }
#endif
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "kv_proj_all launch");
  return 0;
}

// kv_ready: this layer's context K / V^T already sit in w.kf_all / w.vtf_all (run_fusion_kv_all)
int run_encoder_layer(const gmf_ctx* ctx, int li, Work& w, const float* feat_in, const float* image_feat, int B, int N, int T,
                      float* feat_out, cudaStream_t st, bool in_img = false, bool out_img = false, bool kv_ready = false) {
  const LayerW& lw = ctx->layers[li];
  {   // PointCN (conv + folded BN + ReLU) chained with the Q/K/V projections
    PcnQkvArgs a{};
    a.x = feat_in; a.x_img = in_img ? w.feat_img : nullptr; a.L = N; a.tiles = cdiv(N, 128); a.w_packed = lw.pq_w; a.pcn_bias = lw.pcn_b; a.qkv_bias = lw.qkv_b;
    a.feat1 = w.feat1; a.tq = w.qs; a.tk = w.ks; a.tv = w.vts;
#ifdef GMF_FFN_TRACE
    static int n_pcn = 0;
    const bool pcn_trace = (++n_pcn == 20);
    if (pcn_trace) { cudaMalloc(&a.trace, 64 * 8); cudaMemsetAsync(a.trace, 0, 64 * 8, st); }
#endif
    ProfScope ps(CAT_QKV, st);
    cudaError_t e = launch_pcn_qkv(a, B, st);
#ifdef GMF_FFN_TRACE
    if (pcn_trace) {
      cudaStreamSynchronize(st);
      long long h[64];
      cudaMemcpy(h, a.trace, sizeof(h), cudaMemcpyDeviceToHost);
      const long long t0 = h[0];
      fprintf(stderr, "pcn trace worker: start 0 | acc0 got %lld | f1 back %lld |", h[1] - t0, h[2] - t0);
      for (int w = 0; w < 3; ++w) fprintf(stderr, " blk%d wait %lld got %lld image done %lld |", w, h[3 + 3 * w] - t0, h[4 + 3 * w] - t0, h[5 + 3 * w] - t0);
      fprintf(stderr, "\npcn trace mma: tile start %lld img got %lld |", h[32 + 16] - t0, h[32 + 17] - t0);
      for (int c = 0; c < 8; ++c) fprintf(stderr, " c%d start %lld W got %lld |", c, h[32 + 2 * c] - t0, h[32 + 2 * c + 1] - t0);
      fprintf(stderr, "\n");
    }
#endif
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "pcn_qkv launch");
  }
  // fc_message.6(m2) + fusion_layer_2 output (PointDSC.py:73) are folded into the fused FFN kernel's tail; the query side of fusion_layer_2
  // (position encoding, LayerNorm, to_q) is the prologue of the attention kernel
  const __nv_bfloat16 *kf = w.kf_all + (size_t)li * w.kv_stride, *vtf = w.vtf_all + (size_t)li * w.kv_stride;
  if (kv_ready && ctx->side && !ctx->prof.on) {                 // per-launch profiling times every kernel alone: single-stream schedule
    // SC attention (writes m2) on the main stream, fusion attention (writes x1) next to it on the side stream; the FFN joins them
    CU(cudaEventRecord(ctx->ev_fork, st));
    CU(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    TRY(run_sc_attention(ctx, lw, w, w.feat1, B, N, nullptr, st, w.m2, true));
    TRY(run_fusion_attn(lw.f2, w, w.feat1, kf, vtf, B, N, T, ctx->side));
    CU(cudaEventRecord(ctx->ev_join, ctx->side));
    CU(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    return run_fusion_ffn(lw.f2, w, B, N, feat_out, st, w.m2, lw.fc3_w, lw.fc3_b, out_img ? w.feat_img : nullptr);
  }
  // SC attention with fc_message.0/.3 as the kernel's tail: writes m2
  TRY(run_sc_attention(ctx, lw, w, w.feat1, B, N, nullptr, st, w.m2, true));
  if (kv_ready)
    return run_fusion_core(ctx, lw.f2, w, w.feat1, w.kf_all + (size_t)li * w.kv_stride, w.vtf_all + (size_t)li * w.kv_stride, B, N, T, feat_out, st, w.m2,
                           lw.fc3_w, lw.fc3_b, out_img ? w.feat_img : nullptr);
  return run_fusion(ctx, lw.f2, w, w.feat1, image_feat, B, N, T, feat_out, st, w.m2, lw.fc3_w, lw.fc3_b);
}

int run_classify(const gmf_ctx* ctx, const float* feat, long long rows, float* normed, float* conf, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  CU(ensure_dyn_smem(classify_normalize_kernel, kClsSmem, configured));
  ProfScope ps(CAT_CLASSIFY, st);
  classify_normalize_kernel<<<(unsigned)std::min<long long>((rows + 63) / 64, 148 * 3 * 2), 256, kClsSmem, st>>>(feat, rows, ctx->cls, normed, conf);
  LAUNCHED();
  return 0;
}

int run_pick_seeds(const gmf_ctx* ctx, Work& w, const float* conf, int B, int N, int S, int use_nms, int* seeds, cudaStream_t st) {
  int np2 = 1;
  while (np2 < N) np2 <<= 1;
  if (np2 > 16384) return fail(GMF_ERR_INVALID, "pick_seeds supports N <= 16384");
  ProfScope ps(CAT_SEEDS, st);
  static std::atomic<unsigned long long> configured{0};
  CU(ensure_dyn_smem(topk_sort_kernel, 16384 * 8, configured));
  const bool windowed = use_nms && N >= 1024;              // x-sorted sweep: far tiles are skipped (tail.cuh, nms_key_kernel)
  if (windowed) {
    topk_sort_kernel<<<B, 1024, (size_t)np2 * 8, st>>>(reinterpret_cast<const float*>(w.src4), 4, N, np2, N, w.perm);
    LAUNCHED();
  }
  nms_key_kernel<<<dim3(cdiv(N, 256), B), 256, 0, st>>>(w.src4, conf, N, ctx->cfg.nms_radius, use_nms, windowed ? w.perm : nullptr, w.key);
  LAUNCHED();
  topk_sort_kernel<<<B, 1024, (size_t)np2 * 8, st>>>(w.key, 1, N, np2, S, seeds);
  LAUNCHED();
  return 0;
}

int launch_select(const float* dist, int B, int N, int S, int k, int* knn, cudaStream_t st) {
  constexpr int SPC = 8;
  seed_select_kernel<SPC><<<dim3(cdiv(S, SPC), B), SPC * 32, 0, st>>>(dist, N, S, k, knn);
  LAUNCHED();
  return 0;
}

int run_seed_hypotheses(const gmf_ctx* ctx, Work& w, const float* normed, const float* src, const float* tgt, const int* seeds, int B,
                        int N, int S, int k, int* knn, float* seed_w, float* seed_trans, cudaStream_t st) {
  if (k < 1 || k > 40) return fail(GMF_ERR_INVALID, "k must be in [1, 40]");
  {
    ProfScope ps(CAT_KNN, st);
    {   // tensor pipe, error-compensated tf32 (K = 384): dgr_head.cuh knn_operand_kernel + img_gemm_kernel<128, DE_DIST>
      const int st_ = cdiv(S, 128), nt_ = cdiv(N, 128);
      knn_operand_kernel<<<dim3(st_ * 16, B), 256, 0, st>>>(normed, seeds, 1, N, S, st_, w.knn_a);
      LAUNCHED();
      knn_operand_kernel<<<dim3(nt_ * 16, B), 256, 0, st>>>(normed, nullptr, 0, N, N, nt_, w.knn_b);
      LAUNCHED();
      ImgGemmArgs a{};
      a.a_img = w.knn_a; a.w_packed = w.knn_b; a.K = 384; a.L = S; a.tiles = st_; a.out = w.dist; a.ld = N; a.ncols = N;
      a.a_pair_stride = (size_t)st_ * 12 * 4096; a.w_pair_stride = (size_t)nt_ * 12 * 4096; a.out_pair_stride = (size_t)S * N;
      cudaError_t e = launch_img_gemm<128, DE_DIST>(a, nt_, st, B);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (e != cudaSuccess) return fail_cuda(e, "seed kNN distance GEMM launch");
    }
    TRY(launch_select(w.dist, B, N, S, k, knn, st));
  }
  ProfScope ps(CAT_SPECTRAL, st);
  CU(cudaMemsetAsync(w.pair_mask, 0xff, (size_t)B * sizeof(unsigned), st));
  seed_spectral_kernel<0><<<dim3(S, B), 128, 0, st>>>(normed, src, tgt, knn, N, S, k, ctx->sigma, ctx->sigma_spat,
                                                     ctx->cfg.num_iterations, w.pair_mask, nullptr, nullptr, w.seedM);
  LAUNCHED();
  seed_spectral_kernel<1><<<dim3(S, B), 32, 0, st>>>(normed, src, tgt, knn, N, S, k, ctx->sigma, ctx->sigma_spat,
                                                     ctx->cfg.num_iterations, w.pair_mask, seed_w, seed_trans, w.seedM);
  LAUNCHED();
  return 0;
}

int run_score(const gmf_ctx* ctx, Work& w, const float* seed_trans, int B, int N, int S, int refine, float* final_trans, float* labels,
              int* counts, int* best, float* pre_refine, cudaStream_t st) {
  ProfScope ps(CAT_SCORE, st);
  CU(cudaMemsetAsync(counts, 0, (size_t)B * S * sizeof(int), st));
  score_kernel<<<dim3(cdiv(S, kScoreSeeds), cdiv(N, 256 * kScorePPT), B), 256, 0, st>>>(w.src4, w.tgt4, seed_trans, N, S, ctx->cfg.inlier_threshold, counts);
  LAUNCHED();
  const float rtau = (ctx->cfg.inlier_threshold == 0.10f) ? 0.10f : 1.2f;   // PointDSC.py:505-508
  select_refine_kernel<<<B, 1024, 0, st>>>(w.src4, w.tgt4, seed_trans, counts, N, S, ctx->cfg.inlier_threshold, rtau, 20, refine,
                                           pre_refine, final_trans, labels, best);
  LAUNCHED();
  return 0;
}

int forward_chunk(gmf_ctx* ctx, Work& w, const float* corr, const float* src, const float* tgt, const float* p_tok, const float* q_tok,
                  int B, int N, int T, int testing, float* final_trans, float* labels, float* conf_out, int* seeds_out, float* feat_out,
                  cudaStream_t st) {
  const int S = num_seeds(ctx, N), k = eff_k(ctx, N);
  TRY(run_prep(ctx, w, src, tgt, B, N, st));
  {
    ProfScope ps(CAT_PREP, st);
    // written straight as the split fp16 tile image the first encoder layer's persistent PointCN / QKV kernel consumes
    const int tiles = cdiv(N, 128);
    const long long rows = (long long)B * tiles * 128;
    layer0_kernel<<<(unsigned)std::min<long long>((rows + 7) / 8, 148 * 16), 256, 0, st>>>(corr, ctx->l0_w, ctx->l0_b, nullptr, rows, 6, w.feat_img, N, tiles);
    LAUNCHED();
  }
  // Fusion-1: queries = q-image tokens, context = p-image tokens (PointDSC.py:137)
  TRY(run_fusion(ctx, ctx->f1, w, q_tok, p_tok, B, T, T, w.imgfeat, st));
  const int L = ctx->cfg.num_layers;
  // every layer's context K / V^T depends only on the Fusion-1 output: one persistent kernel projects them all, reading the context once
  TRY(run_fusion_kv_all(ctx, w, w.imgfeat, B, T, st));
  // between layers the features travel as the split fp16 tile image the next PointCN/QKV kernel consumes (last layer: row-major)
  for (int li = 0; li < L; ++li)
    TRY(run_encoder_layer(ctx, li, w, w.featA, w.imgfeat, B, N, T, w.featA, st, true, li + 1 < L, true));
  if (feat_out) CU(cudaMemcpyAsync(feat_out, w.featA, (size_t)B * N * 128 * 4, cudaMemcpyDeviceToDevice, st));
  TRY(run_classify(ctx, w.featA, (long long)B * N, w.normed, conf_out, st));
  TRY(run_pick_seeds(ctx, w, conf_out, B, N, S, testing ? 1 : 0, seeds_out, st));
  TRY(run_seed_hypotheses(ctx, w, w.normed, src, tgt, seeds_out, B, N, S, k, w.knn, w.seed_w, w.seed_trans, st));
  TRY(run_score(ctx, w, w.seed_trans, B, N, S, testing ? 1 : 0, final_trans, labels, w.counts, w.best, w.pre_refine, st));
  return 0;
}

int require_loaded(const gmf_ctx* ctx) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (!ctx->loaded) return fail(GMF_ERR_STATE, "weights not loaded (gmf_load_weights)");
  g_prof = const_cast<Prof*>(&ctx->prof);
  return 0;
}

// fp32 row-major [L][D] -> bf16 UMMA tile images (debug / test path only)
__global__ void pack_tiles_kernel(const float* __restrict__ src, int L, int D, int tiles, float scale, int transpose,
                                  __nv_bfloat16* __restrict__ dst) {
  const int pair = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)tiles * 128 * D;
  if (e >= total) return;
  const int row = (int)(e / D), c = (int)(e % D);
  const int tile = row >> 7, r = row & 127;
  const float v = row < L ? src[((size_t)pair * L + row) * D + c] * scale : 0.f;
  uint8_t* base = (uint8_t*)(dst + ((size_t)pair * tiles + tile) * 128 * D);
  size_t off;
  if (!transpose) off = (size_t)(c >> 6) * 16384 + swz_off(r, (c & 63) >> 3) + (c & 7) * 2;
  else off = (size_t)(r >> 6) * (D * 128) + swz_off(c, (r & 63) >> 3) + (r & 7) * 2;
  if (transpose) *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16_rn(v);                 // V^T: bf16
  else *reinterpret_cast<__half*>(base + off) = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));     // Q / K: fp16
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* gmf_last_error(void) { return g_err.c_str(); }
const char* gmf_version(void) { return "gmf_b200 0.1 (sm_100a, tcgen05)"; }

int64_t gmf_launch_count(int reset) {
  const long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int gmf_profile_enable(gmf_ctx* ctx, int enable) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  ctx->prof.on = enable != 0;
  ctx->prof.used = 0;
  ctx->prof.cat.clear();
  return 0;
}

int gmf_profile_read(gmf_ctx* ctx, int category, double* total_ms, int64_t* launches) {
  if (!ctx || category < 0 || category >= CAT_COUNT) return fail(GMF_ERR_INVALID, "bad profile category");
  CU(cudaSetDevice(ctx->device));
  CU(cudaDeviceSynchronize());
  double ms = 0;
  int64_t n = 0;
  for (size_t i = 0; i < ctx->prof.cat.size(); ++i)
    if (ctx->prof.cat[i] == category) {
      float t = 0;
      CU(cudaEventElapsedTime(&t, ctx->prof.ev[2 * i], ctx->prof.ev[2 * i + 1]));
      ms += t; ++n;
    }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  return 0;
}

int gmf_weight_count(int num_layers) { return (int)build_spec(num_layers).size(); }
int gmf_weight_spec(int num_layers, int index, char* name, int cap, int64_t* numel) {
  const auto s = build_spec(num_layers);
  if (index < 0 || index >= (int)s.size()) return fail(GMF_ERR_INVALID, "weight index out of range");
  if (name && cap > 0) {
    strncpy(name, s[index].name.c_str(), cap - 1);
    name[cap - 1] = 0;
  }
  if (numel) *numel = s[index].numel;
  return 0;
}

int gmf_create(gmf_ctx** out, int device, const gmf_config* cfg) {
  if (!out || !cfg) return fail(GMF_ERR_INVALID, "NULL argument");
  if (cfg->num_layers < 1 || cfg->k < 1 || cfg->k > 40 || cfg->num_iterations < 1 || cfg->num_iterations > 31)
    return fail(GMF_ERR_INVALID, "unsupported config (need num_layers>=1, 1<=k<=40, 1<=num_iterations<=31)");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(GMF_ERR_INVALID, "no such CUDA device");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(GMF_ERR_INVALID, std::string("gmf_b200 needs an sm_100 (Blackwell B200) device, found ") + prop.name);
  gmf_ctx* c = new gmf_ctx();
  c->device = device;
  c->cfg = *cfg;
  if (const char* e = getenv("GMF_CHUNK_PAIRS")) c->chunk_pairs = std::max(1, atoi(e));
  if (const char* e = getenv("GMF_MATCH_IMPL")) c->match_impl = atoi(e);
  *out = c;
  return 0;
}

void gmf_destroy(gmf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->blob) cudaFree(ctx->blob);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->side) { cudaStreamDestroy(ctx->side); cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); }
  if (ctx->copy_stream) {
    cudaStreamDestroy(ctx->copy_stream);
    for (auto e : ctx->copy_ev) if (e) cudaEventDestroy(e);
    cudaEventDestroy(ctx->start_ev);
    for (auto e : ctx->in_done) if (e) cudaEventDestroy(e);
  }
  for (auto e : ctx->prof.ev) cudaEventDestroy(e);
  if (g_prof == &ctx->prof) g_prof = nullptr;
  delete ctx;
}

int gmf_load_weights(gmf_ctx* ctx, const float* host, int64_t numel) {
  if (!ctx || !host) return fail(GMF_ERR_INVALID, "NULL argument");
  const int L = ctx->cfg.num_layers;
  const auto spec = build_spec(L);
  int64_t total = 0;
  std::vector<int64_t> offs(spec.size());
  for (size_t i = 0; i < spec.size(); ++i) { offs[i] = total; total += spec[i].numel; }
  if (numel != total) return fail(GMF_ERR_INVALID, "flat weight buffer has " + std::to_string(numel) + " elements, expected " + std::to_string(total));
  size_t cursor = 0;
  auto next = [&](const std::string& suffix) -> const float* {
    const Spec& s = spec[cursor];
    if (s.name.size() < suffix.size() || s.name.compare(s.name.size() - suffix.size(), suffix.size(), suffix) != 0) {
      fprintf(stderr, "gmf_load_weights: internal spec mismatch at %s (wanted *%s)\n", s.name.c_str(), suffix.c_str());
      abort();
    }
    return host + offs[cursor++];
  };
  auto vec = [](const float* p, size_t n) { return std::vector<float>(p, p + n); };

  Blob blob;
  struct FusionOff { bool pe; size_t cqw, cqb, ccw, ccb, lqg, lqb, lcg, lcb, lfg, lfb, wq, wq16, wkv, wkv16, wo, bo, b1, b2, w1f, w2f; };
  auto pack_fusion = [&](bool pe) {
    FusionOff o{};
    o.pe = pe;
    if (pe) {
      o.cqw = blob.push(next("cpe.proj_q.weight"), 384); o.cqb = blob.push(next("cpe.proj_q.bias"), 128);
      o.ccw = blob.push(next("cpe.proj_content.weight"), 384); o.ccb = blob.push(next("cpe.proj_content.bias"), 128);
    }
    o.lqg = blob.push(next("0.norm.weight"), 128); o.lqb = blob.push(next("0.norm.bias"), 128);
    o.lcg = blob.push(next("norm_context.weight"), 128); o.lcb = blob.push(next("norm_context.bias"), 128);
    std::vector<float> wq = vec(next("to_q.weight"), 64 * 128);
    const float qs = kLog2e / 8.0f;   // dim_head ** -0.5 (fusion_layer.py:76) in log2 units
    for (auto& v : wq) v *= qs;
    o.wq = blob.push(pack_linear(wq, 64, 128, 32, 64));
    o.wq16 = blob.push(pack_linear_f16(wq, 64, 128, 64));
    {
      const std::vector<float> Wkv = vec(next("to_kv.weight"), 128 * 128);
      o.wkv = blob.push(pack_linear(Wkv, 128, 128, 32, 128));
      o.wkv16 = blob.push(pack_linear_f16(Wkv, 128, 128, 128));
    }
    o.wo = blob.push(pack_linear(vec(next("to_out.weight"), 128 * 64), 128, 64, 64, 128));
    o.bo = blob.push(next("to_out.bias"), 128);
    o.lfg = blob.push(next("1.norm.weight"), 128); o.lfb = blob.push(next("1.norm.bias"), 128);
    const std::vector<float> W1 = vec(next("net.0.weight"), 1024 * 128);
    std::vector<int> rowmap8(1024);     // fused FFN: pass p = [value rows 64p.., gate rows 512+64p..]
    for (int p = 0; p < 8; ++p)
      for (int n = 0; n < 128; ++n) rowmap8[p * 128 + n] = n < 64 ? p * 64 + n : 512 + p * 64 + (n - 64);
    o.w1f = blob.push(pack_linear_f16(W1, 1024, 128, 128, &rowmap8));
    o.b1 = blob.push(next("net.0.bias"), 1024);
    const std::vector<float> W2 = vec(next("net.2.weight"), 128 * 512);
    o.w2f = blob.push(pack_linear_f16(W2, 128, 512, 128));          // 8 chunks of [128 out rows x 64 hidden] fp16
    o.b2 = blob.push(next("net.2.bias"), 128);
    return o;
  };
  // conv (k=1) followed by eval BatchNorm folded into (W', b'):  PointDSC.py:104-111, 13-21
  auto fold_bn = [&](std::vector<float>& W, std::vector<float>& b, int nout, int k, const float* g, const float* beta, const float* mu,
                     const float* var) {
    for (int o = 0; o < nout; ++o) {
      const float s = g[o] / std::sqrt(var[o] + 1e-5f);
      for (int i = 0; i < k; ++i) W[(size_t)o * k + i] *= s;
      b[o] = (b[o] - mu[o]) * s + beta[o];
    }
  };

  const float sigma = *next("sigma");
  const float sigma_spat = *next("sigma_spat");
  const size_t l0w = blob.push(next("layer0.weight"), 768), l0b = blob.push(next("layer0.bias"), 128);
  const FusionOff f1 = pack_fusion(false);
  struct LayerOff { size_t pb, qw, qb, f1w, f1b, f2w, f2b, f3w, f3b, pqw; FusionOff f2; };
  std::vector<LayerOff> lo(L);
  for (int i = 0; i < L; ++i) {
    LayerOff& o = lo[i];
    std::vector<float> pcn64;
    {
      std::vector<float> W = vec(next("0.weight"), 128 * 128), b = vec(next("0.bias"), 128);
      const float *g = next("1.weight"), *be = next("1.bias"), *mu = next("1.running_mean"), *va = next("1.running_var");
      fold_bn(W, b, 128, 128, g, be, mu, va);
      o.pb = blob.push(b);
      pcn64 = pack_linear_split16(W, 128, 128, 128);
    }
    {
      std::vector<float> W = vec(next("fc_message.0.weight"), 64 * 128), b = vec(next("fc_message.0.bias"), 64);
      const float *g = next("fc_message.1.weight"), *be = next("fc_message.1.bias"), *mu = next("fc_message.1.running_mean"),
                  *va = next("fc_message.1.running_var");
      fold_bn(W, b, 64, 128, g, be, mu, va);
      o.f1w = blob.push(pack_linear(W, 64, 128, 32, 64)); o.f1b = blob.push(b);
    }
    {
      std::vector<float> W = vec(next("fc_message.3.weight"), 64 * 64), b = vec(next("fc_message.3.bias"), 64);
      const float *g = next("fc_message.4.weight"), *be = next("fc_message.4.bias"), *mu = next("fc_message.4.running_mean"),
                  *va = next("fc_message.4.running_var");
      fold_bn(W, b, 64, 64, g, be, mu, va);
      o.f2w = blob.push(pack_linear(W, 64, 64, 64, 64)); o.f2b = blob.push(b);
    }
    o.f3w = blob.push(pack_linear_f16(vec(next("fc_message.6.weight"), 128 * 64), 128, 64, 128));   // ninth W2 chunk of the fused FFN
    o.f3b = blob.push(next("fc_message.6.bias"), 128);
    {
      std::vector<float> W(384 * 128), b(384);
      const float qs = kLog2e / std::sqrt(128.0f);   // 1/sqrt(num_channels) (PointDSC.py:60) in log2 units
      for (int q = 0; q < 3; ++q) {
        const float* wq = next(".weight");
        const float* bq = next(".bias");
        const float s = q == 0 ? qs : 1.0f;
        for (int j = 0; j < 128 * 128; ++j) W[(size_t)q * 128 * 128 + j] = wq[j] * s;
        for (int j = 0; j < 128; ++j) b[q * 128 + j] = bq[j] * s;
      }
      const std::vector<float> qkv64 = pack_linear(W, 384, 128, 64, 128);
      o.qw = blob.push(qkv64); o.qb = blob.push(b);
      const std::vector<float> qkv_split = pack_linear_split16(W, 384, 128, 128);
      pcn64.insert(pcn64.end(), qkv_split.begin(), qkv_split.end());
      o.pqw = blob.push(pcn64);
    }
    o.f2 = pack_fusion(true);
  }
  const size_t c1w = blob.push(next("classification.0.weight"), 32 * 128), c1b = blob.push(next("classification.0.bias"), 32);
  const size_t c2w = blob.push(next("classification.2.weight"), 32 * 32), c2b = blob.push(next("classification.2.bias"), 32);
  const size_t c3w = blob.push(next("classification.4.weight"), 32), c3b = blob.push(next("classification.4.bias"), 1);
  if (cursor != spec.size()) return fail(GMF_ERR_STATE, "internal: weight table not fully consumed");

  CU(cudaSetDevice(ctx->device));
  if (ctx->blob) { CU(cudaDeviceSynchronize()); cudaFree(ctx->blob); ctx->blob = nullptr; }
  CU(cudaMalloc(&ctx->blob, blob.h.size() * sizeof(float)));
  CU(cudaMemcpy(ctx->blob, blob.h.data(), blob.h.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* d = ctx->blob;
  auto fuse = [&](const FusionOff& o) {
    FusionW f;
    f.pe = o.pe;
    if (o.pe) { f.cpe_q_w = d + o.cqw; f.cpe_q_b = d + o.cqb; f.cpe_c_w = d + o.ccw; f.cpe_c_b = d + o.ccb; }
    f.lnq_g = d + o.lqg; f.lnq_b = d + o.lqb; f.lnc_g = d + o.lcg; f.lnc_b = d + o.lcb; f.lnf_g = d + o.lfg; f.lnf_b = d + o.lfb;
    f.wq = d + o.wq; f.wkv = d + o.wkv; f.wo = d + o.wo; f.bo = d + o.bo; f.b1 = d + o.b1; f.b2 = d + o.b2;
    f.w1f = d + o.w1f; f.w2f = d + o.w2f; f.wkv16 = d + o.wkv16; f.wq16 = d + o.wq16;
    return f;
  };
  ctx->sigma = sigma; ctx->sigma_spat = sigma_spat;
  ctx->l0_w = d + l0w; ctx->l0_b = d + l0b;
  ctx->f1 = fuse(f1);
  ctx->layers.resize(L);
  for (int i = 0; i < L; ++i) {
    LayerW& w = ctx->layers[i];
    const LayerOff& o = lo[i];
    w.pcn_b = d + o.pb; w.qkv_w = d + o.qw; w.qkv_b = d + o.qb; w.pq_w = d + o.pqw;
    w.fc1_w = d + o.f1w; w.fc1_b = d + o.f1b; w.fc2_w = d + o.f2w; w.fc2_b = d + o.f2b; w.fc3_w = d + o.f3w; w.fc3_b = d + o.f3b;
    w.f2 = fuse(o.f2);
  }
  ctx->cls = ClsWeights{d + c1w, d + c1b, d + c2w, d + c2b, d + c3w, d + c3b};
  ctx->loaded = true;
  return 0;
}

size_t gmf_workspace_bytes(const gmf_ctx* ctx, int B, int N, int T) {
  if (!ctx || B < 1 || N < 2) return 0;
  Work w;
  const int Bc = std::min(B, ctx->chunk_pairs);
  const int S = std::max(num_seeds(ctx, N), 1), k = std::max(eff_k(ctx, N), 1), L = ctx->cfg.num_layers;
  return carve(w, nullptr, Bc, N, T, S, k, L) + 2048;
}

int gmf_pointdsc_forward(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok, const float* q_tok,
                         int B, int N, int T, int testing, float* final_trans, float* final_labels, float* confidence, int32_t* seeds,
                         float* feat, void* workspace, size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (B < 1 || N < 2 || T < 1) return fail(GMF_ERR_INVALID, "need B >= 1, N >= 2, T >= 1");
  if (!corr_pos || !src || !tgt || !p_tok || !q_tok || !final_trans || !final_labels || !confidence || !seeds)
    return fail(GMF_ERR_INVALID, "NULL tensor argument");
  if (num_seeds(ctx, N) < 1) return fail(GMF_ERR_INVALID, "int(N * ratio) must be >= 1");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int S = num_seeds(ctx, N);
  const int Bc = std::min(B, ctx->chunk_pairs);
  const int L = ctx->cfg.num_layers;
  if (!ctx->side && !getenv("GMF_NO_SIDE")) {
    CU(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  }
  Work w;
  {
    if (!workspace) return fail(GMF_ERR_INVALID, "workspace is NULL");
    const int Sw = std::max(S, 1), kw = std::max(eff_k(ctx, N), 1);
    const size_t need = carve(w, nullptr, Bc, N, T, Sw, kw, L) + 1024;
    if (workspace_bytes < need) return fail(GMF_ERR_STATE, "workspace too small: need " + std::to_string(need) + " bytes");
    carve(w, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), Bc, N, T, Sw, kw, L);
  }
  for (int b0 = 0; b0 < B; b0 += Bc) {
    const int nb = std::min(Bc, B - b0);
    TRY(forward_chunk(ctx, w, corr_pos + (size_t)b0 * N * 6, src + (size_t)b0 * N * 3, tgt + (size_t)b0 * N * 3,
                      p_tok + (size_t)b0 * T * 128, q_tok + (size_t)b0 * T * 128, nb, N, T, testing, final_trans + (size_t)b0 * 16,
                      final_labels + (size_t)b0 * N, confidence + (size_t)b0 * N, seeds + (size_t)b0 * S,
                      feat ? feat + (size_t)b0 * N * 128 : nullptr, st));
  }
  return 0;
}

}  // extern "C"

namespace {
// Chunk boundaries of the host entry point.  The first chunk is small (its upload is the only exposed one), later chunks double
// (the forward of a chunk takes ~2.6x its upload at cfg#2, so the copy stream stays ahead); within +25 % of each target the size is
// chosen so that the attention grids (cdiv(N,128) and cdiv(N,256) CTAs per pair) fill whole waves of the SMs.
std::vector<int> plan_host_chunks(int B, int N, int cap, int sms) {
  std::vector<int> cuts{0};
  cap = std::max(cap, 1);
  if (B <= 8) {
    for (int b0 = 0; b0 < B; b0 += cap) if (b0) cuts.push_back(b0);
    cuts.push_back(B);
    return cuts;
  }
  const int tp = cdiv(N, 128), tp2 = cdiv(tp, 2);
  auto eff = [&](int k) {
    const double w = (double)tp * k / sms, w2 = (double)tp2 * k / sms;
    return (w / std::ceil(w)) * (w2 / std::ceil(w2));
  };
  int done = 0, target = std::max(1, B / 9);
  while (done < B) {
    const int rem = B - done;
    const int hi = std::min(std::min(rem, cap), target + target / 4 + 1), lo = std::min(target, hi);
    int best = lo;
    double be = -1.0;
    for (int k = lo; k <= hi; ++k) {
      const double e = eff(k);
      if (e > be + 1e-9) { be = e; best = k; }
    }
    if (rem - best < target) best = std::min(rem, cap);      // no short tail chunk
    done += best;
    cuts.push_back(done);
    target = std::min(2 * target, cap);
  }
  return cuts;
}
}  // namespace

extern "C" {

static int forward_host_impl(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok,
                             const float* q_tok, int B, int N, int T, int testing, float* final_trans, float* final_labels,
                             float* confidence, void* stream, bool sync) {
  TRY(require_loaded(ctx));
  if (B < 1 || N < 2 || T < 1) return fail(GMF_ERR_INVALID, "need B >= 1, N >= 2, T >= 1");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int S = std::max(num_seeds(ctx, N), 1);
  const size_t ws = gmf_workspace_bytes(ctx, B, N, T);
  Bump b{nullptr};
  auto layout = [&](Bump& bb, float*& d_corr, float*& d_src, float*& d_tgt, float*& d_p, float*& d_q, float*& d_tr, float*& d_lab,
                    float*& d_conf, int*& d_seeds, uint8_t*& d_ws) {
    // two input staging sets: the uploads of call k+1 may start while call k still computes from the other set
    for (int set = 0; set < 2; ++set) {
      float* c_ = bb.take<float>((size_t)B * N * 6); float* s_ = bb.take<float>((size_t)B * N * 3); float* t_ = bb.take<float>((size_t)B * N * 3);
      float* p_ = bb.take<float>((size_t)B * T * 128); float* q_ = bb.take<float>((size_t)B * T * 128);
      if (set == ctx->in_set) { d_corr = c_; d_src = s_; d_tgt = t_; d_p = p_; d_q = q_; }
    }
    d_tr = bb.take<float>((size_t)B * 16); d_lab = bb.take<float>((size_t)B * N); d_conf = bb.take<float>((size_t)B * N);
    d_seeds = bb.take<int>((size_t)B * S); d_ws = bb.take<uint8_t>(ws);
  };
  float *d_corr, *d_src, *d_tgt, *d_p, *d_q, *d_tr, *d_lab, *d_conf; int* d_seeds; uint8_t* d_ws;
  layout(b, d_corr, d_src, d_tgt, d_p, d_q, d_tr, d_lab, d_conf, d_seeds, d_ws);
  const size_t need = b.off + 2048;
  bool realloc_stage = false;
  if (need > ctx->stage_bytes) {
    realloc_stage = true;
    CU(cudaDeviceSynchronize());
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr; ctx->stage_bytes = 0;
    CU(cudaMalloc(&ctx->stage, need));
    ctx->stage_bytes = need;
  }
  Bump bb{(uint8_t*)(((uintptr_t)ctx->stage + 1023) & ~(uintptr_t)1023)};
  layout(bb, d_corr, d_src, d_tgt, d_p, d_q, d_tr, d_lab, d_conf, d_seeds, d_ws);
  // Pipelined in chunks of pairs: the inputs of chunk c+1 travel on a copy stream while chunk c computes, so only the first
  // (small) chunk's upload is exposed (plan_host_chunks: ~B/9 pairs first, then doubling, sizes snapped to whole waves of attention CTAs).
  if (!ctx->copy_stream) {
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->copy_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->start_ev, cudaEventDisableTiming));
    for (auto& e : ctx->in_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventRecord(ctx->in_done[0], st)); CU(cudaEventRecord(ctx->in_done[1], st));
  }
  cudaStream_t cs = ctx->copy_stream;
  if (realloc_stage) {                                         // fresh buffers: nothing in flight refers to them
    CU(cudaEventRecord(ctx->in_done[0], st)); CU(cudaEventRecord(ctx->in_done[1], st));
  }
  CU(cudaStreamWaitEvent(cs, ctx->in_done[ctx->in_set], 0));   // the forward that last read this input set has finished
  if (ctx->in_shape[0] != B || ctx->in_shape[1] != N || ctx->in_shape[2] != T) {
    CU(cudaStreamWaitEvent(cs, ctx->in_done[ctx->in_set ^ 1], 0));   // the layout moved: nothing earlier may still be reading
    ctx->in_shape[0] = B; ctx->in_shape[1] = N; ctx->in_shape[2] = T;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  // chunk boundaries.  The asynchronous entry point is meant for back-to-back submission: the whole upload of call k+1 already overlaps
  // the kernels of call k, so its batch is not split (every chunk pays the fixed cost of the small seed / pose kernels again).
  // ... unless nothing is in flight (first call of a burst: the forward that used the other input set has completed): then there is no
  // compute to hide the upload behind and the call is cut like the synchronous one, which exposes ~1.5 ms of upload instead of ~13 ms.
  bool chunked = sync || realloc_stage;                        // (a re-allocation synchronised the device: idle as well)
  if (!chunked) {
    const cudaError_t q = cudaEventQuery(ctx->in_done[ctx->in_set ^ 1]);
    if (q == cudaSuccess) chunked = true;
    else if (q == cudaErrorNotReady) (void)cudaGetLastError();
    else return fail_cuda(q, "cudaEventQuery(in_done)");
  }
  std::vector<int> cuts = chunked ? plan_host_chunks(B, N, std::min(48, ctx->chunk_pairs), sms) : std::vector<int>{0};
  if (!chunked) {
    for (int b0 = ctx->chunk_pairs; b0 < B; b0 += ctx->chunk_pairs) cuts.push_back(b0);
    cuts.push_back(B);
  }
  const int nchunk = (int)cuts.size() - 1;
  if (nchunk > (int)(sizeof(ctx->copy_ev) / sizeof(ctx->copy_ev[0]))) return fail(GMF_ERR_INVALID, "batch too large for the host entry point (max 64 chunks)");
  for (int c = 0; c < nchunk; ++c) {
    const size_t b0 = cuts[c], nb = cuts[c + 1] - cuts[c];
    CU(cudaMemcpyAsync(d_corr + b0 * N * 6, corr_pos + b0 * N * 6, nb * N * 6 * 4, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(d_src + b0 * N * 3, src + b0 * N * 3, nb * N * 3 * 4, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(d_tgt + b0 * N * 3, tgt + b0 * N * 3, nb * N * 3 * 4, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(d_p + b0 * T * 128, p_tok + b0 * T * 128, nb * T * 128 * 4, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(d_q + b0 * T * 128, q_tok + b0 * T * 128, nb * T * 128 * 4, cudaMemcpyHostToDevice, cs));
    CU(cudaEventRecord(ctx->copy_ev[c], cs));
  }
  for (int c = 0; c < nchunk; ++c) {
    const size_t b0 = cuts[c];
    const int nb = cuts[c + 1] - cuts[c];
    CU(cudaStreamWaitEvent(st, ctx->copy_ev[c], 0));
    TRY(gmf_pointdsc_forward(ctx, d_corr + b0 * N * 6, d_src + b0 * N * 3, d_tgt + b0 * N * 3, d_p + b0 * T * 128, d_q + b0 * T * 128, nb, N, T, testing,
                             d_tr + b0 * 16, d_lab + b0 * N, d_conf + b0 * N, d_seeds + b0 * S, nullptr, d_ws, ws, stream));
  }
  CU(cudaMemcpyAsync(final_trans, d_tr, (size_t)B * 16 * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(final_labels, d_lab, (size_t)B * N * 4, cudaMemcpyDeviceToHost, st));
  if (confidence) CU(cudaMemcpyAsync(confidence, d_conf, (size_t)B * N * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(ctx->in_done[ctx->in_set], st));
  ctx->in_set ^= 1;
  if (sync) CU(cudaStreamSynchronize(st));
  return 0;
}

int gmf_pointdsc_forward_host(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok,
                              const float* q_tok, int B, int N, int T, int testing, float* final_trans, float* final_labels,
                              float* confidence, void* stream) {
  return forward_host_impl(ctx, corr_pos, src, tgt, p_tok, q_tok, B, N, T, testing, final_trans, final_labels, confidence, stream, true);
}

int gmf_pointdsc_forward_host_async(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok,
                                    const float* q_tok, int B, int N, int T, int testing, float* final_trans, float* final_labels,
                                    float* confidence, void* stream) {
  return forward_host_impl(ctx, corr_pos, src, tgt, p_tok, q_tok, B, N, T, testing, final_trans, final_labels, confidence, stream, false);
}

int gmf_stream_synchronize(gmf_ctx* ctx, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int gmf_fusion_layer(gmf_ctx* ctx, int layer, const float* queries, const float* context, int B, int Lq, int Lk, float* out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (layer >= ctx->cfg.num_layers) return fail(GMF_ERR_INVALID, "layer out of range");
  if (B > ctx->chunk_pairs) return fail(GMF_ERR_INVALID, "stage entry points take B <= chunk_pairs");
  CU(cudaSetDevice(ctx->device));
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, std::max(Lq, 2), Lk));
  return run_fusion(ctx, layer < 0 ? ctx->f1 : ctx->layers[layer].f2, w, queries, context, B, Lq, Lk, out, (cudaStream_t)stream);
}

int gmf_sc_attention(gmf_ctx* ctx, int layer, const float* feat, const float* src, const float* tgt, int B, int N, float* msg,
                     void* workspace, size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (layer < 0 || layer >= ctx->cfg.num_layers) return fail(GMF_ERR_INVALID, "layer out of range");
  if (B > ctx->chunk_pairs) return fail(GMF_ERR_INVALID, "stage entry points take B <= chunk_pairs");
  CU(cudaSetDevice(ctx->device));
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, N, 1));
  TRY(run_prep(ctx, w, src, tgt, B, N, (cudaStream_t)stream));
  return run_sc_attention(ctx, ctx->layers[layer], w, feat, B, N, msg, (cudaStream_t)stream);
}

int gmf_encoder_layer(gmf_ctx* ctx, int layer, const float* feat_in, const float* src, const float* tgt, const float* image_feat, int B,
                      int N, int T, float* feat_out, void* workspace, size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (layer < 0 || layer >= ctx->cfg.num_layers) return fail(GMF_ERR_INVALID, "layer out of range");
  if (B > ctx->chunk_pairs) return fail(GMF_ERR_INVALID, "stage entry points take B <= chunk_pairs");
  CU(cudaSetDevice(ctx->device));
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, N, T));
  TRY(run_prep(ctx, w, src, tgt, B, N, (cudaStream_t)stream));
  return run_encoder_layer(ctx, layer, w, feat_in, image_feat, B, N, T, feat_out, (cudaStream_t)stream);
}

int gmf_classify(gmf_ctx* ctx, const float* feat, int B, int N, float* normed, float* confidence, void* stream) {
  TRY(require_loaded(ctx));
  CU(cudaSetDevice(ctx->device));
  return run_classify(ctx, feat, (long long)B * N, normed, confidence, (cudaStream_t)stream);
}

int gmf_pick_seeds(gmf_ctx* ctx, const float* src, const float* confidence, int B, int N, int use_nms, int32_t* seeds, void* workspace,
                   size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (B > ctx->chunk_pairs) return fail(GMF_ERR_INVALID, "stage entry points take B <= chunk_pairs");
  CU(cudaSetDevice(ctx->device));
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, N, 1));
  TRY(run_prep(ctx, w, src, src, B, N, (cudaStream_t)stream));
  return run_pick_seeds(ctx, w, confidence, B, N, num_seeds(ctx, N), use_nms, seeds, (cudaStream_t)stream);
}

int gmf_seed_hypotheses(gmf_ctx* ctx, const float* normed, const float* src, const float* tgt, const int32_t* seeds, int B, int N, int S,
                        float* seed_trans, int32_t* knn_idx, float* seed_weight, void* workspace, size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (B > ctx->chunk_pairs) return fail(GMF_ERR_INVALID, "stage entry points take B <= chunk_pairs");
  if (S < 1 || S > num_seeds(ctx, N)) return fail(GMF_ERR_INVALID, "S must be in [1, int(N*ratio)]");
  CU(cudaSetDevice(ctx->device));
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, N, 1));
  return run_seed_hypotheses(ctx, w, normed, src, tgt, seeds, B, N, S, eff_k(ctx, N), knn_idx ? knn_idx : w.knn, seed_weight,
                             seed_trans ? seed_trans : w.seed_trans, (cudaStream_t)stream);
}

int gmf_score_hypotheses(gmf_ctx* ctx, const float* seed_trans, const float* src, const float* tgt, int B, int N, int S, int refine,
                         float* final_trans, float* final_labels, int32_t* fitness_counts, int32_t* best, float* pre_refine, void* workspace,
                         size_t workspace_bytes, void* stream) {
  TRY(require_loaded(ctx));
  if (B > ctx->chunk_pairs) return fail(GMF_ERR_INVALID, "stage entry points take B <= chunk_pairs");
  if (S < 1 || S > num_seeds(ctx, N)) return fail(GMF_ERR_INVALID, "S must be in [1, int(N*ratio)]");
  CU(cudaSetDevice(ctx->device));
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, N, 1));
  TRY(run_prep(ctx, w, src, tgt, B, N, (cudaStream_t)stream));
  return run_score(ctx, w, seed_trans, B, N, S, refine, final_trans, final_labels, fitness_counts ? fitness_counts : w.counts, best,
                   pre_refine, (cudaStream_t)stream);
}

int gmf_rigid_transform_3d(gmf_ctx* ctx, const float* A, const float* Bp, const float* weights, int M, int k, float* T, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (M < 1 || k < 1) return fail(GMF_ERR_INVALID, "need M >= 1, k >= 1");
  CU(cudaSetDevice(ctx->device));
  rigid_transform_kernel<<<cdiv(M, 4), 128, 0, (cudaStream_t)stream>>>(A, Bp, weights, M, k, T);
  LAUNCHED();
  return 0;
}

int gmf_global_registration(gmf_ctx* ctx, const float* X, const float* Y, const float* w, int B, int N, float quantization_size, int max_iter,
                            int max_break_count, float break_threshold_ratio, float* R, float* t, float* info, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (!X || !Y || !w || !R || !t) return fail(GMF_ERR_INVALID, "gmf_global_registration: NULL argument");
  if (B < 1 || N < 1 || max_iter < 0 || !(quantization_size > 0.f)) return fail(GMF_ERR_INVALID, "gmf_global_registration: need B >= 1, N >= 1, max_iter >= 0, quantization_size > 0");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const float eps = 1.1920928955078125e-07f;                   // np.finfo(np.float32).eps: HighDimSmoothL1Loss.eps, also passed to weighted_procrustes (:160)
  weighted_procrustes_kernel<<<B, 256, 0, st>>>(X, Y, w, N, eps, R, t);     // initialisation (:159-161); refined in place
  LAUNCHED();
  se3_refine_kernel<<<B, 512, 0, st>>>(X, Y, w, N, quantization_size, eps, max_iter, max_break_count, break_threshold_ratio, R, t, R, t, info);
  LAUNCHED();
  return 0;
}

// host-only helper (no device needed): chunk sizes gmf_pointdsc_forward_host would use; returns the number of chunks
int gmf_debug_plan_host_chunks(int B, int N, int cap, int sms, int* sizes, int max_sizes) {
  if (B < 1 || N < 1 || sms < 1) return fail(GMF_ERR_INVALID, "gmf_debug_plan_host_chunks: bad argument");
  const std::vector<int> cuts = plan_host_chunks(B, N, cap, sms);
  const int n = (int)cuts.size() - 1;
  for (int i = 0; i < n && i < max_sizes; ++i) sizes[i] = cuts[i + 1] - cuts[i];
  return n;
}

int gmf_weighted_procrustes(gmf_ctx* ctx, const float* X, const float* Y, const float* w, int B, int N, float eps, float* R, float* t, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  if (!X || !Y || !w || !R || !t) return fail(GMF_ERR_INVALID, "gmf_weighted_procrustes: NULL argument");
  if (B < 1 || N < 1) return fail(GMF_ERR_INVALID, "need B >= 1, N >= 1");
  CU(cudaSetDevice(ctx->device));
  weighted_procrustes_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(X, Y, w, N, eps, R, t);
  LAUNCHED();
  return 0;
}

int gmf_debug_linear(gmf_ctx* ctx, const float* x, const float* w_host, const float* bias_host, const float* residual, int rows, int k,
                     int nout, int relu, float* out, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = nout >= 128 ? 128 : nout;
  std::vector<float> W(w_host, w_host + (size_t)nout * k);
  std::vector<float> packed = pack_linear(W, nout, k, k >= 128 ? 32 : k, nb);
  float *dw = nullptr, *db = nullptr;
  CU(cudaMalloc(&dw, packed.size() * 4));
  CU(cudaMalloc(&db, (size_t)nout * 4));
  CU(cudaMemcpy(dw, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(db, bias_host, (size_t)nout * 4, cudaMemcpyHostToDevice));
  LinArgs a = lin(x, rows, dw, db);
  a.out = out; a.residual = residual;
  int rc = 0;
  if (k == 128 && nout == 128 && relu && !residual) rc = run_linear<128, 128, PRO_NONE, EPI_BIAS_RELU>(a, 1, st);
  else if (k == 128 && nout == 64 && relu && !residual) rc = run_linear<128, 64, PRO_NONE, EPI_BIAS_RELU>(a, 1, st);
  else if (k == 64 && nout == 64 && relu && !residual) rc = run_linear<64, 64, PRO_NONE, EPI_BIAS_RELU>(a, 1, st);
  else if (k == 64 && nout == 128 && !relu && residual) rc = run_linear<64, 128, PRO_NONE, EPI_BIAS_RES>(a, 1, st);
  else rc = fail(GMF_ERR_INVALID, "gmf_debug_linear: unsupported (k, nout, relu, residual) combination");
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dw); cudaFree(db);
  if (rc) return rc;
  if (e != cudaSuccess) return fail_cuda(e, "gmf_debug_linear sync");
  return 0;
}

int gmf_debug_attention(gmf_ctx* ctx, const float* q, const float* k, const float* v, const float* src, const float* tgt, int B, int Lq,
                        int Lk, int D, float scale, float sigma_d, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return fail(GMF_ERR_INVALID, "ctx is NULL");
  const bool sc = src && tgt;
  if (!((D == 64 && !sc) || (D == 128 && sc && Lq == Lk))) return fail(GMF_ERR_INVALID, "gmf_debug_attention: D=64 plain or D=128 SC (Lq==Lk)");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  Work w;
  TRY(check_ws(ctx, w, workspace, workspace_bytes, B, std::max(Lq, Lk), std::max(Lq, Lk)));
  const int qt = cdiv(Lq, 128), kt = cdiv(Lk, 128);
  __nv_bfloat16 *Q = sc ? w.qs : w.qf, *K = sc ? w.ks : w.kf, *V = sc ? w.vts : w.vtf;
  const long long qe = (long long)qt * 128 * D, ke = (long long)kt * 128 * D;
  pack_tiles_kernel<<<dim3((unsigned)((qe + 255) / 256), B), 256, 0, st>>>(q, Lq, D, qt, scale * kLog2e, 0, Q); LAUNCHED();
  pack_tiles_kernel<<<dim3((unsigned)((ke + 255) / 256), B), 256, 0, st>>>(k, Lk, D, kt, 1.f, 0, K); LAUNCHED();
  pack_tiles_kernel<<<dim3((unsigned)((ke + 255) / 256), B), 256, 0, st>>>(v, Lk, D, kt, 1.f, 1, V); LAUNCHED();
  AttnArgs a{};
  a.q_t = Q; a.k_t = K; a.vt_t = V; a.out = out; a.Lq = Lq; a.Lk = Lk; a.q_tiles = qt; a.k_tiles = kt;
  cudaError_t e;
  if (sc) {
    TRY(run_prep(ctx, w, src, tgt, B, Lk, st, sigma_d));
    ScAttnArgs sa{};
    sa.q_t = Q; sa.k_t = K; sa.vt_t = V; sa.aq_t = w.aq; sa.bd_t = w.bd; sa.out = out; sa.N = Lk; sa.tiles = kt;
    e = launch_sc_any(ctx, sa, B, st);
  } else {
    e = launch_fus_attn_v2(a, B, st);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail_cuda(e, "attn launch");
  return 0;
}

}  // extern "C"

#include "dgr_head_api.inl"
#include "dgr_train_api.inl"
#include "matcher_api.inl"
#include "compat_api.inl"
#include "sm_api.inl"
#include "pdsc_train_api.inl"
