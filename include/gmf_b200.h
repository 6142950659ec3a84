/*
 * gmf_b200 — C ABI of the B200-native GMF-PointDSC correspondence outlier-rejection forward path.
 *
 * The reference (XiaoshuiHuang/GMF, GMF_PointDSC/) has no FFI: its boundary is the Python nn.Module
 * `PointDSC.forward(data) -> dict` (models/PointDSC.py:146-266).  This header is the boundary a
 * maintainer binds instead (ctypes stub in INTEGRATION.md; `gmf_b200/module.py` is that binding):
 * everything between "image tokens exist" (PointDSC.py:135) and the returned dict (:261-266) runs
 * behind `gmf_pointdsc_forward`, and each reference method on the path has a per-stage entry point so
 * that parity can be checked stage by stage with teacher forcing.
 *
 * Conventions: plain pointers and sizes only.  Unless a function says "host", every data pointer is
 * a DEVICE pointer owned by the caller; tensors are dense row-major fp32 (indices int32).  All
 * functions return 0 on success or a negative gmf_status; the message is available from
 * gmf_last_error() (thread local).  Work is enqueued asynchronously on `stream` (a cudaStream_t
 * passed as void*; NULL = default stream); nothing synchronises the device unless stated.  A context
 * is bound to one device and may be used from one thread at a time; distinct contexts are
 * independent (one process / context per GPU when sharding pairs).
 */
#ifndef GMF_B200_H
#define GMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmf_ctx gmf_ctx;

enum gmf_status {
  GMF_OK = 0,
  GMF_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  GMF_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
  GMF_ERR_STATE = -3,     /* weights not loaded, workspace too small, ... */
};

/* Hyper-parameters of PointDSC.__init__ (models/PointDSC.py:147-167). */
typedef struct gmf_config {
  int32_t num_layers;       /* 12 */
  int32_t num_iterations;   /* 10, power iteration cap (:160) */
  int32_t k;                /* 40, seed neighbourhood (:166), <= 40 */
  double ratio;             /* 0.1, seeds = int(N * ratio) in double precision, as Python evaluates it (:161, :246, :270) */
  float inlier_threshold;   /* 0.10 (3DMatch) / 1.2 (KITTI) (:163) */
  float nms_radius;         /* (:167) */
} gmf_config;

const char* gmf_last_error(void);
const char* gmf_version(void);

/* ---- context & weights ------------------------------------------------------------------- */
int gmf_create(gmf_ctx** out, int device, const gmf_config* cfg);
void gmf_destroy(gmf_ctx* ctx);

/* The hot-path tensors of the reference state_dict, in canonical order (gmf_b200/weights.py mirrors
 * this table).  gmf_weight_spec writes the state_dict key into name[cap] and the element count. */
int gmf_weight_count(int num_layers);
int gmf_weight_spec(int num_layers, int index, char* name, int cap, int64_t* numel);
/* host pointer: all tensors concatenated in that order (fp32).  Folds eval-mode BatchNorm into the
 * preceding 1x1 conv, folds the softmax scales into the query projections, rounds GEMM weights to
 * TF32 and uploads them as pre-swizzled UMMA tile images.  May be called again after an update. */
int gmf_load_weights(gmf_ctx* ctx, const float* host_flat, int64_t numel);

/* ---- whole path --------------------------------------------------------------------------- */
size_t gmf_workspace_bytes(const gmf_ctx* ctx, int B, int N, int T);

/* PointDSC.forward after the image backbone (PointDSC.py:137-266), testing-mode semantics per pair
 * (B > 1 == the reference looped over pairs: it asserts bs == 1, :279,:504).
 *   corr_pos [B,N,6]  src,tgt [B,N,3]  p_tok,q_tok [B,T,128] (backbone tokens, PointDSC.py:129-135)
 * outputs: final_trans [B,4,4], final_labels [B,N] (0/1), confidence [B,N] (classifier logits),
 *          seeds [B,int(N*ratio)] int32; optional (may be NULL): feat [B,N,128] un-normalised encoder
 *          features.  testing == 0 gives the training-mode selections (top-S seeds without NMS, no
 *          post-refinement, :246,:256) — final_labels stay 0/1 from the best hypothesis. */
int gmf_pointdsc_forward(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok,
                         const float* q_tok, int B, int N, int T, int testing, float* final_trans, float* final_labels,
                         float* confidence, int32_t* seeds, float* feat, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Same call with HOST buffers (pinned or pageable): H2D of the inputs, the forward, D2H of
 * final_trans / final_labels / confidence, then a stream synchronise.  Uses context-owned device
 * staging (grown on demand). */
int gmf_pointdsc_forward_host(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok,
                              const float* q_tok, int B, int N, int T, int testing, float* final_trans,
                              float* final_labels, float* confidence, void* stream);

/* Same, without the final synchronise: the call returns once everything is enqueued (uploads on a context-owned copy stream,
 * kernels and the D2H copies on `stream`).  Input staging is double-buffered, so the uploads of call k+1 overlap the kernels of
 * call k; the host input buffers must stay valid and the outputs must not be read until gmf_stream_synchronize(ctx, stream)
 * (or any later synchronising call on that stream) returns.  Use one stream per context for the host entry points. */
int gmf_pointdsc_forward_host_async(gmf_ctx* ctx, const float* corr_pos, const float* src, const float* tgt, const float* p_tok,
                                    const float* q_tok, int B, int N, int T, int testing, float* final_trans,
                                    float* final_labels, float* confidence, void* stream);
int gmf_stream_synchronize(gmf_ctx* ctx, void* stream);

/* ---- per-stage entry points (teacher-forced parity) ---------------------------------------- */
/* FusionLayer.forward (models/fusion_layer.py:172-201), depth 0.  layer < 0: encoder.fusion_layer_1 (pe=False);
 * layer >= 0: NonLocal_layer_{layer}.fusion_layer_2 (pe=True).  queries [B,Lq,128], context [B,Lk,128] -> out [B,Lq,128]. */
int gmf_fusion_layer(gmf_ctx* ctx, int layer, const float* queries, const float* context, int B, int Lq, int Lk, float* out,
                     void* workspace, size_t workspace_bytes, void* stream);
/* SC-guided non-local attention core (PointDSC.py:56-64): Q/K/V projections + softmax(compat * QK^T/sqrt(C)) V,
 * compat recomputed on the fly from src/tgt.  feat [B,N,128] (PointCN output, token-major) -> msg [B,N,128]. */
int gmf_sc_attention(gmf_ctx* ctx, int layer, const float* feat, const float* src, const float* tgt, int B, int N, float* msg,
                     void* workspace, size_t workspace_bytes, void* stream);
/* One encoder layer = PointCN_layer_i + NonLocal_layer_i (PointDSC.py:140-142, 40-74).  feat_in/out [B,N,128] token-major. */
int gmf_encoder_layer(gmf_ctx* ctx, int layer, const float* feat_in, const float* src, const float* tgt, const float* image_feat,
                      int B, int N, int T, float* feat_out, void* workspace, size_t workspace_bytes, void* stream);
/* F.normalize + classification MLP (PointDSC.py:229,241).  feat [B,N,128] -> normed [B,N,128], confidence [B,N]. */
int gmf_classify(gmf_ctx* ctx, const float* feat, int B, int N, float* normed, float* confidence, void* stream);
/* pick_seeds (PointDSC.py:268-286) when use_nms != 0, else argsort(confidence, descending)[:S] (:246). */
int gmf_pick_seeds(gmf_ctx* ctx, const float* src, const float* confidence, int B, int N, int use_nms, int32_t* seeds,
                   void* workspace, size_t workspace_bytes, void* stream);
/* cal_seed_trans up to the per-seed transforms (PointDSC.py:323-407): kNN, 40x40 compat, power iteration with the
 * reference's global allclose exit, weighted Kabsch.  Optional outputs may be NULL. */
int gmf_seed_hypotheses(gmf_ctx* ctx, const float* normed, const float* src, const float* tgt, const int32_t* seeds, int B, int N,
                        int S, float* seed_trans /*[B,S,4,4]*/, int32_t* knn_idx /*[B,S,k]*/, float* seed_weight /*[B,S,k]*/,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Hypothesis scoring (PointDSC.py:413-427) + post_refinement (:493-528, when refine != 0).  fitness_counts [B,S] int32,
 * best [B] int32 and pre_refine [B,4,4] are optional. */
int gmf_score_hypotheses(gmf_ctx* ctx, const float* seed_trans, const float* src, const float* tgt, int B, int N, int S, int refine,
                         float* final_trans, float* final_labels, int32_t* fitness_counts, int32_t* best, float* pre_refine,
                         void* workspace, size_t workspace_bytes, void* stream);
/* rigid_transform_3d (models/common.py:10-50): A,B [M,k,3], weights [M,k] or NULL -> T [M,4,4]. */
int gmf_rigid_transform_3d(gmf_ctx* ctx, const float* A, const float* B, const float* weights, int M, int k, float* T, void* stream);

/* DGR's weighted Procrustes (GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/core/registration.py:91-113, called per pair
 * from core/trainer.py:594-614 with a host-side double SVD): X, Y [B,N,3], w [B,N] -> R [B,3,3], t [B,3] with
 * w_norm = w / (sum|w| + eps), R = U diag(1,1,sign) V^T of Sxy = (Y - mu_y)^T diag(w_norm) (X - mu_x), t = mu_y - R mu_x. */
int gmf_weighted_procrustes(gmf_ctx* ctx, const float* X, const float* Y, const float* w, int B, int N, float eps, float* R, float* t,
                            void* stream);

/* DGR's pose solver `GlobalRegistration` (core/registration.py:135-194; called from core/deep_global_registration.py:336-341 with the predicted
 * inlier weights): weighted Procrustes initialisation (:159-161), then up to max_iter (reference: 1000) Adam iterations (lr 0.1, ExponentialLR 0.999)
 * on the 6-D rotation parameterisation + translation (`Transformation`, :116-132) of the weighted HighDimSmoothL1Loss with quantization_size
 * (core/loss.py:42-61); early exits: loss < 1e-7, or max_break_count (20) iterations with |dloss| < break_threshold_ratio x loss (:172,:183-186).
 * One persistent CTA per pair runs the whole loop on the device (the reference does every iteration on the host with autograd).
 * X, Y [B,N,3], w [B,N] (non-negative) -> R [B,3,3], t [B,3]; info [B,3] = (exit iteration index, final loss, break count), may be NULL. */
int gmf_global_registration(gmf_ctx* ctx, const float* X, const float* Y, const float* w, int B, int N, float quantization_size, int max_iter,
                            int max_break_count, float break_threshold_ratio, float* R, float* t, float* info, void* stream);

/* ---- classical SM baseline ------------------------------------------------ */
/* `SM` of GMF_PointDSC/baseline_scripts/baseline_3DMatch.py:19-53: M_ij = max(0, 4.5 - (|s_i-s_j| - |t_i-t_j|)^2 / (2 sigma^2)) with
 * sigma = inlier_threshold / 3 and a zero diagonal, `iters` (reference: 10) power iterations v <- M v / (|M v| + 1e-6) from v = 1, labels =
 * the top int(N * top_ratio) entries of v, pose = rigid_transform_3d(src, tgt, v * labels).  The N x N matrix is recomputed from the points
 * inside every fused mat-vec and never stored.  src, tgt [B,N,3] -> trans [B,4,4], labels [B,N] (0/1), leading_eig [B,N] (optional, may be
 * NULL).  No weights needed.  N <= 16384. */
size_t gmf_sm_workspace_bytes(int B, int N, double top_ratio);
int gmf_sm_baseline(gmf_ctx* ctx, const float* src, const float* tgt, int B, int N, float inlier_threshold, double top_ratio, int iters,
                    float* trans, float* labels, float* leading_eig, void* workspace, size_t workspace_bytes, void* stream);

/* ---- DGR bottleneck fusion head --------------------------------------------- */
/* PerceiverIO(depth=0, dim=128, latent_dim=256, cross_heads=1, cross_dim_head=128, pe) of the DGR inlier network
 * (GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py:140-221; built at model/resunet_new.py:516-525, called from
 * ResUNet2.transformer :694-705).  Replaces `self.perceiver_io(image, queries_encoder=P_att)`: ConvPosEnc on latents and
 * context, PreNorm cross-attention (1 head, d=128) + to_out + residual, PreNorm GEGLU feed-forward (256 -> 2048 -> 256) +
 * residual.  The handle owns its packed weights and a grow-on-demand device workspace; one handle per device/thread.
 * (The other PerceiverIO instance of that network, `image_fusion` (latent_dim=128, cross_dim_head=64, pe=False,
 * resunet_new.py:618-626), has the shape of PointDSC's fusion_layer_1 and runs through gmf_fusion_layer.) */
typedef struct gmf_dgr_head gmf_dgr_head;
int gmf_dgr_head_create(gmf_dgr_head** out, int device, int latent_dim /*256*/, int context_dim /*128*/, int head_dim /*128*/, int pe);
void gmf_dgr_head_destroy(gmf_dgr_head* h);
/* state_dict tensors of that module in canonical order (gmf_b200/dgr_head.py mirrors the table) */
int gmf_dgr_head_weight_count(int pe);
int gmf_dgr_head_weight_spec(int pe, int index, char* name, int cap, int64_t* numel);
/* host pointer: the tensors concatenated in that order (fp32) */
int gmf_dgr_head_load_weights(gmf_dgr_head* h, const float* host_flat, int64_t numel);
/* latents [M,256] (all active bottleneck voxels of the batch as one sequence), image_feat [T,128] -> out [M,256]; device pointers */
int gmf_dgr_head_forward(gmf_dgr_head* h, const float* latents, const float* image_feat, int M, int T, float* out, void* stream);

/* Training-mode output `M` of PointDSC.forward (models/PointDSC.py:231-234): feat [B,N,128] (un-normalised encoder features, the
 * optional `feat` output of gmf_pointdsc_forward) -> M [B,N,N] = clamp(1 - (1 - Fn Fn^T) / sigma^2, 0, 1) with a zero diagonal,
 * Fn = F.normalize(feat), sigma = the loaded `sigma` parameter.  Tensor-pipe GEMM at fp32 accuracy (error-compensated tf32). */
size_t gmf_feature_compat_workspace_bytes(int B, int N);
int gmf_feature_compat(gmf_ctx* ctx, const float* feat, int B, int N, float* M, void* workspace, size_t workspace_bytes, void* stream);

/* ---- DGR head training step (BASELINE.json configs[4]) ---------------------------------------- */
/* Reference loop: GMF_DeepGlobalRegistration_fcgf/core/trainer.py:226-300 (forward :236, loss.backward() :271, optimizer.step() :300) around
 * model/perceiver_io.py:187-221, optimiser optim.SGD(lr, momentum, weight_decay) (core/trainer.py:75-79).  params / grads / momentum_buf: FLAT fp32
 * device buffers of gmf_dgr_head_param_count(pe) floats in gmf_dgr_head_weight_spec order, owned by the caller (so that the gradient can be
 * all-reduced in one NCCL call).  train_forward saves its activations in `workspace`; train_backward must follow with the same workspace, M, T,
 * params and inputs.  grads is overwritten; d_latents [M,256] / d_image_feat [T,128] may be NULL.  Products run in TF32 on the tensor pipe. */
size_t gmf_dgr_head_train_workspace_bytes(int M, int T);
int64_t gmf_dgr_head_param_count(int pe);
int gmf_dgr_head_train_forward(gmf_dgr_head* h, const float* params, const float* latents, const float* image_feat, int M, int T, float* out,
                               void* workspace, size_t workspace_bytes, void* stream);
int gmf_dgr_head_train_backward(gmf_dgr_head* h, const float* params, const float* latents, const float* image_feat, const float* d_out, int M, int T,
                                float* d_latents, float* d_image_feat, float* grads, void* workspace, size_t workspace_bytes, void* stream);
/* torch.optim.SGD (dampening 0, no Nesterov): g = grad_scale * grad + weight_decay * p; buf = first ? g : momentum * buf + g; p -= lr * buf. */
int gmf_sgd_step(float* params, const float* grads, float* momentum_buf, int64_t n, float lr, float momentum, float weight_decay, float grad_scale,
                 int first, void* stream);

/* ---- PointDSC training step -------------------------------------------------- */
/* Reference loop: GMF_PointDSC/libs/trainer.py:123-168 (forward :134, ClassificationLoss + SpectralMatchingLoss :137-146, loss.backward() :160,
 * optimizer.step() :168) around models/PointDSC.py:191-266 in TRAINING mode (batch-statistics BatchNorm with running-statistics update, logits
 * returned as `final_labels`, feature compatibility M consumed by the SM loss), losses libs/loss.py:66-139 (both `balanced` settings).
 * params / grads: FLAT fp32 device buffers of gmf_pointdsc_param_count(num_layers) floats in gmf_weight_spec order (the reference state_dict
 * without the image backbone), owned by the caller so that the gradient is all-reduced in one NCCL call.  Device inputs: corr_pos [B,N,6],
 * src_keypts / tgt_keypts [B,N,3], p_tokens / q_tokens [B,T,128] (image backbone output), gt_labels [B,N] (0 / 1 as fp32).
 * train_forward: losses[3] (device) = {classification loss, spectral-matching loss, w_class * cls + w_sm * sm}; logits [B,N] and features
 * [B,N,128] (un-normalised encoder output) may be NULL; the running statistics inside `params` are updated.  The N x N matrix M and its gradient
 * only ever exist as 64 x 64 tiles in shared memory (fused SM loss).  train_backward must follow with the same workspace, shapes, params and
 * inputs; grads is overwritten (entries of running statistics and sigma_spat stay 0); d_p_tokens / d_q_tokens [B,T,128] (the gradient that
 * flows on into the PyTorch image backbone) may be NULL.  Matrix products run on the tensor pipe: tf32x3 = 0 plain TF32 (gradients within a few
 * per cent of fp32 autograd through 12 layers), tf32x3 = 1 error-compensated (hi hi + hi lo + lo hi, fp32-level products, three times the tensor
 * work and twice the operand-image workspace); everything else is fp32.  The same tf32x3 must be passed to all three calls of a step.
 * Autograd entry (the reference's own trainer computes its losses in Python on `final_labels` and `M`): train_forward with gt_labels = losses =
 * NULL skips the fused loss head and can write the materialised M [B,N,N] (`M`, may be NULL otherwise too); train_backward then takes the
 * caller's d_logits [B,N] / d_M [B,N,N] / d_features [B,N,128] (any of them NULL = zero; all three NULL = use the fused loss head). */
size_t gmf_pointdsc_train_workspace_bytes(int num_layers, int B, int N, int T, int tf32x3);
int64_t gmf_pointdsc_param_count(int num_layers);
int gmf_pointdsc_train_forward(int device, int num_layers, float* params, const float* corr_pos, const float* src_keypts, const float* tgt_keypts,
                               const float* p_tokens, const float* q_tokens, const float* gt_labels, int B, int N, int T, int balanced, float w_class,
                               float w_sm, int tf32x3, float* losses, float* logits, float* features, float* M, void* workspace,
                               size_t workspace_bytes, void* stream);
int gmf_pointdsc_train_backward(int device, int num_layers, const float* params, const float* corr_pos, const float* p_tokens, const float* q_tokens,
                                int B, int N, int T, float w_class, float w_sm, int tf32x3, const float* d_logits, const float* d_M,
                                const float* d_features, float* grads, float* d_p_tokens, float* d_q_tokens, void* workspace,
                                size_t workspace_bytes, void* stream);
/* torch.optim.Adam (amsgrad off; train_3DMatch.py:52-58): g = grad_scale * grad + weight_decay * p; m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
 * p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps).  step >= 1.  mask (one byte per element, NULL = all) selects the trainable
 * entries: running statistics and sigma_spat share the flat buffer and must be left alone. */
int gmf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const unsigned char* mask, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float grad_scale, int step, void* stream);

/* ---- correspondence construction (the step right before the path) ------------- */
/* Nearest-neighbour matcher in descriptor space + network input assembly, NumPy in the reference's datasets
 * (GMF_PointDSC/datasets/ThreeDMatch.py:384-391, 401-402, 411-414; datasets/KITTI.py:94-102):
 *   distance = sqrt(2 - 2 * src_desc @ tgt_desc.T + 1e-6); source_idx = argmin(distance, axis=1);
 *   use_mutual: keep i only if argmin(distance, axis=0)[source_idx[i]] == i;  corr = [i, source_idx[i]] in increasing i;
 *   src_sel = src_keypts[corr[:,0]], tgt_sel = tgt_keypts[corr[:,1]], corr_pos = [src_sel | tgt_sel] - mean over rows.
 * src_desc [B,Ns,D], tgt_desc [B,Nt,D] (unit-norm descriptors), src_keypts [B,Ns,3], tgt_keypts [B,Nt,3]  ->
 * source_idx [B,Ns] int32, corr [B,Ns,2] int32, n_corr [B] int32, src_sel / tgt_sel [B,Ns,3], corr_pos [B,Ns,6]; rows at and past
 * n_corr[b] are zero (corr: -1).  The Ns x Nt distance matrix is never materialised.  All pointers device. */
size_t gmf_match_workspace_bytes(int B, int Ns, int Nt, int D);
int gmf_build_correspondences(gmf_ctx* ctx, const float* src_desc, const float* tgt_desc, const float* src_keypts, const float* tgt_keypts,
                              int B, int Ns, int Nt, int D, int use_mutual, int32_t* source_idx, int32_t* corr, int32_t* n_corr,
                              float* src_sel, float* tgt_sel, float* corr_pos, void* workspace, size_t workspace_bytes, void* stream);

/* ---- introspection / debugging ------------------------------------------------------------- */
/* Per-launch CUDA-event timing (bench.py's roofline leg).  While enabled every kernel launch is bracketed by events on
 * the launching stream; gmf_profile_read synchronises the device and returns the summed device time and launch count
 * of one category since the last gmf_profile_enable call.  Categories: 0 PointCN + QKV projection (chained), 1 fusion Q
 * projection (CPE + LN + to_q), 2 fusion K/V projection (CPE + LN + to_kv), 3 fused GEGLU FFN (+ fc_message.6 tail),
 * 4 fusion flash attention (+ to_out + residual), 5 SC-guided flash attention (+ fc_message.0/.3 tail), 6 prep + layer0,
 * 7 classifier, 8 seed picking, 9 seed kNN, 10 spectral matching + Kabsch, 11 scoring + refinement, 12 other (stage / debug
 * entry points only). */
#define GMF_PROFILE_CATEGORIES 13
int gmf_profile_enable(gmf_ctx* ctx, int enable);
int gmf_profile_read(gmf_ctx* ctx, int category, double* total_ms, int64_t* launches);
/* number of gmf kernels launched by this process since the last reset (bench.py's gpu_launches) */
int64_t gmf_launch_count(int reset);
/* host-only: the chunk sizes gmf_pointdsc_forward_host cuts a batch of B pairs into (first chunk small, then doubling, snapped to whole
 * waves of attention CTAs on `sms` SMs, at most `cap` pairs each); writes up to max_sizes entries, returns the chunk count. */
int gmf_debug_plan_host_chunks(int B, int N, int cap, int sms, int* sizes, int max_sizes);
/* out[rows,nout] = act(x[rows,k] . W[nout,k]^T + bias) (+ residual) through the tcgen05 TF32 linear kernel.
 * W, bias are HOST pointers (packed on the fly); x/residual/out device.  (k,nout) in {(128,128),(128,64),(64,64),(64,128)}. */
int gmf_debug_linear(gmf_ctx* ctx, const float* x, const float* w_host, const float* bias_host, const float* residual, int rows, int k,
                     int nout, int relu, float* out, void* stream);
/* out[B,Lq,D] = softmax(q k^T * scale [* compat]) v through the tcgen05 attention kernel; q,k,v row-major fp32 device
 * [B,L,D], D in {64,128}; src/tgt non-NULL (with D == 128, Lq == Lk) switches the SC-guided variant on. */
int gmf_debug_attention(gmf_ctx* ctx, const float* q, const float* k, const float* v, const float* src, const float* tgt, int B, int Lq,
                        int Lk, int D, float scale, float sigma_d, float* out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GMF_B200_H */
