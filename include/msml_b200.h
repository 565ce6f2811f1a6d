/*
 * msml_b200 — C ABI of the B200-native (sm_100a) MSML hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference
 * (ygtxr1997/MSML) has no native code on this path — every entry point below replaces a chain
 * of ATen / cuBLAS calls made from the reference's Python operators, cited per function as
 * ref <file>:<lines>.  The Python host (msml_b200/) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns all memory; the library borrows pointers for the duration of the enqueue,
 *     allocates nothing persistent and never frees caller memory;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on that stream;
 *   - return 0 on success; a negative msml_status or a positive cudaError_t otherwise, with
 *     msml_last_error() (thread-local) describing it.  Shape/dtype/alignment violations are
 *     errors — there is no fallback path of any kind.
 */
#ifndef MSML_B200_H_
#define MSML_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSML_B200_ABI_VERSION 1

typedef enum { MSML_OK = 0, MSML_EINVAL = -1, MSML_EALIGN = -2, MSML_EWORKSPACE = -3, MSML_EUNSUPPORTED = -4, MSML_ENCCL = -5 } msml_status;
typedef enum { MSML_F32 = 0, MSML_BF16 = 1, MSML_F16 = 2 } msml_dtype;
typedef enum { MSML_ACT_TANH = 0, MSML_ACT_SIGMOID = 1 } msml_act;       /* ref fmoperator.py:113-117 */
typedef enum { MSML_ARITH_ADD = 0, MSML_ARITH_SUB = 1, MSML_ARITH_DIV = 2, MSML_ARITH_MUL = 3 } msml_arith; /* :71-81 */
typedef enum { MSML_MARGIN_ARC = 0, MSML_MARGIN_COS = 1 } msml_margin;   /* ref margin_losses.py:318,203 */
typedef enum { MSML_RESIZE_NEAREST = 0 } msml_resize;

int msml_abi_version(void);
const char* msml_last_error(void);
/* Number of kernel launches this library enqueued (process-wide) since the last reset. */
int64_t msml_launch_count(void);
void msml_launch_count_reset(void);
/* Optional per-launch timing with CUDA events recorded on the launching stream.  collect()
 * synchronises and writes "name launches total_ms total_algorithmic_work\n" per kernel family
 * (work = bytes for HBM-bound kernels, flops for the tcgen05 contractions). */
void msml_profile_enable(int on);
int64_t msml_profile_collect(char* buf_host, int64_t capacity);

/* ------------------------------------------------------------------------------------------
 * K-A  mask fusion: FMCnn elementwise tail.           ref backbones/fm/fmoperator.py:288,304-310
 *   out = arith(yf, act(z)) [+ f_out] + yf            (default act=sigmoid, arith=mul)
 * All tensors share one shape AND one memory format (channels-last NHWC in the drop-in), so the
 * op is a flat stream of n elements.  16-byte alignment of every pointer is required.
 * bwd:  dyf = direct gradient wrt yf (the cat->conv path is autograd's), dz = gradient wrt z.
 * ------------------------------------------------------------------------------------------ */
int msml_fm_gate_fwd(const void* yf, const void* z, const void* f_out /*nullable*/, void* out,
                     int64_t n, int dtype, int act, int arith, void* stream);
int msml_fm_gate_bwd(const void* dout, const void* yf, const void* z, void* dyf, void* dz,
                     int64_t n, int dtype, int act, int arith, void* stream);
/* One launch over up to 8 segments (the 4 feature scales of the backbone).  Pointer / size
 * arrays are HOST arrays of length nseg. */
#define MSML_MAX_SEGMENTS 8
int msml_fm_gate_fwd_multi(int nseg, const void* const* yf_host, const void* const* z_host,
                           void* const* out_host, const int64_t* n_host,
                           int dtype, int act, int arith, void* stream);
int msml_fm_gate_bwd_multi(int nseg, const void* const* dout_host, const void* const* yf_host,
                           const void* const* z_host, void* const* dyf_host, void* const* dz_host,
                           const int64_t* n_host, int dtype, int act, int arith, void* stream);

/* Extension named by BASELINE.json north_star (not taken by the reference):
 * a low-resolution / single-channel mask  m (B,Hm,Wm,Cm), Cm in {1,C}, resized (nearest) to the
 * feature resolution, normalised into a gate and fused into yf (B,H,W,C), all NHWC.
 * bwd: dyf (B,H,W,C) and dm (B,Hm,Wm,Cm) in FP32 (reduced over broadcast channels by warp
 * shuffles and over the resize fan-out); dm must be zero-filled by the caller. */
int msml_fm_mask_fwd(const void* yf, const void* m, void* out, int64_t B, int64_t H, int64_t W,
                     int64_t C, int64_t Hm, int64_t Wm, int64_t Cm, int dtype, int act, int arith,
                     void* stream);
int msml_fm_mask_bwd(const void* dout, const void* yf, const void* m, void* dyf, float* dm,
                     int64_t B, int64_t H, int64_t W, int64_t C, int64_t Hm, int64_t Wm, int64_t Cm,
                     int dtype, int act, int arith, void* stream);

/* K-C  input assembly of the FM operator: channel concat + zero pad over NHWC tensors viewed as (P = N*H*W) pixel rows.
 *   ref backbones/fm/fmoperator.py:277-279   x = torch.cat((yf, yo), dim=1)  -> same_conv
 * fwd: cat (P, Ct) = [ yf (P, C) | yo (P, Co) | zeros ], Ct >= C + Co.  C and Ct are multiples of the 16-byte vector
 *      width (8 for bf16 / fp16, 4 for fp32); Co is arbitrary (the 18 occlusion maps).  All tensors share `dtype`.
 * bwd: dyf (P, C) = dcat[:, 0:C] [+ dadd]   where dadd (nullable, (P, C)) is the gradient reaching yf through its other
 *      consumer (the fused tail, msml_fm_gate_bwd): the sum autograd would do in one more pass over a strided slice.
 *      dyo (nullable, (P, Co)) = dcat[:, C:C+Co].
 * One launch each; replaces three strided ATen copy kernels (fwd) and a strided add (bwd). */
int msml_fm_cat_fwd(const void* yf, const void* yo, void* cat, int64_t P, int64_t C, int64_t Co, int64_t Ct,
                    int dtype, void* stream);
int msml_fm_cat_bwd(const void* dcat, const void* dadd /*nullable*/, void* dyf, void* dyo /*nullable*/,
                    int64_t P, int64_t C, int64_t Co, int64_t Ct, int dtype, void* stream);

/* K-S  structure-via-consensus segmentation criterion.
 *   ref tricks/consensus_loss.py:63-178  StructureConsensuLossFunction(alpha, beta, reduce_pixel, reduce_pixel_kl)
 *   .forward(logit, blobs, target), used as seg_criterion(final_seg, msk, msk) in ref train.py:228-229,258.
 * logit (N, C, H*W) in `dtype`, NCHW (channels_last = 0) or NHWC (1), 2 <= C <= 4; blobs / target (N, H*W) int64, blob
 * ids in [0, K), K <= 32, or -1 for a pixel that belongs to no blob (any other id, labels that differ inside a blob, or
 * a label outside [0, C) make the loss NaN).  fwd: one pass over the logits + a one-CTA finalize; writes the scalar loss (device) and `coef`
 * (2*K*N*C floats, device) that bwd needs.  bwd: dlogit = gout * dloss/dlogit (gout: device scalar, nullable = 1),
 * one elementwise pass.  reduce_*_all != 0 selects the reference's 'all' normalisations ('idx' otherwise); with
 * reduce_pixel_all a sample that lacks a blob gets a zero gradient from that blob (the reference yields NaN there). */
size_t msml_consensus_workspace(int64_t N, int64_t C, int64_t HW, int64_t K);
int msml_consensus_fwd(const void* logit, const int64_t* blobs, const int64_t* target, int64_t N, int64_t C, int64_t HW,
                       int64_t K, int channels_last, int dtype, float alpha, float beta, int reduce_pixel_all,
                       int reduce_kl_all, float* loss, float* coef, void* ws, size_t ws_bytes, void* stream);
int msml_consensus_bwd(const void* logit, const int64_t* blobs, const float* coef, const float* gout /*nullable*/,
                       void* dlogit, int64_t N, int64_t C, int64_t HW, int64_t K, int channels_last, int dtype,
                       void* stream);

/* K-O  sharded-weight SGD of the PartialFC head.
 *   ref train.py:188-191,299-300 (opt_pfc.step(); module_partial_fc.update()), headers/partial_fc.py:93-94,101-104,112-115.
 * One pass over the n_s sampled rows of the shard (row = index[r], or r when index is NULL: sample_rate 1):
 *   d = dw[r] + weight_decay * w[row];  mom[row] = momentum * mom[row] + (1 - dampening) * d;
 *   w[row] -= lr * (nesterov ? d + momentum * mom[row] : mom[row])          (torch.optim.SGD with an existing momentum buffer)
 * in place in weight / weight_mom (num_local, D): the gather of ref :93-94 and the scatter of ref :101-104 disappear.
 * lr is read from lr_dev (device scalar) when given, else taken by value.  wn_bf16 (nullable, (n_s, D)) and inv_norm
 * (nullable, (n_s)) receive normalize(w') and 1 / max(||w'||, 1e-12), i.e. ref :115 for the NEXT step (same rows:
 * sample_rate 1).  index entries must be distinct (they are: sorted unique sample); entries outside [0, num_local) are
 * skipped.  D is a multiple of 128, at most 1024. */
int msml_pfc_sgd_update(float* weight, float* weight_mom, const float* dw, const int64_t* index /*nullable*/, int64_t n_s,
                        int64_t num_local, int64_t D, const float* lr_dev /*nullable*/, float lr, float momentum,
                        float weight_decay, float dampening, int nesterov, void* wn_bf16 /*nullable*/,
                        float* inv_norm /*nullable*/, void* stream);
/* Same update from the RAW gradient dwn = dcos^T X of msml_head_bwd_raw / msml_head_step_raw: the backward of
 * normalize(sub_weight) (ref partial_fc.py:115) is applied to the row in registers first,
 *   n = max(||w||, 1e-12);  dw = dwn / n - w * <w, dwn> / n^3,
 * with the fp32 master row (as the reference's autograd), then the update above. */
int msml_pfc_sgd_update_raw(float* weight, float* weight_mom, const float* dwn, const int64_t* index /*nullable*/, int64_t n_s,
                            int64_t num_local, int64_t D, const float* lr_dev /*nullable*/, float lr, float momentum,
                            float weight_decay, float dampening, int nesterov, void* wn_bf16 /*nullable*/,
                            float* inv_norm /*nullable*/, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-P  Peer-guided branch of the Feature-Masking operator, elementwise ends.
 * Replaces ref backbones/fm/fmoperator.py:293-302
 *     m_bar = conv_m(gate);  f_out = conv1(m_bar * yf);  f_occ = conv2(m_bar * yt);  l2 = MSELoss()(f_occ, f_out)
 * where the reference runs one ATen multiply per product (plus activation and `1 - x` for mask_trans 'invert',
 * ref :160-166) and sub / pow / mean for the loss.
 * msml_fm_peer_mul_fwd: pf = m_bar * yf and (yt non-null) pt = m_bar * yt in one pass over n elements of `dtype`, all
 *   operands in one shared physical layout.  mode 0: src IS m_bar.  mode 1: src is the PRE-activation z and
 *   m_bar = 1 - act(z) (act: MSML_ACT_*), so neither the gate nor its inverse is materialised.
 * msml_fm_peer_mul_bwd: dsrc = d m_bar (mode 0) or -d m_bar * act'(z) (mode 1) with d m_bar = dpf*yf + dpt*yt, and
 *   dyf = dpf * m_bar.  yt (the frozen teacher's feature map, ref peer/arcface.py:176-190) receives no gradient.
 * msml_mse_fwd: *out (fp32, device) = mean((a - b)^2) accumulated in fp32 from `dtype` storage (autocast runs mse_loss in
 *   fp32); per-CTA partials in `workspace` (msml_mse_workspace() bytes) summed in a fixed order: deterministic.
 * msml_mse_bwd: da = 2 (a - b) / n * *gout (gout: fp32 device scalar), db = -da (db nullable).
 * ------------------------------------------------------------------------------------------ */
int msml_fm_peer_mul_fwd(const void* src, const void* yf, const void* yt, void* pf, void* pt, int64_t n, int dtype, int mode,
                         int act, void* stream);
int msml_fm_peer_mul_bwd(const void* dpf, const void* dpt, const void* src, const void* yf, const void* yt, void* dsrc, void* dyf,
                         int64_t n, int dtype, int mode, int act, void* stream);
size_t msml_mse_workspace(void);
int msml_mse_fwd(const void* a, const void* b, int64_t n, int dtype, float* out, void* workspace, size_t workspace_bytes,
                 void* stream);
int msml_mse_bwd(const void* a, const void* b, const float* gout, void* da, void* db, int64_t n, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-N  fused BatchNorm (+ residual add) (+ PReLU) over an NHWC tensor viewed as (P = N*H*W, C):
 *        y = prelu( (x - mean) * invstd * gamma + beta [+ res] )
 *   ref backbones/frb/iresnet.py:56-67, backbones/osb/unet.py:80-91, backbones/fm/fmoperator.py:52-68
 *   (torch.nn.BatchNorm2d + torch.nn.PReLU + `out += identity`).
 * training != 0: batch statistics (biased variance), running stats updated with `momentum`
 * (unbiased variance), num_batches_tracked += 1 (nullable).  training == 0: running statistics.
 * res / prelu are nullable.  save_mean / save_invstd (C) are outputs of fwd and inputs of bwd.
 * bwd: dres is written only when BOTH res and prelu are fused (otherwise d res == dy);
 *      dgamma / dbeta / dprelu are fp32 (C), overwritten, or added to when accumulate_param_grads != 0
 *      (the caller passes the parameters' .grad storage: what torch's AccumulateGrad would do in a
 *      separate kernel per parameter).  C must be a multiple of the 16-byte vector width with
 *      (C / width) dividing 256.  dadd (nullable, same layout as x) is added to dx: the gradient that reaches x
 *      through its OTHER consumer (the skip connection of ref iresnet.py:56-67: `identity = x; out = bn1(x)`),
 *      which autograd would otherwise add in a separate pass.
 * Training-mode fwd and bwd are three plain launches each (slab statistics, per-channel finalize, apply), the last two
 * with programmatic dependent launch; all of them are enqueued on `stream` by one call.  (MSML_BN_FUSED=1 in the
 * environment selects the single cooperative launch with two grid barriers that these replaced; it is slower.)
 * ------------------------------------------------------------------------------------------ */
size_t msml_bn_workspace(int64_t P, int64_t C);
int msml_bn_fwd(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                float* running_mean, float* running_var, int64_t* num_batches_tracked, float* save_mean,
                float* save_invstd, int64_t P, int64_t C, int dtype, int training, float momentum, float eps,
                void* workspace, size_t workspace_bytes, void* stream);
/* msml_bn_fwd with the statistics pass of a FOLLOWING BatchNorm folded in (training mode, split launches only).  In the
 * residual unit of ref backbones/frb/iresnet.py:56-67 the output of `bn3(out) + identity` is the input of the next unit's
 * `bn1`; torch reads that tensor once more just for bn1's batch statistics.
 *   next_ws  (nullable, a msml_bn_workspace(P, C) buffer distinct from `workspace`): the apply pass also leaves the slab
 *            statistics of the y it writes (rounded to `dtype`, i.e. what the next BN would read) there;
 *   stats_ready != 0: `workspace` is such a buffer, filled by the call that produced x with next_ws == workspace here; this
 *            call starts at the per-channel merge and never reads x for statistics.
 * Everything else as msml_bn_fwd.  Results equal the unchained call up to fp32 summation order of the slab partials. */
int msml_bn_fwd_ex(const void* x, const void* res, void* y, const float* gamma, const float* beta, const float* prelu,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float* save_mean,
                   float* save_invstd, int64_t P, int64_t C, int dtype, int training, float momentum, float eps,
                   void* workspace, size_t workspace_bytes, void* next_ws, size_t next_ws_bytes, int stats_ready,
                   void* stream);
int msml_bn_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* beta,
                const float* prelu, const float* save_mean, const float* save_invstd, void* dx, void* dres,
                const void* dadd /* nullable: dx = bn_bwd(..) + dadd */, float* dgamma, float* dbeta, float* dprelu,
                int64_t P, int64_t C, int dtype, int training, int accumulate_param_grads, void* workspace,
                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Gradient plumbing of the training step (no reference counterpart: replaces one ATen mixed-dtype add per weight,
 * i.e. torch's AccumulateGrad after the bf16 -> fp32 cast of ref train.py:283-300 under autocast).
 * dst[i] (fp32, n[i] elements) += src[i] (bf16, same element order), all segments in one launch per
 * MSML_ACCUM_MAX_SEGMENTS segments.  Pointers may be unaligned (scalar path), 16-byte alignment is fastest.
 * ------------------------------------------------------------------------------------------ */
#define MSML_ACCUM_MAX_SEGMENTS 96
int msml_accum_bf16_multi(int nseg, float* const* dst, const void* const* src, const int64_t* n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backbone optimizer of the training step: ref train.py:186-191 (torch.optim.SGD over backbone.parameters(), momentum 0.9,
 * weight decay 5e-4), stepped at train.py:299 after clip_grad_norm_(…, 5).  All parameters, momentum buffers and
 * gradients are views of three flat fp32 buffers with identical offsets (engine.FlatSGD lays them out), n elements each,
 * n a multiple of 4, padding lanes zero:
 *     g = grad / *grad_scale_dev (skipped when null) + weight_decay * w;   m = momentum * m + g;
 *     w -= *lr_dev * (nesterov ? g + momentum * m : m);                    shadow_bf16 (nullable) = bf16(w)
 * i.e. torch.optim.SGD with dampening 0 (a zero momentum buffer makes the first step m = g, as torch's does), with the
 * GradScaler-style division torch's fused SGD applies (the engine passes clip coefficient x world size there), plus the
 * bf16 copy of the new weights that the next step's autocast convolutions read.  One launch, HBM-bound
 * (20 B read/written per element + 2 B shadow).  lr and grad_scale are DEVICE scalars: capturable in a CUDA graph.
 * ------------------------------------------------------------------------------------------ */
int msml_sgd_flat(float* weight, float* momentum_buf, const float* grad, void* shadow_bf16, int64_t n, const float* lr_dev,
                  const float* grad_scale_dev, float momentum, float weight_decay, int nesterov, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-B  DAP head of the segmentation branch + argmax mask.
 *   ref backbones/osb/unet.py:158-161,223 (PixelShuffle(k)+AvgPool2d(k) == mean over k*k channel
 *   groups), train.py:357 / eval/qeval_mxnet.py:347 (final_seg[b].max(0)[1]).
 * x (B, G*kk, H, W) -> y (B, G, H, W) [+ mask (B,H,W) int64 = argmax over G, first index on ties;
 * nullable].  channels_last != 0 means the PHYSICAL layout of x and y is NHWC.
 * ------------------------------------------------------------------------------------------ */
int msml_dap_fwd(const void* x, void* y, int64_t* mask /*nullable*/, int64_t B, int64_t G, int64_t kk,
                 int64_t H, int64_t W, int channels_last, int dtype, void* stream);
int msml_dap_bwd(const void* dy, void* dx, int64_t B, int64_t G, int64_t kk, int64_t H, int64_t W,
                 int channels_last, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-D  PartialFC label remap + class sampling.          ref headers/partial_fc.py:77-94
 * ------------------------------------------------------------------------------------------ */
/* :79-81  in place: off-shard -> -1, on-shard -= class_start */
int msml_pfc_remap(int64_t* total_label, int64_t n, int64_t class_start, int64_t num_local, void* stream);
/* :86  perm[label] = 2.0 for every label != -1 (duplicates are idempotent: no unique needed) */
int msml_pfc_mark_positive(float* perm, const int64_t* total_label, int64_t n, int64_t num_local, void* stream);
/* :87-90  index = sort(topk(perm, k).indices) with k = max(num_sample, #positives): exact radix
 * select of the k-th largest value + ordered compaction (ties at the k-th value: lowest class
 * index first).  n_index (device scalar) receives k.  index must hold max(num_sample, n_labels)
 * entries. */
size_t msml_pfc_select_workspace(int64_t num_local);
int msml_pfc_select(const float* perm, int64_t num_local, int64_t num_sample, int64_t* index,
                    int64_t* n_index, void* workspace, size_t workspace_bytes, void* stream);
/* :92  in place: label != -1 -> lower_bound(index, label) */
int msml_pfc_searchsorted(int64_t* total_label, int64_t n, const int64_t* index, const int64_t* n_index,
                          void* stream);
/* :93-94 gather  dst[i,:] = src[index[i],:]   and  :103-104 scatter  dst[index[i],:] = src[i,:]
 * (fp32 rows of d elements, d % 4 == 0) */
int msml_gather_rows_f32(const float* src, const int64_t* index, float* dst, int64_t n_rows, int64_t d, void* stream);
int msml_scatter_rows_f32(float* dst, const int64_t* index, const float* src, int64_t n_rows, int64_t d, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-E..K-H  PartialFC head.                              ref headers/partial_fc.py:96-99,115,118-177
 * Per rank: X (B_tot, D) gathered features, W (n_s, D) sub-weights, tl (B_tot) remapped labels.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int   kind;   /* msml_margin */
  float s, m, a, k;
} msml_margin_params;

/* :115  wn = bf16(w / max(||w||, 1e-12)), inv_norm = 1 / max(||w||, 1e-12).  wn_t (nullable) gets
 * the transposed copy (D, ld_t) used by the dX contraction.  D % 8 == 0. */
int msml_wnorm_cast(const float* w, void* wn_bf16, void* wn_t_bf16 /*nullable*/, int64_t ld_t,
                    float* inv_norm, int64_t n, int64_t D, void* stream);
/* fp32 (rows, D) -> bf16 (rows, D) and optionally the transposed bf16 copy (D, ld_t) */
int msml_cast_bf16(const float* x, void* x_bf16, void* x_t_bf16 /*nullable*/, int64_t ld_t,
                   int64_t rows, int64_t D, void* stream);

/* bf16 (rows, cols) row-major -> (cols, ld_t) row-major (ld_t >= rows, ld_t % 8 == 0) */
int msml_transpose_bf16(const void* src, void* dst, int64_t rows, int64_t cols, int64_t ld_t, void* stream);

/* Workspace (bytes) for one head step of the given geometry. */
size_t msml_head_workspace(int64_t B_tot, int64_t n_s, int64_t D);

/* :98,132,135,139-140  logits = margin(X Wn^T) never materialised: tcgen05 GEMM whose epilogue
 * applies margin + scale and reduces per-row (max, sum exp, target logit) of THIS rank's shard.
 * stats (3, B_tot) fp32 = [rowmax | rowsum (relative to rowmax) | target logit or -inf]. */
int msml_head_fwd(const void* x_bf16, const void* wn_bf16, const int64_t* tl, int64_t B_tot,
                  int64_t n_s, int64_t D, const msml_margin_params* margin_host, float* stats,
                  void* workspace, size_t workspace_bytes, void* stream);
/* :136,141,144,162-163  merge the stats of W ranks (gathered: (W, 3, B_tot)) into the global
 * row max / log-sum and the loss  -mean(log(max(p_target, 1e-30))).
 * gstats (2, B_tot) = [global max | global sum]; loss: device scalar. */
int msml_head_merge_stats(const float* gathered, int64_t W, int64_t B_tot, float* gstats, float* loss,
                          void* stream);
/* :144-169  recompute the logits tile by tile, form grad = (softmax - smoothed one-hot) / B_tot,
 * chain through the margin, and contract: dX_full = dcos Wn (B_tot, D) fp32 (this rank's partial,
 * to be reduce-scattered), dW = normalize_bwd(dcos^T X) (n_s, D) fp32.  dcos, Wn and X are read in
 * place by MN-major UMMA operands: no transposed copies. */
int msml_head_bwd(const void* x_bf16, const void* wn_bf16, const float* inv_norm, const int64_t* tl,
                  int64_t B_tot, int64_t n_s, int64_t D, const msml_margin_params* margin_host,
                  const float* gstats, float* dx_full, float* dw,
                  void* workspace, size_t workspace_bytes, void* stream);

/* Raw mode (pairs with msml_pfc_sgd_update_raw): same as msml_head_bwd but dwn (n_s, D) receives dWn = dcos^T X, the gradient
 * with respect to the NORMALISED centres; the backward of `normalize(sub_weight)` (ref :115) is left to the fused optimizer,
 * which applies it to the row it already holds.  Removes the <Wn, dWn> reduction from the dcos epilogue and the Wn stream
 * from the dW epilogue. */
int msml_head_bwd_raw(const void* x_bf16, const void* wn_bf16, const int64_t* tl, int64_t B_tot, int64_t n_s, int64_t D,
                      const msml_margin_params* margin_host, const float* gstats, float* dx_full, float* dwn,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * The class-sharded step over NCCL, owned by the library.
 * ref headers/partial_fc.py:110,126,136,141,162,174 issues six collectives per step; here three, enqueued from C on
 * the caller's stream between the kernels.  NCCL is bound at run time (dlopen "libnccl.so.2": inside a PyTorch process
 * that is the copy torch loaded).  comm == NULL means world size 1: same calls, no collective.
 * ------------------------------------------------------------------------------------------ */
typedef struct msml_comm msml_comm;
/* rank 0: a fresh 128-byte ncclUniqueId to broadcast to the other ranks by any means */
int msml_nccl_unique_id(void* out128);
/* every rank, with its CUDA device current: ncclCommInitRank (blocking; the only synchronous call of the library) */
int msml_nccl_init(const void* unique_id128, int rank, int world, msml_comm** out);
int msml_nccl_destroy(msml_comm* comm);
int msml_comm_world(const msml_comm* comm);
int msml_comm_rank(const msml_comm* comm);
/* :110 + :126 + :79-81 in ONE all-gather: this rank's (B, D) fp32 embeddings are cast to bf16 and packed with its B int64
 * labels into one message; after the gather, x_bf16 (B*W, D) is the contiguous GEMM operand and total_label (B*W) holds
 * the labels remapped to this rank's shard (off-shard -> -1). */
size_t msml_head_gather_workspace(int64_t B, int64_t W, int64_t D);
int msml_head_gather(msml_comm* comm, const float* feat, const int64_t* label, int64_t B, int64_t D, int64_t class_start,
                     int64_t num_local, void* x_bf16, int64_t* total_label, void* workspace, size_t workspace_bytes, void* stream);
/* :132-175 on one stream: msml_head_fwd -> all-gather of the per-row (max, sum-exp, target logit) [replaces the MAX, SUM
 * and loss all-reduces] -> msml_head_merge_stats -> msml_head_bwd -> reduce-scatter(dX) -> x world_size.
 * x_grad (B, D) fp32, dw (n_s, D) fp32, loss: device scalar (identical on every rank). */
size_t msml_head_step_workspace(int64_t B, int64_t W, int64_t n_s, int64_t D);
int msml_head_step(msml_comm* comm, const void* x_bf16, const void* wn_bf16, const float* inv_norm, const int64_t* tl,
                   int64_t B, int64_t n_s, int64_t D, const msml_margin_params* margin_host, float* x_grad, float* dw,
                   float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* msml_head_step in raw mode: dwn = dcos^T X (see msml_head_bwd_raw). */
int msml_head_step_raw(msml_comm* comm, const void* x_bf16, const void* wn_bf16, const int64_t* tl, int64_t B, int64_t n_s,
                       int64_t D, const msml_margin_params* margin_host, float* x_grad, float* dwn, float* loss,
                       void* workspace, size_t workspace_bytes, void* stream);

/* In-model full-FC margin heads (ref headers/margin_losses.py:275-303, 390-418) on a
 * materialised cosine matrix (B, C) fp32, in place:  fwd applies margin + scale, bwd multiplies
 * dlogits by d logit / d cos (needs the pre-margin cosine). */
int msml_margin_fwd(float* cos_inout, const int64_t* label, int64_t B, int64_t C, int64_t ld,
                    const msml_margin_params* margin_host, void* stream);
int msml_margin_bwd(float* dlogits_inout, const float* cos, const int64_t* label, int64_t B, int64_t C,
                    int64_t ld, const msml_margin_params* margin_host, void* stream);

/* Plain tcgen05 GEMM used by the tests to validate the tensor-core mainloop in isolation:
 * C (M, N) fp32 = A (M, K) bf16 * B (N, K)^T bf16, K-major operands, lda/ldb in elements (%8). */
int msml_gemm_bf16_tn(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc,
                      int64_t M, int64_t N, int64_t K, void* stream);
/* Same contraction with per-operand storage: a_mn != 0 => A stored (K, M) row-major (M contiguous),
 * b_mn != 0 => B stored (K, N) row-major.  block_n in {256, 512} selects the accumulator tile. */
int msml_gemm_bf16(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, float* c,
                   int64_t ldc, int64_t M, int64_t N, int64_t K, int block_n, void* stream);
/* Same contract on the CTA-pair mainloop (clusters of two CTAs, tcgen05.mma.cta_group::2, 256 x block_n tiles;
 * block_n in {128, 256}): the mainloop of the head's three tensor-bound GEMMs, exposed for the exact-integer tests. */
int msml_gemm_bf16_pair(const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn, float* c,
                        int64_t ldc, int64_t M, int64_t N, int64_t K, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSML_B200_H_ */
