/*
 * leod_b200 — C ABI of the B200-native LEOD hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer named *_dev* / documented as
 * "device" must point to memory accessible from the current CUDA device; `stream` is a
 * cudaStream_t passed as void*.  All functions return 0 on success and a negative value on
 * failure; leod_last_error() then returns a thread-local, NUL-terminated description.  Nothing in
 * this library aborts the process or falls back to the CPU.
 *
 * Each entry point cites the reference interface (paths relative to the LEOD tree) it replaces.
 * Layouts: activations are channels-last ("NHWC", tokens x channels, row-major) in the handle's
 * storage dtype (LEOD_F32 or LEOD_BF16); parameters and gradients are fp32.
 */
#ifndef LEOD_B200_H_
#define LEOD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEOD_ABI_VERSION 1

enum { LEOD_F32 = 0, LEOD_BF16 = 1, LEOD_U8 = 2 };

const char *leod_last_error(void);
int leod_abi_version(void);
/* Kernels launched by this library so far in this process. */
unsigned long long leod_launch_count(void);
/* Per-kernel-class timing with CUDA events on the launching stream (measurement aid for bench.py;
 * off by default).  Kinds: 0 gemm_nt, 1 gemm_tn, 2 attention fwd, 3 attention bwd, 4 layernorm,
 * 5 lstm gates, 6 patch gather/scatter, 7 other, 8 neck/head convolution GEMMs.  collect(): out[kind*4 + {0,1,2,3}] = launches,
 * total ms, algorithmic FLOPs, algorithmic bytes since the last collect. */
int leod_profile_enable(int on);
/* Debug: route bf16 attention through the SIMT kernel instead of the tensor-core one. */
int leod_debug_force_simt_attention(int on);
int leod_profile_collect(double *out, int n_kinds);
/* Additionally log every launch (kind,us,flops,bytes,d0,d1,d2) to a CSV file at the next collect; NULL closes it. */
int leod_profile_csv(const char *path);

/* ------------------------------------------------------------------ recurrent backbone
 * Replaces models/detection/recurrent_backbone/maxvit_rnn.py:23-115 (RNNDetector) and everything
 * below it: maxvit.py:143-182, 185-270, 328-354, 85-118, 45-53 and models/layers/rnn.py:7-70. */
typedef struct leod_backbone leod_backbone_t;

typedef struct {
  int32_t in_channels;  /* backbone.input_channels (20)                         */
  int32_t embed_dim;    /* backbone.embed_dim 32/48/64; stage dims = x1,x2,x4,x8 */
  int32_t dim_head;     /* stage.attention.dim_head                              */
  int32_t part_h;       /* stage.attention.partition_size[0]                     */
  int32_t part_w;       /* stage.attention.partition_size[1]                     */
  int32_t mlp_ratio;    /* stage.attention.mlp_ratio (4)                         */
  int32_t in_h;         /* padded input height, backbone.in_res_hw[0]            */
  int32_t in_w;         /* padded input width,  backbone.in_res_hw[1]            */
  int32_t dtype;        /* LEOD_F32 or LEOD_BF16: activation storage / GEMM operand type */
  float ln_eps;         /* 1e-5                                                  */
} leod_backbone_cfg;

int leod_backbone_create(const leod_backbone_cfg *cfg, leod_backbone_t **out);
/* Same, without any device allocation: only leod_backbone_param_info / _param_count / _save_bytes and
 * _destroy are valid on the result (lets host code lay out parameters on a machine without a GPU). */
int leod_backbone_layout_only(const leod_backbone_cfg *cfg, leod_backbone_t **out);
void leod_backbone_destroy(leod_backbone_t *h);

/* Flat fp32 parameter buffer layout.  Entry i has the reference's state_dict name (without the
 * leading "backbone."), an element offset into the flat buffer and a shape (ndim <= 4).
 * Returns the number of entries when i < 0. */
int leod_backbone_param_info(const leod_backbone_t *h, int i, char *name, size_t name_cap, int64_t *offset,
                             int32_t *ndim, int64_t shape[4]);
int64_t leod_backbone_param_count(const leod_backbone_t *h); /* floats in the flat buffer */

/* Bind the flat parameter / gradient buffers (device, fp32, leod_backbone_param_count floats). */
int leod_backbone_bind(leod_backbone_t *h, float *params_dev, float *grads_dev);
/* Re-derive the operand-typed, layout-prepared weight copies from the bound parameters.  Call after
 * every parameter update (optimizer step, load_state_dict) and before the next forward. */
int leod_backbone_prepare(leod_backbone_t *h, void *stream);

int64_t leod_backbone_save_bytes(const leod_backbone_t *h, int B);

/* One timestep, all four stages (maxvit_rnn.py:97-115).
 *  x        : device, [B, in_channels, x_h, x_w] channel-first, dtype x_dtype (LEOD_F32/BF16/U8);
 *             zero-padded on the fly to in_h x in_w (utils/padding.py:33-58, detection.py:132).
 *  h_prev/c_prev : 4 device pointers (NHWC, storage dtype) or NULL entries = zero state (rnn.py:45-50)
 *  h_out/c_out   : 4 device pointers (NHWC, storage dtype), written.  h_out[s] is feature s+1.
 *  save     : device scratch of leod_backbone_save_bytes(B) bytes that keeps the activations needed
 *             by leod_backbone_step_bwd, or NULL for inference. */
int leod_backbone_step_fwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B,
                           const void *const h_prev[4], const void *const c_prev[4], void *const h_out[4],
                           void *const c_out[4], void *save, void *stream);

/* Backward of one timestep.  h_prev/c_prev/h_out/c_out: the tensors of the forward call.  dh_out/dc_out: gradients w.r.t. the step's outputs (NULL = zero);
 * dh_prev/dc_prev: written (gradients w.r.t. the incoming state).  Parameter gradients are
 * ACCUMULATED into the bound gradient buffer (and internal LayerScale scratch). */
int leod_backbone_step_bwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B,
                           const void *const h_prev[4], const void *const c_prev[4], const void *const h_out[4],
                           const void *const c_out[4], const void *save, const void *const dh_out[4], const void *const dc_out[4],
                           void *const dh_prev[4], void *const dc_prev[4], void *stream);
/* A whole BPTT window (L timesteps) in one call — the fast path; replaces the time loops of modules/detection.py:188-224
 * and modules/pseudo_labeler.py:676-704.  Stage-major schedule: stage s at time t depends only on stage s-1 at time t and on
 * its own state at t-1, so every stage runs over all L timesteps before the next one starts; everything except the
 * hidden-state half of the ConvLSTM is batched over the L*B frames, the recurrence itself is one persistent kernel per
 * stage (bf16), and weight gradients are computed by one GEMM per layer over all timesteps.  Activations live in a
 * library-owned arena (leod_backbone_seq_arena_bytes), so a forward must be followed by at most one backward with the
 * same (B, L) before the next forward.  Results equal L calls of leod_backbone_step_fwd / _step_bwd.
 *  x      : device [L, B, in_channels, x_h, x_w], dtype x_dtype
 *  h0/c0  : initial state per stage (NHWC, storage dtype) or NULL
 *  h_all  : out, per stage [L, B, h, w, C] hidden states of every timestep (= the features)
 *  c_last : out, per stage [B, h, w, C] cell state after the last timestep
 * Backward: dh_all[s] ([L,B,h,w,C] or NULL) and dc_last[s] ([B,h,w,C] or NULL) are the output gradients;
 * dh0/dc0 (NULL entries allowed) receive the gradients w.r.t. the initial state. */
int64_t leod_backbone_seq_arena_bytes(const leod_backbone_t *h, int B, int L);
int leod_backbone_seq_fwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B, int L,
                          const void *const h0[4], const void *const c0[4], void *const h_all[4], void *const c_last[4],
                          void *stream);
int leod_backbone_seq_bwd(leod_backbone_t *h, const void *x, int x_dtype, int x_h, int x_w, int B, int L,
                          const void *const h0[4], const void *const c0[4], const void *const h_all[4],
                          const void *const dh_all[4], const void *const dc_last[4], void *const dh0[4], void *const dc0[4],
                          void *stream);
/* Fold the internal scratch into the bound gradient buffer; call once after the last step_bwd of a
 * backward pass. */
int leod_backbone_grads_finalize(leod_backbone_t *h, void *stream);
/* Pre-size the internal workspace for batches up to B (call outside CUDA-graph capture). */
int leod_backbone_reserve(leod_backbone_t *h, int B);
/* 0 = SIMT GEMMs, 1 = tcgen05/TMA GEMMs (LEOD_BF16 handles only; their default). */
int leod_backbone_set_gemm_impl(leod_backbone_t *h, int impl);

/* ------------------------------------------------------------------ building blocks (exported for tests)
 * C[M,N] = epilogue(A[M,K] * B[N,K]^T + bias).  A may be split in two sources along K
 * (A for k < K1, A2 for k >= K1; pass A2 = NULL, K1 = K otherwise).  dtype: operand/output type.
 * impl: 0 = SIMT fp32-accumulate kernel, 1 = tcgen05/TMA tensor-core kernel (LEOD_BF16 only).
 * epi: 0 none, 1 C = gelu(v), aux = gelu'(v)   2 C = R + v   3 C = v * aux  (the backward of 1, fed with its aux). */
int leod_gemm_nt(int impl, int dtype, const void *A, int lda, const void *A2, int lda2, int K1, const void *B, int ldb,
                 void *C, int ldc, int M, int N, int K, const float *bias, int epi, const void *R, int ldr, void *aux,
                 int ldaux, void *stream);
/* dW[N,K] (fp32, ld ldw) += dY[M,N]^T * X[M,K];  dbias[N] += column sums of dY (if non-NULL). */
int leod_gemm_tn(int impl, int dtype, const void *dY, int ldy, const void *X, int ldx, float *dW, int ldw, float *dbias,
                 int M, int N, int K, void *stream);

/* The stem convolution (maxvit.py:143-182: bias-free 7x7, stride 4, padding 3) applied to the dataloader's uint8 event tensor
 * without a patch matrix (implicit GEMM: operand tiles are built on chip).  x [nimg, Cin, xh, xw] uint8; the frame may be smaller
 * than 4*Ho x 4*Wo, the missing rows/columns are the zero padding of utils/padding.py:33-58.  Weights / weight gradients are in
 * patch order: k = (cin*7 + ky)*8 + slot, slot 0 <-> a zero weight, slot 1+kx <-> W[c, cin, ky, kx].
 *   fwd  : y[nimg*Ho*Wo, C] (bf16) = conv(x, W);  W_f16 [C, ldw >= Cin*56] IEEE half
 *   wgrad: dW[C, ldw] (fp32) += dY[nimg*Ho*Wo, C]^T (bf16) * patches(x)
 * Needs Ho % 8 == 0, Wo % 16 == 0, xw % 16 == 0, x 16-byte aligned, C % 16 == 0, C <= 64, 3 <= Cin <= 32 (else an error). */
int leod_stem_conv_fwd(const void *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *W_f16, int ldw, void *y,
                       void *stream);
int leod_stem_conv_wgrad(const void *x, int nimg, int Cin, int xh, int xw, int Ho, int Wo, int C, const void *dY, float *dW, int ldw,
                         void *stream);

/* Window / grid multi-head attention over an NHWC token matrix (maxvit.py:273-304, 343-354 minus the
 * two Linear layers).  qkv [B*H*W, 3C] with per-head [q|k|v] column blocks; out [B*H*W, C]. */
int leod_attention_fwd(int dtype, const void *qkv, void *out, int B, int H, int W, int C, int dim_head, int ph, int pw,
                       int window, void *stream);
int leod_attention_bwd(int dtype, const void *qkv, const void *dout, void *dqkv, int B, int H, int W, int C, int dim_head,
                       int ph, int pw, int window, void *stream);

/* LayerNorm over the channel dimension of a token matrix (layers/norm.py:44-56 -> F.layer_norm) and its
 * backward: dx = (dres ? dres : 0) + dLN(dy); dw += sum dy*xhat; db += sum dy (fp32 accumulators). */
int leod_layernorm_fwd(int dtype, const void *x, const float *w, const float *b, void *y, int M, int C, float eps,
                       void *stream);
int leod_layernorm_bwd(int dtype, const void *x, const float *w, const void *dy, const void *dres, void *dx, float *dw,
                       float *db, int M, int C, float eps, void *stream);
/* ConvLSTM gate math (models/layers/rnn.py:57-68).  gates [M,4C] = pre-activations (f|i|o|g) on entry,
 * activations on exit; c_prev may be NULL (zero state).  Backward: dh2 is an optional second gradient
 * w.r.t. h that is added to dh. */
int leod_lstm_gates_fwd(int dtype, void *gates, const void *c_prev, void *h_out, void *c_out, int M, int C, void *stream);
int leod_lstm_gates_bwd(int dtype, const void *gates, const void *c_prev, const void *c_out, const void *dh, const void *dh2,
                        const void *dc, void *dgates, void *dc_prev, int M, int C, void *stream);

/* ------------------------------------------------------------------ PAFPN neck + YOLOX head + SimOTA loss
 * Replaces models/detection/yolox_extension/models/yolo_pafpn.py:18-140 (YOLOPAFPN), models/detection/yolox/models/
 * network_blocks.py:29-142 (BaseConv / Bottleneck / CSPLayer), models/detection/yolox/models/yolo_head.py:21-332 (YOLOXHead
 * forward + decode) and :382-1148 (get_losses / get_assignments / SimOTA, plain and ignore-label variants), i.e. everything
 * behind YoloXDetector.forward_detect (models/detection/yolox_extension/models/detector.py:55-77). */
typedef struct leod_detect leod_detect_t;

typedef struct {
  int32_t in_channels[3];   /* backbone.get_stage_dims(fpn.in_stages), stride 8 / 16 / 32 level first to last */
  int32_t strides[3];       /* backbone.get_strides(fpn.in_stages): 8, 16, 32                                */
  int32_t num_classes;      /* head.num_classes (1..3)                                                        */
  int32_t n_bottleneck;     /* round(3 * fpn.depth): bottlenecks per CSPLayer                                 */
  int32_t in_h, in_w;       /* backbone.in_res_hw (level l is in_h/strides[l] x in_w/strides[l])              */
  int32_t dtype;            /* LEOD_F32 or LEOD_BF16: activation storage / GEMM operand type                  */
  float bn_eps, bn_momentum;/* nn.BatchNorm2d defaults 1e-5, 0.1 (network_blocks.py:46)                       */
  float ignore_label;       /* head.ignore_label (1024)                                                       */
  int32_t n_ignore_thresh;  /* len(head.ignore_bbox_thresh) or 0                                              */
  float ignore_thresh[8];
  float reg_weight, obj_weight, cls_weight;   /* 5, 1, 1 (yolo_head.py:67-69)                                 */
} leod_detect_cfg;

int leod_detect_create(const leod_detect_cfg *cfg, leod_detect_t **out);
/* Without any device allocation: only the *_info / *_count queries and _destroy are valid on the result. */
int leod_detect_layout_only(const leod_detect_cfg *cfg, leod_detect_t **out);
void leod_detect_destroy(leod_detect_t *h);
/* Flat buffers.  Parameters (fp32): entry names are the reference state_dict keys ("fpn.lateral_conv0.conv.weight",
 * "yolox_head.cls_preds.0.bias", ...).  Buffers (fp32): BatchNorm running_mean / running_var.  Counters (int64):
 * num_batches_tracked.  Each returns the number of entries when i < 0. */
int leod_detect_param_info(const leod_detect_t *h, int i, char *name, size_t name_cap, int64_t *offset, int32_t *ndim, int64_t shape[4]);
int leod_detect_buffer_info(const leod_detect_t *h, int i, char *name, size_t name_cap, int64_t *offset, int32_t *ndim, int64_t shape[4]);
int leod_detect_counter_info(const leod_detect_t *h, int i, char *name, size_t name_cap, int64_t *offset);
int64_t leod_detect_param_count(const leod_detect_t *h);
int64_t leod_detect_buffer_count(const leod_detect_t *h);
int64_t leod_detect_counter_count(const leod_detect_t *h);
int leod_detect_num_anchors(const leod_detect_t *h);
/* grads may be NULL for inference-only handles; counters may be NULL. */
int leod_detect_bind(leod_detect_t *h, float *params_dev, float *grads_dev, float *buffers_dev, int64_t *counters_dev);
/* Re-derive the operand-typed weight copies; call after every parameter update and before the next forward. */
int leod_detect_prepare(leod_detect_t *h, void *stream);
/* Pre-size the activation arena for up to B images (call outside CUDA-graph capture). */
int leod_detect_reserve(leod_detect_t *h, int B);
/* SyncBatchNorm (train.py:247): fn(ctx, buf, n, stream) must sum the n doubles at device pointer buf over all ranks, in place,
 * ordered on `stream`.  It is called once per dependency level of the network (forward: sums / sums of squares / counts,
 * backward: the two BatchNorm gradient sums), not once per layer.  NULL = single process. */
typedef int (*leod_allreduce_fn)(void *ctx, double *buf, int64_t n, void *stream);
int leod_detect_set_allreduce(leod_detect_t *h, leod_allreduce_fn fn, void *ctx);

/* yolo_pafpn.py:109-140 + yolo_head.py:195-287, 310-332.
 *  feats[l] : device, dense channels-last [B, h_l, w_l, in_channels[l]] in the handle's dtype
 *  training : != 0 -> BatchNorm batch statistics (+ running-statistics update) and activations kept for the backward
 *  preds    : device fp32 [B, A, 5 + num_classes] = decoded (cx, cy, w, h), sigmoid(obj), sigmoid(cls); A = leod_detect_num_anchors */
int leod_fpn_head_fwd(leod_detect_t *h, const void *const feats[3], int B, int training, float *preds, void *stream);
/* yolo_head.py:403-597 / 776-972 on the outputs of the preceding training-mode leod_fpn_head_fwd.
 *  labels     : device fp32 [B, nmax, 7] rows (cls, cx, cy, w, h, obj_conf, cls_conf), all-zero rows = padding
 *  losses_out : device fp32 [6] = loss, iou_loss, conf_loss, cls_loss, l1_loss (always 0), num_fg / num_gts */
int leod_simota_loss_fwd(leod_detect_t *h, const float *labels, int nmax, float *losses_out, void *stream);
/* Diagnostics: the assignment of the last leod_simota_loss_fwd.  assign_out: device int32 [B, A] label row per anchor (-1 =
 * background, yolo_head.py:768-774 matched_gt_inds); miou_out (NULL allowed): device fp32 [B, A] pred_ious_this_matching. */
int leod_simota_assignment(leod_detect_t *h, int32_t *assign_out, float *miou_out, void *stream);
/* d(gscale * loss)/d(raw head outputs); gscale: device fp32 scalar or NULL (= 1).  Result stays inside the handle. */
int leod_simota_loss_bwd(leod_detect_t *h, const float *labels, int nmax, const float *gscale, void *stream);
/* Raw (undecoded) head outputs / their gradients as dense device fp32 [B, A, 8] tensors (columns: reg x4, obj logit, class logits,
 * zero padding; yolo_head.py:236 `output` before get_output_and_grid).  _get_raw: outputs of the last forward.  _get_raw_grad: the
 * result of leod_simota_loss_bwd.  _set_raw_grad: replace it with the gradient of a caller-defined loss before leod_fpn_head_bwd. */
int leod_detect_get_raw(leod_detect_t *h, float *out, void *stream);
int leod_detect_get_raw_grad(leod_detect_t *h, float *out, void *stream);
int leod_detect_set_raw_grad(leod_detect_t *h, const float *in, void *stream);
/* Backward of the neck + head.  Parameter gradients are ACCUMULATED into the bound gradient buffer; dfeats[l] (same layout as
 * feats[l], NULL entries allowed) are written.  Fails if another forward overwrote the activations since the training forward. */
int leod_fpn_head_bwd(leod_detect_t *h, void *const dfeats[3], void *stream);

/* ------------------------------------------------------------------ detection post-processing
 * Replaces models/detection/yolox/utils/boxes.py:32-86 (postprocess) including the
 * torchvision.ops.batched_nms call at :73, batched over images with no host round trip.
 *  pred  : device fp32 [B, A, 5+num_classes] = (cx, cy, w, h, obj, cls...) — NOT modified
 *  out   : device fp32 [B, max_det, 7] rows (x1,y1,x2,y2,obj,cls_conf,cls_idx) in descending score order
 *  count : device int32 [B] number of valid rows per image
 * max_det <= A. */
int leod_postprocess(const float *pred, int B, int A, int num_classes, float conf_thre, float nms_thre,
                     int class_agnostic, float *out, int32_t *count, int max_det, void *stream);

/* Replaces modules/utils/ssod.py:147-188 (pred2label) + :113-133 (filter_pred_boxes) on the packed
 * output of leod_postprocess.  labels: device fp32 [B, max_det, 8] rows (t=0, x, y, w, h, cls,
 * cls_conf, obj_conf), compacted per image; lab_count int32 [B].  obj_thresh/cls_thresh: host arrays
 * of num_classes floats.  frame_h/frame_w <= 0 disables the box filter. */
int leod_pred2label(const float *dets, const int32_t *count, int B, int max_det, int num_classes,
                    const float *obj_thresh, const float *cls_thresh, int frame_h, int frame_w, float *labels,
                    int32_t *lab_count, void *stream);

/* Replaces modules/pseudo_labeler.py:37-91 (tta_postprocess, also modules/utils/tta.py:18-61): the second NMS over
 * the concatenated test-time-augmentation views of a frame, batched over frames.
 *  labels    : device fp32 [F, nmax, 8] ObjectLabels rows (t, x, y, w, h, cls, cls_conf, obj_conf), corner format
 *  count     : device int32 [F] valid rows per frame
 *  out/out_count : same layout; kept rows in descending obj*cls_conf order with w,h = (x+w)-x, (y+h)-y as the
 *              reference's xyxy round trip produces them.  Frames holding ground truth (any t > 0) are copied through. */
int leod_tta_merge(const float *labels, const int32_t *count, int F, int nmax, float conf_thre, float nms_thre,
                   int class_agnostic, float *out, int32_t *out_count, void *stream);

/* ------------------------------------------------------------------ pseudo-label sequence post-processing
 * Replaces EventSeqData._track / _track_filter (modules/pseudo_labeler.py:201-333) with modules/tracking/linear.py:10-292 and
 * modules/tracking/utils.py:7-96 behind it, for S sequences in one call (one CTA per sequence and direction), bit-exact.
 *  rows       : device fp32 [total_rows, 8] ObjectLabels rows (t, x, y, w, h, cls, cls_conf, obj) of every labelled frame, sequence after
 *               sequence, frames ascending;  frame_ptr: device int32 [total_frames + 1] first row of each labelled frame;  frame_idx:
 *               device int32 [total_frames] index of the frame inside its sequence;  seq_ptr: device int32 [S + 1] first labelled frame of
 *               each sequence;  hw: device int32 [S, 2] frame height / width
 *  qpow_host  : HOST doubles q^0 .. q^(npow-1) (npow > longest run of frames a track can age); q 0.9, min_conf 0.55, iou_thr 0.45,
 *               min_track_len 6 are the reference's defaults (linear.py:200-206, config/predict.yaml)
 *  use_backward : 'forward or backward' track_method — a box is ignored only if both directions put it on a short track
 *  hole_cap   : capacity for in-painted boxes per sequence;  ws: device scratch of leod_track_workspace_bytes()
 *  outputs    : out_rows + 8 * out_row_base[s] (capacity rows of s + hole_cap): the final rows, frame after frame, short-track boxes
 *               and in-painted boxes carrying ignore_label;  out_frame_idx / out_frame_start + out_frame_base[s]: frame index and first
 *               row of each output frame;  out_counts [S, 2] = rows, frames;  status [S]: 0 ok, 1 more than 64 live tracks or boxes per
 *               frame, 2 hole_cap exceeded, 3 qpow too short. */
int64_t leod_track_workspace_bytes(int64_t total_rows, int S, int hole_cap, int npow);
int leod_track_filter(const float *rows, const int32_t *frame_ptr, const int32_t *frame_idx, const int32_t *seq_ptr, const int32_t *hw,
                      int64_t total_rows, int total_frames, int S, const double *qpow_host, int npow, double q, double min_conf, float iou_thr,
                      int min_track_len, int inpaint, int use_backward, float ignore_label, int hole_cap, void *ws,
                      const int32_t *out_row_base, const int32_t *out_frame_base, float *out_rows, int32_t *out_frame_idx,
                      int32_t *out_frame_start, int32_t *out_counts, int32_t *status, void *stream);
/* ObjectLabels rows -> label-file records (data/genx_utils/labels.py:12-16 BBOX_DTYPE, :312-325 to_structured_array).
 * stride 40 = the declared dtype, 36 = the packed form EventSeqData._summarize (modules/pseudo_labeler.py:179-199) stores. */
int leod_pack_bbox(const float *rows, int64_t n, void *out, int stride, void *stream);

/* ------------------------------------------------------------------ input side: small uploads, augmentation
 * leod_upload_small: host->device copy of a small tensor (label rows, index lists, first-sample masks) done by an SM kernel that reads
 * the PINNED host buffer through its device mapping, so it never queues on the copy engine behind a bulk batch upload.  The host
 * buffer must stay valid until the stream has passed this point.  Replaces the `.to(device)` calls of modules/detection.py:134-147,
 * 226-236 (labels / selected indices) inside the step. */
int leod_upload_small(void *dst, const void *src_pinned_host, int64_t nbytes, void *stream);

/* Augmentation state of ONE sequence of the batch (data/utils/augmentor.py:26-60 AugmentationState after randomize_augmentation /
 * _zoom_in_and_rescale chose the window; rotation has probability 0 in every shipped config and is not supported).
 * The integer fields drive the event-tensor gather, the fp32 fields are the constants of the label arithmetic exactly as torch
 * rounds the reference's Python doubles when they meet a float32 tensor. */
typedef struct leod_augm_state {
  int32_t h_flip;           /* apply_h_flip (augmentor.py:404-410, th.flip over x) */
  int32_t t_flip;           /* DataType.IS_REVERSED (sequence_base.py:208-225): frames reversed in time AND channels reversed */
  int32_t zoom_mode;        /* 0 none, 1 zoom-in (augmentor.py:310-330), 2 zoom-out (:228-247) */
  int32_t x0, y0;           /* top-left corner of the zoom window */
  int32_t win_h, win_w;     /* int(H / factor), int(W / factor) */
  float flip_c;             /* W - 1 (labels.py:499-502) */
  float lo_x, hi_x, lo_y, hi_y; /* zoom-in clamp window z_x0, z_x1 - 1, z_y0, z_y1 - 1 (labels.py:389-397) */
  float mul;                /* scale_ multiplier (labels.py:482-497): zoom_in_factor, or 1 / zoom_out_factor */
  float cap_x, cap_y;       /* new_img_wd - 1, new_img_ht - 1 of that scale_ call */
} leod_augm_state;
#define LEOD_AUGM_MAX_SEQ 32   /* sequences per launch (larger batches are processed in several launches) */

/* in/out: device uint8 [L, B, C, H, W] (the loader's layout, data/utils/types.py EV_REPR stacked over time), out != in.
 * states: HOST array of B records.  hflip -> zoom-in | zoom-out in the reference's order (augmentor.py:457-478) plus the time flip,
 * nearest-exact resampling bit-identical to torch.nn.functional.interpolate(mode='nearest-exact'). */
int leod_augment_ev_repr(const void *in, void *out, int L, int B, int C, int H, int W, const leod_augm_state *states, void *stream);
/* rows: device fp32 [n, 8] ObjectLabels rows (t, x, y, w, h, cls, cls_conf, obj), transformed IN PLACE; row_seq: device int32 [n]
 * batch index of each row; keep: device uint8 [n], 0 = removed by remove_flat_labels_ (labels.py:67-69).  Bit-identical to
 * ObjectLabels.flip_lr_ / zoom_in_and_rescale_ / zoom_out_and_rescale_ (labels.py:371-411, 437-459, 482-502). */
int leod_augment_labels(float *rows, const int32_t *row_seq, int64_t n, int B, const leod_augm_state *states, uint8_t *keep, void *stream);

/* ------------------------------------------------------------------ evaluation
 * Prophesee box filter + COCO bounding-box evaluation of per-frame buffers, replacing evaluate_list
 * (utils/evaluation/prophesee/evaluation.py:5-42) -> filter_boxes (io/box_filtering.py:18-36) -> evaluate_detection
 * (metrics/coco_eval.py:32-120) -> pycocotools COCOeval.evaluate()/accumulate() as PropheseeEvaluator.evaluate_buffer
 * (utils/evaluation/prophesee/evaluator.py:73-110) drives them: frame f owns gt rows gt_ptr[f]..gt_ptr[f+1] and detection rows
 * dt_ptr[f]..dt_ptr[f+1] (all device arrays): t int64 (us), xywh fp32 [n,4] top-left corner format, cls int32, score fp32
 * (= class_confidence, coco_eval.py:176).  only_class >= 0 evaluates that class alone (evaluator.py:95-105), -1 all classes.
 * iou_thrs / rec_thrs: HOST arrays of 10 / 101 doubles (np.linspace(.5,.95,10), np.linspace(0,1,101) of pycocotools Params).
 * Outputs (device fp64): precision [10, 101, K, 4, 3] (-1 = undefined), recall [10, K, 4, 3] in COCOeval.eval's layout
 * (IoU thr, recall thr, class, area range all/small/medium/large, maxDets 1/10/100); counts int32 [3] = images, detections kept,
 * status (1: more than 128 gt boxes or 2048 detections of one class in a frame).  The summary numbers (AP, AP50, ...) are means
 * over these arrays (COCOeval.summarize), taken by the caller. */
int64_t leod_coco_eval_workspace_bytes(int F, int num_classes);
int leod_coco_eval(const int64_t *gt_t, const float *gt_xywh, const int32_t *gt_cls, const int32_t *gt_ptr, const int64_t *dt_t,
                   const float *dt_xywh, const int32_t *dt_cls, const float *dt_score, const int32_t *dt_ptr, int F, int num_classes,
                   int64_t skip_ts, int min_box_diag, int min_box_side, int only_class, const double *iou_thrs, const double *rec_thrs,
                   void *ws, double *precision, double *recall, int32_t *counts, void *stream);

/* ------------------------------------------------------------------ event binning
 * Replaces data/utils/representations.py:78-123 (StackedHistogram.construct).
 * x,y,p: device int32 [n]; t: device int64 [n] (sorted); out: device uint8 [2*bins, H, W].
 * fastmode != 0 reproduces the uint8 wrap-around accumulation. */
int leod_voxel_bin(const int32_t *x, const int32_t *y, const int32_t *p, const int64_t *t, int64_t n, int bins, int H,
                   int W, int count_cutoff, int fastmode, uint8_t *out, void *stream);

/* ------------------------------------------------------------------ optimizer / teacher update
 * Fused AdamW (modules/detection.py:485-518; torch.optim.AdamW semantics) with gradient
 * clip-by-value (train.py:236-237; clip_value <= 0 disables) over a flat fp32 buffer, `step` 1-based.
 * If ema != NULL also applies modules/utils/ssod.py:429-438: ema = a*ema + (1-a)*p with the caller's a. */
int leod_adamw_ema(float *p, const float *g, float *m, float *v, float *ema, int64_t n, int step, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float clip_value, float ema_alpha, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LEOD_B200_H_ */
