/* dadetect_b200 — C ABI of the B200 (sm_100a) DA Faster R-CNN training hot path.
 *
 * Drop-in boundary: the reference reaches native code through the pybind
 * module `maskrcnn_benchmark._C` (maskrcnn_benchmark/csrc/vision.cpp:7-15) and through
 * ATen/cuDNN/cuBLAS library calls made by its Python layers.  This header is the C-ABI
 * replacement for BOTH: every entry point takes plain device pointers, sizes and a CUDA
 * stream (as void*), no torch types.  The Python binding (`da-detect_b200/_lib.py`, ctypes)
 * and the `_C`-compatible shim (`da-detect_b200/_C.py`) sit on top; INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - all pointers are device pointers unless named h_*; float = IEEE fp32;
 *   - activations are NHWC ([N][H][W][C], C fastest); conv weights are OHWI
 *     ([Cout][KH][KW][Cin]); the *_nchw entry points accept the reference's NCHW layout;
 *   - boxes are xyxy in pixels with the reference's inclusive "+1" width convention;
 *   - every function returns 0 on success, otherwise a cudaError_t value (or -1 for an
 *     argument error); dd_last_error() returns a static description of the last failure;
 *   - work is enqueued on `stream` and is asynchronous w.r.t. the host unless stated.
 */
#ifndef DADETECT_B200_H_
#define DADETECT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* dd_last_error(void);
int dd_abi_version(void);
/* number of kernels launched by this library since process start (bench.py's gpu_launches) */
long long dd_launch_count(void);
/* Limits the persistent dense kernels (conv / dgrad / wgrad on tcgen05) launched from now on to `sms` CTAs (0 or
 * >= 148: all SMs); returns the previous limit.  Host-side state read at launch time, so inside a stream capture
 * it is baked into the captured launches.  Used while few-CTA latency-bound kernels (top-k, NMS scan, samplers)
 * run beside dense work on another stream. */
int dd_set_sm_budget(int sms);
/* 1 when the tcgen05/TMA arm of the dense tier is compiled into this library, else 0 */
int dd_tcgen05_built(void);

/* ---------------------------------------------------------------- maskrcnn_benchmark._C ops */

/* Replaces ROIAlign_forward (csrc/ROIAlign.h:11-25 -> cuda/ROIAlign_cuda.cu:64-122,257-299).
 * feat [N,H,W,C] NHWC, rois [K,5] = (batch_idx, x1, y1, x2, y2), out [K, PH/bin_step, PW/bin_step, C].
 * bin_step = 1 is the reference op.  bin_step = 2 computes only the even (ph,pw) bins of the
 * PH x PW bin geometry — exactly what res5's stride-2 1x1 convs consume. */
int dd_roi_align_forward(const float* feat, const float* rois, float* out, int N, int H, int W, int C,
                         int K, float spatial_scale, int PH, int PW, int sampling_ratio, int bin_step,
                         void* stream);
/* Replaces ROIAlign_backward (csrc/ROIAlign.h:27-45 -> cuda/ROIAlign_cuda.cu:177-254,302-346).
 * grad_feat [N,H,W,C] must be zero-filled by the caller (as the reference does, :316) or hold a
 * gradient to accumulate into. */
int dd_roi_align_backward(const float* grad_out, const float* rois, float* grad_feat, int N, int H, int W,
                          int C, int K, float spatial_scale, int PH, int PW, int sampling_ratio,
                          int bin_step, void* stream);
/* NCHW wrappers with the reference's exact tensor layout ([N,C,H,W] in, [K,C,PH,PW] out). */
int dd_roi_align_forward_nchw(const float* feat, const float* rois, float* out, int N, int C, int H, int W,
                              int K, float spatial_scale, int PH, int PW, int sampling_ratio, void* stream);
int dd_roi_align_backward_nchw(const float* grad_out, const float* rois, float* grad_feat, int N, int C,
                               int H, int W, int K, float spatial_scale, int PH, int PW, int sampling_ratio,
                               void* stream);

/* Replaces nms (csrc/nms.h:10-28 -> cuda/nms.cu:70-131), GPU semantics: suppress when IoU > thresh.
 * boxes [n,4], scores [n] in ANY order; keep_out int64[n] receives the kept ORIGINAL indices in
 * ascending order, *keep_count (device int) their number.  Entirely on device: no D2H mask copy,
 * no host scan.  n <= 16384.  workspace: dd_nms_workspace_bytes(n) bytes. */
size_t dd_nms_workspace_bytes(int n);
int dd_nms(const float* boxes, const float* scores, int n, float thresh, int64_t* keep_out, int* keep_count,
           void* workspace, void* stream);
/* Same, for boxes already sorted by descending score (the RPN path): keep_out holds positions in
 * that order (ascending).  max_keep > 0 stops after that many kept boxes — identical to
 * boxlist_nms(max_proposals) (structures/boxlist_ops.py:30-33) because positions ARE score order. */
int dd_nms_sorted(const float* boxes_sorted, int n, float thresh, int max_keep, int64_t* keep_out,
                  int* keep_count, void* workspace, void* stream);
/* The same for a batch of images in two launches: boxes_sorted [images, n_cap, 4], image g holds n_dev[g]
 * (device int32, <= n_cap) boxes; keep_out + g * keep_stride receives image g's kept positions, keep_count[g]
 * their number.  No host knowledge of the counts is needed (rpn/inference.py:87-127 loops over images on the
 * host and synchronises per image).  workspace: dd_nms_batched_workspace_bytes(images, n_cap). */
size_t dd_nms_batched_workspace_bytes(int images, int n_cap);
int dd_nms_sorted_batched(const float* boxes_sorted, const int* n_dev, int images, int n_cap, float thresh,
                          int max_keep, int64_t* keep_out, int keep_stride, int* keep_count, void* workspace,
                          void* stream);

/* ---------------------------------------------------------------- RPN proposal generation */

/* AnchorGenerator.grid_anchors + add_visibility_to (rpn/anchor_generator.py:73-111):
 * anchors [FH*FW*A,4] in (y, x, a) order, visibility uint8. */
int dd_anchor_grid(const float* cell_anchors, int A, int FH, int FW, int stride, int img_w, int img_h,
                   int straddle_thresh, float* anchors, uint8_t* visibility, void* stream);
/* RPNPostProcessor.forward_for_single_feature_map up to (not including) NMS
 * (rpn/inference.py:87-115): per image sigmoid -> top-k (sorted, ties by lower index) ->
 * gather deltas+anchors -> BoxCoder.decode(1,1,1,1) -> clip_to_image -> small-box filter.
 * logits [N,FH,FW,A] NHWC, deltas [N,FH,FW,4A]; boxes [N,k,4], scores [N,k], topk_idx int32[N,k];
 * valid[N] = boxes surviving the min_size filter (compacted to the front, order preserved).
 * k <= 16384.  workspace: dd_rpn_topk_workspace_bytes(N, FH*FW*A). */
size_t dd_rpn_topk_workspace_bytes(int N, int num_anchors);
int dd_rpn_topk_decode(const float* logits, const float* deltas, const float* anchors, int N, int FH, int FW,
                       int A, int k, int img_w, int img_h, float min_size, float* boxes, float* scores,
                       int32_t* topk_idx, int32_t* valid, void* workspace, void* stream);

/* ---------------------------------------------------------------- matching / box coding */

/* boxlist_iou + Matcher (structures/boxlist_ops.py:56-91, modeling/matcher.py:42-112) fused: never
 * materialises the [M,N] matrix.  gt [M,4], pred [N,4]; matches int64[N] in {-2,-1,0..M-1};
 * matched_vals float[N] (may be NULL).  gt_best is workspace float[M].
 * m_dev (may be NULL): device int32[1], the number of live GT rows when gt is padded to the capacity M (a
 * training step captured once as a CUDA graph then serves batches with any number of boxes per image,
 * data/datasets/coco.py:96-97); rows at and beyond *m_dev do not exist for the Matcher. */
int dd_match(const float* gt, int M, const int32_t* m_dev, const float* pred, int N, float high, float low,
             int allow_low_quality, int64_t* matches, float* matched_vals, float* gt_best, void* stream);
/* RPN anchor labels (rpn/loss.py:57-89) from dd_match's result and the anchor visibility mask (uint8 [N]):
 * 1 matched, 0 background, -1 ignored (between the thresholds, or straddling the image border). */
int dd_rpn_anchor_labels(const int64_t* matches, const uint8_t* visibility, int N, int32_t* labels, void* stream);
/* RPNLossComputation's loss tail (rpn/loss.py:118-141) over the anchors sampled by dd_balanced_sample for the S source
 * images of a batch, forward and gradient in one launch: losses[0] = BCE-with-logits / #sampled, losses[1] =
 * smooth-L1(beta) over the sampled positives against BoxCoder(1,1,1,1) targets / #sampled.  logits [n_img*A],
 * deltas [n_img*A,4] (rows s*A + a belong to source image s), anchors [A,4], sel int64 [S,B], counts int32 [S,2],
 * labels int32 [S,A], matches int64 [S,A], gt_cat [G,4] with gt_offsets int32 [n_img+1], src_img int32 [S] (batch
 * index of each source image).  dlogits / ddeltas: dense gradients, zero-filled by the caller. */
int dd_rpn_sampled_losses(const float* logits, const float* deltas, const float* anchors, int A, int S, int B,
                          const int64_t* sel, const int32_t* counts, const int32_t* labels, const int64_t* matches,
                          const float* gt_cat, const int32_t* gt_offsets, const int32_t* src_img, float beta,
                          float* losses, float* dlogits, float* ddeltas, void* stream);
/* Box-head proposal labels of one image (box_head/loss.py:55-99) from dd_match's result (matches int64 [cap]) and
 * the image's ground-truth classes (int64): class of the match, 0 background, -1 ignored (between the thresholds, or
 * a buffer row at / beyond *n_prop); is_source == 0: every proposal is background (:84-85). */
int dd_roi_labels(const int64_t* matches, const int64_t* gt_labels, int is_source, const int32_t* n_prop, int cap,
                  int32_t* labels, void* stream);
/* The sampled ROIs of a batch after dd_balanced_sample (box_head/loss.py:100-130): boxes [n_img,cap,4],
 * objectness [n_img,cap], sel int64 [n_img,B], counts int32 [n_img,2], labels int32 / matches int64 [n_img,cap],
 * gt_cat [G,4] + gt_offsets int32 [n_img+1] (+ gt_counts int32 [n_img] or NULL: live rows of a padded GT buffer),
 * is_source uint8 [n_img] -> rois [n_img*B,5] (batch index, box), labels int64 (0 in slots beyond the sampled
 * count), BoxCoder(wx,wy,ww,wh) regression targets [n_img*B,4] (negative matches wrap for target-domain images,
 * :47-51), domain / valid uint8 [n_img*B], objectness [n_img*B]. */
int dd_roi_gather_sampled(const float* boxes, const float* objectness, const int64_t* sel, const int32_t* counts,
                          const int32_t* labels, const int64_t* matches, const float* gt_cat,
                          const int32_t* gt_offsets, const int32_t* gt_counts, const uint8_t* is_source, int n_img,
                          int cap, int B, float wx, float wy, float ww, float wh, float* rois, int64_t* out_labels,
                          float* reg_targets, uint8_t* domain, uint8_t* valid, float* out_objectness, void* stream);
/* BoxCoder.encode (box_coder.py:22-50) of gt[matches[i] clamped at 0] against pred[i];
 * wrap_negative != 0 reproduces the reference's negative-index wrap for target images
 * (box_head/loss.py:47-51) — relative to the live row count (*m_dev when given, else M). */
int dd_box_encode(const float* gt, int M, const int32_t* m_dev, const float* pred, const int64_t* matches, int N,
                  float wx, float wy, float ww, float wh, int wrap_negative, float* targets, void* stream);
/* BoxCoder.decode (box_coder.py:52-95) for codes [R, 4*k] against boxes [R,4]. */
int dd_box_decode(const float* codes, const float* boxes, int R, int k, float wx, float wy, float ww, float wh,
                  float* out, void* stream);

/* ---------------------------------------------------------------- device-resident proposals / sampling */

/* RPNPostProcessor tail (rpn/inference.py:116-127 + add_gt_proposals :51-74) without host reads: image g keeps
 * its first min(keep_count[g], post) NMS survivors boxes[g, keep[g, r]] and, when append_gt[g] != 0, is extended
 * with its ground-truth boxes gt[gt_offsets[g] .. gt_offsets[g+1]) (objectness 1) — only the first gt_counts[g]
 * of them when gt_counts (device int32 [N], may be NULL) is given: GT rows padded to a fixed capacity.
 * out_boxes [N, cap, 4], out_objectness [N, cap] (rows beyond out_count[g] are zero). */
int dd_proposals_gather(const float* boxes, const float* scores, const int64_t* keep, const int* keep_count,
                        const float* gt, const int* gt_offsets, const int* gt_counts, const uint8_t* append_gt, int N,
                        int k, int post, int cap, float* out_boxes, float* out_objectness, int* out_count,
                        void* stream);
/* BalancedPositiveNegativeSampler (balanced_positive_negative_sampler.py:27-76) per image on the device:
 * labels int32 [images, n_cap] (>= 1 positive, 0 negative, < 0 ignored), image g has n_dev[g] candidates
 * (n_dev NULL: n_cap), keys float [images, n_cap] the random draw.  Chooses min(#pos, max_pos) positives and
 * min(#neg, batch - that) negatives with the smallest keys (ties: lower index).  sel_idx int64 [images, batch]
 * = chosen candidate indices in ascending order (rows beyond the total are 0); counts int32 [images, 2] =
 * {positives chosen, total chosen}. */
int dd_balanced_sample(const int* labels, const int* n_dev, const float* keys, int images, int n_cap, int batch,
                       int max_pos, int64_t* sel_idx, int* counts, void* stream);

/* ---------------------------------------------------------------- dense layers (implicit GEMM) */

/* impl: 0 = fp32 SIMT tiles (bit-faithful fp32 accumulate), 1 = tcgen05 TF32 (TMA-staged, TMEM accum). */
#define DD_IMPL_SIMT 0
#define DD_IMPL_TCGEN05 1
/* 2 = tcgen05 "3xTF32": every operand is split into TF32-exact high and low parts inside the kernel and
 * D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi — fp32-grade products on the tensor cores (forward, data and weight
 * gradient). */
#define DD_IMPL_TCGEN05_X3 2
#define DD_ACT_NONE 0
#define DD_ACT_RELU 1

/* y = act( conv(x, w) * scale[co] + bias[co] + residual ).  Conv2d + FrozenBatchNorm2d (+ residual
 * add + ReLU) of resnet.py:294-314 / batch_norm.py:19-24 in one kernel; scale/bias/residual may be
 * NULL.  x [N,H,W,Cin], w [Cout,KH,KW,Cin], y [N,OH,OW,Cout], OH = (H+2p-KH)/s+1. */
int dd_conv2d_forward(const float* x, const float* w, const float* scale, const float* bias,
                      const float* residual, float* y, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                      int stride, int pad, int act, int impl, void* workspace, void* stream);
/* 3xTF32 arm with the hi / lo weight planes prepared beforehand: dd_conv2d_forward_prepare_batch splits the weights of
 * n layers (workspaces[i] of dd_conv2d_forward_workspace_bytes bytes each) in one launch per 32 layers — e.g. once per
 * training step for a whole model — and dd_conv2d_forward_prepared is dd_conv2d_forward minus its per-call split. */
int dd_conv2d_forward_prepare_batch(int n, const float* const* w, void* const* workspaces, const int* Cin,
                                    const int* Cout, const int* KH, const int* KW, int impl, void* stream);
int dd_conv2d_forward_prepared(const float* x, const float* w, const float* scale, const float* bias,
                               const float* residual, float* y, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                               int stride, int pad, int act, int impl, void* workspace, void* stream);
/* workspace: dd_conv2d_forward_workspace_bytes(...) bytes (0 unless impl == DD_IMPL_TCGEN05_X3: the hi / lo
 * planes of the weights), 16-byte aligned; may be NULL when 0. */
size_t dd_conv2d_forward_workspace_bytes(int Cin, int Cout, int KH, int KW, int impl);
/* The ResNet stem (BaseStem, resnet.py:317-336: 7x7 stride-2 pad-3 conv of the 3-channel image + FrozenBN
 * + ReLU) on the tensor cores, reading the reference's NCHW image directly: x_nchw [N,3,H,W] (H, W even),
 * w [Cout,7,7,3] (OHWI), y [N,H/2,W/2,Cout] NHWC, Cout <= 64.  workspace: dd_stem_workspace_bytes(...) bytes,
 * 256-byte aligned (zero-haloed NHWC4 copy of the image + re-packed weights). */
size_t dd_stem_workspace_bytes(int N, int H, int W, int Cout);
int dd_stem_conv7x7s2_forward(const float* x_nchw, const float* w_ohwi, const float* scale, const float* bias,
                              float* y, int N, int H, int W, int Cout, int act, int impl, void* workspace,
                              void* stream);
/* gx = conv_transpose(gy, w * scale[co]) (+ addend) (* (mask_act > 0) if mask_act).  gx [N,H,W,Cin]
 * is fully written (positions a strided conv never read receive addend or 0).
 * workspace: dd_conv2d_dgrad_workspace_bytes(...) bytes (16-byte aligned) — the tcgen05 arm keeps the
 * BN-scaled, tap-flipped, transposed weights W'[ci,kh',kw',co] there.  prepared != 0: the workspace already
 * holds W' for this (w, scale) from an earlier call and the preparation kernel is skipped. */
size_t dd_conv2d_dgrad_workspace_bytes(int Cin, int Cout, int KH, int KW);
int dd_conv2d_dgrad(const float* gy, const float* w, const float* scale, const float* addend,
                    const float* mask_act, float* gx, int N, int H, int W, int Cin, int Cout, int KH, int KW,
                    int stride, int pad, int impl, void* workspace, int prepared, void* stream);
/* Prepare the dgrad workspaces of n layers at once (one launch per 16 layers instead of one per layer); layer i
 * then calls dd_conv2d_dgrad with workspaces[i] and prepared = 1.  A no-op for the SIMT arm. */
int dd_conv2d_dgrad_prepare_batch(int n, const float* const* w, const float* const* scale, void* const* workspaces,
                                  const int* Cin, const int* Cout, const int* KH, const int* KW, int impl,
                                  void* stream);
/* gw[co,kh,kw,ci] (+)= scale[co] * sum_{n,oh,ow} gy[n,oh,ow,co] * x[n,oh*s+kh-p,ow*s+kw-p,ci].
 * workspace: dd_conv2d_wgrad_workspace_bytes(...) (split-K partials). */
size_t dd_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int dd_conv2d_wgrad(const float* gy, const float* x, const float* scale, float* gw, int N, int H, int W, int Cin,
                    int Cout, int KH, int KW, int stride, int pad, int accumulate, int impl, void* workspace,
                    void* stream);
/* gb[c] (+)= sum over rows of gy [rows, C]. */
int dd_bias_grad(const float* gy, float* gb, int rows, int C, int accumulate, void* stream);

/* ---------------------------------------------------------------- pooling / elementwise */
int dd_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, void* stream);
int dd_nhwc_to_nchw(const float* x, float* y, int N, int C, int H, int W, void* stream);
/* F.max_pool2d(k=3, s=2, p=1) of the stem (resnet.py:335), NHWC; forward only (stem is frozen). */
int dd_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, void* stream);
/* nn.AvgPool2d(7) on [K,HW,C] -> [K,C] and its backward. */
int dd_avgpool_forward(const float* x, float* y, int K, int HW, int C, void* stream);
int dd_avgpool_backward(const float* gy, float* gx, int K, int HW, int C, void* stream);
/* AvgPool2d backward fused with the ReLU mask of the pooled map: gx[k,i,c] = act[k,i,c] > 0 ? gy[k,c] / HW : 0
 * (the 7x7 average of the res5 output feeds the predictors; its gradient enters res5 through that ReLU). */
int dd_avgpool_relu_backward(const float* gy, const float* act, float* gx, int K, int HW, int C, void* stream);
/* out = g * (act > 0 ? 1 : 0); ReLU backward applied to a gradient that fans in from several consumers. */
int dd_relu_backward(const float* g, const float* act, float* out, long long n, void* stream);
/* GradientScalarLayer backward (layers/gradient_scalar_layer.py:11-13): out = w * g (no clone);
 * accumulate != 0: out += w * g. */
int dd_grl_backward(const float* g, float w, float* out, long long n, int accumulate, void* stream);
/* Same with the weight read from device memory at execution time (AdvGRL: the weight depends on a loss
 * value that is never brought to the host). */
int dd_grl_backward_dev(const float* g, const float* w_dev, float* out, long long n, int accumulate, void* stream);
/* AdvGRL weight (da_heads/da_heads.py:173-195; intent per README of the reference):
 * *w_out = (*loss <= bce) ? -lam_adv * min(threshold, 1 / *loss) : -lam.  loss, w_out: device float[1]. */
int dd_adv_grl_weight(const float* loss, float bce, float lam, float lam_adv, float threshold, float* w_out,
                      void* stream);
/* x * keep * 2 (F.dropout p=0.5 with a caller-supplied keep mask, da_heads.py:63,65); same op is its backward. */
int dd_dropout_apply(const float* x, const float* keep, float* out, long long n, void* stream);

/* ---------------------------------------------------------------- losses (fused forward + gradient) */

/* F.binary_cross_entropy_with_logits(x, t) mean over n.  Targets: `targets` float[n] if non-NULL, else
 * segment labels: element i belongs to segment i / seg_len and takes seg_labels[segment] (uint8) — the
 * per-image domain label of da_heads/loss.py:153-167.  loss: device float[1]; grad float[n] = dL/dx. */
int dd_bce_logits_mean(const float* x, const float* targets, const uint8_t* seg_labels, long long seg_len,
                       long long n, float* loss, float* grad, void* stream);
/* F.cross_entropy(logits[rows,C], labels) mean over rows with row_mask != 0 (box_head/loss.py:193-200).
 * grad [rows,C] (zero for masked-out rows). */
int dd_softmax_ce_mean(const float* logits, const int64_t* labels, const uint8_t* row_mask, int rows, int C,
                       float* loss, float* grad, void* stream);
/* smooth_l1_loss(x, t, beta, size_average=False) / divisor (layers/smooth_l1_loss.py:6-16). */
int dd_smooth_l1_sum(const float* x, const float* t, long long n, float beta, float divisor, float* loss,
                     float* grad, void* stream);
/* Box-head regression loss (box_head/loss.py:202-219): for rows with mask && label>0 gathers the 4
 * columns 4*label..4*label+3 of box_reg [rows, 4*C], smooth-L1(beta=1) sum / (#masked rows). */
int dd_box_reg_loss(const float* box_reg, const float* reg_targets, const int64_t* labels, const uint8_t* row_mask,
                    int rows, int C, float* loss, float* grad, void* stream);
/* consistency_loss (layers/consistency_loss.py:3-27) on LOGITS: img_logits [2, hw], ins_logits [K];
 * first n_src ROIs belong to image 0.  Returns loss and gradients w.r.t. both logit tensors
 * (sigmoid folded in). */
int dd_consistency_loss(const float* img_logits, long long hw, const float* ins_logits, int K, int n_src,
                        const uint8_t* row_valid, float* loss, float* grad_img, float* grad_ins, float* workspace2,
                        void* stream);
/* row_valid (nullable, uint8 [K]): ROIs of the fixed-capacity layout that do not exist are skipped and the
 * mean runs over the existing ones. */
/* nn.TripletMarginLoss(margin, p=2) (da_heads/loss.py:198-200): a,p,n [rows, D] with the norm over D
 * computed on rows gathered with element stride `inner` (NHWC feature maps: D = W taken with stride C);
 * see SURVEY §2.3/§9.8.  rows = number of distance vectors.  Layout: element (r, d) of a tensor lives at
 * base[(r / inner) * D * inner + d * inner + (r % inner)].  grads may be NULL.
 * margin_dev (may be NULL): the margin is read from device memory instead (the adaptive margin below). */
int dd_triplet_margin_loss(const float* a, const float* p, const float* n, long long rows, int D, long long inner,
                           float margin, const float* margin_dev, float* loss, float* grad_a, float* grad_p,
                           float* grad_n, void* stream);
/* The adaptive margin of DALossComputation_Component.triplet_img_loss (da_heads/loss.py:182-200) kept on the
 * device: state (double[1], the reference's Python float; 0 = not yet initialised) becomes margin_cfg when it is
 * 0 and grows by lr when the PREVIOUS step's triplet loss (prev_loss, device float[1], may be NULL) was exactly 0
 * and int(state) != int(max_margin); margin_out (float[1]) receives the value the loss kernel consumes.  Replaces
 * the reference's host read of the previous loss (da_heads.py:325 `.cpu()`), so that the step stays capturable. */
int dd_adaptive_margin_update(double* state, const float* prev_loss, double margin_cfg, double lr, double max_margin,
                              float* margin_out, void* stream);

/* ---------------------------------------------------------------- optimiser */
/* torch.optim.SGD step over a flat segment: g' = g*grad_scale + wd*p; buf = momentum*buf + g'; p -= lr*buf
 * (solver/build.py:7-20; first step semantics buf = g' selected by first_step != 0). */
int dd_sgd_momentum(float* p, const float* g, float* buf, long long n, float lr, float momentum, float wd,
                    float grad_scale, int first_step, void* stream);
/* The same with lr = lr_factor * (*lr_dev) read on the device, so that a captured CUDA graph of the whole step
 * follows the LR schedule; buf must be zero before the first step (then buf = g' falls out of the recurrence). */
int dd_sgd_momentum_dev(float* p, const float* g, float* buf, long long n, const float* lr_dev, float lr_factor,
                        float momentum, float wd, float grad_scale, void* stream);

/* ---------------------------------------------------------------- FPN */

/* F.interpolate(x, scale_factor=2, mode="nearest") (modeling/backbone/fpn.py:62): x [N,H,W,C] -> y [N,2H,2W,C];
 * backward: gx [N,H,W,C] = 2x2 block sums of gy [N,2H,2W,C].  C % 4 == 0, 16-byte aligned pointers. */
int dd_upsample2x_forward(const float* x, float* y, int N, int H, int W, int C, void* stream);
int dd_upsample2x_backward(const float* gy, float* gx, int N, int H, int W, int C, void* stream);
/* LastLevelMaxPool = F.max_pool2d(x, 1, 2, 0) (fpn.py:80-82): x [N,H,W,C] -> y [N,(H-1)/2+1,(W-1)/2+1,C]; backward
 * writes EVERY element of gx [N,H,W,C] (gy at the even pixels, zeros elsewhere). */
int dd_subsample2_forward(const float* x, float* y, int N, int H, int W, int C, void* stream);
int dd_subsample2_backward(const float* gy, float* gx, int N, int H, int W, int C, void* stream);
/* LevelMapper (modeling/poolers.py:11-42): levels[k] = clamp(floor(canonical_level + log2(sqrt(area_k) /
 * canonical_scale + eps)), k_min, k_max) - k_min for rois [K,5] = (batch_idx, x1, y1, x2, y2), +1 areas. */
int dd_fpn_level_map(const float* rois, int K, int k_min, int k_max, float canonical_scale, int canonical_level,
                     float eps, int* levels, void* stream);
/* Multi-level Pooler.forward (poolers.py:104-121) without index lists: one call per feature level over ALL K rois;
 * only rois with roi_level[k] == level are pooled from this map, into their own rows of out [K,PH,PW,C] (the other
 * rows are left untouched — after the calls for every level each row has been written exactly once).  Backward
 * likewise scatters only those rois' gradients into this level's grad_feat (zero-filled by the caller). */
int dd_roi_align_level_forward(const float* feat, const float* rois, const int* roi_level, int level, float* out,
                               int N, int H, int W, int C, int K, float spatial_scale, int PH, int PW,
                               int sampling_ratio, void* stream);
int dd_roi_align_level_backward(const float* grad_out, const float* rois, const int* roi_level, int level,
                                float* grad_feat, int N, int H, int W, int C, int K, float spatial_scale, int PH,
                                int PW, int sampling_ratio, void* stream);

/* ---------------------------------------------------------------- input pipeline */

/* Pillow's bilinear resampling coefficients for one axis (Resample.c precompute_coeffs + normalize_coeffs_8bpc,
 * the arithmetic behind torchvision F.resize on a PIL image — data/transforms/transforms.py:65-69).  HOST
 * functions, no CUDA work: h_bounds int[out_size][2] = (first source index, count), h_kk int[out_size][ksize]
 * fixed-point weights with 22 fractional bits.  dd_resample_ksize returns ksize (-1 on bad sizes). */
int dd_resample_ksize(int in_size, int out_size);
int dd_resample_coeffs(int in_size, int out_size, int* h_bounds, int* h_kk);
/* One sample of Compose([Resize, RandomHorizontalFlip, ToTensor, Normalize]) (data/transforms/build.py:22-31,
 * transforms.py:35-98) fused with the zero padding of to_image_list (structures/image_list.py:66-88), one launch:
 * src uint8 [src_h][src_w] RGB pixels of pixel_stride (3 or 4) bytes, rows row_stride bytes apart (device memory)
 * -> dst float [3][Hp][Wp], the image's slot of the batch tensor: rows < out_h / columns < out_w hold the resized
 * (two integer passes, horizontal first, uint8 intermediate: bit-exact Pillow), optionally mirrored image as
 * ((u8 / 255) [* 255, channels reversed when to_bgr255] - mean) / std in correctly rounded fp32 steps, the rest
 * zeros.  xbounds/xk/ybounds/yk: DEVICE copies of dd_resample_coeffs(src_w, out_w) / (src_h, out_h);
 * h_mean / h_std: 3 host floats each, in output-plane order. */
int dd_preprocess_image(const uint8_t* src, int src_h, int src_w, int pixel_stride, long long row_stride,
                        const int* xbounds, const int* xk, int xksize, const int* ybounds, const int* yk, int yksize,
                        int out_h, int out_w, int flip, int to_bgr255, const float* h_mean, const float* h_std,
                        float* dst, int Hp, int Wp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DADETECT_B200_H_ */
