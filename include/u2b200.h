/* libu2b200.so — C ABI of the B200-native U2Seg hot paths (sm_100a only).
 *
 * The reference (u2seg/U2Seg, a Detectron2 fork) has no native boundary on these paths: every
 * device op is reached through torch / torchvision / pykeops Python calls. Each entry point below
 * names the reference call site (file:line under the reference root) it replaces; the Python
 * host code in u2seg_b200/ mirrors those call sites' signatures and calls these functions with
 * raw device pointers (tensor.data_ptr()), shapes and the caller's CUDA stream.
 *
 * Conventions
 *  - return 0 on success; >0 = cudaError_t; <0 = library error (U2B_ERR_*). u2b_last_error()
 *    returns a thread-local message. No C++ exception crosses the boundary.
 *  - the caller owns every buffer (inputs, outputs, workspace); the library never allocates,
 *    frees or retains device pointers. All work is enqueued on `stream`; no hidden sync.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef U2B200_H_
#define U2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U2B200_VERSION 100

typedef struct CUstream_st* u2b_stream_t; /* == cudaStream_t */

const char* u2b_last_error(void);
int u2b_version(void);
int u2b_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * k-means (Lloyd) — u2seg/Instance_Clustering/shared/utils/nn_utils.py:304-379 (KMeans)
 * x16: (N, D) fp16 row-major embeddings, D % 64 == 0, D <= 384. Centroids c32: (K, D) fp32.
 * ------------------------------------------------------------------------------------------ */

/* rows of the padded fp16 centroid operand (multiple of the 160-wide accumulator tile) */
int64_t u2b_kmeans_kpad(int64_t K);
size_t u2b_kmeans_workspace_bytes(int64_t N, int64_t D, int64_t K);

/* max_i |x_i| -> *xmax (device float). Once per data set; feeds the refinement bound. */
int u2b_kmeans_xnorm_max(const void* x16, int64_t N, int64_t D, float* xmax, u2b_stream_t stream);

/* c32 -> c16 (kpad x D fp16, zero rows beyond K), cnorm (kpad fp32, +inf beyond K),
 * *cmax2 = max finite |c|^2 (device float). Run before every E-step. */
int u2b_kmeans_prepare(const float* c32, int64_t K, int64_t D, void* c16, float* cnorm,
                       float* cmax2, u2b_stream_t stream);

/* E-step, nn_utils.py:353-355: labels[i] = argmin_j sum_d (x_i - c_j)^2 (first minimum; a NaN
 * centroid is never selected). tcgen05 fp16 pass + exact fp32 refinement of undecidable rows.
 * amb_count_out (device int, may be NULL) receives the number of refined rows. */
int u2b_kmeans_assign(const void* x16, int64_t N, int64_t D, int64_t K, const void* c16,
                      const float* c32, const float* cnorm, const float* xmax, const float* cmax2,
                      int32_t* labels, int32_t* amb_count_out, void* workspace,
                      size_t workspace_bytes, u2b_stream_t stream);

/* thread-block cluster size of the E-step kernel (1, 2 or 4): centroid tiles are TMA-multicast across the cluster */
int u2b_kmeans_set_cluster(int cluster);
/* M-step implementation: 1 (default) = rows counting-sorted by label, then atomics-free segment sums (X read once, at
 * whole-row granularity); 0 = round-1 shared-memory accumulators (fp32 shared atomics). Same results up to fp32
 * summation order. */
int u2b_kmeans_set_mstep(int sort_by_label);

/* M-step part 1, nn_utils.py:359-363: sums[k, 0:D] = sum of rows with label k, sums[k, D] =
 * count (fp32). (K, D+1) fp32 — the buffer a row-sharded job all-reduces across ranks. */
int u2b_kmeans_accumulate(const void* x16, const int32_t* labels, int64_t N, int64_t D, int64_t K,
                          float* sums, void* workspace, size_t workspace_bytes,
                          u2b_stream_t stream);

/* M-step part 2, nn_utils.py:364: c32[k] = sums[k, 0:D] / sums[k, D] (NaN when empty). */
int u2b_kmeans_finalize(const float* sums, int64_t K, int64_t D, float* c32, u2b_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Exact k-nearest neighbours (squared L2) - u2seg/Instance_Clustering/shared/utils/nn_utils.py:203-224 kNN()
 * (KeOps Kmin_argKmin over ((x_i - y_j)^2).sum(-1)) and :227-299 partitioned_kNN(). csrc/knn.cu.
 * Candidate pass on tcgen05 (fp16 operands, fp32 accumulate) keeping 48 candidates per query, then an exact fp32 pass
 * that returns the K smallest distances (ascending; ties: lower index) and certifies them against the rounding bound
 * `eps` of the candidate pass; rows that cannot be certified are listed in `flagged` for exhaustive recomputation.
 * x16 / y16: fp16 rows, ynorm = |y|^2 (+inf beyond N2): u2b_kmeans_prepare() produces them (kpad = u2b_knn_npad).
 * ------------------------------------------------------------------------------------------ */
int u2b_knn_candidates_per_row(void);
int64_t u2b_knn_npad(int64_t N2);
int u2b_knn_set_cluster(int cluster);
int u2b_knn_candidates(const void* x16, int64_t N1, const void* y16, const float* ynorm, int64_t N2, int64_t D,
                       int32_t* cand_idx, float* cand_val, float* cand_thr, u2b_stream_t stream);
int u2b_knn_refine(const float* x, const float* y, const int32_t* cand_idx, const float* cand_val, const float* cand_thr,
                   const float* xnorm, int64_t N1, int64_t D, int K, float eps, float* d_out, int64_t* i_out,
                   int32_t* flagged, int32_t* n_flagged, u2b_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Detector ops. Feature maps are NHWC ("channels_last"); dtype codes: 0 = fp32, 1 = fp16, 2 = bf16.
 * rois5: (K, 5) fp32 rows (batch_index, x0, y0, x1, y1) — detectron2/modeling/poolers.py:72-98.
 * ------------------------------------------------------------------------------------------ */

/* poolers.py:23-59 assign_boxes_to_levels: levels[i] = clamp(floor(canonical_level +
 * log2(sqrt(area)/canonical_size + 1e-8)), min_level, max_level) - min_level. */
int u2b_assign_levels(const float* rois5, int64_t K, int min_level, int max_level,
                      float canonical_size, int canonical_level, int32_t* levels,
                      u2b_stream_t stream);

/* poolers.py:206-263 ROIPooler.forward / layers/roi_align.py:49-65 ROIAlign.forward (ROIAlignV2:
 * aligned=True, sampling_ratio=0). feats/hs/ws/scales are HOST arrays of num_levels (1..4) entries;
 * feats[l] is a device pointer to (N, H_l, W_l, C). levels (device, K) may be NULL when num_levels==1.
 * out: (K, P, P, C) of the same dtype. */
int u2b_roi_align_fwd(int dtype, int num_levels, const void* const* feats, const int32_t* hs,
                      const int32_t* ws, const float* scales, int64_t C, const float* rois5,
                      const int32_t* levels, int64_t K, int P, void* out, u2b_stream_t stream);

/* backward of the above: accumulates grad_scale * d(out)/d(feats) into grad_feats[l] (N, H_l, W_l, C) fp32, zeroed
 * by the caller. grad_scale folds cascade_rcnn.py:20-28 _ScaleGradient (1/num_stages) into the kernel; 1 otherwise. */
int u2b_roi_align_bwd(int dtype, int num_levels, float* const* grad_feats, const int32_t* hs,
                      const int32_t* ws, const float* scales, int64_t C, const float* rois5,
                      const int32_t* levels, int64_t K, int P, const void* grad_out, float grad_scale,
                      u2b_stream_t stream);
/* Implementation switch for u2b_roi_align_fwd (bit 0) / _bwd (bit 1): bit set = one CTA per ROI, separable weight tables in
 * shared memory (forward gathers (g+1)^2 pixels per bin, backward issues one vector atomic per footprint pixel); clear = one
 * warp per output bin (torchvision's per-sample order). Default 2 (backward only). Same results up to fp32 summation order. */
int u2b_roi_align_set_impl(int impl);

/* Channel-major variants (round-2 draft): out / grad_out are (K, C, P, P) contiguous, the layout torch.flatten(x, 1)
 * of the box head reads (box_head.py:99-106), so no transposing copy is needed on either side of the FC layers. */
int u2b_roi_align_chw_supported(int64_t C, int P);
int u2b_roi_align_fwd_chw(int dtype, int num_levels, const void* const* feats, const int32_t* hs, const int32_t* ws,
                          const float* scales, int64_t C, const float* rois5, const int32_t* levels, int64_t K, int P,
                          void* out, u2b_stream_t stream);
int u2b_roi_align_bwd_chw(int dtype, int num_levels, float* const* grad_feats, const int32_t* hs, const int32_t* ws,
                          const float* scales, int64_t C, const float* rois5, const int32_t* levels, int64_t K, int P,
                          const void* grad_out, float grad_scale, u2b_stream_t stream);

/* layers/mask_ops.py:74-147 paste_masks_in_image: masks (N, M, M) fp32 probabilities, boxes (N, 4)
 * fp32 -> out (N, H, W) bytes in {0,1} (= `img >= threshold`). */
int u2b_paste_masks(const float* masks, const float* boxes, int64_t N, int M, int H, int W,
                    float threshold, uint8_t* out, u2b_stream_t stream);

/* structures/masks.py:191-222 BitMasks.crop_and_resize fused with the gt_masks[matched_idxs] gather
 * of roi_heads.py:286-288. masks (G, H, W) bool bytes; gt_index (M) int64 or NULL (identity);
 * boxes (M, 4). out_bool (M, P, P) bytes (value >= 0.5) and/or out_val (M, P, P) fp32; either may be NULL. */
int u2b_crop_resize_masks(const uint8_t* masks, const int64_t* gt_index, const float* boxes,
                          int64_t M, int H, int W, int P, uint8_t* out_bool, float* out_val,
                          u2b_stream_t stream);

/* modeling/backbone/resnet.py:338-362 BasicStem.conv1: 7x7 stride-2 pad-3 convolution, 3 -> 64 channels, bf16 NHWC,
 * no bias (a norm layer follows). x (N, H, W, 3); w (64, 7, 7, 3); y (N, OH, OW, 64), OH = (H - 1) / 2 + 1.
 * Weight gradient (the image needs none): partials (u2b_stem_conv_wgrad_num_partials(N,H,W), 64, 160) fp32 receives
 * one partial per persistent CTA; dW[co][r][s][ci] = sum over partials of [co][(r*7+s)*3+ci] (columns >= 147 are
 * padding). */
int u2b_stem_conv_supported(int Cin, int Cout, int R, int S, int stride, int pad);
int u2b_stem_conv_fwd(const void* x, int64_t N, int H, int W, const void* w, void* y, u2b_stream_t stream);
int u2b_stem_conv_wgrad_num_partials(int64_t N, int H, int W);
int u2b_stem_conv_wgrad(const void* x, const void* dy, int64_t N, int H, int W, float* partials,
                        u2b_stream_t stream);

/* modeling/meta_arch/semantic_seg.py:255-267 SemSegFPNHead.losses: bilinear upsampling by `scale`
 * (F.interpolate, align_corners=False) of the stride-`scale` logits fused with F.cross_entropy(ignore_index) and
 * with the backward of both. logits (N, h, w, C) NHWC of dtype 0 = fp32 / 1 = fp16 / 2 = bf16; targets
 * (N, h*scale, w*scale) int64. partials (u2b_upsample_ce_num_partials(N, H, W) x 2 fp32) receives per-tile
 * [sum of pixel losses, number of non-ignored pixels]: loss = sum(partials[:,0]) / sum(partials[:,1]).
 * grad_logits (N, h, w, C) fp32, ZEROED by the caller, nullable: receives d(sum of pixel losses)/d(logits); the
 * gradient of the mean loss is grad_logits * grad_out / count. */
int64_t u2b_upsample_ce_num_partials(int64_t N, int H, int W);
int u2b_upsample_ce_supported(int C, int scale);
int u2b_upsample_ce(int dtype, const void* logits, const int64_t* targets, int64_t N, int h, int w, int C,
                    int scale, int64_t ignore_index, float* grad_logits, float* partials, u2b_stream_t stream);

/* structures/boxes.py:336-358 pairwise_iou + modeling/matcher.py:62-127 Matcher.__call__, fused:
 * gt (G, 4), pred (A, 4) fp32 -> matches int64 (A) (first maximum), matched_vals fp32 (A),
 * out_labels int8 (A). thresholds: nthr device floats [-inf, t.., +inf]; labels: nthr-1 device ints.
 * allow_low_quality needs gt_max_scratch (G uint32). G must be in 1..1024. gt_valid (G bytes, nullable) marks the
 * live rows of a fixed-capacity GT buffer; padded rows are never matched. */
int u2b_iou_match(const float* gt, int64_t G, const uint8_t* gt_valid, const float* pred, int64_t A,
                  const float* thresholds, const int32_t* labels, int nthr, int allow_low_quality,
                  int64_t* matches, float* matched_vals, int8_t* out_labels, uint32_t* gt_max_scratch,
                  u2b_stream_t stream);

/* layers/nms.py:9-21 batched_nms (torchvision class-by-class semantics; IoU > threshold suppresses).
 * order = indices of the boxes sorted by score, descending, stable. keep (n) int64 receives the kept
 * original indices in score order, *num_keep (device) their number; the scan stops after max_keep kept boxes
 * (max_keep < 0: keep all). valid (n bytes in the ORIGINAL box order, nullable): boxes with valid[i] == 0 are
 * neither kept nor suppress anything (fixed-capacity buffers). No host synchronisation. */
size_t u2b_nms_workspace_bytes(int64_t n);
/* developer instrumentation of the scan kernel (tools/nms_profile.py); not part of the drop-in surface */
int u2b_debug_nms_profile(int enable, uint64_t* out6);
int u2b_batched_nms(const float* boxes, const int64_t* cats, const int64_t* order, const uint8_t* valid, int64_t n,
                    float iou_threshold, int64_t max_keep, int64_t* keep, int32_t* num_keep, void* workspace,
                    size_t workspace_bytes, u2b_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Convolution / Linear on tcgen05 tensor cores (implicit GEMM, TMA-fed, fp32 accumulation in TMEM).
 * Replaces the F.conv2d call of detectron2/layers/wrappers.py:127 (and, as a 1x1 conv over a (1,1,M,K)
 * image, nn.Linear of roi_heads/box_head.py:70 and fast_rcnn.py:236-239) for shapes with Cin % 64 == 0,
 * Cout % 64 == 0, kernel 1x1 (pad 0) or 3x3 (pad 1), stride 1 or 2. dtype: 1 = fp16, 2 = bf16.
 * ------------------------------------------------------------------------------------------ */
int u2b_conv2d_supported(int Cin, int Cout, int R, int S, int stride, int pad);
/* thread-block cluster size of the kernel (1, 2 or 4): with >1 the filter tile is TMA-multicast across the cluster */
int u2b_conv2d_set_cluster(int cluster);

/* x (N,H,W,Cin) NHWC; w (Cout,R,S,Cin); out (N,OH,OW,Cout) NHWC, OH = (H + 2*pad - R)/stride + 1.
 * Epilogue: out = [relu]( acc + bias[c] + residual ), bias (Cout fp32) / residual (as out) optional. */
int u2b_conv2d_nhwc_fwd(int dtype, const void* x, int N, int H, int W, int Cin, const void* w, int Cout,
                        int R, int S, int stride, int pad, const float* bias, const void* residual, int relu,
                        void* out, u2b_stream_t stream);

/* Second-generation kernel (csrc/conv2.cu): a cluster of TWO CTAs (one SM pair) computes a 256 x {64,128,256} tile with
 * tcgen05.mma.cta_group::2 (each SM stages half of the filter tile), batched TMEM reads, TMA-store epilogue. Same
 * shapes and layouts as u2b_conv2d_nhwc_fwd; no residual operand. `stats` (optional) receives, per 128-pixel output
 * tile, the per-channel sum and sum of squares of the ROUNDED outputs: (u2b_conv2_stats_rows(...), 2*Cout) fp32, the
 * partial-row layout u2b_bn_finalize consumes (S = rows) - the statistics pass of the SyncBN that follows the conv
 * (layers/batch_norm.py:187 via layers/wrappers.py:127-134) without re-reading the activation. */
int u2b_conv2_supported(int Cin, int Cout, int R, int S, int stride, int pad);
int64_t u2b_conv2_stats_rows(int N, int H, int W, int R, int S, int stride, int pad);
/* Programmatic dependent launch between consecutive libu2b200 kernels of a stream (batch-norm and convolution kernels):
 * 1 (default) lets a kernel become resident and run its set-up while its predecessor drains; 0 = plain stream order. */
int u2b_set_pdl(int on);
/* SMs the persistent tcgen05 kernels (conv2, conv_wgrad2) size their grids for; 0 = all. Data-parallel training sets it
 * below the SM count during the backward pass so that NCCL's resident all-reduce CTAs do not force a second wave. */
int u2b_set_sm_budget(int sms);
/* 0 = choose the tile width per problem (default); 64 / 128 / 256 force it (benchmarking) */
int u2b_conv2_set_tile_n(int bn);
/* 1: short-K layers (<= 12 k-blocks of 64) run a shorter operand ring and 2-3 epilogue staging chunks per half (several TMA
 * stores in flight); 0 (default: measured no faster): full ring, one chunk. Developer switch for A/B timing. */
int u2b_conv2_set_staging(int on);
int u2b_conv2_nhwc_fwd(int dtype, const void* x, int N, int H, int W, int Cin, const void* w, int Cout, int R, int S,
                       int stride, int pad, const float* bias, int relu, void* out, float* stats,
                       u2b_stream_t stream);
/* Input gradient of a stride-1 'same' convolution, dX = conv(dY, rot180(W)^T), on the same kernel: the FORWARD filter
 * (Cout,R,S,Cin) is read in place as an MN-major UMMA operand (no transposed copy), taps flipped by index arithmetic.
 * dy (N,H,W,Cout) -> dx (N,H,W,Cin), NHWC; Cin % 128 == 0, Cout % 64 == 0. */
int u2b_conv2_dgrad_supported(int Cin, int Cout, int R, int S, int stride, int pad);
int u2b_conv2_nhwc_dgrad(int dtype, const void* dy, int N, int H, int W, int Cout, const void* w, int Cin, int R, int S,
                         int pad, void* dx, u2b_stream_t stream);

/* ConvTranspose2d(kernel 2, stride 2) forward, NHWC, on the 2-CTA kernel (roi_heads/mask_head.py:256 `deconv` + ReLU):
 * y[n,2h+i,2w+j,co] = [relu](bias[co] + sum_ci x[n,h,w,ci] * w[ci,co,i,j]); w is the channels_last weight, physical
 * (Cin,2,2,Cout), read in place. Its input gradient is u2b_conv2_nhwc_fwd(dy, w as a (Cin,2,2,Cout) OHWI filter, 2x2,
 * stride 2, pad 0); its weight gradient u2b_conv_wgrad2 of that convolution. Cin % 64 == 0, Cout % 128 == 0. */
int u2b_deconv2x2_supported(int Cin, int Cout);
int u2b_deconv2x2_nhwc_fwd(int dtype, const void* x, int N, int H, int W, int Cin, const void* w, int Cout,
                           const float* bias, int relu, void* y, u2b_stream_t stream);

/* Weight gradient on the 2-CTA tcgen05 kernel (csrc/conv_wgrad2.cu): dW[co,r,s,ci] = sum over output pixels of
 * dY[n,oh,ow,co] * X[n,oh*stride+r-pad,ow*stride+s-pad,ci] - the backward of the F.conv2d at layers/wrappers.py:127 and
 * (1x1 over a (1,1,M,K) image) of nn.Linear at roi_heads/box_head.py:70. Shapes: (Cout % 256 == 0 and Cin % 128 == 0)
 * or (Cin % 256 == 0 and Cout % 128 == 0); kernel 1x1 (pad 0) or 3x3 (pad 1); stride 1 or 2. x, dy fp16 (1) / bf16 (2)
 * NHWC; dw (Cout,R,S,Cin) in out_dtype (0 fp32, 1 fp16, 2 bf16); workspace holds the split-K partials, summed in a fixed
 * order (deterministic). */
int u2b_conv_wgrad2_supported(int Cin, int Cout, int R, int S, int stride, int pad);
int64_t u2b_conv_wgrad2_workspace_floats(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad);
int u2b_conv_wgrad2(int dtype, const void* x, const void* dy, int N, int H, int W, int Cin, int Cout, int R, int S,
                    int stride, int pad, float* workspace, int out_dtype, void* dw, u2b_stream_t stream);

/* Mask-head predictor + loss restricted to the class the loss reads (csrc/mask_loss.cu): replaces the 1x1 predictor conv,
 * the class gather and binary_cross_entropy_with_logits of detectron2/modeling/roi_heads/mask_head.py:33-112 (and their
 * backward) for fixed-capacity ROI slots. x (R, P, C): ROI features after the deconv+ReLU, NHWC rows, fp16 (1) / bf16 (2);
 * w (K, C) in x's dtype, bias (K) fp32; classes (R) int64; target (R, P) bool; ok (R) bool (dead slots contribute 0).
 * fwd: sum_j loss_per_roi[r][j] = sum_p bce(z[r,p], t[r,p]) * ok[r], g[r,p] = (sigmoid(z) - t) * ok[r].
 * bwd: upstream = device scalar d/d(loss sum); dx (R, P, C); dw (K, C), db (K) fp32 zero-filled by the caller;
 *      workspace R * (C + 1) floats. Deterministic (no atomics). C % 256 == 0, C <= 1024. */
int u2b_mask_loss_supported(int C);
int u2b_mask_loss_num_partials(void);   /* loss_per_roi is (R, this many) partial sums */
int u2b_mask_loss_fwd(int dtype, const void* x, const void* w, const float* bias, const int64_t* classes,
                      const uint8_t* target, const uint8_t* ok, int64_t R, int P, int C, float* g, float* loss_per_roi,
                      u2b_stream_t stream);
int u2b_mask_loss_bwd(int dtype, const void* x, const void* w, const int64_t* classes, const float* g,
                      const float* upstream, int64_t R, int P, int C, void* dx, float* dw, float* db, float* workspace,
                      u2b_stream_t stream);

/* GeneralizedRCNN.preprocess_image + ImageList.from_tensors (meta_arch/rcnn.py:223-234, structures/image_list.py:59-129)
 * for a batch of same-size uint8 NHWC images: (x - mean[c]) / std[c] in fp32 (subtract, IEEE divide), zero padding up to
 * (Hp, Wp), one rounding to out_dtype (0 fp32, 1 fp16, 2 bf16). mean3 / std3: host arrays of 3 floats. */
int u2b_preprocess_u8_nhwc(const uint8_t* img, int N, int H, int W, int Hp, int Wp, const float* mean3, const float* std3,
                           int out_dtype, void* out, u2b_stream_t stream);

/* NHWC pooling (csrc/pool.cu), fp16 (1) / bf16 (2), C % 8 == 0.
 * max pool 3x3 / stride 2 / pad 1 of the ResNet stem (backbone/resnet.py:358): forward also records, per pooled element,
 * the window position (kh*3+kw, 1 byte) of the first maximum in scan order (ATen's tie rule); backward gathers with it.
 * sum2x2: gradient of a nearest x2 upsampling (backbone/fpn.py:153): y[n,h,w] = sum of x over the 2x2 block. */
int u2b_maxpool3x3s2_fwd(int dtype, const void* x, int N, int H, int W, int C, void* y, uint8_t* idx, u2b_stream_t stream);
int u2b_maxpool3x3s2_bwd(int dtype, const void* dy, const uint8_t* idx, int N, int H, int W, int C, void* dx,
                         u2b_stream_t stream);
int u2b_sum2x2_nhwc(int dtype, const void* x, int N, int H, int W, int C, void* y, u2b_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Training-mode (Sync)BatchNorm on NHWC activations (P = N*H*W pixels, C % 8 == 0 channels), fused with the
 * residual add and ReLU that follow it. Replaces nn.SyncBatchNorm (detectron2/layers/batch_norm.py:187) inside
 * Conv2d.forward (layers/wrappers.py:87-134) and the `out += shortcut; relu` of backbone/resnet.py:194-210.
 * dtype: 0 fp32, 1 fp16, 2 bf16. Reductions write per-strip partial rows (no atomics, deterministic); a
 * data-parallel job sums them (u2b_bn_sum_partials), all-reduces the (2C) sums and passes them on with S = 1.
 * ------------------------------------------------------------------------------------------ */
int u2b_bn_supported(int C);
/* rows of the partial-sum buffer the reduce kernels write: partials is (num_strips, 2C) fp32 */
int u2b_bn_num_strips(int64_t P, int C);
/* partials[s] = (sum x | sum x^2) over pixel strip s */
int u2b_bn_stats(int dtype, const void* x, int64_t P, int C, float* partials, u2b_stream_t stream);
/* sums[0:C2] = sum of the S partial rows (for the all-reduce of a data-parallel job) */
int u2b_bn_sum_partials(const float* partials, int S, int C2, float* sums, u2b_stream_t stream);
/* stats (4C): mean | invstd | scale = w*invstd | shift = b - mean*scale, from S partial rows (S = 1: summed /
 * all-reduced) over n_total pixels; running statistics updated with momentum (unbiased variance) */
int u2b_bn_finalize(const float* partials, int S, double n_total, const float* w, const float* b, float eps,
                    float momentum, float* running_mean, float* running_var, float* stats, int C,
                    u2b_stream_t stream);
/* y = [relu](x * scale[c] + shift[c] [+ residual]) */
int u2b_bn_apply(int dtype, const void* x, const float* stats, const void* residual, int relu, void* y, int64_t P,
                 int C, u2b_stream_t stream);
/* y = [relu](x * scale[c] + shift[c] + nearest_upsample_x2(residual)): the FPN top-down sum (backbone/fpn.py:153-156)
 * folded into the lateral conv's SyncBN pass. x, y (N,H,W,C); residual (N,H/2,W/2,C); H, W even. */
int u2b_bn_apply_resup(int dtype, const void* x, const float* stats, const void* residual, int relu, void* y, int N, int H,
                       int W, int C, u2b_stream_t stream);

/* Backward of y = relu(bn(x)) WITHOUT residual, ReLU mask recomputed from x (y is not read: one tensor read less in each
 * of the two passes): dz = dy * (stored(x * scale + shift) > 0) with scale / shift = stats[2C:4C] and `stored` = rounding
 * to the activation dtype, i.e. exactly the mask y > 0. Otherwise as u2b_bn_bwd_reduce / u2b_bn_bwd_apply. */
int u2b_bn_bwd_reduce_relu_x(int dtype, const void* dy, const void* x, const float* stats, int64_t P, int C,
                             float* partials, u2b_stream_t stream);
int u2b_bn_bwd_apply_relu_x(int dtype, const void* dy, const void* x, const float* stats, const float* coeff, void* dx,
                            int64_t P, int C, u2b_stream_t stream);

/* partials[s] = (sum dz | sum dz*xhat), dz = dy * (y > 0) when y != NULL (fused ReLU backward) */
int u2b_bn_bwd_reduce(int dtype, const void* dy, const void* x, const void* y, const float* stats, int64_t P, int C,
                      float* partials, u2b_stream_t stream);
/* coeff (3C): dx = A*dz + B*x + K; gw_gb (2C, nullable) = dgamma | dbeta (the local sums) */
int u2b_bn_bwd_coeff(const float* partials, int S, double n_total, const float* stats, const float* w, float* coeff,
                     float* gw_gb, int C, u2b_stream_t stream);
/* nn.GroupNorm(G, C) (layers/batch_norm.py get_norm "GN"; semantic_seg.py:180-200 head convs) on the same kernels:
 * statistics are per image, so the caller runs u2b_bn_stats / u2b_bn_apply / u2b_bn_bwd_reduce / u2b_bn_bwd_apply on
 * ONE image (P = H*W pixels) and these two fold the per-channel partial sums across each group of C/G channels.
 * stats (4C) and coeff (3C) have the BN layout; gw_gb (2C) = dgamma | dbeta, summed over the batch when
 * accumulate != 0. */
int u2b_gn_supported(int C, int G);
int u2b_gn_finalize(const float* partials, int S, int64_t HW, int G, const float* w, const float* b, float eps,
                    float* stats, int C, u2b_stream_t stream);
int u2b_gn_bwd_coeff(const float* partials, int S, int64_t HW, int G, const float* stats, const float* w,
                     float* coeff, float* gw_gb, int accumulate, int C, u2b_stream_t stream);
/* Data-parallel (SyncBN) variants with the cross-GPU reduction fused into the kernel: `sums` (2C, this rank's summed
 * partials) is stored into every peer's symmetric buffer over NVLink, published with release/acquire flags, and
 * reduced locally - no NCCL call. peers: device array of `world` device pointers (every rank's buffer of
 * u2b_bn_xchg_buffer_bytes, zero-initialised once); epoch_ctr: device uint32 (start 0) advanced inside the kernel,
 * one per exchange, so every rank sees the same sequence and CUDA-graph replays stay valid. */
size_t u2b_bn_xchg_buffer_bytes(int world, int slot_floats);
int u2b_bn_xchg_finalize(const float* sums, const void* peers, int world, int rank, uint32_t* epoch_ctr, int slot_floats,
                         double n_total, const float* w, const float* b, float eps, float momentum,
                         float* running_mean, float* running_var, float* stats, int C, u2b_stream_t stream);
int u2b_bn_xchg_bwd_coeff(const float* sums, const void* peers, int world, int rank, uint32_t* epoch_ctr, int slot_floats,
                          double n_total, const float* stats, const float* w, float* coeff, float* gw_gb, int C,
                          u2b_stream_t stream);
/* Multi-CTA variants (one CTA per 32 channels) that take the (S, 2C) partial rows directly: partial-row summation, NVLink
 * exchange and the statistics / coefficients in ONE launch per BN direction. epoch_ctrs: device uint32[u2b_bn_xchg2_max_ctas()],
 * zero-initialised once. Same symmetric buffers (u2b_bn_xchg_buffer_bytes covers both flag areas). Reference:
 * detectron2/layers/batch_norm.py:187-229 (NaiveSyncBatchNorm: all-reduce of [mean, meansqr] / of the backward sums). */
int u2b_bn_xchg2_max_ctas(void);
int u2b_bn_xchg2_finalize(const float* partials, int S, const void* peers, int world, int rank, uint32_t* epoch_ctrs,
                          int slot_floats, double n_total, const float* w, const float* b, float eps, float momentum,
                          float* running_mean, float* running_var, float* stats, int C, cudaStream_t stream);
int u2b_bn_xchg2_bwd_coeff(const float* partials, int S, const void* peers, int world, int rank, uint32_t* epoch_ctrs,
                           int slot_floats, double n_total, const float* stats, const float* w, float* coeff, float* gw_gb,
                           int C, cudaStream_t stream);
/* dx = A*dz + B*x + K; dres = dz when dres != NULL */
int u2b_bn_bwd_apply(int dtype, const void* dy, const void* x, const void* y, const float* coeff, void* dx,
                     void* dres, int64_t P, int C, u2b_stream_t stream);

/* nn.Upsample(scale_factor=s, mode="bilinear", align_corners=False) on NHWC activations (semantic_seg.py:195-199),
 * round-2 draft. dir 0: out (N, h*s, w*s, C) = upsample(in (N, h, w, C)); dir 1: out (N, h, w, C) = gradient w.r.t. the
 * input given in = gradient of the output (N, h*s, w*s, C) (gather form: deterministic). Even scales, C % 8 == 0. */
int u2b_upsample_bilinear_supported(int C, int scale);
int u2b_upsample_bilinear(int dtype, int dir, const void* in, void* out, int64_t N, int h, int w, int C, int scale,
                          u2b_stream_t stream);

/* Weight gradient of the NHWC convolutions on tcgen05 (round-2 draft, csrc/conv_wgrad_tc.cu): x (N,H,W,Cin),
 * dy (N,OH,OW,Cout), fp16 (1) / bf16 (2). partials (u2b_conv2d_wgrad_ksplit(...), Cout, R, S, Cin) fp32 receives one
 * partial per K split; dW (Cout,R,S,Cin) = their sum. Cin % 64 == 0, Cout % 128 == 0, 1x1 or 3x3 pad 1, stride 1|2. */
int u2b_conv2d_wgrad_supported(int Cin, int Cout, int R, int S, int stride, int pad);
int u2b_conv2d_wgrad_ksplit(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad);
int u2b_conv2d_nhwc_wgrad(int dtype, const void* x, const void* dy, int N, int H, int W, int Cin, int Cout, int R, int S,
                          int stride, int pad, float* partials, u2b_stream_t stream);

/* Fused detection losses (value + closed-form gradient of the SUMMED loss in one pass; the caller applies the scalar
 * normaliser). dtype 0 = fp32 / 1 = fp16 / 2 = bf16 for the head outputs; gradients are always fp32.
 * u2b_rpn_losses - proposal_generator/rpn.py:365-429: logits (N, A), deltas (N, A, 4), anchors (A, 4) fp32, labels
 *   (N, A) int8 in {-1 ignore, 0, 1}, matched (N, A) int64 index into gt_boxes (N, G, 4) fp32, weights4 (host) =
 *   Box2BoxTransform weights. partials (u2b_rpn_losses_num_partials(N*A), 2) = [BCE sum, L1 sum] per CTA.
 *   grad_logits / grad_deltas nullable.
 * u2b_box_losses - roi_heads/fast_rcnn.py:307-352 (class-agnostic regression): scores (R, C), classes (R) int64
 *   (-100 = ignored slot, K = background), deltas (R, 4), proposals / gt_boxes (R, 4) fp32. partials
 *   (u2b_box_losses_num_partials(R), 2) = [CE sum, L1 sum over foreground rows]. refined (R, 4) fp32 (nullable) =
 *   apply_deltas(deltas, proposals), the next cascade stage's boxes (cascade_rcnn.py:271-299). */
int64_t u2b_rpn_losses_num_partials(int64_t total);
int64_t u2b_box_losses_num_partials(int64_t R);
int u2b_rpn_losses(int dtype, const void* logits, const void* deltas, const float* anchors, const int8_t* labels,
                   const int64_t* matched, const float* gt_boxes, int64_t N, int64_t A, int G, const float* weights4,
                   float* grad_logits, float* grad_deltas, float* partials, u2b_stream_t stream);
int u2b_box_losses(int dtype, const void* scores, const int64_t* classes, const void* deltas, const float* proposals,
                   const float* gt_boxes, int64_t R, int C, int K, const float* weights4, float scale_clamp,
                   float* grad_scores, float* grad_deltas, float* refined, float* partials, u2b_stream_t stream);

/* rpn.py:497-533 + proposal_utils.py:85-121 for the anchors kept by the per-level top-k (round-2 draft): decode, clip to
 * the image, validity (finite, both sides > min_size). sel (N, Ksel) int64 anchor indices into deltas (N, A, 4) /
 * anchors (A, 4); scores (N, Ksel) fp32. boxes (N, Ksel, 4) fp32, valid (N, Ksel) bytes, *nonfinite (device int, caller
 * zeroes it) set when a selected box / score is not finite. */
int u2b_rpn_decode_selected(int dtype, const void* deltas, const float* anchors, const int64_t* sel, const float* scores,
                            int64_t N, int64_t A, int Ksel, const float* weights4, float scale_clamp, float img_h,
                            float img_w, float min_size, float* boxes, uint8_t* valid, int* nonfinite,
                            u2b_stream_t stream);

/* cascade_rcnn.py:193-236,271-299 relabelling of the refined boxes for cascade stage k > 0 on fixed-capacity slots
 * (round-2 draft): clip, dead-slot handling, IoU matching against the image's valid GT boxes with threshold iou_thr
 * (first maximum; foreground iff IoU >= thr), class K = background, -100 = dead slot, matched GT box. */
int u2b_cascade_relabel(const float* refined, const uint8_t* ok_prev, const float* gt_boxes, const int64_t* gt_classes,
                        const uint8_t* gt_valid, int64_t N, int R, int G, float img_h, float img_w, float iou_thr, int K,
                        float* boxes, int64_t* classes, uint8_t* ok, float* gtb, u2b_stream_t stream);

/* solver/build.py:63-73 (per-parameter gradient-norm clipping) + solver/build.py:119-139 (torch.optim.SGD: weight
 * decay, momentum, optional Nesterov) + the refresh of the bf16 compute weights, fused over flat buffers.
 * grad / master / mom: n fp32 each, same offsets, every parameter starting on a 64-element boundary.
 * seg_of_chunk (n/64 int32): parameter index owning each 64-element chunk, -1 for padding. seg_wd / seg_coef: per
 * parameter weight decay and clip coefficient min(1, max_norm/(norm+1e-6)) (seg_coef nullable: no clipping).
 * lr: device scalar. w16 (nullable): bf16 copy of master[w16_begin : n], refreshed in the same pass. */
int u2b_sgd_step_segments(const float* grad, float* master, float* mom, void* w16, int64_t w16_begin,
                          const int32_t* seg_of_chunk, const float* seg_wd, const float* seg_coef, const float* lr,
                          float momentum, int nesterov, int64_t n, u2b_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* U2B200_H_ */
