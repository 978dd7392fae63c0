// Fused detection losses of the Panoptic-FPN step (forward value AND the closed-form gradient in one pass, like
// semseg_loss.cu): each kernel reads the head outputs once, writes per-CTA partial sums of the loss and the gradient of
// the SUMMED loss with respect to the head outputs; the caller applies the scalar normaliser (1/count, loss weight,
// upstream gradient) with one tiny multiply. Gradients are written in fp32 whatever the head dtype: the reference's
// loss arithmetic is fp32 (autocast promotes the losses), the cast to bf16 happens after the scalar multiply. They replace ~80 (RPN) and ~60 per cascade stage (box head) library
// launches of 2-4 us each in forward plus as many in backward.
//
//   u2b_rpn_losses: proposal_generator/rpn.py:365-429 RPN.losses
//       loss_rpn_cls = BCE-with-logits(objectness[valid], labels[valid]) summed   (valid: label >= 0)
//       loss_rpn_loc = L1(pred_deltas[pos], get_deltas(anchors, matched_gt)[pos]) summed   (SMOOTH_L1_BETA 0 -> L1)
//   u2b_box_losses: roi_heads/fast_rcnn.py:307-352 FastRCNNOutputLayers.losses (class-agnostic regression)
//       loss_cls     = cross_entropy(scores, classes, ignore_index=-100) summed
//       loss_box_reg = L1(deltas[fg], get_deltas(proposals, gt)[fg]) summed        (fg: 0 <= class < K)
//       plus the refined boxes apply_deltas(deltas, proposals) the next cascade stage starts from
//       (cascade_rcnn.py:271-299), so that no second pass over the head outputs is needed.
// get_deltas / apply_deltas: modeling/box_regression.py:43-116, same operation order in fp32.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

// box_regression.py:43-75 get_deltas(src, target) with weights w
__device__ __forceinline__ void get_deltas(const float4 s, const float4 t, const float4 w, float (&d)[4]) {
  const float sw = s.z - s.x, sh = s.w - s.y;
  const float sx = s.x + 0.5f * sw, sy = s.y + 0.5f * sh;
  const float tw = t.z - t.x, th = t.w - t.y;
  const float tx = t.x + 0.5f * tw, ty = t.y + 0.5f * th;
  d[0] = w.x * (tx - sx) / sw;
  d[1] = w.y * (ty - sy) / sh;
  d[2] = w.z * logf(tw / sw);
  d[3] = w.w * logf(th / sh);
}

// CTA sum of two floats -> partials[2*blockIdx.x + {0,1}] (fixed order: deterministic)
__device__ __forceinline__ void block_sum2(float a, float b, float* __restrict__ partials) {
  __shared__ float red[2][32];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    red[0][warp] = a;
    red[1][warp] = b;
  }
  __syncthreads();
  if (warp == 0) {
    a = lane < nw ? red[0][lane] : 0.f;
    b = lane < nw ? red[1][lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      partials[2 * blockIdx.x] = a;
      partials[2 * blockIdx.x + 1] = b;
    }
  }
}

// one thread per (image, anchor)
template <typename T>
__global__ void __launch_bounds__(256)
rpn_losses_kernel(const T* __restrict__ logits, const T* __restrict__ deltas, const float4* __restrict__ anchors,
                  const int8_t* __restrict__ labels, const int64_t* __restrict__ matched, const float4* __restrict__ gt,
                  int G, long long A, long long total, float4 w, float* __restrict__ g_logits, float* __restrict__ g_deltas,
                  float* __restrict__ partials) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float l_cls = 0.f, l_loc = 0.f;
  if (i < total) {
    const int lab = labels[i];
    float gx = 0.f, gd[4] = {0.f, 0.f, 0.f, 0.f};
    if (lab >= 0) {
      const float x = ldf<T>(logits + i), y = static_cast<float>(lab);
      // F.binary_cross_entropy_with_logits: (1 - y) * x + max(-x, 0) + log1p(exp(-|x|))
      l_cls = (1.f - y) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
      gx = 1.f / (1.f + expf(-x)) - y;
    }
    if (lab == 1) {
      const long long n = i / A, a = i - n * A;
      float t[4];
      get_deltas(anchors[a], gt[n * G + matched[i]], w, t);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float d = ldf<T>(deltas + i * 4 + k) - t[k];
        l_loc += fabsf(d);
        gd[k] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      }
    }
    if (g_logits) g_logits[i] = gx;
    if (g_deltas) {
#pragma unroll
      for (int k = 0; k < 4; ++k) g_deltas[i * 4 + k] = gd[k];
    }
  }
  block_sum2(l_cls, l_loc, partials);
}

// one warp per sampled ROI row: softmax cross-entropy over C = K+1 scores, L1 on the 4 class-agnostic deltas,
// refined box for the next cascade stage.
template <typename T>
__global__ void __launch_bounds__(256)
box_losses_kernel(const T* __restrict__ scores, const int64_t* __restrict__ classes, const T* __restrict__ deltas,
                  const float4* __restrict__ props, const float4* __restrict__ gtb, int R, int C, int K, float4 w,
                  float scale_clamp, float* __restrict__ g_scores, float* __restrict__ g_deltas, float4* __restrict__ refined,
                  float* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float l_ce = 0.f, l_l1 = 0.f;
  if (row < R) {
    const long long cls = classes[row];
    const bool counted = cls >= 0 && cls < C;            // ignore_index (-100) rows contribute nothing
    const T* s = scores + static_cast<size_t>(row) * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, ldf<T>(s + c));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(ldf<T>(s + c) - m);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lse = m + logf(sum);
    if (counted && lane == 0) l_ce = lse - ldf<T>(s + cls);
    if (g_scores) {
      float* g = g_scores + static_cast<size_t>(row) * C;
      for (int c = lane; c < C; c += 32)
        g[c] = counted ? expf(ldf<T>(s + c) - lse) - (c == cls ? 1.f : 0.f) : 0.f;
    }
    if (lane == 0) {
      const float4 p = props[row];
      float dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) dv[k] = ldf<T>(deltas + static_cast<size_t>(row) * 4 + k);
      float gd[4] = {0.f, 0.f, 0.f, 0.f};
      if (cls >= 0 && cls < K) {                         // foreground: box regression target
        float t[4];
        get_deltas(p, gtb[row], w, t);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = dv[k] - t[k];
          l_l1 += fabsf(d);
          gd[k] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        }
      }
      if (g_deltas) {
#pragma unroll
        for (int k = 0; k < 4; ++k) g_deltas[static_cast<size_t>(row) * 4 + k] = gd[k];
      }
      if (refined) {                                     // box_regression.py:77-116 apply_deltas
        const float bw = p.z - p.x, bh = p.w - p.y;
        const float cx = p.x + 0.5f * bw, cy = p.y + 0.5f * bh;
        const float dx = dv[0] / w.x, dy = dv[1] / w.y;
        const float dw = fminf(dv[2] / w.z, scale_clamp), dh = fminf(dv[3] / w.w, scale_clamp);
        const float pcx = dx * bw + cx, pcy = dy * bh + cy;
        const float pw = expf(dw) * bw, ph = expf(dh) * bh;
        refined[row] = make_float4(pcx - 0.5f * pw, pcy - 0.5f * ph, pcx + 0.5f * pw, pcy + 0.5f * ph);
      }
    }
  }
  block_sum2(l_ce, l_l1, partials);
}

// proposal_generator/rpn.py:497-533 _decode_proposals + proposal_utils.py:85-121 (clip to the image, drop boxes
// with a side <= min_size or non-finite coordinates) for the anchors selected by the per-level top-k only:
// one thread per (image, selected anchor). boxes out (N, Ksel, 4) fp32 clipped, valid (N, Ksel) bytes; *nonfinite is
// set to 1 if any selected box or score is not finite (the reference raises FloatingPointError there).
template <typename T>
__global__ void __launch_bounds__(256)
rpn_decode_selected_kernel(const T* __restrict__ deltas, const float4* __restrict__ anchors,
                           const int64_t* __restrict__ sel, const float* __restrict__ scores, long long A, int Ksel,
                           long long total, float4 w, float scale_clamp, float img_h, float img_w, float min_size,
                           float4* __restrict__ boxes, uint8_t* __restrict__ valid, int* __restrict__ nonfinite) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / Ksel, a = sel[i];
  const float4 p = anchors[a];
  const T* d = deltas + (n * A + a) * 4;
  const float dv0 = ldf<T>(d), dv1 = ldf<T>(d + 1), dv2 = ldf<T>(d + 2), dv3 = ldf<T>(d + 3);
  const float bw = p.z - p.x, bh = p.w - p.y;
  const float cx = p.x + 0.5f * bw, cy = p.y + 0.5f * bh;
  const float dx = dv0 / w.x, dy = dv1 / w.y;
  const float dw = fminf(dv2 / w.z, scale_clamp), dh = fminf(dv3 / w.w, scale_clamp);
  const float pcx = dx * bw + cx, pcy = dy * bh + cy;
  const float pw = expf(dw) * bw, ph = expf(dh) * bh;
  float x0 = pcx - 0.5f * pw, y0 = pcy - 0.5f * ph, x1 = pcx + 0.5f * pw, y1 = pcy + 0.5f * ph;
  const bool fin = isfinite(x0) && isfinite(y0) && isfinite(x1) && isfinite(y1) && isfinite(scores[i]);
  if (!fin) *nonfinite = 1;
  x0 = fminf(fmaxf(x0, 0.f), img_w);      // Boxes.clip: x in [0, w], y in [0, h]
  x1 = fminf(fmaxf(x1, 0.f), img_w);
  y0 = fminf(fmaxf(y0, 0.f), img_h);
  y1 = fminf(fmaxf(y1, 0.f), img_h);
  boxes[i] = make_float4(x0, y0, x1, y1);
  valid[i] = fin && (x1 - x0) > min_size && (y1 - y0) > min_size;
}

// Cascade stage k > 0 relabelling (cascade_rcnn.py:271-299 _create_proposals_from_boxes + :193-236
// _match_and_label_boxes) on fixed-capacity slots, one thread per (image, slot): clip the refined box to the image,
// drop empty boxes (keep the slot, mark it dead), match against the image's <= G ground-truth boxes with the stage's IoU
// threshold (Matcher([thr], [0, 1], allow_low_quality_matches=False): first maximum, foreground iff IoU >= thr) and
// emit class (K = background, -100 = dead slot) and the matched GT box. structures/boxes.py:336-358 pairwise_iou.
__global__ void __launch_bounds__(256)
cascade_relabel_kernel(const float4* __restrict__ refined, const uint8_t* __restrict__ ok_prev,
                       const float4* __restrict__ gt_boxes, const int64_t* __restrict__ gt_classes,
                       const uint8_t* __restrict__ gt_valid, int R, int G, long long total, float img_h, float img_w,
                       float iou_thr, int K, float4* __restrict__ boxes, int64_t* __restrict__ classes,
                       uint8_t* __restrict__ ok_out, float4* __restrict__ gtb) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / R;
  float4 b = refined[i];
  b.x = fminf(fmaxf(b.x, 0.f), img_w);
  b.z = fminf(fmaxf(b.z, 0.f), img_w);
  b.y = fminf(fmaxf(b.y, 0.f), img_h);
  b.w = fminf(fmaxf(b.w, 0.f), img_h);
  const bool ok = ok_prev[i] != 0 && (b.z - b.x) > 0.f && (b.w - b.y) > 0.f;
  if (!ok) b = make_float4(0.f, 0.f, 1.f, 1.f);            // placeholder box of dead slots
  // same arithmetic as boxes.cu iou_match_kernel (areas rounded on their own: no fma contraction with the sum)
  const float area_b = __fmul_rn(b.z - b.x, b.w - b.y);
  float best = -1.f;
  int best_j = 0;
  bool any_gt = false;
  const float4* g = gt_boxes + n * G;
  for (int j = 0; j < G; ++j) {
    if (!gt_valid[n * G + j]) continue;
    any_gt = true;
    const float4 t = g[j];
    const float iou = ptx_free::iou_ref(t, __fmul_rn(t.z - t.x, t.w - t.y), b, area_b);
    if (iou > best) {          // first maximum
      best = iou;
      best_j = j;
    }
  }
  long long cls = K;
  if (any_gt && best >= iou_thr) cls = gt_classes[n * G + best_j];
  boxes[i] = b;
  classes[i] = ok ? cls : -100;
  ok_out[i] = ok;
  gtb[i] = g[best_j];
}

}  // namespace

extern "C" {

int64_t u2b_rpn_losses_num_partials(int64_t total) { return (total + 255) / 256; }
int64_t u2b_box_losses_num_partials(int64_t R) { return (R + 7) / 8; }

int u2b_rpn_losses(int dtype, const void* logits, const void* deltas, const float* anchors, const int8_t* labels,
                   const int64_t* matched, const float* gt_boxes, int64_t N, int64_t A, int G, const float* weights4,
                   float* grad_logits, float* grad_deltas, float* partials, cudaStream_t stream) {
  const long long total = N * A;
  if (total == 0) return 0;
  U2B_CHECK_ARG(logits && deltas && anchors && labels && matched && gt_boxes && weights4 && partials && G > 0,
                "rpn_losses: bad arguments");
  const float4 w = make_float4(weights4[0], weights4[1], weights4[2], weights4[3]);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
#define U2B_RPN(T)                                                                                                  \
  rpn_losses_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(logits), static_cast<const T*>(deltas),       \
                                                 reinterpret_cast<const float4*>(anchors), labels, matched,          \
                                                 reinterpret_cast<const float4*>(gt_boxes), G, A, total, w,          \
                                                 grad_logits, grad_deltas, partials)
  if (dtype == 0) U2B_RPN(float);
  else if (dtype == 1) U2B_RPN(__half);
  else if (dtype == 2) U2B_RPN(__nv_bfloat16);
  else {
    u2b_set_error("rpn_losses: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
#undef U2B_RPN
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_box_losses(int dtype, const void* scores, const int64_t* classes, const void* deltas, const float* proposals,
                   const float* gt_boxes, int64_t R, int C, int K, const float* weights4, float scale_clamp,
                   float* grad_scores, float* grad_deltas, float* refined, float* partials, cudaStream_t stream) {
  if (R == 0) return 0;
  U2B_CHECK_ARG(scores && classes && deltas && proposals && gt_boxes && weights4 && partials && C > 0 && K >= 0 && K <= C,
                "box_losses: bad arguments");
  const float4 w = make_float4(weights4[0], weights4[1], weights4[2], weights4[3]);
  const unsigned grid = static_cast<unsigned>((R + 7) / 8);
#define U2B_BOX(T)                                                                                                   \
  box_losses_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(scores), classes, static_cast<const T*>(deltas), \
                                                 reinterpret_cast<const float4*>(proposals),                          \
                                                 reinterpret_cast<const float4*>(gt_boxes), (int)R, C, K, w,          \
                                                 scale_clamp, grad_scores, grad_deltas,                                \
                                                 reinterpret_cast<float4*>(refined), partials)
  if (dtype == 0) U2B_BOX(float);
  else if (dtype == 1) U2B_BOX(__half);
  else if (dtype == 2) U2B_BOX(__nv_bfloat16);
  else {
    u2b_set_error("box_losses: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
#undef U2B_BOX
  U2B_LAUNCH_CHECK();
  return 0;
}


// sel (N, Ksel) int64: index of each selected anchor in [0, A); deltas (N, A, 4); anchors (A, 4); scores (N, Ksel) fp32.
int u2b_rpn_decode_selected(int dtype, const void* deltas, const float* anchors, const int64_t* sel, const float* scores,
                            int64_t N, int64_t A, int Ksel, const float* weights4, float scale_clamp, float img_h,
                            float img_w, float min_size, float* boxes, uint8_t* valid, int* nonfinite,
                            cudaStream_t stream) {
  const long long total = N * Ksel;
  if (total == 0) return 0;
  U2B_CHECK_ARG(deltas && anchors && sel && scores && weights4 && boxes && valid && nonfinite, "rpn_decode_selected: null pointer");
  const float4 w = make_float4(weights4[0], weights4[1], weights4[2], weights4[3]);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
#define U2B_DEC(T)                                                                                                    \
  rpn_decode_selected_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(deltas),                                \
                                                          reinterpret_cast<const float4*>(anchors), sel, scores, A, Ksel, \
                                                          total, w, scale_clamp, img_h, img_w, min_size,                 \
                                                          reinterpret_cast<float4*>(boxes), valid, nonfinite)
  if (dtype == 0) U2B_DEC(float);
  else if (dtype == 1) U2B_DEC(__half);
  else if (dtype == 2) U2B_DEC(__nv_bfloat16);
  else {
    u2b_set_error("rpn_decode_selected: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
#undef U2B_DEC
  U2B_LAUNCH_CHECK();
  return 0;
}


// refined (N, R, 4) fp32, ok_prev (N, R) bytes, gt_boxes (N, G, 4), gt_classes (N, G) int64, gt_valid (N, G) bytes.
// Outputs: boxes (N, R, 4), classes (N, R) int64, ok (N, R) bytes, gtb (N, R, 4).
int u2b_cascade_relabel(const float* refined, const uint8_t* ok_prev, const float* gt_boxes, const int64_t* gt_classes,
                        const uint8_t* gt_valid, int64_t N, int R, int G, float img_h, float img_w, float iou_thr, int K,
                        float* boxes, int64_t* classes, uint8_t* ok, float* gtb, cudaStream_t stream) {
  const long long total = N * R;
  if (total == 0) return 0;
  U2B_CHECK_ARG(refined && ok_prev && gt_boxes && gt_classes && gt_valid && boxes && classes && ok && gtb && G > 0,
                "cascade_relabel: bad arguments");
  cascade_relabel_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(refined), ok_prev, reinterpret_cast<const float4*>(gt_boxes), gt_classes, gt_valid, R, G,
      total, img_h, img_w, iou_thr, K, reinterpret_cast<float4*>(boxes), classes, ok, reinterpret_cast<float4*>(gtb));
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
