// Mask kernels of the detector path (HBM-bound byte work).
//  * paste_masks_in_image — detectron2/layers/mask_ops.py:74-147 (+ _do_paste_mask :17-69):
//    the reference builds an (N,H,W,2) fp32 grid, grid_samples the MxM probabilities
//    (bilinear, align_corners=False, zero padding), thresholds and index_puts: ~2.9 GB of traffic
//    for a 0.107 GB result at N=100, 800x1333. Here: one pass, the mask staged in shared memory,
//    16 output pixels per thread written as one 16-byte store; algorithmic bytes = N*H*W written.
//  * crop_and_resize of bit masks — detectron2/structures/masks.py:191-222 (ROIAlign 28x28,
//    scale 1, sampling_ratio 0, aligned) fused with the `gt_masks[matched_idxs]` gather of
//    roi_heads.py:286-288: reads the bool masks through the indirection (no 537 MB/image
//    materialisation), lanes split the adaptive sample grid and reduce with warp shuffles.
#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

constexpr int PASTE_MAX_M = 56;  // mask side staged in smem (28 in the reference configs)

// One block = one instance x PASTE_ROWS output rows = one contiguous span of the output. ~98 % of an instance's image
// lies outside its box, where every bilinear corner is out of bounds and grid_sample yields exactly 0, so the block
// first FILLS its span with the constant (0 >= threshold) using 16-byte stores (memset speed), then - only if some of
// its rows reach into the mask's support - builds the x-dependent half of the bilinear sample once (source column and
// the two column weights per output column: identical for every row; structure of arrays, conflict-free 16-byte reads)
// and recomputes the pixels of the support rectangle, 4 consecutive pixels per thread from the tables, the mask in shared
// memory and the row's two y weights. Degenerate boxes (x1 <= x0 or y1 <= y0: NaN / inf coordinates) take the per-pixel
// path for the whole span, as the reference's formula would.
constexpr int PASTE_ROWS = 32;
constexpr int PASTE_MAX_W = 4096;

__global__ void __launch_bounds__(256)
paste_masks_kernel(const float* __restrict__ masks, const float* __restrict__ boxes, int N, int M,
                   int H, int W, int Wpad, float threshold, uint8_t* __restrict__ out) {
  __shared__ float sm[PASTE_MAX_M * PASTE_MAX_M];
  __shared__ int s_xlo, s_xhi;
  extern __shared__ __align__(16) uint8_t col_raw[];
  float* cwx0 = reinterpret_cast<float*>(col_raw);  // [Wpad]
  float* cwx1 = cwx0 + Wpad;                        // [Wpad]
  int* cxi0 = reinterpret_cast<int*>(cwx1 + Wpad);  // [Wpad]
  const int n = blockIdx.y;
  const int row0 = blockIdx.x * PASTE_ROWS;
  const int rows = min(PASTE_ROWS, H - row0);
  const float x0 = boxes[n * 4 + 0], y0 = boxes[n * 4 + 1], x1 = boxes[n * 4 + 2], y1 = boxes[n * 4 + 3];
  const float fM = static_cast<float>(M);
  const bool regular = (x1 > x0) && (y1 > y0);
  const uint8_t fill = (0.f >= threshold) ? 1 : 0;

  // rows of this block inside the mask's vertical support (same arithmetic as the sampling code below)
  bool row_in = false;
  if (threadIdx.x < rows) {
    const int y = row0 + threadIdx.x;
    const float gy = ((static_cast<float>(y) + 0.5f) - y0) / (y1 - y0) * 2.f - 1.f;
    const float iy = ((gy + 1.f) * fM - 1.f) / 2.f;
    const float iy_nw = floorf(iy);
    row_in = iy_nw >= -1.f && iy_nw <= fM - 1.f;   // yi0 or yi0 + 1 inside [0, M)
  }
  const bool any_row = __syncthreads_or(row_in || !regular);

  if (regular) {   // constant fill of the whole span
    uint8_t* p = out + (static_cast<size_t>(n) * H + row0) * W;
    const size_t len = static_cast<size_t>(rows) * W;
    size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > len) head = len;
    const size_t body = (len - head) / 16;
    const uint32_t w32 = fill * 0x01010101u;
    const uint4 v = make_uint4(w32, w32, w32, w32);
    for (size_t i = threadIdx.x; i < head; i += blockDim.x) p[i] = fill;
    uint4* pb = reinterpret_cast<uint4*>(p + head);
    for (size_t i = threadIdx.x; i < body; i += blockDim.x) pb[i] = v;
    const size_t tail0 = head + body * 16;
    for (size_t i = tail0 + threadIdx.x; i < len; i += blockDim.x) p[i] = fill;
    if (!any_row) return;
  }

  const float* mk = masks + static_cast<size_t>(n) * M * M;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) sm[i] = mk[i];
  if (threadIdx.x == 0) {
    s_xlo = Wpad;
    s_xhi = -1;
  }
  __syncthreads();   // also orders the fill before the recomputation below (same block, same addresses)
  int my_lo = Wpad, my_hi = -1;
  for (int x = threadIdx.x; x < Wpad; x += blockDim.x) {
    // mask_ops.py:51-54: img_x = (arange + 0.5 - x0) / (x1 - x0) * 2 - 1; grid_sample unnormalize
    // (align_corners=False): ((g + 1) * size - 1) / 2
    const float gx = ((static_cast<float>(x) + 0.5f) - x0) / (x1 - x0) * 2.f - 1.f;
    const float ix = ((gx + 1.f) * fM - 1.f) / 2.f;
    const float ix_nw = floorf(ix);
    // out-of-range coordinates are clamped to sentinels whose two corners are both invalid:
    // left of the mask -> -2, right of it -> M + 1, NaN (degenerate box) -> -3
    const int xi = (ix_nw != ix_nw) ? -3 : (ix_nw < -1.f ? -2 : (ix_nw > fM ? M + 1 : static_cast<int>(ix_nw)));
    cxi0[x] = xi;
    cwx1[x] = ix - ix_nw;
    cwx0[x] = (ix_nw + 1.f) - ix;
    if (xi >= -1 && xi < M) {   // a valid corner exists
      my_lo = min(my_lo, x);
      my_hi = max(my_hi, x);
    }
  }
  if (my_hi >= 0) {
    atomicMin(&s_xlo, my_lo);
    atomicMax(&s_xhi, my_hi);
  }
  __syncthreads();
  int g_lo = 0, g_hi = Wpad / 4;                     // 4-pixel groups to recompute: [g_lo, g_hi)
  if (regular) {
    if (s_xhi < 0) return;                           // no column touches the mask
    g_lo = s_xlo / 4;
    g_hi = s_xhi / 4 + 1;
  }
  const int ng = g_hi - g_lo;
  for (int t = threadIdx.x; t < rows * ng; t += blockDim.x) {
    const int y = row0 + t / ng;
    const int xg = (g_lo + t % ng) * 4;
    const float gy = ((static_cast<float>(y) + 0.5f) - y0) / (y1 - y0) * 2.f - 1.f;
    const float iy = ((gy + 1.f) * fM - 1.f) / 2.f;
    const float iy_nw = floorf(iy);
    const int yi0 = (iy_nw != iy_nw) ? -3 : (iy_nw < -1.f ? -2 : (iy_nw > fM ? M + 1 : static_cast<int>(iy_nw)));
    const int yi1 = yi0 + 1;
    const float wy1 = iy - iy_nw, wy0 = (iy_nw + 1.f) - iy;
    const bool y0ok = yi0 >= 0 && yi0 < M, y1ok = yi1 >= 0 && yi1 < M;
    if (regular && !(y0ok || y1ok)) continue;        // the fill already wrote this row
    const int4 xi = *reinterpret_cast<const int4*>(cxi0 + xg);
    const float4 w0 = *reinterpret_cast<const float4*>(cwx0 + xg);
    const float4 w1 = *reinterpret_cast<const float4*>(cwx1 + xg);
    const int xis[4] = {xi.x, xi.y, xi.z, xi.w};
    const float w0s[4] = {w0.x, w0.y, w0.z, w0.w}, w1s[4] = {w1.x, w1.y, w1.z, w1.w};
    uint8_t res[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xi0 = xis[j], xi1 = xis[j] + 1;
      const bool x0ok = xi0 >= 0 && xi0 < M, x1ok = xi1 >= 0 && xi1 < M;
      float v = 0.f;
      // ATen grid_sampler_2d bilinear: nw*(ix_se-ix)*(iy_se-iy) + ne*(ix-ix_sw)*(iy_sw-iy) + sw*... + se*...
      if (y0ok && x0ok) v += sm[yi0 * M + xi0] * (w0s[j] * wy0);
      if (y0ok && x1ok) v += sm[yi0 * M + xi1] * (w1s[j] * wy0);
      if (y1ok && x0ok) v += sm[yi1 * M + xi0] * (w0s[j] * wy1);
      if (y1ok && x1ok) v += sm[yi1 * M + xi1] * (w1s[j] * wy1);
      // NaN (degenerate box: 0/0) compares false, as `img >= threshold` does in the reference
      res[j] = (v >= threshold) ? 1 : 0;
    }
    uint8_t* o = out + (static_cast<size_t>(n) * H + y) * W + xg;
    if (xg + 4 <= W && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
      *reinterpret_cast<uchar4*>(o) = make_uchar4(res[0], res[1], res[2], res[3]);
    } else {
      for (int j = 0; j < 4 && xg + j < W; ++j) o[j] = res[j];
    }
  }
}

// One warp per output bin (m, ph, pw). masks: (G, H, W) bool bytes; gt_index[m] selects the mask.
__global__ void __launch_bounds__(256)
crop_resize_masks_kernel(const uint8_t* __restrict__ masks, const int64_t* __restrict__ gt_index,
                         const float* __restrict__ boxes, int Mrois, int H, int W, int P,
                         uint8_t* __restrict__ out_bool, float* __restrict__ out_val) {
  const int lane = threadIdx.x & 31;
  const long long bin = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bin >= static_cast<long long>(Mrois) * P * P) return;
  const int m = static_cast<int>(bin / (P * P));
  const int ph = static_cast<int>((bin / P) % P), pw = static_cast<int>(bin % P);
  const uint8_t* mk = masks + static_cast<size_t>(gt_index ? gt_index[m] : m) * H * W;
  const float* b = boxes + static_cast<size_t>(m) * 4;
  // torchvision roi_align, spatial_scale 1, aligned=True
  const float start_w = b[0] - 0.5f, start_h = b[1] - 0.5f;
  const float roi_w = (b[2] - 0.5f) - start_w, roi_h = (b[3] - 0.5f) - start_h;
  const float bin_h = roi_h / static_cast<float>(P), bin_w = roi_w / static_cast<float>(P);
  const int grid_h = static_cast<int>(ceilf(roi_h / P)), grid_w = static_cast<int>(ceilf(roi_w / P));
  const float count = fmaxf(static_cast<float>(grid_h) * static_cast<float>(grid_w), 1.f);
  float acc = 0.f;
  const int total = grid_h * grid_w;
  for (int s = lane; s < total; s += 32) {
    const int iy = s / grid_w, ix = s % grid_w;
    float y = start_h + ph * bin_h + (iy + 0.5f) * bin_h / static_cast<float>(grid_h);
    float x = start_w + pw * bin_w + (ix + 0.5f) * bin_w / static_cast<float>(grid_w);
    if (y < -1.0f || y > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W)) continue;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int y_low = static_cast<int>(y), x_low = static_cast<int>(x), y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else { y_high = y_low + 1; }
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else { x_high = x_low + 1; }
    const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
    const float v1 = mk[static_cast<size_t>(y_low) * W + x_low] ? 1.f : 0.f;
    const float v2 = mk[static_cast<size_t>(y_low) * W + x_high] ? 1.f : 0.f;
    const float v3 = mk[static_cast<size_t>(y_high) * W + x_low] ? 1.f : 0.f;
    const float v4 = mk[static_cast<size_t>(y_high) * W + x_high] ? 1.f : 0.f;
    acc += hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float v = acc / count;
    if (out_bool) out_bool[bin] = (v >= 0.5f) ? 1 : 0;  // masks.py:221 `output >= 0.5`
    if (out_val) out_val[bin] = v;
  }
}

}  // namespace

extern "C" {

int u2b_paste_masks(const float* masks, const float* boxes, int64_t N, int M, int H, int W,
                    float threshold, uint8_t* out, cudaStream_t stream) {
  if (N == 0) return 0;
  U2B_CHECK_ARG(masks && boxes && out && N > 0 && H > 0 && W > 0, "paste_masks: bad arguments");
  U2B_CHECK_ARG(M > 0 && M <= PASTE_MAX_M, "paste_masks: mask side %d unsupported (<= %d)", M, PASTE_MAX_M);
  U2B_CHECK_ARG(N <= 65535, "paste_masks: N too large for one launch");
  U2B_CHECK_ARG(W <= PASTE_MAX_W, "paste_masks: W=%d too wide (<= %d)", W, PASTE_MAX_W);
  dim3 grid((H + PASTE_ROWS - 1) / PASTE_ROWS, static_cast<unsigned>(N));
  const int Wpad = (W + 3) / 4 * 4;
  const size_t smem = static_cast<size_t>(Wpad) * 12;
  static bool attr = false;
  if (!attr) {
    U2B_CUDA(cudaFuncSetAttribute(paste_masks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PASTE_MAX_W * 12));
    attr = true;
  }
  paste_masks_kernel<<<grid, 256, smem, stream>>>(masks, boxes, (int)N, M, H, W, Wpad, threshold, out);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_crop_resize_masks(const uint8_t* masks, const int64_t* gt_index, const float* boxes,
                          int64_t M, int H, int W, int P, uint8_t* out_bool, float* out_val,
                          cudaStream_t stream) {
  if (M == 0) return 0;
  U2B_CHECK_ARG(masks && boxes && (out_bool || out_val) && H > 0 && W > 0 && P > 0,
                "crop_resize_masks: bad arguments");
  const long long bins = static_cast<long long>(M) * P * P;
  crop_resize_masks_kernel<<<static_cast<unsigned>((bins + 7) / 8), 256, 0, stream>>>(
      masks, gt_index, boxes, (int)M, H, W, P, out_bool, out_val);
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
