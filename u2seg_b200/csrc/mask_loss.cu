// Mask-head predictor + loss restricted to the one class the loss reads, forward and backward.
// Reference: detectron2/modeling/roi_heads/mask_head.py:33-112 (mask_rcnn_loss): the 1x1 predictor produces
// (R, num_classes, S, S) logits, the loss indexes the ground-truth class of every ROI (:95-97) and takes
// binary_cross_entropy_with_logits against the cropped GT masks (:111). With 800 pseudo-classes 799/800 of the predictor
// output is never read, so the predictor runs as a per-ROI dot product with the selected filter row (see
// MaskRCNNConvUpsampleHead.forward_selected) - here fused with the loss and with its closed-form gradient:
//   fwd : z[r,p] = x[r,p,:] . W[cls_r,:] + b[cls_r];  loss_r = sum_p bce(z, t);  g[r,p] = (sigmoid(z) - t) * ok_r
//   bwd : dX[r,p,:] = s * g[r,p] * W[cls_r,:];  dWr[r,:] = s * sum_p g[r,p] * x[r,p,:];  dbr[r] = s * sum_p g[r,p]
//         (s = upstream gradient of the loss SUM, a device scalar), then dW[k,:] = sum of dWr[r,:] over the ROIs of class k
//         in index order (deterministic, no atomics).
// HBM-bound: x (R, S*S, C) is read once per direction, dX written once: 103 MB each at R=256, S=28, C=256, bf16.
// One CTA per ROI, one warp per pixel: a lane owns 8 consecutive channels (one 16-byte vector) of each 256-channel group.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

constexpr int ML_THREADS = 256;
constexpr int ML_MAXG = 4;  // channel groups of 256: C <= 1024

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& r, float (&v)[8]);
template <>
__device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void unpack8<__half>(const uint4& r, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&v)[8]);
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return r;
}
template <>
__device__ __forceinline__ uint4 pack8<__half>(const float (&v)[8]) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  return r;
}

// grid (R, ML_SPLIT): CTA (r, s) takes pixels [s*chunk, (s+1)*chunk) of ROI r. A warp walks 32 pixels at a time: the 32
// dot products are reduced one after the other (independent loads, pipelined), lane j keeps pixel j's logit, then all
// lanes evaluate the loss / gradient of their pixel together (no single-lane transcendental chain) and store g coalesced.
constexpr int ML_SPLIT = 4;

template <typename T, int G>
__global__ void __launch_bounds__(ML_THREADS)
mask_loss_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                     const int64_t* __restrict__ classes, const uint8_t* __restrict__ target,
                     const uint8_t* __restrict__ ok, int P, float* __restrict__ g, float* __restrict__ loss_part) {
  constexpr int C = 256 * G;
  __shared__ float red[ML_THREADS / 32];
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = (P + ML_SPLIT - 1) / ML_SPLIT;
  const int p_begin = blockIdx.y * chunk, p_end = min(P, p_begin + chunk);
  const int64_t cls = classes[r];
  const float live = ok[r] ? 1.f : 0.f;
  float wr[G][8];
#pragma unroll
  for (int gI = 0; gI < G; ++gI)
    unpack8<T>(*reinterpret_cast<const uint4*>(w + cls * C + gI * 256 + lane * 8), wr[gI]);
  const float b = bias ? bias[cls] : 0.f;
  const T* xr = x + static_cast<size_t>(r) * P * C;
  float acc = 0.f;
  for (int p0 = p_begin + warp * 32; p0 < p_end; p0 += (ML_THREADS / 32) * 32) {
    float z = 0.f;
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
      const int p = p0 + j;
      if (p >= p_end) break;            // warp-uniform
      float d = 0.f;
#pragma unroll
      for (int gI = 0; gI < G; ++gI) {
        float v[8];
        unpack8<T>(*reinterpret_cast<const uint4*>(xr + static_cast<size_t>(p) * C + gI * 256 + lane * 8), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(v[i], wr[gI][i], d);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      if (lane == j) z = d + b;
    }
    const int p = p0 + lane;
    if (p < p_end) {
      const float t = target[static_cast<size_t>(r) * P + p] ? 1.f : 0.f;
      // binary_cross_entropy_with_logits (ATen): max(z,0) - z*t + log1p(exp(-|z|))
      acc += fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
      g[static_cast<size_t>(r) * P + p] = (1.f / (1.f + expf(-z)) - t) * live;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < ML_THREADS / 32; ++i) s += red[i];
    loss_part[static_cast<size_t>(r) * ML_SPLIT + blockIdx.y] = s * live;
  }
}

template <typename T, int G>
__global__ void __launch_bounds__(ML_THREADS)
mask_loss_bwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const int64_t* __restrict__ classes,
                     const float* __restrict__ g, const float* __restrict__ upstream, int P, T* __restrict__ dx,
                     float* __restrict__ dw_roi, float* __restrict__ db_roi) {
  constexpr int C = 256 * G;
  __shared__ float red[ML_THREADS / 32][C + 1];
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t cls = classes[r];
  const float s = *upstream;
  float wr[G][8], dwa[G][8];
#pragma unroll
  for (int gI = 0; gI < G; ++gI) {
    unpack8<T>(*reinterpret_cast<const uint4*>(w + cls * C + gI * 256 + lane * 8), wr[gI]);
#pragma unroll
    for (int i = 0; i < 8; ++i) dwa[gI][i] = 0.f;
  }
  const T* xr = x + static_cast<size_t>(r) * P * C;
  T* dxr = dx + static_cast<size_t>(r) * P * C;
  float dba = 0.f;
  for (int p = warp; p < P; p += ML_THREADS / 32) {
    const float gg = g[static_cast<size_t>(r) * P + p] * s;
    dba += gg;
#pragma unroll
    for (int gI = 0; gI < G; ++gI) {
      float v[8], o[8];
      unpack8<T>(*reinterpret_cast<const uint4*>(xr + static_cast<size_t>(p) * C + gI * 256 + lane * 8), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dwa[gI][i] = fmaf(gg, v[i], dwa[gI][i]);
        o[i] = gg * wr[gI][i];
      }
      *reinterpret_cast<uint4*>(dxr + static_cast<size_t>(p) * C + gI * 256 + lane * 8) = pack8<T>(o);
    }
  }
#pragma unroll
  for (int gI = 0; gI < G; ++gI)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][gI * 256 + lane * 8 + i] = dwa[gI][i];
  if (lane == 0) red[warp][C] = dba;
  __syncthreads();
  for (int c = threadIdx.x; c <= C; c += ML_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < ML_THREADS / 32; ++k) t += red[k][c];
    if (c < C) dw_roi[static_cast<size_t>(r) * C + c] = t;
    else db_roi[r] = t;
  }
}

// dW[k,:] = sum over ROIs r with classes[r] == k (index order) of dw_roi[r,:]; same for db. One CTA per ROI: the first
// ROI of each class owns the sum (the others exit after one block-wide vote). dW / db must be zero-filled by the caller
// (classes without an ROI). Deterministic.
__global__ void __launch_bounds__(256)
mask_loss_scatter_kernel(const int64_t* __restrict__ classes, const float* __restrict__ dw_roi,
                         const float* __restrict__ db_roi, int R, int C, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ int same[1024];
  __shared__ int n_same;
  const int r = blockIdx.x;
  const int64_t cls = classes[r];
  bool earlier = false;
  for (int j = threadIdx.x; j < r; j += blockDim.x) earlier |= classes[j] == cls;
  if (__syncthreads_or(earlier)) return;   // an earlier ROI owns this class
  if (threadIdx.x == 0) {                  // the (few) later ROIs of the same class, in index order
    int n = 0;
    for (int j = r; j < R && n < 1024; ++j)
      if (classes[j] == cls) same[n++] = j;
    n_same = n;
  }
  __syncthreads();
  const int n = n_same;
  for (int c = threadIdx.x; c <= C; c += blockDim.x) {
    float t = 0.f;
    for (int k = 0; k < n; ++k) t += (c < C) ? dw_roi[static_cast<size_t>(same[k]) * C + c] : db_roi[same[k]];
    if (c < C) dw[cls * C + c] = t;
    else db[cls] = t;
  }
}

}  // namespace

extern "C" {

int u2b_mask_loss_supported(int C) { return C > 0 && C % 256 == 0 && C <= 256 * ML_MAXG; }
int u2b_mask_loss_num_partials(void) { return ML_SPLIT; }

// dtype 1 = fp16, 2 = bf16. x (R, P, C) pooled-and-convolved ROI features (NHWC rows, P = S*S); w (K, C) predictor filter
// in x's dtype; bias (K) fp32 or NULL; classes (R) int64 in [0, K); target (R, P) bool; ok (R) bool (fixed-capacity
// slots: dead ROIs contribute nothing). Outputs: g (R, P) fp32 = d loss_sum / d logit, loss_per_roi (R, 4) fp32 partial
// sums (u2b_mask_loss_num_partials() per ROI; the caller adds them up).
int u2b_mask_loss_fwd(int dtype, const void* x, const void* w, const float* bias, const int64_t* classes,
                      const uint8_t* target, const uint8_t* ok, int64_t R, int P, int C, float* g, float* loss_per_roi,
                      cudaStream_t stream) {
  if (R == 0) return 0;
  U2B_CHECK_ARG(x && w && classes && target && ok && g && loss_per_roi && P > 0, "mask_loss_fwd: bad arguments");
  U2B_CHECK_ARG(u2b_mask_loss_supported(C) && (dtype == 1 || dtype == 2), "mask_loss_fwd: C=%d dtype=%d unsupported", C, dtype);
  const dim3 grid(static_cast<unsigned>(R), ML_SPLIT);
#define U2B_ML_FWD(T, G)                                                                                              \
  mask_loss_fwd_kernel<T, G><<<grid, ML_THREADS, 0, stream>>>(static_cast<const T*>(x), static_cast<const T*>(w), bias, \
                                                              classes, target, ok, P, g, loss_per_roi)
  const int G = C / 256;
  if (dtype == 2) {
    if (G == 1) U2B_ML_FWD(__nv_bfloat16, 1); else if (G == 2) U2B_ML_FWD(__nv_bfloat16, 2);
    else if (G == 3) U2B_ML_FWD(__nv_bfloat16, 3); else U2B_ML_FWD(__nv_bfloat16, 4);
  } else {
    if (G == 1) U2B_ML_FWD(__half, 1); else if (G == 2) U2B_ML_FWD(__half, 2);
    else if (G == 3) U2B_ML_FWD(__half, 3); else U2B_ML_FWD(__half, 4);
  }
#undef U2B_ML_FWD
  U2B_LAUNCH_CHECK();
  return 0;
}

// upstream: device scalar, gradient of the loss SUM. dx (R, P, C) in x's dtype; dw (K, C), db (K) fp32, ZERO-FILLED by the
// caller; workspace: R * (C + 1) floats.
int u2b_mask_loss_bwd(int dtype, const void* x, const void* w, const int64_t* classes, const float* g,
                      const float* upstream, int64_t R, int P, int C, void* dx, float* dw, float* db, float* workspace,
                      cudaStream_t stream) {
  if (R == 0) return 0;
  U2B_CHECK_ARG(x && w && classes && g && upstream && dx && dw && db && workspace && P > 0, "mask_loss_bwd: bad arguments");
  U2B_CHECK_ARG(u2b_mask_loss_supported(C) && (dtype == 1 || dtype == 2), "mask_loss_bwd: C=%d dtype=%d unsupported", C, dtype);
  float* dw_roi = workspace;
  float* db_roi = workspace + static_cast<size_t>(R) * C;
  const unsigned grid = static_cast<unsigned>(R);
#define U2B_ML_BWD(T, G)                                                                                               \
  mask_loss_bwd_kernel<T, G><<<grid, ML_THREADS, 0, stream>>>(static_cast<const T*>(x), static_cast<const T*>(w), classes, g, \
                                                              upstream, P, static_cast<T*>(dx), dw_roi, db_roi)
  const int G = C / 256;
  if (dtype == 2) {
    if (G == 1) U2B_ML_BWD(__nv_bfloat16, 1); else if (G == 2) U2B_ML_BWD(__nv_bfloat16, 2);
    else if (G == 3) U2B_ML_BWD(__nv_bfloat16, 3); else U2B_ML_BWD(__nv_bfloat16, 4);
  } else {
    if (G == 1) U2B_ML_BWD(__half, 1); else if (G == 2) U2B_ML_BWD(__half, 2);
    else if (G == 3) U2B_ML_BWD(__half, 3); else U2B_ML_BWD(__half, 4);
  }
#undef U2B_ML_BWD
  U2B_LAUNCH_CHECK();
  mask_loss_scatter_kernel<<<grid, 256, 0, stream>>>(classes, dw_roi, db_roi, static_cast<int>(R), C, dw, db);
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
