// NHWC max pooling of the ResNet stem and the 2x2 gradient fold of the FPN top-down path.
//  * max_pool 3x3 / stride 2 / pad 1 (detectron2/modeling/backbone/resnet.py:358 F.max_pool2d): forward writes the pooled
//    map and a 1-byte window position of the FIRST maximum in (kh, kw) scan order (ATen's tie rule: `val > max`), backward
//    GATHERS: every input pixel looks at the <= 4 windows that contain it and takes their gradient where it is the
//    recorded position - no atomics, no int64 index tensor (ATen writes 8 bytes of index per pooled element).
//  * sum2x2: d(prev) of `lateral + nearest_upsample_x2(prev)` (backbone/fpn.py:153-156) = sum of the gradient over each
//    2x2 block (the forward add + upsample is folded into the SyncBN apply pass of the lateral conv, csrc/batchnorm.cu).
// HBM-bound; a thread owns 8 consecutive channels (one 16-byte vector) of one pixel.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

template <typename T>
struct H8;
template <>
struct H8<__nv_bfloat16> {
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return r;
  }
};
template <>
struct H8<__half> {
  static __device__ __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[8]) {
    uint4 r;
    __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    return r;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int OH, int OW, T* __restrict__ y,
                        uint8_t* __restrict__ idx) {
  const int vecs = C / 8;
  const long long total = static_cast<long long>(N) * OH * OW * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vecs);
    long long pix = i / vecs;
    const int ow = static_cast<int>(pix % OW); pix /= OW;
    const int oh = static_cast<int>(pix % OH);
    const int n = static_cast<int>(pix / OH);
    float best[8];
    int pos[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      best[k] = -INFINITY;
      pos[k] = 0;
    }
    bool first = true;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = oh * 2 - 1 + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = ow * 2 - 1 + kw;
        if (w < 0 || w >= W) continue;
        float v[8];
        H8<T>::unpack(*reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + h) * W + w) * C + cv * 8), v);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (first || v[k] > best[k] || v[k] != v[k]) {   // ATen: (val > maxval) || isnan(val); first valid tap seeds
            best[k] = v[k];
            pos[k] = kh * 3 + kw;
          }
        first = false;
      }
    }
    const size_t o = ((static_cast<size_t>(n) * OH + oh) * OW + ow) * C + cv * 8;
    *reinterpret_cast<uint4*>(y + o) = H8<T>::pack(best);
    uint2 pk;
    pk.x = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (pos[3] << 24);
    pk.y = pos[4] | (pos[5] << 8) | (pos[6] << 16) | (pos[7] << 24);
    *reinterpret_cast<uint2*>(idx + o) = pk;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, int N, int H, int W, int C, int OH,
                        int OW, T* __restrict__ dx) {
  const int vecs = C / 8;
  const long long total = static_cast<long long>(N) * H * W * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vecs);
    long long pix = i / vecs;
    const int w = static_cast<int>(pix % W); pix /= W;
    const int h = static_cast<int>(pix % H);
    const int n = static_cast<int>(pix / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int oh0 = h / 2, oh1 = min(OH - 1, (h + 1) / 2), ow0 = w / 2, ow1 = min(OW - 1, (w + 1) / 2);
    for (int oh = oh0; oh <= oh1; ++oh)
      for (int ow = ow0; ow <= ow1; ++ow) {
        const int me = (h - (oh * 2 - 1)) * 3 + (w - (ow * 2 - 1));
        const size_t o = ((static_cast<size_t>(n) * OH + oh) * OW + ow) * C + cv * 8;
        const uint2 pk = *reinterpret_cast<const uint2*>(idx + o);
        float g[8];
        H8<T>::unpack(*reinterpret_cast<const uint4*>(dy + o), g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int p = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xff;
          if (p == me) acc[k] += g[k];
        }
      }
    *reinterpret_cast<uint4*>(dx + ((static_cast<size_t>(n) * H + h) * W + w) * C + cv * 8) = H8<T>::pack(acc);
  }
}

// y[n, h, w, :] = x[n, 2h, 2w, :] + x[n, 2h, 2w+1, :] + x[n, 2h+1, 2w, :] + x[n, 2h+1, 2w+1, :]  (fp32 sum, one rounding)
template <typename T>
__global__ void __launch_bounds__(256)
sum2x2_kernel(const T* __restrict__ x, int N, int H, int W, int C, T* __restrict__ y) {
  const int vecs = C / 8, OH = H / 2, OW = W / 2;
  const long long total = static_cast<long long>(N) * OH * OW * vecs;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vecs);
    long long pix = i / vecs;
    const int ow = static_cast<int>(pix % OW); pix /= OW;
    const int oh = static_cast<int>(pix % OH);
    const int n = static_cast<int>(pix / OH);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        float v[8];
        H8<T>::unpack(*reinterpret_cast<const uint4*>(
                          x + ((static_cast<size_t>(n) * H + oh * 2 + dh) * W + ow * 2 + dw) * C + cv * 8), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
      }
    *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * OH + oh) * OW + ow) * C + cv * 8) = H8<T>::pack(acc);
  }
}

inline unsigned pool_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = static_cast<long long>(u2b_num_sms()) * 16;
  return static_cast<unsigned>(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" {

// dtype 1 = fp16, 2 = bf16. x (N,H,W,C) NHWC, C % 8 == 0. y, idx: (N,OH,OW,C), OH = (H - 1) / 2 + 1.
int u2b_maxpool3x3s2_fwd(int dtype, const void* x, int N, int H, int W, int C, void* y, uint8_t* idx, cudaStream_t stream) {
  U2B_CHECK_ARG(x && y && idx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "maxpool3x3s2_fwd: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "maxpool3x3s2_fwd: dtype must be fp16(1) or bf16(2)");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(N) * OH * OW * (C / 8);
  if (dtype == 2)
    maxpool3x3s2_fwd_kernel<__nv_bfloat16><<<pool_grid(total), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(x), N, H, W, C, OH, OW, static_cast<__nv_bfloat16*>(y), idx);
  else
    maxpool3x3s2_fwd_kernel<__half><<<pool_grid(total), 256, 0, stream>>>(static_cast<const __half*>(x), N, H, W, C, OH, OW,
                                                                         static_cast<__half*>(y), idx);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_maxpool3x3s2_bwd(int dtype, const void* dy, const uint8_t* idx, int N, int H, int W, int C, void* dx,
                         cudaStream_t stream) {
  U2B_CHECK_ARG(dy && idx && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "maxpool3x3s2_bwd: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "maxpool3x3s2_bwd: dtype must be fp16(1) or bf16(2)");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total = static_cast<long long>(N) * H * W * (C / 8);
  if (dtype == 2)
    maxpool3x3s2_bwd_kernel<__nv_bfloat16><<<pool_grid(total), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(dy), idx, N, H, W, C, OH, OW, static_cast<__nv_bfloat16*>(dx));
  else
    maxpool3x3s2_bwd_kernel<__half><<<pool_grid(total), 256, 0, stream>>>(static_cast<const __half*>(dy), idx, N, H, W, C, OH,
                                                                         OW, static_cast<__half*>(dx));
  U2B_LAUNCH_CHECK();
  return 0;
}

// x (N,H,W,C) -> y (N,H/2,W/2,C): sums over 2x2 blocks; H, W even.
int u2b_sum2x2_nhwc(int dtype, const void* x, int N, int H, int W, int C, void* y, cudaStream_t stream) {
  U2B_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0, "sum2x2_nhwc: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "sum2x2_nhwc: dtype must be fp16(1) or bf16(2)");
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  if (dtype == 2)
    sum2x2_kernel<__nv_bfloat16><<<pool_grid(total), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), N, H, W, C,
                                                                      static_cast<__nv_bfloat16*>(y));
  else
    sum2x2_kernel<__half><<<pool_grid(total), 256, 0, stream>>>(static_cast<const __half*>(x), N, H, W, C, static_cast<__half*>(y));
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Image preprocessing, detectron2/modeling/meta_arch/rcnn.py:223-234 + structures/image_list.py:59-129 for a batch of
// same-size uint8 images already stored NHWC: out[n,h,w,c] = (float(img[n,h,w,c]) - mean[c]) / std[c] for h < H, w < W,
// zero in the padding up to (Hp, Wp) (size_divisibility). fp32 arithmetic in the reference's op order (subtract, then
// IEEE divide), then one rounding to the output dtype (what autocast's cast of the stem input does).
namespace {
template <typename OutT>
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ img, int N, int H, int W, int Hp, int Wp, float m0, float m1, float m2,
                     float s0, float s1, float s2, OutT* __restrict__ out) {
  const long long total = static_cast<long long>(N) * Hp * Wp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % Wp);
    const int h = static_cast<int>((i / Wp) % Hp);
    const long long n = i / (static_cast<long long>(Wp) * Hp);
    float r = 0.f, g = 0.f, b = 0.f;
    if (h < H && w < W) {
      const uint8_t* p = img + ((n * H + h) * W + w) * 3;
      r = __fdiv_rn(static_cast<float>(p[0]) - m0, s0);
      g = __fdiv_rn(static_cast<float>(p[1]) - m1, s1);
      b = __fdiv_rn(static_cast<float>(p[2]) - m2, s2);
    }
    OutT* o = out + i * 3;
    o[0] = static_cast<OutT>(r);
    o[1] = static_cast<OutT>(g);
    o[2] = static_cast<OutT>(b);
  }
}
}  // namespace

extern "C" {
// img (N,H,W,3) uint8 NHWC; out (N,Hp,Wp,3) NHWC, out_dtype 0 = fp32, 1 = fp16, 2 = bf16; mean/std: 3 host floats each.
int u2b_preprocess_u8_nhwc(const uint8_t* img, int N, int H, int W, int Hp, int Wp, const float* mean3, const float* std3,
                           int out_dtype, void* out, cudaStream_t stream) {
  U2B_CHECK_ARG(img && out && mean3 && std3 && N > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, "preprocess_u8_nhwc: bad arguments");
  const long long total = static_cast<long long>(N) * Hp * Wp;
  const unsigned grid = pool_grid(total);
  if (out_dtype == 0)
    preprocess_u8_kernel<float><<<grid, 256, 0, stream>>>(img, N, H, W, Hp, Wp, mean3[0], mean3[1], mean3[2], std3[0],
                                                          std3[1], std3[2], static_cast<float*>(out));
  else if (out_dtype == 1)
    preprocess_u8_kernel<__half><<<grid, 256, 0, stream>>>(img, N, H, W, Hp, Wp, mean3[0], mean3[1], mean3[2], std3[0],
                                                           std3[1], std3[2], static_cast<__half*>(out));
  else if (out_dtype == 2)
    preprocess_u8_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(img, N, H, W, Hp, Wp, mean3[0], mean3[1], mean3[2],
                                                                  std3[0], std3[1], std3[2], static_cast<__nv_bfloat16*>(out));
  else {
    u2b_set_error("preprocess_u8_nhwc: out_dtype %d", out_dtype);
    return U2B_ERR_BAD_ARG;
  }
  U2B_LAUNCH_CHECK();
  return 0;
}
}  // extern "C"
