// Bilinear upsampling by an integer factor on NHWC activations, forward and backward
// (nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False) of the semantic head, semantic_seg.py:195-199).
// ROUND-2 DRAFT. The library's bf16 NHWC kernels reach 0.4 TB/s on these shapes (82 us forward / 106 us backward per
// call on 2x128x256x256 outputs, 6 calls per step); the op is a pure streaming one: forward writes N*H*W*C elements and
// reads a quarter of that, backward the reverse. One thread per (pixel, 8-channel vector), 16-byte accesses, fp32
// arithmetic with PyTorch's source-index formula src = max(0, (dst + 0.5) / s - 0.5) and operation order
// (upsample_bilinear2d_out_frame), so fp32 results agree to rounding and bf16 results are rounded once.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

template <typename T>
struct V8;
template <>
struct V8<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct V8<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};
template <>
struct V8<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint4 r;
    __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

__device__ __forceinline__ void src_coord(int dst, float rscale, int in_size, int& i0, int& i1, float& l1) {
  float s = rscale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

// y (N, h*s, w*s, C) <- x (N, h, w, C)
template <typename T>
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int h, int w, int C, int s, long long total_vec) {
  constexpr int VN = V8<T>::N;
  const int vecs = C / VN, H = h * s, W = w * s;
  const float rs = 1.0f / static_cast<float>(s);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vecs);
    long long p = i / vecs;
    const int X = static_cast<int>(p % W);
    p /= W;
    const int Y = static_cast<int>(p % H);
    const long long n = p / H;
    int y0, y1, x0, x1;
    float ly, lx;
    src_coord(Y, rs, h, y0, y1, ly);
    src_coord(X, rs, w, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const T* xn = x + n * h * w * C + v * VN;
    float a[VN], b[VN], c[VN], d[VN], o[VN];
    V8<T>::load(xn + (static_cast<long long>(y0) * w + x0) * C, a);
    V8<T>::load(xn + (static_cast<long long>(y0) * w + x1) * C, b);
    V8<T>::load(xn + (static_cast<long long>(y1) * w + x0) * C, c);
    V8<T>::load(xn + (static_cast<long long>(y1) * w + x1) * C, d);
#pragma unroll
    for (int k = 0; k < VN; ++k) o[k] = hy * (hx * a[k] + lx * b[k]) + ly * (hx * c[k] + lx * d[k]);
    V8<T>::store(y + i * VN, o);
  }
}

// dx (N, h, w, C) <- dy (N, h*s, w*s, C): gather form of the transposed interpolation (no atomics, deterministic)
template <typename T>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int h, int w, int C, int s, long long total_vec) {
  constexpr int VN = V8<T>::N;
  const int vecs = C / VN, H = h * s, W = w * s;
  const float rs = 1.0f / static_cast<float>(s);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vecs);
    long long p = i / vecs;
    const int xi = static_cast<int>(p % w);
    p /= w;
    const int yi = static_cast<int>(p % h);
    const long long n = p / h;
    // output rows with weight on input row yi lie in [s*yi - s/2, s*yi + s + s/2 - 1] (+ the clamped border rows)
    int Ya = s * yi - s / 2, Yb = s * yi + s + s / 2 - 1, Xa = s * xi - s / 2, Xb = s * xi + s + s / 2 - 1;
    if (yi == 0) Ya = 0;
    if (xi == 0) Xa = 0;
    if (yi == h - 1) Yb = H - 1;
    if (xi == w - 1) Xb = W - 1;
    Ya = max(Ya, 0); Xa = max(Xa, 0); Yb = min(Yb, H - 1); Xb = min(Xb, W - 1);
    const T* gn = dy + n * H * W * C + v * VN;
    float acc[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[k] = 0.f;
    for (int Y = Ya; Y <= Yb; ++Y) {
      int y0, y1;
      float ly;
      src_coord(Y, rs, h, y0, y1, ly);
      const float wy = (y0 == yi ? 1.f - ly : 0.f) + (y1 == yi ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int X = Xa; X <= Xb; ++X) {
        int x0, x1;
        float lx;
        src_coord(X, rs, w, x0, x1, lx);
        const float wx = (x0 == xi ? 1.f - lx : 0.f) + (x1 == xi ? lx : 0.f);
        if (wx == 0.f) continue;
        float g[VN];
        V8<T>::load(gn + (static_cast<long long>(Y) * W + X) * C, g);
        const float wgt = wy * wx;
#pragma unroll
        for (int k = 0; k < VN; ++k) acc[k] = fmaf(wgt, g[k], acc[k]);
      }
    }
    V8<T>::store(dx + i * VN, acc);
  }
}

inline unsigned grid_for(long long total_vec) {
  long long b = (total_vec + 255) / 256;
  const long long cap = static_cast<long long>(u2b_num_sms()) * 16;
  return static_cast<unsigned>(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" {

int u2b_upsample_bilinear_supported(int C, int scale) { return C > 0 && C % 8 == 0 && scale >= 2 && scale % 2 == 0; }

// dir 0: y (N, h*s, w*s, C) = upsample(x (N, h, w, C));  dir 1: dx (N, h, w, C) = transposed op of dy (N, h*s, w*s, C)
int u2b_upsample_bilinear(int dtype, int dir, const void* in, void* out, int64_t N, int h, int w, int C, int scale,
                          cudaStream_t stream) {
  if (N == 0) return 0;
  U2B_CHECK_ARG(in && out && N > 0 && h > 0 && w > 0 && u2b_upsample_bilinear_supported(C, scale) && (dir == 0 || dir == 1),
                "upsample_bilinear: bad arguments");
  U2B_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "upsample_bilinear: buffers must be 16-byte aligned");
  const int vn = dtype == 0 ? 4 : 8;
  const long long px = dir == 0 ? static_cast<long long>(N) * h * scale * w * scale : static_cast<long long>(N) * h * w;
  const long long tv = px * (C / vn);
  const unsigned grid = grid_for(tv);
#define U2B_UP(T)                                                                                              \
  do {                                                                                                         \
    if (dir == 0)                                                                                              \
      upsample_fwd_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(in), static_cast<T*>(out), h, w, C, \
                                                       scale, tv);                                              \
    else                                                                                                       \
      upsample_bwd_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(in), static_cast<T*>(out), h, w, C, \
                                                       scale, tv);                                              \
  } while (0)
  if (dtype == 0) U2B_UP(float);
  else if (dtype == 1) U2B_UP(__half);
  else if (dtype == 2) U2B_UP(__nv_bfloat16);
  else {
    u2b_set_error("upsample_bilinear: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
#undef U2B_UP
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
