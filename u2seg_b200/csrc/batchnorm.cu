// Training-mode (Sync)BatchNorm for NHWC activations, fused with the residual add and ReLU that follow it in
// the ResNet/FPN blocks. Reference: detectron2/layers/batch_norm.py:187 (nn.SyncBatchNorm) as used by
// layers/wrappers.py:87-134 (Conv2d.forward: conv -> norm -> activation) and backbone/resnet.py:194-210
// (out += shortcut; relu).
//
// HBM-bound. Forward = 2 reads + 1 write of the activation (library path: 2 reads + 1 write for BN, then 1 read +
// 1 write each for the add and the ReLU); backward = 5 reads + 1-2 writes, no separate ReLU/add backward kernels.
//   reduce (MODE 0: sum x, sum x^2; MODE 1: sum dz, sum dz*xhat with dz = dy*(y>0)):
//       grid = (pixel strips, channel groups of 256); a thread owns 8 channels (one 16-byte vector) and a pixel
//       lane, keeps 4 independent loads in flight, block-reduces through shared memory and writes ONE partial row
//       per strip - no atomics, deterministic. Partials are summed by the finalize / coefficient kernels (or by
//       sum_partials when a data-parallel job has to all-reduce the (2C) sums in between).
//   finalize  : mean / invstd / scale / shift, running statistics (momentum, unbiased variance)
//   apply     : y = relu(x * scale[c] + shift[c] + residual); per-thread channel coefficients live in registers
//   bwd_coeff : dx = A[c]*dz + B[c]*x + K[c] with A = w*invstd, B = -A*invstd*S2/n, K = -A*S1/n - B*mean;
//               also emits dgamma = S2, dbeta = S1 (local sums)
//   bwd_apply : dx (and dres = dz), coefficients in registers
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

template <typename T>
struct V8;  // 8 channels per thread
template <>
struct V8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw ldraw(const float* p) {
    Raw r;
    r.a = reinterpret_cast<const float4*>(p)[0];
    r.b = reinterpret_cast<const float4*>(p)[1];
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <>
struct V8<__half> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw ldraw(const __half* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
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
template <>
struct V8<__nv_bfloat16> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void unpack(const Raw& r, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
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

// x * scale + shift as the forward pass STORED it (rounded to T): the ReLU mask y > 0 recomputed without reading y
template <typename T>
__device__ __forceinline__ float as_stored(float z);
template <>
__device__ __forceinline__ float as_stored<float>(float z) { return z; }
template <>
__device__ __forceinline__ float as_stored<__half>(float z) { return __half2float(__float2half_rn(z)); }
template <>
__device__ __forceinline__ float as_stored<__nv_bfloat16>(float z) { return __bfloat162float(__float2bfloat16_rn(z)); }

constexpr int BN_THREADS = 256;
#ifndef BN_STRIPS_PER_SM
#define BN_STRIPS_PER_SM 4
#endif
constexpr int BN_VPB = 32;  // channel vectors (of 8) per block -> 256 channels per channel group

inline int vecs_per_block(int C) { return (C / 8 < BN_VPB) ? C / 8 : BN_VPB; }
inline int channel_groups(int C) { return (C / 8 + BN_VPB - 1) / BN_VPB; }
inline int num_strips(long long P, int C) {
  const int lanes = BN_THREADS / vecs_per_block(C);
  long long s = P / (static_cast<long long>(lanes) * 8);
  // r2 draft: 4 strips per SM (was 2). ncu (profiles/r01_ncu_full_stem_fwd_bn_reduce.txt): 23 % occupancy and 2.1 TB/s
  // on 67 MB tensors with 2 CTAs/SM; more CTAs in flight hide the DRAM latency of the short strips of small layers.
  const long long cap = static_cast<long long>(u2b_num_sms()) * BN_STRIPS_PER_SM / channel_groups(C);
  if (s > cap) s = cap;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS, (MODE == 0 ? 4 : 2))
bn_reduce_kernel(const T* __restrict__ a, const T* __restrict__ x, const T* __restrict__ y,
                 const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int C,
                 int vpb, float* __restrict__ partials, const float* __restrict__ relu_scale_shift = nullptr) {
  ptx_free::pdl_prologue();
  __shared__ float red[BN_THREADS * 16];
  const int lanes = BN_THREADS / vpb;
  const int tv = threadIdx.x % vpb, tp = threadIdx.x / vpb;
  const int vec = blockIdx.y * vpb + tv;
  const bool active = vec * 8 < C && tp < lanes;
  const int c0 = vec * 8;
  float s0[8], s1[8], m[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
  float rsc[8], rsh[8];   // relu_scale_shift != NULL: mask = (stored(x * scale + shift) > 0), y is not read
  if (MODE == 1 && active) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      m[i] = mean[c0 + i];
      is[i] = invstd[c0 + i];
      rsc[i] = relu_scale_shift ? relu_scale_shift[c0 + i] : 0.f;
      rsh[i] = relu_scale_shift ? relu_scale_shift[C + c0 + i] : 0.f;
    }
  }
  const long long strip = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * strip;
  const long long p1 = (p0 + strip < P) ? p0 + strip : P;
  if (active) {
    constexpr int U = 4;
    using Raw = typename V8<T>::Raw;
    for (long long pb = p0 + tp; pb < p1; pb += static_cast<long long>(lanes) * U) {
      Raw ra[U], rx[U], ry[U];   // packed loads stay in flight; unpacked one at a time below
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = pb + static_cast<long long>(u) * lanes;
        ok[u] = p < p1;
        if (ok[u]) {
          ra[u] = V8<T>::ldraw(a + p * C + c0);
          if (MODE == 1) {
            rx[u] = V8<T>::ldraw(x + p * C + c0);
            if (y) ry[u] = V8<T>::ldraw(y + p * C + c0);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        float va[8];
        V8<T>::unpack(ra[u], va);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            s0[i] += va[i];
            s1[i] = fmaf(va[i], va[i], s1[i]);
          }
        } else {
          float vx[8];
          V8<T>::unpack(rx[u], vx);
          if (y) {
            float vy[8];
            V8<T>::unpack(ry[u], vy);
#pragma unroll
            for (int i = 0; i < 8; ++i) va[i] = vy[i] > 0.f ? va[i] : 0.f;
          } else if (relu_scale_shift) {
#pragma unroll
            for (int i = 0; i < 8; ++i) va[i] = as_stored<T>(fmaf(vx[i], rsc[i], rsh[i])) > 0.f ? va[i] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            s0[i] += va[i];
            s1[i] = fmaf(va[i], (vx[i] - m[i]) * is[i], s1[i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x * 16 + i] = s0[i];
    red[threadIdx.x * 16 + 8 + i] = s1[i];
  }
  __syncthreads();
  if (tp == 0 && vec * 8 < C) {
    float t0[8], t1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t0[i] = t1[i] = 0.f;
    for (int q = 0; q < lanes; ++q) {
      const float* r = red + (q * vpb + tv) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t0[i] += r[i];
        t1[i] += r[8 + i];
      }
    }
    float* out = partials + static_cast<size_t>(blockIdx.x) * 2 * C;
    V8<float>::store(out + c0, t0);
    V8<float>::store(out + C + c0, t1);
  }
}

// block = 8 columns x 32 strip lanes (like bn_finalize_kernel): the S rows are summed by 32 lanes in parallel and
// combined in a fixed order, instead of one thread walking all S rows serially.
__global__ void __launch_bounds__(256)
bn_sum_partials_kernel(const float* __restrict__ partials, int S, int C2, float* __restrict__ sums) {
  ptx_free::pdl_prologue();
  __shared__ float sh[32][8];
  const int cl = threadIdx.x & 7, sl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  float s = 0.f;
  if (c < C2)
    for (int i = sl; i < S; i += 32) s += partials[static_cast<size_t>(i) * C2 + c];
  sh[sl][cl] = s;
  __syncthreads();
  if (sl != 0 || c >= C2) return;
  for (int q = 1; q < 32; ++q) s += sh[q][cl];
  sums[c] = s;
}

__global__ void bn_finalize_kernel(const float* __restrict__ partials, int S, double n_total,
                                   const float* __restrict__ w, const float* __restrict__ b, float eps,
                                   float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ stats, int C) {
  ptx_free::pdl_prologue();
  // block = 8 channels x 32 strip lanes; lanes sum strided partial rows, then combine through shared memory
  __shared__ double sh[2][32][8];
  const int cl = threadIdx.x & 7, sl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  double s = 0.0, ss = 0.0;
  if (c < C)
    for (int i = sl; i < S; i += 32) {
      s += partials[static_cast<size_t>(i) * 2 * C + c];
      ss += partials[static_cast<size_t>(i) * 2 * C + C + c];
    }
  sh[0][sl][cl] = s;
  sh[1][sl][cl] = ss;
  __syncthreads();
  if (sl != 0 || c >= C) return;
  for (int q = 1; q < 32; ++q) {
    s += sh[0][q][cl];
    ss += sh[1][q][cl];
  }
  const double mu = s / n_total;
  double var = ss / n_total - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = (w ? w[c] : 1.f) * is;
  stats[c] = static_cast<float>(mu);
  stats[C + c] = is;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = (b ? b[c] : 0.f) - static_cast<float>(mu) * sc;
  if (running_mean) {
    const double unbiased = n_total > 1.0 ? var * n_total / (n_total - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void bn_bwd_coeff_kernel(const float* __restrict__ partials, int S, double n_total,
                                    const float* __restrict__ stats, const float* __restrict__ w,
                                    float* __restrict__ coeff, float* __restrict__ gw_gb, int C) {
  ptx_free::pdl_prologue();
  __shared__ float sh[2][32][8];
  const int cl = threadIdx.x & 7, sl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  float s1 = 0.f, s2 = 0.f;
  if (c < C)
    for (int i = sl; i < S; i += 32) {
      s1 += partials[static_cast<size_t>(i) * 2 * C + c];
      s2 += partials[static_cast<size_t>(i) * 2 * C + C + c];
    }
  sh[0][sl][cl] = s1;
  sh[1][sl][cl] = s2;
  __syncthreads();
  if (sl != 0 || c >= C) return;
  for (int q = 1; q < 32; ++q) {
    s1 += sh[0][q][cl];
    s2 += sh[1][q][cl];
  }
  if (gw_gb) {
    gw_gb[c] = s2;      // dgamma = sum dz * xhat
    gw_gb[C + c] = s1;  // dbeta  = sum dz
  }
  const float mu = stats[c], is = stats[C + c];
  const float inv_n = static_cast<float>(1.0 / n_total);
  const float A = (w ? w[c] : 1.f) * is;
  const float B = -A * is * s2 * inv_n;
  coeff[c] = A;
  coeff[C + c] = B;
  coeff[2 * C + c] = -A * s1 * inv_n - B * mu;
}

// ---- GroupNorm (layers/batch_norm.py get_norm "GN" -> nn.GroupNorm(32, C), sem-seg head) on the same per-channel
// machinery: statistics are per IMAGE and per GROUP of C/G channels, so the strip partial sums of one image
// (bn_reduce kernels with P = H*W) are folded across each group's channels here and written back as per-channel
// mean | invstd | scale | shift (forward) or dx coefficients (backward); the apply kernels are the BN ones.
// One CTA of 1024 threads: 1024/C strip lanes per channel. C a power of two <= 1024.
template <typename Acc>
__device__ __forceinline__ void gn_channel_sums(const float* __restrict__ partials, int S, int C, Acc* sh1, Acc* sh2) {
  const int lanes = 1024 / C;
  const int c = threadIdx.x % C, l = threadIdx.x / C;
  Acc a = 0, b = 0;
  for (int i = l; i < S; i += lanes) {
    a += partials[static_cast<size_t>(i) * 2 * C + c];
    b += partials[static_cast<size_t>(i) * 2 * C + C + c];
  }
  sh1[threadIdx.x] = a;
  sh2[threadIdx.x] = b;
  __syncthreads();
  if (l == 0) {
    for (int q = 1; q < lanes; ++q) {
      a += sh1[q * C + c];
      b += sh2[q * C + c];
    }
  }
  __syncthreads();
  if (l == 0) {
    sh1[c] = a;
    sh2[c] = b;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
gn_finalize_kernel(const float* __restrict__ partials, int S, double m, int G, const float* __restrict__ w,
                   const float* __restrict__ b, float eps, float* __restrict__ stats, int C) {
  __shared__ double sh1[1024], sh2[1024];
  gn_channel_sums<double>(partials, S, C, sh1, sh2);
  const int c = threadIdx.x;
  if (c >= C) return;
  const int cpg = C / G, g0 = (c / cpg) * cpg;
  double s = 0.0, ss = 0.0;
  for (int k = 0; k < cpg; ++k) {
    s += sh1[g0 + k];
    ss += sh2[g0 + k];
  }
  const double mu = s / m;
  double var = ss / m - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = (w ? w[c] : 1.f) * is;
  stats[c] = static_cast<float>(mu);
  stats[C + c] = is;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = (b ? b[c] : 0.f) - static_cast<float>(mu) * sc;
}

// partials: (sum dz | sum dz*xhat) per channel of ONE image. coeff (3C): dx = A*dz + B*x + K.
// gw_gb (2C): dgamma | dbeta, accumulated over the images of the batch when `accumulate`.
__global__ void __launch_bounds__(1024)
gn_bwd_coeff_kernel(const float* __restrict__ partials, int S, double m, int G, const float* __restrict__ stats,
                    const float* __restrict__ w, float* __restrict__ coeff, float* __restrict__ gw_gb,
                    int accumulate, int C) {
  __shared__ float sh1[1024], sh2[1024];
  __shared__ float gam[1024];
  gn_channel_sums<float>(partials, S, C, sh1, sh2);
  const int c = threadIdx.x;
  if (c < C) gam[c] = w ? w[c] : 1.f;
  __syncthreads();
  if (c >= C) return;
  const float s1 = sh1[c], s2 = sh2[c];
  if (gw_gb) {
    gw_gb[c] = (accumulate ? gw_gb[c] : 0.f) + s2;
    gw_gb[C + c] = (accumulate ? gw_gb[C + c] : 0.f) + s1;
  }
  const int cpg = C / G, g0 = (c / cpg) * cpg;
  float ds = 0.f, db = 0.f;
  for (int k = 0; k < cpg; ++k) {
    ds += gam[g0 + k] * sh1[g0 + k];
    db += gam[g0 + k] * sh2[g0 + k];
  }
  const float mu = stats[c], is = stats[C + c];
  const float inv_m = static_cast<float>(1.0 / m);
  const float A = gam[c] * is;
  const float B = -is * is * db * inv_m;
  coeff[c] = A;
  coeff[C + c] = B;
  coeff[2 * C + c] = -is * ds * inv_m - B * mu;
}

// total_vec vectors of 8 channels; (blockDim*gridDim) % (C/8) == 0 so a thread's channels never change
// up_h/up_w > 0: `residual` is a (N, up_h/2, up_w/2, C) map added through a nearest-neighbour x2 upsampling (the FPN
// top-down sum, backbone/fpn.py:153-156, folded into the lateral conv's SyncBN pass); x / y are (N, up_h, up_w, C).
template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ stats, const T* __restrict__ residual, int relu,
                T* __restrict__ y, long long total_vec, int C, int up_h = 0, int up_w = 0) {
  ptx_free::pdl_prologue();
  const int vecs = C / 8;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c0 = static_cast<int>(i0 % vecs) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = stats[2 * C + c0 + k];
    sh[k] = stats[3 * C + c0 + k];
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = i0; i < total_vec; i += stride) {
    float v[8];
    V8<T>::load(x + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], sc[k], sh[k]);
    if (residual) {
      float r[8];
      long long ri = i;
      if (up_w > 0) {
        long long pix = i / vecs;
        const int w = static_cast<int>(pix % up_w);
        pix /= up_w;
        const int h = static_cast<int>(pix % up_h);
        const long long n = pix / up_h;
        ri = ((n * (up_h / 2) + h / 2) * (up_w / 2) + w / 2) * vecs + (i % vecs);
      }
      V8<T>::load(residual + ri * 8, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += r[k];
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    V8<T>::store(y + i * 8, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                    const float* __restrict__ coeff, T* __restrict__ dx, T* __restrict__ dres,
                    long long total_vec, int C, const float* __restrict__ relu_scale_shift = nullptr) {
  ptx_free::pdl_prologue();
  const int vecs = C / 8;
  const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c0 = static_cast<int>(i0 % vecs) * 8;
  float A[8], B[8], K[8], rsc[8], rsh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    A[k] = coeff[c0 + k];
    B[k] = coeff[C + c0 + k];
    K[k] = coeff[2 * C + c0 + k];
    rsc[k] = relu_scale_shift ? relu_scale_shift[c0 + k] : 0.f;
    rsh[k] = relu_scale_shift ? relu_scale_shift[C + c0 + k] : 0.f;
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = i0; i < total_vec; i += stride) {
    float g[8], vx[8];
    V8<T>::load(dy + i * 8, g);
    V8<T>::load(x + i * 8, vx);
    if (y) {
      float vy[8];
      V8<T>::load(y + i * 8, vy);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = vy[k] > 0.f ? g[k] : 0.f;
    } else if (relu_scale_shift) {
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = as_stored<T>(fmaf(vx[k], rsc[k], rsh[k])) > 0.f ? g[k] : 0.f;
    }
    if (dres) V8<T>::store(dres + i * 8, g);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], g[k], fmaf(B[k], vx[k], K[k]));
    V8<T>::store(dx + i * 8, o);
  }
}

// ---------------- cross-GPU exchange over NVLink peer memory (SyncBN without NCCL launches) ----------------
// Symmetric buffer of every rank (same layout, allocated once through torch symmetric memory):
//   float data[2][world][slot_floats];  uint32 flags[world]   (flags at byte offset 2*world*slot_floats*4)
// Call `epoch` (1, 2, 3, ... identical on all ranks, one per exchange) uses data slot epoch & 1. Rank r stores its
// sums into data[slot][r] of EVERY rank (remote stores over NVLink), fences, then publishes `epoch` in flags[r] of
// every rank; it then waits until its own flags[q] >= epoch for all q and reduces data[slot][0..world) locally in a
// fixed order (all ranks get bit-identical sums). A rank can be at most one call ahead of a peer (it needs the
// peer's flag of the previous call to proceed), so two slots are enough.
#ifndef U2B_XCHG_TIMEOUT_CYCLES
#define U2B_XCHG_TIMEOUT_CYCLES (240000000000LL)  // ~2 min: with 8 processes loading kernels from a cold file system, ranks were seen
                                                     // to reach an exchange more than 20 s apart (lazy module loading); a dead peer still traps
#endif

__device__ __forceinline__ void xchg_all_reduce(float* __restrict__ vals /*smem [n]*/, int n,
                                                const unsigned long long* __restrict__ peers, int world, int rank,
                                                unsigned int epoch, int slot_floats) {
  const int slot = epoch & 1;
  const size_t flag_off = static_cast<size_t>(2) * world * slot_floats * sizeof(float);
  for (int p = 0; p < world; ++p) {
    float* dst = reinterpret_cast<float*>(peers[p]) + (static_cast<size_t>(slot) * world + rank) * slot_floats;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
  }
  __syncthreads();  // the release store below is cumulative over the block's writes observed through this barrier
  if (threadIdx.x < world) {
    unsigned int* f = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(peers[threadIdx.x]) + flag_off) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
  if (threadIdx.x < world) {
    const unsigned int* f =
        reinterpret_cast<const unsigned int*>(reinterpret_cast<const char*>(peers[rank]) + flag_off) + threadIdx.x;
    const long long t0 = clock64();
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (static_cast<int>(v - epoch) >= 0) break;
      if (clock64() - t0 > U2B_XCHG_TIMEOUT_CYCLES) {
        printf("u2b: SyncBN peer exchange timeout rank %d waiting for rank %d epoch %u (have %u)\n", rank,
               threadIdx.x, epoch, v);
        __trap();
      }
    } while (true);
  }
  __syncthreads();
  const float* mine = reinterpret_cast<const float*>(peers[rank]) + static_cast<size_t>(slot) * world * slot_floats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int q = 0; q < world; ++q) acc += __ldcv(mine + static_cast<size_t>(q) * slot_floats + i);
    vals[i] = acc;
  }
  __syncthreads();
}

// one CTA: local (2C) sums -> exchange -> mean/invstd/scale/shift + running statistics
__global__ void __launch_bounds__(512)
bn_xchg_finalize_kernel(const float* __restrict__ sums, const unsigned long long* __restrict__ peers, int world,
                        int rank, unsigned int* __restrict__ epoch_ctr, int slot_floats, double n_total,
                        const float* __restrict__ w,
                        const float* __restrict__ b, float eps, float momentum, float* __restrict__ running_mean,
                        float* __restrict__ running_var, float* __restrict__ stats, int C) {
  extern __shared__ float xv[];  // [2C]
  // the exchange number lives in device memory (advanced here), so a CUDA-graph replay uses fresh epochs
  const unsigned int epoch = *epoch_ctr + 1u;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) xv[i] = sums[i];
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ctr = epoch;
  xchg_all_reduce(xv, 2 * C, peers, world, rank, epoch, slot_floats);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mu = static_cast<double>(xv[c]) / n_total;
    double var = static_cast<double>(xv[C + c]) / n_total - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = (w ? w[c] : 1.f) * is;
    stats[c] = static_cast<float>(mu);
    stats[C + c] = is;
    stats[2 * C + c] = sc;
    stats[3 * C + c] = (b ? b[c] : 0.f) - static_cast<float>(mu) * sc;
    if (running_mean) {
      const double unbiased = n_total > 1.0 ? var * n_total / (n_total - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
  }
}

// one CTA: local (2C) backward sums -> dgamma/dbeta (local) -> exchange -> dx coefficients
__global__ void __launch_bounds__(512)
bn_xchg_bwd_coeff_kernel(const float* __restrict__ sums, const unsigned long long* __restrict__ peers, int world,
                         int rank, unsigned int* __restrict__ epoch_ctr, int slot_floats, double n_total,
                         const float* __restrict__ stats, const float* __restrict__ w, float* __restrict__ coeff,
                         float* __restrict__ gw_gb, int C) {
  extern __shared__ float xv[];  // [2C]
  const unsigned int epoch = *epoch_ctr + 1u;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) xv[i] = sums[i];
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ctr = epoch;
  if (gw_gb)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      gw_gb[c] = xv[C + c];
      gw_gb[C + c] = xv[c];
    }
  xchg_all_reduce(xv, 2 * C, peers, world, rank, epoch, slot_floats);
  const float inv_n = static_cast<float>(1.0 / n_total);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mu = stats[c], is = stats[C + c];
    const float A = (w ? w[c] : 1.f) * is;
    const float B = -A * is * xv[C + c] * inv_n;
    coeff[c] = A;
    coeff[C + c] = B;
    coeff[2 * C + c] = -A * xv[c] * inv_n - B * mu;
  }
}

// ---- multi-CTA exchange: partial rows -> sums -> NVLink exchange -> statistics, ONE kernel per BN direction ----------
// One CTA per 32 channels (<= 64 CTAs). Each CTA sums its 64 columns of the (S, 2C) partial rows (8 warps stride the rows,
// 128-byte coalesced reads, fixed combination order), exchanges those 64 floats with every peer through its own flag
// (flag[src rank][cta], epoch sequence per cta index: every rank runs the same layer sequence, so the per-cta sequences
// agree across ranks), and finishes its 32 channels. Replaces bn_sum_partials_kernel + the single-CTA exchange kernel.
constexpr int X2_CH = 32;
constexpr int X2_MAXCTAS = 64;
constexpr int X2_WARPS = 32;   // 1024 threads per CTA
constexpr size_t X2_FLAG_SKIP = 256;   // the single-CTA kernels' flags (world <= 32 words) live in front of ours

__device__ __forceinline__ void xchg2_reduce(const float* __restrict__ partials, int S, int C,
                                             const unsigned long long* __restrict__ peers, int world, int rank,
                                             unsigned int* __restrict__ epoch_ctrs, int slot_floats,
                                             float* __restrict__ local /*smem[64]*/, float* __restrict__ tot /*smem[64]*/) {
  // X2_WARPS warps stride the partial rows (up to 1024 of them after a res2 convolution): the row loop is a chain of
  // L2 latencies, so its length - not bandwidth - sets the kernel's duration
  __shared__ float red[X2_WARPS][2 * X2_CH];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, t = threadIdx.x;
  const int c0 = blockIdx.x * X2_CH;
  const bool live = c0 + l < C;
  float s1 = 0.f, s2 = 0.f;
  if (live) {
#pragma unroll 4
    for (int row = w; row < S; row += X2_WARPS) {
      s1 += partials[static_cast<size_t>(row) * 2 * C + c0 + l];
      s2 += partials[static_cast<size_t>(row) * 2 * C + C + c0 + l];
    }
  }
  red[w][l] = s1;
  red[w][X2_CH + l] = s2;
  const unsigned int epoch = epoch_ctrs[blockIdx.x] + 1u;
  __syncthreads();
  if (t == 0) epoch_ctrs[blockIdx.x] = epoch;
  const int slot = epoch & 1;
  const bool col_live = t < 2 * X2_CH && (c0 + (t & (X2_CH - 1))) < C;
  // A word's location is (cta, t), not the channel column: every location is then written by ONE cta index only, whose epoch
  // sequence it follows (a column-indexed layout lets a stale word of another layer's cta carry the awaited epoch).
  const int cell = static_cast<int>(blockIdx.x) * 2 * X2_CH + t;
  // Low-latency exchange (the "LL" idea of NCCL's small-message protocol): every value travels as ONE 8-byte store
  // {float bits, epoch}; 8-byte scalar stores are single-copy atomic, so the reader polls the value words themselves and
  // needs neither a separate flag nor a release fence (which would cost a second NVLink round trip per exchange).
  const size_t ll_off = static_cast<size_t>(2) * world * slot_floats * sizeof(float) + X2_FLAG_SKIP +
                        static_cast<size_t>(world) * X2_MAXCTAS * 4 + 64;
  if (t < 2 * X2_CH) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < X2_WARPS; ++q) v += red[q][t];
    local[t] = v;
    if (col_live) {
      const unsigned long long word = (static_cast<unsigned long long>(epoch) << 32) | __float_as_uint(v);
      const size_t idx = (static_cast<size_t>(slot) * world + rank) * slot_floats + cell;
      for (int p = 0; p < world; ++p) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peers[p]) + ll_off) + idx;
        asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
      }
    }
    float acc = 0.f;
    if (col_live) {
      const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(peers[rank]) + ll_off) +
                                       static_cast<size_t>(slot) * world * slot_floats + cell;
      const long long t0 = clock64();
      for (int q = 0; q < world; ++q) {        // fixed order: every rank adds the same values in the same order
        unsigned long long w64;
        do {
          asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w64) : "l"(mine + static_cast<size_t>(q) * slot_floats) : "memory");
          if (static_cast<unsigned int>(w64 >> 32) == epoch) break;
          if (clock64() - t0 > U2B_XCHG_TIMEOUT_CYCLES) {
            printf("u2b: SyncBN peer exchange timeout rank %d cta %d cell %d waiting for rank %d epoch %u (have %u)\n", rank,
                   static_cast<int>(blockIdx.x), cell, q, epoch, static_cast<unsigned int>(w64 >> 32));
            __trap();
          }
        } while (true);
        acc += __uint_as_float(static_cast<unsigned int>(w64));
      }
    }
    tot[t] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(X2_WARPS * 32)
bn_xchg2_finalize_kernel(const float* __restrict__ partials, int S, const unsigned long long* __restrict__ peers, int world,
                         int rank, unsigned int* __restrict__ epoch_ctrs, int slot_floats, double n_total,
                         const float* __restrict__ w, const float* __restrict__ b, float eps, float momentum,
                         float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ stats,
                         int C) {
  ptx_free::pdl_prologue();
  __shared__ float local[2 * X2_CH], tot[2 * X2_CH];
  xchg2_reduce(partials, S, C, peers, world, rank, epoch_ctrs, slot_floats, local, tot);
  const int c = blockIdx.x * X2_CH + threadIdx.x;
  if (threadIdx.x >= X2_CH || c >= C) return;
  const double mu = static_cast<double>(tot[threadIdx.x]) / n_total;
  double var = static_cast<double>(tot[X2_CH + threadIdx.x]) / n_total - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = (w ? w[c] : 1.f) * is;
  stats[c] = static_cast<float>(mu);
  stats[C + c] = is;
  stats[2 * C + c] = sc;
  stats[3 * C + c] = (b ? b[c] : 0.f) - static_cast<float>(mu) * sc;
  if (running_mean) {
    const double unbiased = n_total > 1.0 ? var * n_total / (n_total - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void __launch_bounds__(X2_WARPS * 32)
bn_xchg2_bwd_coeff_kernel(const float* __restrict__ partials, int S, const unsigned long long* __restrict__ peers, int world,
                          int rank, unsigned int* __restrict__ epoch_ctrs, int slot_floats, double n_total,
                          const float* __restrict__ stats, const float* __restrict__ w, float* __restrict__ coeff,
                          float* __restrict__ gw_gb, int C) {
  ptx_free::pdl_prologue();
  __shared__ float local[2 * X2_CH], tot[2 * X2_CH];
  xchg2_reduce(partials, S, C, peers, world, rank, epoch_ctrs, slot_floats, local, tot);
  const int c = blockIdx.x * X2_CH + threadIdx.x;
  if (threadIdx.x >= X2_CH || c >= C) return;
  if (gw_gb) {   // LOCAL sums: the gradient all-reduce averages them like every other parameter gradient
    gw_gb[c] = local[X2_CH + threadIdx.x];
    gw_gb[C + c] = local[threadIdx.x];
  }
  const float inv_n = static_cast<float>(1.0 / n_total);
  const float mu = stats[c], is = stats[C + c];
  const float A = (w ? w[c] : 1.f) * is;
  const float B = -A * is * tot[X2_CH + threadIdx.x] * inv_n;
  coeff[c] = A;
  coeff[C + c] = B;
  coeff[2 * C + c] = -A * tot[threadIdx.x] * inv_n - B * mu;
}

template <typename T, int MODE>
int launch_reduce(const void* a, const void* x, const void* y, const float* mean, const float* invstd,
                  long long P, int C, float* partials, cudaStream_t stream, const float* relu_scale_shift = nullptr) {
  dim3 grid(num_strips(P, C), channel_groups(C));
  u2b_launch_pdl(bn_reduce_kernel<T, MODE>, dim3(grid), dim3(BN_THREADS), 0, stream, static_cast<const T*>(a), static_cast<const T*>(x),
                                                             static_cast<const T*>(y), mean, invstd, P, C,
                                                             vecs_per_block(C), partials, relu_scale_shift);
  U2B_LAUNCH_CHECK();
  return 0;
}

// grid*256 must be a multiple of C/8 (a power of two <= 256 here): any grid works when 256 % (C/8) == 0
inline unsigned ew_grid(long long total_vec) {
  long long b = (total_vec + 255) / 256;
  const long long cap = static_cast<long long>(u2b_num_sms()) * 8;
  return static_cast<unsigned>(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

#define U2B_BN_DISPATCH(CALL_F, CALL_H, CALL_B)                 \
  if (dtype == 0) { CALL_F; } else if (dtype == 1) { CALL_H; } \
  else if (dtype == 2) { CALL_B; } else { u2b_set_error("batchnorm: unknown dtype %d", dtype); return U2B_ERR_BAD_ARG; }

extern "C" {

// channel counts handled: multiples of 8 whose vector count C/8 divides 256 (64, 128, 256, 512, 1024, 2048, ...)
int u2b_bn_supported(int C) { return C >= 8 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0; }

// number of partial rows (strips) the reduce kernels write for a (P, C) activation
int u2b_bn_num_strips(int64_t P, int C) { return num_strips(P, C); }

// partials[s][0:C] = sum x, partials[s][C:2C] = sum x^2 over strip s of the P pixels of x (P, C) NHWC
int u2b_bn_stats(int dtype, const void* x, int64_t P, int C, float* partials, cudaStream_t stream) {
  U2B_CHECK_ARG(x && partials && P > 0 && u2b_bn_supported(C), "bn_stats: bad arguments");
  U2B_BN_DISPATCH(return (launch_reduce<float, 0>(x, nullptr, nullptr, nullptr, nullptr, P, C, partials, stream)),
                  return (launch_reduce<__half, 0>(x, nullptr, nullptr, nullptr, nullptr, P, C, partials, stream)),
                  return (launch_reduce<__nv_bfloat16, 0>(x, nullptr, nullptr, nullptr, nullptr, P, C, partials, stream)))
}

// sums[0:C2] = sum over the S partial rows (used when the sums are all-reduced across ranks before finalize/coeff)
int u2b_bn_sum_partials(const float* partials, int S, int C2, float* sums, cudaStream_t stream) {
  U2B_CHECK_ARG(partials && sums && S > 0 && C2 > 0, "bn_sum_partials: bad arguments");
  u2b_launch_pdl(bn_sum_partials_kernel, dim3((C2 + 7) / 8), dim3(256), 0, stream, partials, S, C2, sums);
  U2B_LAUNCH_CHECK();
  return 0;
}

// From S partial rows (S = 1: already summed / all-reduced) over n_total pixels: stats[0:C] mean, [C:2C] invstd,
// [2C:3C] scale = w*invstd, [3C:4C] shift = b - mean*scale; running statistics updated (unbiased variance).
int u2b_bn_finalize(const float* partials, int S, double n_total, const float* w, const float* b, float eps,
                    float momentum, float* running_mean, float* running_var, float* stats, int C,
                    cudaStream_t stream) {
  U2B_CHECK_ARG(partials && stats && S > 0 && C > 0 && n_total > 0, "bn_finalize: bad arguments");
  u2b_launch_pdl(bn_finalize_kernel, dim3((C + 7) / 8), dim3(256), 0, stream, partials, S, n_total, w, b, eps, momentum, running_mean,
                                                          running_var, stats, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

// y = [relu](x * scale[c] + shift[c] [+ residual]); stats from u2b_bn_finalize
int u2b_bn_apply(int dtype, const void* x, const float* stats, const void* residual, int relu, void* y, int64_t P,
                 int C, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(x && y && stats && u2b_bn_supported(C), "bn_apply: bad arguments");
  const long long tv = static_cast<long long>(P) * C / 8;
  U2B_BN_DISPATCH(
      (u2b_launch_pdl(bn_apply_kernel<float>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const float*)x, stats, (const float*)residual, relu, (float*)y, tv, C, 0, 0)),
      (u2b_launch_pdl(bn_apply_kernel<__half>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __half*)x, stats, (const __half*)residual, relu, (__half*)y, tv, C, 0, 0)),
      (u2b_launch_pdl(bn_apply_kernel<__nv_bfloat16>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __nv_bfloat16*)x, stats, (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)y, tv, C, 0, 0)))
  U2B_LAUNCH_CHECK();
  return 0;
}

// y = [relu](x * scale[c] + shift[c] + up2x(residual)): x, y (N,H,W,C); residual (N,H/2,W/2,C); H, W even
int u2b_bn_apply_resup(int dtype, const void* x, const float* stats, const void* residual, int relu, void* y, int N, int H,
                       int W, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(x && y && stats && residual && u2b_bn_supported(C) && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
                "bn_apply_resup: bad arguments");
  const long long tv = static_cast<long long>(N) * H * W * C / 8;
  U2B_BN_DISPATCH(
      (u2b_launch_pdl(bn_apply_kernel<float>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const float*)x, stats, (const float*)residual, relu, (float*)y, tv, C, H, W)),
      (u2b_launch_pdl(bn_apply_kernel<__half>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __half*)x, stats, (const __half*)residual, relu, (__half*)y, tv, C, H, W)),
      (u2b_launch_pdl(bn_apply_kernel<__nv_bfloat16>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __nv_bfloat16*)x, stats, (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)y, tv, C, H, W)))
  U2B_LAUNCH_CHECK();
  return 0;
}

// partials[s] = (sum dz, sum dz*xhat) per strip, dz = dy * (y > 0) when y != NULL (fused ReLU backward)
int u2b_bn_bwd_reduce(int dtype, const void* dy, const void* x, const void* y, const float* stats, int64_t P, int C,
                      float* partials, cudaStream_t stream) {
  U2B_CHECK_ARG(dy && x && stats && partials && P > 0 && u2b_bn_supported(C), "bn_bwd_reduce: bad arguments");
  U2B_BN_DISPATCH(return (launch_reduce<float, 1>(dy, x, y, stats, stats + C, P, C, partials, stream)),
                  return (launch_reduce<__half, 1>(dy, x, y, stats, stats + C, P, C, partials, stream)),
                  return (launch_reduce<__nv_bfloat16, 1>(dy, x, y, stats, stats + C, P, C, partials, stream)))
}

// same with the ReLU mask recomputed from x: dz = dy * (stored(x * scale + shift) > 0), scale / shift = stats[2C:4C]. Valid
// when the forward was y = relu(bn(x)) WITHOUT a residual; saves the read of y (one of three tensors).
int u2b_bn_bwd_reduce_relu_x(int dtype, const void* dy, const void* x, const float* stats, int64_t P, int C,
                             float* partials, cudaStream_t stream) {
  U2B_CHECK_ARG(dy && x && stats && partials && P > 0 && u2b_bn_supported(C), "bn_bwd_reduce_relu_x: bad arguments");
  U2B_BN_DISPATCH(return (launch_reduce<float, 1>(dy, x, nullptr, stats, stats + C, P, C, partials, stream, stats + 2 * C)),
                  return (launch_reduce<__half, 1>(dy, x, nullptr, stats, stats + C, P, C, partials, stream, stats + 2 * C)),
                  return (launch_reduce<__nv_bfloat16, 1>(dy, x, nullptr, stats, stats + C, P, C, partials, stream, stats + 2 * C)))
}

// coefficients of dx = A*dz + B*x + K from S partial rows (S = 1: all-reduced sums); gw_gb (2C, nullable) receives
// dgamma | dbeta = the sums themselves (meaningful when the rows are this rank's local sums)
int u2b_bn_bwd_coeff(const float* partials, int S, double n_total, const float* stats, const float* w, float* coeff,
                     float* gw_gb, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(partials && stats && coeff && S > 0 && C > 0 && n_total > 0, "bn_bwd_coeff: bad arguments");
  u2b_launch_pdl(bn_bwd_coeff_kernel, dim3((C + 7) / 8), dim3(256), 0, stream, partials, S, n_total, stats, w, coeff, gw_gb, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

// GroupNorm glue for ONE image (see gn_finalize_kernel): partials from u2b_bn_stats / u2b_bn_bwd_reduce with P = H*W.
int u2b_gn_supported(int C, int G) {
  return u2b_bn_supported(C) && C <= 1024 && 1024 % C == 0 && G > 0 && C % G == 0;
}

int u2b_gn_finalize(const float* partials, int S, int64_t HW, int G, const float* w, const float* b, float eps,
                    float* stats, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(partials && stats && S > 0 && HW > 0 && u2b_gn_supported(C, G), "gn_finalize: bad arguments");
  gn_finalize_kernel<<<1, 1024, 0, stream>>>(partials, S, static_cast<double>(HW) * (C / G), G, w, b, eps, stats, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_gn_bwd_coeff(const float* partials, int S, int64_t HW, int G, const float* stats, const float* w,
                     float* coeff, float* gw_gb, int accumulate, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(partials && stats && coeff && S > 0 && HW > 0 && u2b_gn_supported(C, G), "gn_bwd_coeff: bad arguments");
  gn_bwd_coeff_kernel<<<1, 1024, 0, stream>>>(partials, S, static_cast<double>(HW) * (C / G), G, stats, w, coeff, gw_gb,
                                              accumulate, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

// Data-parallel variants: `sums` (2C, this rank's summed partials) are exchanged with all peers through the
// symmetric buffers `peers[world]` (device array of device pointers to every rank's buffer) inside the kernel.
// epoch_ctr: device uint32 (start 0), advanced by one per exchange inside the kernel; every rank performs the same
// sequence of exchanges, so the counters stay identical (and a CUDA-graph replay keeps working). slot_floats >= 2C.
size_t u2b_bn_xchg_buffer_bytes(int world, int slot_floats) {
  // [2 slots x world x slot_floats fp32 | flags of the single-CTA kernels | (unused) | 2 slots x world x slot_floats 8-byte
  // {value, epoch} words of the multi-CTA kernels]
  return static_cast<size_t>(2) * world * slot_floats * sizeof(float) + X2_FLAG_SKIP +
         static_cast<size_t>(world) * X2_MAXCTAS * 4 + 64 + static_cast<size_t>(2) * world * slot_floats * 8 + 64;
}

int u2b_bn_xchg2_max_ctas(void) { return X2_MAXCTAS; }

// Multi-CTA variants taking the (S, 2C) partial rows directly (no separate u2b_bn_sum_partials launch). epoch_ctrs: device
// uint32[u2b_bn_xchg2_max_ctas()], zero-initialised once, advanced inside the kernels.
int u2b_bn_xchg2_finalize(const float* partials, int S, const void* peers, int world, int rank, uint32_t* epoch_ctrs,
                          int slot_floats, double n_total, const float* w, const float* b, float eps, float momentum,
                          float* running_mean, float* running_var, float* stats, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(partials && S > 0 && peers && stats && epoch_ctrs && world > 0 && world <= 32 && rank >= 0 && rank < world &&
                    2 * C <= slot_floats && C <= X2_CH * X2_MAXCTAS && slot_floats >= 2 * X2_CH * ((C + X2_CH - 1) / X2_CH),
                "bn_xchg2_finalize: bad arguments");
  u2b_launch_pdl(bn_xchg2_finalize_kernel, dim3((C + X2_CH - 1) / X2_CH), dim3(X2_WARPS * 32), 0, stream, 
      partials, S, static_cast<const unsigned long long*>(peers), world, rank, epoch_ctrs, slot_floats, n_total, w, b, eps,
      momentum, running_mean, running_var, stats, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_bn_xchg2_bwd_coeff(const float* partials, int S, const void* peers, int world, int rank, uint32_t* epoch_ctrs,
                           int slot_floats, double n_total, const float* stats, const float* w, float* coeff, float* gw_gb,
                           int C, cudaStream_t stream) {
  U2B_CHECK_ARG(partials && S > 0 && peers && stats && coeff && epoch_ctrs && world > 0 && world <= 32 && rank >= 0 &&
                    rank < world && 2 * C <= slot_floats && C <= X2_CH * X2_MAXCTAS && slot_floats >= 2 * X2_CH * ((C + X2_CH - 1) / X2_CH),
                "bn_xchg2_bwd_coeff: bad arguments");
  u2b_launch_pdl(bn_xchg2_bwd_coeff_kernel, dim3((C + X2_CH - 1) / X2_CH), dim3(X2_WARPS * 32), 0, stream, 
      partials, S, static_cast<const unsigned long long*>(peers), world, rank, epoch_ctrs, slot_floats, n_total, stats, w,
      coeff, gw_gb, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_bn_xchg_finalize(const float* sums, const void* peers, int world, int rank, uint32_t* epoch_ctr, int slot_floats,
                         double n_total, const float* w, const float* b, float eps, float momentum,
                         float* running_mean, float* running_var, float* stats, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(sums && peers && stats && epoch_ctr && world > 0 && world <= 32 && rank >= 0 && rank < world &&
                    2 * C <= slot_floats,
                "bn_xchg_finalize: bad arguments");
  bn_xchg_finalize_kernel<<<1, 512, static_cast<size_t>(2) * C * sizeof(float), stream>>>(
      sums, static_cast<const unsigned long long*>(peers), world, rank, epoch_ctr, slot_floats, n_total, w, b, eps,
      momentum, running_mean, running_var, stats, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_bn_xchg_bwd_coeff(const float* sums, const void* peers, int world, int rank, uint32_t* epoch_ctr, int slot_floats,
                          double n_total, const float* stats, const float* w, float* coeff, float* gw_gb, int C,
                          cudaStream_t stream) {
  U2B_CHECK_ARG(sums && peers && stats && coeff && epoch_ctr && world > 0 && world <= 32 && rank >= 0 && rank < world &&
                    2 * C <= slot_floats,
                "bn_xchg_bwd_coeff: bad arguments");
  bn_xchg_bwd_coeff_kernel<<<1, 512, static_cast<size_t>(2) * C * sizeof(float), stream>>>(
      sums, static_cast<const unsigned long long*>(peers), world, rank, epoch_ctr, slot_floats, n_total, stats, w, coeff,
      gw_gb, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

// dx = A*dz + B*x + K; dres = dz when dres != NULL
int u2b_bn_bwd_apply(int dtype, const void* dy, const void* x, const void* y, const float* coeff, void* dx,
                     void* dres, int64_t P, int C, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(dy && x && dx && coeff && u2b_bn_supported(C), "bn_bwd_apply: bad arguments");
  const long long tv = static_cast<long long>(P) * C / 8;
  U2B_BN_DISPATCH(
      (u2b_launch_pdl(bn_bwd_apply_kernel<float>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const float*)dy, (const float*)x, (const float*)y, coeff, (float*)dx, (float*)dres, tv, C, nullptr)),
      (u2b_launch_pdl(bn_bwd_apply_kernel<__half>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __half*)dy, (const __half*)x, (const __half*)y, coeff, (__half*)dx, (__half*)dres, tv, C, nullptr)),
      (u2b_launch_pdl(bn_bwd_apply_kernel<__nv_bfloat16>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, coeff, (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, tv, C, nullptr)))
  U2B_LAUNCH_CHECK();
  return 0;
}

// dx = A*dz + B*x + K with dz = dy * (stored(x * scale + shift) > 0): the companion of u2b_bn_bwd_reduce_relu_x
int u2b_bn_bwd_apply_relu_x(int dtype, const void* dy, const void* x, const float* stats, const float* coeff, void* dx,
                            int64_t P, int C, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(dy && x && dx && coeff && stats && u2b_bn_supported(C), "bn_bwd_apply_relu_x: bad arguments");
  const long long tv = static_cast<long long>(P) * C / 8;
  const float* rs = stats + 2 * C;
  U2B_BN_DISPATCH(
      (u2b_launch_pdl(bn_bwd_apply_kernel<float>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const float*)dy, (const float*)x, nullptr, coeff, (float*)dx, nullptr, tv, C, rs)),
      (u2b_launch_pdl(bn_bwd_apply_kernel<__half>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __half*)dy, (const __half*)x, nullptr, coeff, (__half*)dx, nullptr, tv, C, rs)),
      (u2b_launch_pdl(bn_bwd_apply_kernel<__nv_bfloat16>, dim3(ew_grid(tv)), dim3(256), 0, stream, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, nullptr, coeff, (__nv_bfloat16*)dx, nullptr, tv, C, rs)))
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
