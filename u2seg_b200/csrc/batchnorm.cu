// Training-mode (Sync)BatchNorm for NHWC activations, fused with the residual add and ReLU that follow it in
// the ResNet/FPN blocks. Reference: detectron2/layers/batch_norm.py:187 (nn.SyncBatchNorm) as used by
// layers/wrappers.py:87-134 (Conv2d.forward: conv -> norm -> activation) and backbone/resnet.py:194-210
// (out += shortcut; relu).
//
// HBM-bound. Forward = 2 reads + 1 write of the activation (library path: 2 reads + 1 write for BN, then 1 read +
// 1 write each for the add and the ReLU); backward = 3 reads + 1-2 writes.
//   stats      : per-channel sum / sum of squares (fp32), register accumulation over a pixel strip, block reduction
//                in shared memory, one fp32 atomicAdd per channel per block
//   finalize   : mean / invstd / scale / shift, running statistics (momentum, unbiased variance); re-zeroes scratch
//   apply      : y = relu(x * scale[c] + shift[c] + residual)
//   bwd_reduce : sum(dz), sum(dz * xhat) with dz = dy * (y > 0)
//   bwd_apply  : dx = scale * (dz - sum_dy/n - xhat * sum_dy_xhat/n); dres = dz
// With data parallelism the (2C) scratch sums are all-reduced between stats/bwd_reduce and finalize/bwd_apply.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

template <typename T>
struct V8;  // 8 channels per thread
template <>
struct V8<float> {
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

constexpr int BN_THREADS = 256;

// Two per-channel sums over pixels. MODE 0: (x, x^2). MODE 1: (dz, dz*xhat), dz = dy * (y>0 if y given).
// Thread t owns channel vector (t % vecs) and pixels (t / vecs) + k * (BN_THREADS / vecs) of the block's strip.
template <typename T, int MODE>
__global__ void __launch_bounds__(BN_THREADS)
bn_reduce_kernel(const T* __restrict__ a, const T* __restrict__ x, const T* __restrict__ y,
                 const float* __restrict__ mean, const float* __restrict__ invstd, long long P, int C,
                 float* __restrict__ sums) {
  extern __shared__ float red[];  // [BN_THREADS][16]
  const int vecs = C / 8;
  float s0[8], s1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
  const long long strip = (P + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * strip;
  const long long p1 = (p0 + strip < P) ? p0 + strip : P;
  for (int v0 = 0; v0 < vecs; v0 += BN_THREADS) {  // C > 2048: several passes over the channel vectors
    const int pv = (vecs - v0 < BN_THREADS) ? (vecs - v0) : BN_THREADS;  // vectors handled in this pass
    const int ppi = BN_THREADS / pv;                                     // pixels per iteration
    const int tv = threadIdx.x % pv, tp = threadIdx.x / pv;
    const int c0 = (v0 + tv) * 8;
    float m[8], is[8];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        m[i] = mean[c0 + i];
        is[i] = invstd[c0 + i];
      }
    }
    if (tp < ppi) {
      for (long long p = p0 + tp; p < p1; p += ppi) {
        float va[8];
        V8<T>::load(a + p * C + c0, va);
        if (MODE == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            s0[i] += va[i];
            s1[i] = fmaf(va[i], va[i], s1[i]);
          }
        } else {
          float vx[8];
          V8<T>::load(x + p * C + c0, vx);
          if (y) {
            float vy[8];
            V8<T>::load(y + p * C + c0, vy);
#pragma unroll
            for (int i = 0; i < 8; ++i) va[i] = vy[i] > 0.f ? va[i] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            s0[i] += va[i];
            s1[i] = fmaf(va[i], (vx[i] - m[i]) * is[i], s1[i]);
          }
        }
      }
    }
    // block reduction over the ppi pixel lanes of each channel vector
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[threadIdx.x * 16 + i] = s0[i];
      red[threadIdx.x * 16 + 8 + i] = s1[i];
      s0[i] = s1[i] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x < pv) {
      float t0[8], t1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t0[i] = t1[i] = 0.f;
      for (int q = 0; q < ppi; ++q) {
        const float* r = red + (q * pv + threadIdx.x) * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          t0[i] += r[i];
          t1[i] += r[8 + i];
        }
      }
      const int cc = (v0 + threadIdx.x) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(sums + cc + i, t0[i]);
        atomicAdd(sums + C + cc + i, t1[i]);
      }
    }
    __syncthreads();
  }
}

__global__ void bn_finalize_kernel(float* __restrict__ sums, double n_total, const float* __restrict__ w,
                                   const float* __restrict__ b, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s = sums[c], ss = sums[C + c];
  const double mu = s / n_total;
  double var = ss / n_total - mu * mu;
  if (var < 0.0) var = 0.0;
  const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  mean[c] = static_cast<float>(mu);
  invstd[c] = is;
  const float sc = (w ? w[c] : 1.f) * is;
  scale[c] = sc;
  shift[c] = (b ? b[c] : 0.f) - static_cast<float>(mu) * sc;
  if (running_mean) {
    const double unbiased = n_total > 1.0 ? var * n_total / (n_total - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
  sums[c] = 0.f;  // scratch is handed back zeroed
  sums[C + c] = 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                const T* __restrict__ residual, int relu, T* __restrict__ y, long long total_vec, int C) {
  const int vecs = C / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(i % vecs) * 8;
    float v[8];
    V8<T>::load(x + i * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], scale[c0 + k], shift[c0 + k]);
    if (residual) {
      float r[8];
      V8<T>::load(residual + i * 8, r);
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
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ w, const float* __restrict__ sums, float inv_n,
                    T* __restrict__ dx, T* __restrict__ dres, long long total_vec, int C) {
  const int vecs = C / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total_vec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(i % vecs) * 8;
    float g[8], vx[8];
    V8<T>::load(dy + i * 8, g);
    V8<T>::load(x + i * 8, vx);
    if (y) {
      float vy[8];
      V8<T>::load(y + i * 8, vy);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = vy[k] > 0.f ? g[k] : 0.f;
    }
    if (dres) V8<T>::store(dres + i * 8, g);
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      const float is = invstd[c];
      const float xh = (vx[k] - mean[c]) * is;
      o[k] = (w ? w[c] : 1.f) * is * (g[k] - sums[c] * inv_n - xh * sums[C + c] * inv_n);
    }
    V8<T>::store(dx + i * 8, o);
  }
}

template <typename T, int MODE>
int launch_reduce(const void* a, const void* x, const void* y, const float* mean, const float* invstd,
                  long long P, int C, float* sums, cudaStream_t stream) {
  long long blocks = (P + 63) / 64;
  const long long cap = static_cast<long long>(u2b_num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  bn_reduce_kernel<T, MODE><<<static_cast<unsigned>(blocks), BN_THREADS, BN_THREADS * 16 * sizeof(float), stream>>>(
      static_cast<const T*>(a), static_cast<const T*>(x), static_cast<const T*>(y), mean, invstd, P, C, sums);
  U2B_LAUNCH_CHECK();
  return 0;
}

inline unsigned ew_grid(long long total_vec) {
  long long b = (total_vec + 255) / 256;
  const long long cap = static_cast<long long>(u2b_num_sms()) * 16;
  return static_cast<unsigned>(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

#define U2B_BN_DISPATCH(CALL_F, CALL_H, CALL_B)                 \
  if (dtype == 0) { CALL_F; } else if (dtype == 1) { CALL_H; } \
  else if (dtype == 2) { CALL_B; } else { u2b_set_error("batchnorm: unknown dtype %d", dtype); return U2B_ERR_BAD_ARG; }

extern "C" {

// sums (2C fp32) += per-channel (sum x, sum x^2) over the P pixels of x (P, C) NHWC. C % 8 == 0.
int u2b_bn_stats(int dtype, const void* x, int64_t P, int C, float* sums, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(x && sums && C > 0 && C % 8 == 0, "bn_stats: bad arguments (C %% 8 == 0 required)");
  U2B_BN_DISPATCH(return (launch_reduce<float, 0>(x, nullptr, nullptr, nullptr, nullptr, P, C, sums, stream)),
                  return (launch_reduce<__half, 0>(x, nullptr, nullptr, nullptr, nullptr, P, C, sums, stream)),
                  return (launch_reduce<__nv_bfloat16, 0>(x, nullptr, nullptr, nullptr, nullptr, P, C, sums, stream)))
}

// From the (all-reduced) sums over n_total pixels: mean, invstd, scale = w*invstd, shift = b - mean*scale; running
// statistics updated with `momentum` (unbiased variance) when given; sums is zeroed for the next use.
int u2b_bn_finalize(float* sums, double n_total, const float* w, const float* b, float eps, float momentum,
                    float* running_mean, float* running_var, float* mean, float* invstd, float* scale,
                    float* shift, int C, cudaStream_t stream) {
  U2B_CHECK_ARG(sums && mean && invstd && scale && shift && C > 0 && n_total > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, n_total, w, b, eps, momentum, running_mean,
                                                          running_var, mean, invstd, scale, shift, C);
  U2B_LAUNCH_CHECK();
  return 0;
}

// y = [relu](x * scale[c] + shift[c] [+ residual])
int u2b_bn_apply(int dtype, const void* x, const float* scale, const float* shift, const void* residual, int relu,
                 void* y, int64_t P, int C, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(x && y && scale && shift && C % 8 == 0, "bn_apply: bad arguments");
  const long long tv = static_cast<long long>(P) * C / 8;
  U2B_BN_DISPATCH(
      (bn_apply_kernel<float><<<ew_grid(tv), 256, 0, stream>>>((const float*)x, scale, shift, (const float*)residual, relu, (float*)y, tv, C)),
      (bn_apply_kernel<__half><<<ew_grid(tv), 256, 0, stream>>>((const __half*)x, scale, shift, (const __half*)residual, relu, (__half*)y, tv, C)),
      (bn_apply_kernel<__nv_bfloat16><<<ew_grid(tv), 256, 0, stream>>>((const __nv_bfloat16*)x, scale, shift, (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)y, tv, C)))
  U2B_LAUNCH_CHECK();
  return 0;
}

// sums (2C) += (sum dz, sum dz*xhat), dz = dy * (y > 0) when y != NULL (fused ReLU backward)
int u2b_bn_bwd_reduce(int dtype, const void* dy, const void* x, const void* y, const float* mean,
                      const float* invstd, int64_t P, int C, float* sums, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(dy && x && mean && invstd && sums && C % 8 == 0, "bn_bwd_reduce: bad arguments");
  U2B_BN_DISPATCH(return (launch_reduce<float, 1>(dy, x, y, mean, invstd, P, C, sums, stream)),
                  return (launch_reduce<__half, 1>(dy, x, y, mean, invstd, P, C, sums, stream)),
                  return (launch_reduce<__nv_bfloat16, 1>(dy, x, y, mean, invstd, P, C, sums, stream)))
}

// dx = w*invstd*(dz - sums[c]/n - xhat*sums[C+c]/n); dres = dz when dres != NULL. sums = all-reduced bwd sums.
int u2b_bn_bwd_apply(int dtype, const void* dy, const void* x, const void* y, const float* mean,
                     const float* invstd, const float* w, const float* sums, double n_total, void* dx,
                     void* dres, int64_t P, int C, cudaStream_t stream) {
  if (P == 0) return 0;
  U2B_CHECK_ARG(dy && x && dx && mean && invstd && sums && C % 8 == 0 && n_total > 0, "bn_bwd_apply: bad arguments");
  const long long tv = static_cast<long long>(P) * C / 8;
  const float inv_n = static_cast<float>(1.0 / n_total);
  U2B_BN_DISPATCH(
      (bn_bwd_apply_kernel<float><<<ew_grid(tv), 256, 0, stream>>>((const float*)dy, (const float*)x, (const float*)y, mean, invstd, w, sums, inv_n, (float*)dx, (float*)dres, tv, C)),
      (bn_bwd_apply_kernel<__half><<<ew_grid(tv), 256, 0, stream>>>((const __half*)dy, (const __half*)x, (const __half*)y, mean, invstd, w, sums, inv_n, (__half*)dx, (__half*)dres, tv, C)),
      (bn_bwd_apply_kernel<__nv_bfloat16><<<ew_grid(tv), 256, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, mean, invstd, w, sums, inv_n, (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, tv, C)))
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
