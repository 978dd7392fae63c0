// Fused optimizer step of the detector path over FLAT buffers: per-parameter gradient-norm clipping
// (detectron2/solver/build.py:63-73, CLIP_TYPE "norm": clip_grad_norm_ applied to each parameter on its own),
// SGD with momentum / weight decay / optional Nesterov (solver/build.py:119-139 -> torch.optim.SGD) and the refresh
// of the bf16 compute copy of the weights, in ONE pass: the reference's foreach implementation walks the 76 M
// parameters eight times (clip multiply per tensor, weight-decay add, momentum multiply, momentum add, lr multiply,
// parameter subtract, bf16 cast ...), ~250 + 40 launches per step; here every value is read once and written once
// (grad r, master r/w, momentum r/w, bf16 w = 22 B per parameter, HBM-bound).
//
// Layout: all parameters live in one flat fp32 master buffer, each starting on a 64-element boundary; gradients and
// momentum use the same offsets. seg_of_chunk[i] names the parameter that owns elements [64 i, 64 i + 64)
// (-1: padding); seg_wd / seg_coef hold the per-parameter weight decay and clip coefficient.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sgd_segments_kernel(const float4* __restrict__ grad, float4* __restrict__ master, float4* __restrict__ mom,
                    __nv_bfloat16* __restrict__ w16, long long w16_begin, const int32_t* __restrict__ seg_of_chunk,
                    const float* __restrict__ seg_wd, const float* __restrict__ seg_coef,
                    const float* __restrict__ lr_ptr, float momentum, int nesterov, long long n_vec) {
  const float lr = *lr_ptr;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const int seg = seg_of_chunk[i >> 4];           // 16 float4 per 64-element chunk
    if (seg < 0) continue;
    const float wd = seg_wd[seg];
    const float coef = seg_coef ? seg_coef[seg] : 1.f;
    const float4 g4 = grad[i];
    float4 p4 = master[i], m4 = mom[i];
    float g[4] = {g4.x, g4.y, g4.z, g4.w}, p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float d = g[k] * coef;                          // clip_grad_norm_: grad *= min(1, max_norm / (norm + 1e-6))
      d = fmaf(wd, p[k], d);                          // weight decay: d_p = grad + wd * p
      m[k] = fmaf(momentum, m[k], d);                 // buf = momentum * buf + d_p   (dampening 0)
      const float upd = nesterov ? fmaf(momentum, m[k], d) : m[k];
      p[k] = p[k] - lr * upd;
    }
    master[i] = make_float4(p[0], p[1], p[2], p[3]);
    mom[i] = make_float4(m[0], m[1], m[2], m[3]);
    const long long e = i * 4;
    if (w16 && e >= w16_begin) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p[0], p[1]), hi = __floats2bfloat162_rn(p[2], p[3]);
      uint2 v;
      v.x = *reinterpret_cast<uint32_t*>(&lo);
      v.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(w16 + (e - w16_begin)) = v;
    }
  }
}

}  // namespace

extern "C" {

int u2b_sgd_step_segments(const float* grad, float* master, float* mom, void* w16, int64_t w16_begin,
                          const int32_t* seg_of_chunk, const float* seg_wd, const float* seg_coef, const float* lr,
                          float momentum, int nesterov, int64_t n, cudaStream_t stream) {
  if (n == 0) return 0;
  U2B_CHECK_ARG(grad && master && mom && seg_of_chunk && seg_wd && lr && n > 0, "sgd_step_segments: null pointer");
  U2B_CHECK_ARG(n % 64 == 0 && w16_begin % 64 == 0 && w16_begin >= 0, "sgd_step_segments: n and w16_begin must be multiples of 64");
  U2B_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(master) |
                  reinterpret_cast<uintptr_t>(mom)) & 15) == 0 && (reinterpret_cast<uintptr_t>(w16) & 7) == 0,
                "sgd_step_segments: buffers must be 16-byte aligned");
  const long long n_vec = n / 4;
  const int grid = u2b_num_sms() * 8;
  sgd_segments_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(master),
                                                reinterpret_cast<float4*>(mom), static_cast<__nv_bfloat16*>(w16),
                                                w16_begin, seg_of_chunk, seg_wd, seg_coef, lr, momentum, nesterov, n_vec);
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
