// Semantic-segmentation loss of the Panoptic-FPN step, fused.
//   detectron2/modeling/meta_arch/semantic_seg.py:255-267 (SemSegFPNHead.losses):
//       predictions = F.interpolate(predictions.float(), scale_factor=common_stride, mode="bilinear",
//                                   align_corners=False)
//       loss = F.cross_entropy(predictions, targets, reduction="mean", ignore_index=ignore_value)
// The reference materialises the full-resolution logits (N x C x H x W fp32 = 235 MB at 2 x 28 x 1024^2), their
// log-softmax and, in backward, both gradients: ~1.9 GB of HBM traffic for a result that depends on 7.3 MB of
// stride-4 logits and 16.8 MB of labels. Here one kernel reads the low-resolution logits and the labels once and
// produces (a) per-CTA partial sums of the loss and of the number of non-ignored pixels and (b) the gradient with
// respect to the LOW-resolution logits up to the scalar factor grad_out / count (applied by the caller), so nothing
// of full resolution is ever stored. Algorithmic bytes: logits + labels + gradient = N*h*w*C*(b+4) + N*H*W*8.
//
// One CTA = one 32 x 32 tile of full-resolution pixels (1024 threads, one pixel each). The (32/s + 2)^2 window of
// low-resolution logits that the tile's bilinear samples touch is staged in shared memory; every thread
// interpolates its C logits (PyTorch's upsample_bilinear2d formula: src = max(0, (dst + 0.5) / s - 0.5)), takes the
// log-sum-exp, and leaves d(loss_px)/d(logit_c) = softmax_c - [c == target] in a shared-memory tile g[pixel][c].
// The transposed interpolation (gradient of the upsampling) is then a GATHER over that tile: each thread owns
// (window pixel, class) outputs and sums its <= (2s)^2 contributing full-resolution pixels with separable weights
// from two small tables; window outputs are flushed with one coalesced atomicAdd per element (only the one-pixel
// halo shared with neighbouring tiles actually collides).
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

constexpr int TILE = 32;

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// source index / weight of PyTorch's bilinear upsampling for one output coordinate
__device__ __forceinline__ void src_coord(int dst, float rscale, int in_size, int& i0, int& i1, float& l1) {
  float s = rscale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - static_cast<float>(i0);
}

template <typename T>
__global__ void __launch_bounds__(TILE* TILE)
upsample_ce_kernel(const T* __restrict__ z, const int64_t* __restrict__ tgt, int h, int w, int C, int s,
                   int H, int W, int64_t ignore, float* __restrict__ dz, float* __restrict__ partials) {
  extern __shared__ float smem[];
  const int WR = TILE / s + 2;       // window side (low-resolution pixels)
  const int CP = C | 1;              // odd row pitch: conflict-free for both access patterns
  float* zwin = smem;                            // [WR*WR][C]
  float* g = zwin + WR * WR * C;                 // [TILE*TILE][CP]
  float* wyt = g + TILE * TILE * CP;             // [WR][TILE]
  float* wxt = wyt + WR * TILE;                  // [WR][TILE]
  __shared__ float red[2][32];
  const int tid = threadIdx.x;
  const int tx = tid & (TILE - 1), ty = tid / TILE;
  const int n = blockIdx.z;
  const int Y0 = blockIdx.y * TILE, X0 = blockIdx.x * TILE;
  const int wy0 = Y0 / s - 1, wx0 = X0 / s - 1;
  const float rscale = 1.0f / static_cast<float>(s);

  const T* zn = z + static_cast<size_t>(n) * h * w * C;
  for (int i = tid; i < WR * WR * C; i += TILE * TILE) {
    const int c = i % C, q = (i / C) % WR, r = i / (C * WR);
    const int yy = min(max(wy0 + r, 0), h - 1), xx = min(max(wx0 + q, 0), w - 1);
    zwin[i] = to_f32<T>(zn[(static_cast<size_t>(yy) * w + xx) * C + c]);
  }
  for (int i = tid; i < 2 * WR * TILE; i += TILE * TILE) {   // separable weights of the transposed interpolation
    const bool isx = i >= WR * TILE;
    const int j = isx ? i - WR * TILE : i;
    const int r = j / TILE, t = j % TILE;
    const int dst = (isx ? X0 : Y0) + t, lim = isx ? W : H, in = isx ? w : h, k = (isx ? wx0 : wy0) + r;
    float wgt = 0.f;
    if (dst < lim) {
      int i0, i1;
      float l1;
      src_coord(dst, rscale, in, i0, i1, l1);
      if (i0 == k) wgt += 1.f - l1;
      if (i1 == k) wgt += l1;
    }
    (isx ? wxt : wyt)[j] = wgt;
  }
  __syncthreads();

  const int Y = Y0 + ty, X = X0 + tx;
  const bool inside = Y < H && X < W;
  float loss = 0.f, cnt = 0.f;
  float* gp = g + tid * CP;
  if (inside) {
    int y0, y1, x0, x1;
    float ly, lx;
    src_coord(Y, rscale, h, y0, y1, ly);
    src_coord(X, rscale, w, x0, x1, lx);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* z00 = zwin + ((y0 - wy0) * WR + (x0 - wx0)) * C;
    const float* z01 = zwin + ((y0 - wy0) * WR + (x1 - wx0)) * C;
    const float* z10 = zwin + ((y1 - wy0) * WR + (x0 - wx0)) * C;
    const float* z11 = zwin + ((y1 - wy0) * WR + (x1 - wx0)) * C;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float v = hy * (hx * z00[c] + lx * z01[c]) + ly * (hx * z10[c] + lx * z11[c]);
      gp[c] = v;
      m = fmaxf(m, v);
    }
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(gp[c] - m);
    const float lse = m + logf(sum);
    const int64_t t = tgt[(static_cast<size_t>(n) * H + Y) * W + X];
    const bool valid = t != ignore && t >= 0 && t < C;
    if (valid) {
      loss = lse - gp[static_cast<int>(t)];
      cnt = 1.f;
    }
    for (int c = 0; c < C; ++c) gp[c] = valid ? expf(gp[c] - lse) - (c == static_cast<int>(t) ? 1.f : 0.f) : 0.f;
  } else {
    for (int c = 0; c < C; ++c) gp[c] = 0.f;
  }
  // CTA partial sums (fixed order: deterministic)
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = loss;
    red[1][tid >> 5] = cnt;
  }
  __syncthreads();
  if (tid < 32) {
    float a = red[0][tid], b = red[1][tid];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (tid == 0) {
      const size_t cta = (static_cast<size_t>(n) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      partials[2 * cta] = a;
      partials[2 * cta + 1] = b;
    }
  }
  if (dz == nullptr) return;
  // transposed interpolation: gather from g into the window, flush
  const int span = 2 * s;                         // full-resolution rows with a non-zero weight on one window row
  float* dzn = dz + static_cast<size_t>(n) * h * w * C;
  for (int i = tid; i < WR * WR * C; i += TILE * TILE) {
    const int c = i % C, q = (i / C) % WR, r = i / (C * WR);
    const int k = wy0 + r, j = wx0 + q;
    if (k < 0 || k >= h || j < 0 || j >= w) continue;
    // rows Y with weight on low-res row k lie in [s*k - s/2, s*k + s + s/2 - 1]; clamped rows only add to the edges
    int ta = s * k - s / 2 - Y0, tb = ta + span - 1;
    int ua = s * j - s / 2 - X0, ub = ua + span - 1;
    if (k == 0) ta = 0;
    if (j == 0) ua = 0;
    if (k == h - 1) tb = TILE - 1;
    if (j == w - 1) ub = TILE - 1;
    ta = max(ta, 0);
    ua = max(ua, 0);
    tb = min(tb, TILE - 1);
    ub = min(ub, TILE - 1);
    float acc = 0.f;
    for (int a = ta; a <= tb; ++a) {
      const float wy = wyt[r * TILE + a];
      float row = 0.f;
      for (int b = ua; b <= ub; ++b) row += wxt[q * TILE + b] * g[(a * TILE + b) * CP + c];
      acc += wy * row;
    }
    if (acc != 0.f) atomicAdd(dzn + (static_cast<size_t>(k) * w + j) * C + c, acc);
  }
}

}  // namespace

extern "C" {

int64_t u2b_upsample_ce_num_partials(int64_t N, int H, int W) {
  return N * ((H + TILE - 1) / TILE) * ((W + TILE - 1) / TILE);
}

int u2b_upsample_ce_supported(int C, int scale) {
  if (scale < 2 || (scale & 1) || TILE % scale != 0) return 0;
  const int WR = TILE / scale + 2;
  const size_t smem = (static_cast<size_t>(WR) * WR * C + static_cast<size_t>(TILE) * TILE * (C | 1) + 2 * WR * TILE) * 4;
  return C >= 1 && smem <= 220 * 1024;
}

int u2b_upsample_ce(int dtype, const void* logits, const int64_t* targets, int64_t N, int h, int w, int C,
                    int scale, int64_t ignore_index, float* grad_logits, float* partials, cudaStream_t stream) {
  if (N == 0) return 0;
  U2B_CHECK_ARG(logits && targets && partials && N > 0 && h > 0 && w > 0, "upsample_ce: bad arguments");
  U2B_CHECK_ARG(u2b_upsample_ce_supported(C, scale), "upsample_ce: C=%d scale=%d not supported", C, scale);
  const int H = h * scale, W = w * scale;
  const int WR = TILE / scale + 2;
  const size_t smem = (static_cast<size_t>(WR) * WR * C + static_cast<size_t>(TILE) * TILE * (C | 1) + 2 * WR * TILE) * 4;
  const dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, static_cast<unsigned>(N));
  U2B_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "upsample_ce: image too large");
#define U2B_LAUNCH_CE(T)                                                                                         \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      U2B_CUDA(cudaFuncSetAttribute(upsample_ce_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                    220 * 1024));                                                                \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    upsample_ce_kernel<T><<<grid, TILE * TILE, smem, stream>>>(static_cast<const T*>(logits), targets, h, w, C,  \
                                                               scale, H, W, ignore_index, grad_logits, partials); \
  } while (0)
  if (dtype == 0)
    U2B_LAUNCH_CE(float);
  else if (dtype == 1)
    U2B_LAUNCH_CE(__half);
  else if (dtype == 2)
    U2B_LAUNCH_CE(__nv_bfloat16);
  else {
    u2b_set_error("upsample_ce: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
#undef U2B_LAUNCH_CE
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
