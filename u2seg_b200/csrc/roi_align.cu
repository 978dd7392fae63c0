// ROIAlignV2 (aligned=True, sampling_ratio=0 -> adaptive ceil(roi/bin) samples) over an FPN pyramid,
// NHWC feature maps, with the level assignment of ROIPooler fused into the same launch.
// Reference call chain: detectron2/modeling/poolers.py:206-263 (ROIPooler.forward),
// poolers.py:23-59 (assign_boxes_to_levels), detectron2/layers/roi_align.py:49-65 ->
// torchvision.ops.roi_align (CUDA kernel roi_align_forward_kernel_impl / bilinear_interpolate).
//
// HBM/L2-bound. Forward (roi_align_fwd_kernel): one warp owns one output bin (roi, ph, pw) and sweeps the channel
// vector with 16-byte loads (NHWC makes the 4 bilinear corners 4 contiguous channel vectors); fp32 accumulation;
// one 16-byte store per lane. Backward (roi_align_bwd2_kernel): one CTA per ROI factorises the pooling into two small
// per-axis weight tables and issues ONE vectorised fp32 atomic (red.v4.f32) per footprint pixel and 4 channels into
// per-level fp32 gradient maps; roi_align_bwd_kernel (one warp per bin) and roi_align_fwd2_kernel are the alternative
// implementations selected by u2b_roi_align_set_impl.
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

constexpr int MAX_LEVELS = 4;

struct Pyramid {
  const void* feat[MAX_LEVELS];
  float* grad[MAX_LEVELS];
  int H[MAX_LEVELS];
  int W[MAX_LEVELS];
  float scale[MAX_LEVELS];
  int num_levels;
};

// poolers.py:51-59: floor(canonical_level + log2(sqrt(area)/canonical_size + 1e-8)) clamped, minus min_level
__global__ void assign_levels_kernel(const float* __restrict__ rois, int K, int min_level,
                                     int max_level, float canonical_size, int canonical_level,
                                     int32_t* __restrict__ levels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const float* r = rois + static_cast<size_t>(i) * 5;
  const float area = (r[3] - r[1]) * (r[4] - r[2]);
  const float sz = sqrtf(area);
  float lv = floorf(static_cast<float>(canonical_level) + log2f(sz / canonical_size + 1e-8f));
  lv = fminf(fmaxf(lv, static_cast<float>(min_level)), static_cast<float>(max_level));
  // NaN (negative area) -> torch.clamp keeps NaN -> int64 cast is undefined in the reference; map to min level
  levels[i] = (lv == lv) ? static_cast<int>(lv) - min_level : 0;
}

template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  using Raw = float4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__half> {
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
template <>
struct Vec<__nv_bfloat16> {
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

struct Sample {
  int o1, o2, o3, o4;  // pixel offsets (in pixels, row-major) of the 4 corners
  float w1, w2, w3, w4;
  bool valid;
};

// torchvision bilinear_interpolate (roi_align_kernel.cu): same comparisons / clamps / op order.
__device__ __forceinline__ Sample make_sample(float y, float x, int H, int W) {
  Sample s;
  s.valid = !(y < -1.0f || y > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W));
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = static_cast<int>(y), x_low = static_cast<int>(x);
  int y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else { y_high = y_low + 1; }
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else { x_high = x_low + 1; }
  const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
  s.w1 = hy * hx; s.w2 = hy * lx; s.w3 = ly * hx; s.w4 = ly * lx;
  s.o1 = y_low * W + x_low; s.o2 = y_low * W + x_high; s.o3 = y_high * W + x_low; s.o4 = y_high * W + x_high;
  if (!s.valid) { s.o1 = s.o2 = s.o3 = s.o4 = 0; s.w1 = s.w2 = s.w3 = s.w4 = 0.f; }
  return s;
}

struct BinGeom {
  float start_h, start_w, bin_h, bin_w;
  int grid_h, grid_w;
  float count;
};
__device__ __forceinline__ BinGeom bin_geometry(const float* r, float scale, int P) {
  BinGeom g;
  const float offset = 0.5f;  // aligned=True
  g.start_w = r[1] * scale - offset;
  g.start_h = r[2] * scale - offset;
  const float end_w = r[3] * scale - offset, end_h = r[4] * scale - offset;
  const float roi_w = end_w - g.start_w, roi_h = end_h - g.start_h;
  g.bin_h = roi_h / static_cast<float>(P);
  g.bin_w = roi_w / static_cast<float>(P);
  g.grid_h = static_cast<int>(ceilf(roi_h / P));
  g.grid_w = static_cast<int>(ceilf(roi_w / P));
  const float c = static_cast<float>(g.grid_h) * static_cast<float>(g.grid_w);
  g.count = fmaxf(c, 1.f);
  return g;
}

// value of one output bin for the VN channels starting at c0 (roi_align_kernel.cu sampling, adaptive grid)
template <typename T>
__device__ __forceinline__ void bin_forward(const T* __restrict__ feat, int H, int W, int C, const BinGeom& g, int ph,
                                            int pw, int c0, float (&acc)[Vec<T>::N]) {
  constexpr int VN = Vec<T>::N;
#pragma unroll
  for (int i = 0; i < VN; ++i) acc[i] = 0.f;
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h);
    for (int ix = 0; ix < g.grid_w; ++ix) {
      const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w);
      const Sample s = make_sample(y, x, H, W);
      if (!s.valid) continue;
      float v1[VN], v2[VN], v3[VN], v4[VN];
      Vec<T>::load(feat + static_cast<size_t>(s.o1) * C + c0, v1);
      Vec<T>::load(feat + static_cast<size_t>(s.o2) * C + c0, v2);
      Vec<T>::load(feat + static_cast<size_t>(s.o3) * C + c0, v3);
      Vec<T>::load(feat + static_cast<size_t>(s.o4) * C + c0, v4);
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] += s.w1 * v1[i] + s.w2 * v2[i] + s.w3 * v3[i] + s.w4 * v4[i];
    }
  }
#pragma unroll
  for (int i = 0; i < VN; ++i) acc[i] /= g.count;
}

template <typename T>
__global__ void __launch_bounds__(256)
roi_align_fwd_kernel(Pyramid pyr, int C, const float* __restrict__ rois,
                     const int32_t* __restrict__ levels, int K, int P, T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const long long bin = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bin >= static_cast<long long>(K) * P * P) return;
  const int k = static_cast<int>(bin / (P * P));
  const int ph = static_cast<int>((bin / P) % P), pw = static_cast<int>(bin % P);
  const float* r = rois + static_cast<size_t>(k) * 5;
  const int lvl = levels ? levels[k] : 0;
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  const int b = static_cast<int>(r[0]);
  const T* feat = static_cast<const T*>(pyr.feat[lvl]) + static_cast<size_t>(b) * H * W * C;
  const BinGeom g = bin_geometry(r, pyr.scale[lvl], P);
  T* o = out + static_cast<size_t>(bin) * C;
  for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
    float acc[VN];
    bin_forward<T>(feat, H, W, C, g, ph, pw, c0, acc);
    Vec<T>::store(o + c0, acc);
  }
}

// One axis of torchvision's bilinear_interpolate: the two pixels a sample touches and their weights.
struct AxisSample {
  int lo, hi;
  float wlo, whi;
  bool valid;
};
__device__ __forceinline__ AxisSample axis_sample(float v, int size) {
  AxisSample a;
  a.valid = !(v < -1.0f || v > static_cast<float>(size));
  if (v <= 0.f) v = 0.f;
  a.lo = static_cast<int>(v);
  if (a.lo >= size - 1) {
    a.hi = a.lo = size - 1;
    v = static_cast<float>(a.lo);
  } else {
    a.hi = a.lo + 1;
  }
  a.whi = v - a.lo;
  a.wlo = 1.f - a.whi;
  return a;
}

// grad_out: (K, P, P, C) of T; grad feature maps: fp32 NHWC per level (pre-zeroed by the caller).
// One warp per output bin, lanes over channels. A bin's samples form a product grid and the bilinear weight of a
// sample is wy * wx, so the total weight a feature pixel (py, px) receives from the bin factorises into
// WY[py] * WX[px] (sums over the sample rows / columns that touch it). The kernel therefore issues ONE vector
// atomic per touched pixel ((g+1)^2 of them for a g x g sample grid, samples being < 1 px apart) instead of four
// per sample (4 g^2): 1.8x fewer for g = 2, 2.6x for g = 4. WX is computed once per bin, one pixel column per lane,
// and broadcast with shuffles; bins wider than 32 columns (not produced by the level assignment) fall back to
// the per-sample path.
// Scatter the gradient of one output bin into the fp32 feature-gradient map; executed by a whole warp. `load_g(c0, gv)`
// supplies the VN upstream-gradient values of the bin for the lane's channels (from global memory or, for the
// channel-major layout, from the ROI's shared-memory tile).
template <typename T, typename LoadG>
__device__ __forceinline__ void bin_backward(float* __restrict__ gfeat, int H, int W, int C, const BinGeom& g, int ph,
                                             int pw, int lane, float grad_scale, LoadG load_g) {
  constexpr int VN = Vec<T>::N;
  const float y_base = g.start_h + ph * g.bin_h, x_base = g.start_w + pw * g.bin_w;
  // pixel columns touched by the bin: [cx0, cx1]; lane j owns column cx0 + j
  int cx0 = W, cx1 = -1;
  for (int ix = 0; ix < g.grid_w; ++ix) {
    const AxisSample a = axis_sample(x_base + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w), W);
    if (a.valid) {
      cx0 = min(cx0, a.lo);
      cx1 = max(cx1, a.hi);
    }
  }
  int cy0 = H, cy1 = -1;
  for (int iy = 0; iy < g.grid_h; ++iy) {
    const AxisSample a = axis_sample(y_base + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h), H);
    if (a.valid) {
      cy0 = min(cy0, a.lo);
      cy1 = max(cy1, a.hi);
    }
  }
  if (cx1 < cx0 || cy1 < cy0) return;             // every sample outside the map: no gradient
  const int ncols = cx1 - cx0 + 1;
  if (ncols <= 32) {
    float wx_lane = 0.f;
    {
      const int cx = cx0 + lane;
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const AxisSample a = axis_sample(x_base + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w), W);
        if (a.valid) wx_lane += (a.lo == cx ? a.wlo : 0.f) + (a.hi == cx ? a.whi : 0.f);
      }
    }
    const float inv_count = grad_scale / g.count;
    for (int c0 = lane * VN; c0 - lane * VN < C; c0 += 32 * VN) {   // whole warp iterates together
      const bool act = c0 < C;
      float gv[VN];
      if (act) {
        load_g(c0, gv);
#pragma unroll
        for (int i = 0; i < VN; ++i) gv[i] *= inv_count;
      }
      for (int py = cy0; py <= cy1; ++py) {
        float wy = 0.f;
        for (int iy = 0; iy < g.grid_h; ++iy) {
          const AxisSample a = axis_sample(y_base + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h), H);
          if (a.valid) wy += (a.lo == py ? a.wlo : 0.f) + (a.hi == py ? a.whi : 0.f);
        }
        for (int j = 0; j < ncols; ++j) {
          const float w = wy * __shfl_sync(0xffffffffu, wx_lane, j);
          if (!act || w == 0.f) continue;
          float* dst = gfeat + (static_cast<size_t>(py) * W + cx0 + j) * C + c0;
#pragma unroll
          for (int i = 0; i < VN; i += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + i),
                      make_float4(w * gv[i], w * gv[i + 1], w * gv[i + 2], w * gv[i + 3]));
        }
      }
    }
    return;
  }
  for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
    float gv[VN];
    load_g(c0, gv);
#pragma unroll
    for (int i = 0; i < VN; ++i) gv[i] = gv[i] / g.count * grad_scale;
    for (int iy = 0; iy < g.grid_h; ++iy) {
      const float y = y_base + (iy + 0.5f) * g.bin_h / static_cast<float>(g.grid_h);
      for (int ix = 0; ix < g.grid_w; ++ix) {
        const float x = x_base + (ix + 0.5f) * g.bin_w / static_cast<float>(g.grid_w);
        const Sample s = make_sample(y, x, H, W);
        if (!s.valid) continue;
        const int offs[4] = {s.o1, s.o2, s.o3, s.o4};
        const float ws[4] = {s.w1, s.w2, s.w3, s.w4};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float* dst = gfeat + static_cast<size_t>(offs[q]) * C + c0;
#pragma unroll
          for (int i = 0; i < VN; i += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + i),
                      make_float4(ws[q] * gv[i], ws[q] * gv[i + 1], ws[q] * gv[i + 2], ws[q] * gv[i + 3]));
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
roi_align_bwd_kernel(Pyramid pyr, int C, const float* __restrict__ rois,
                     const int32_t* __restrict__ levels, int K, int P, const T* __restrict__ gout, float grad_scale) {
  const int lane = threadIdx.x & 31;
  const long long bin = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bin >= static_cast<long long>(K) * P * P) return;
  const int k = static_cast<int>(bin / (P * P));
  const int ph = static_cast<int>((bin / P) % P), pw = static_cast<int>(bin % P);
  const float* r = rois + static_cast<size_t>(k) * 5;
  const int lvl = levels ? levels[k] : 0;
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  const int b = static_cast<int>(r[0]);
  float* gfeat = pyr.grad[lvl] + static_cast<size_t>(b) * H * W * C;
  const BinGeom g = bin_geometry(r, pyr.scale[lvl], P);
  const T* go = gout + static_cast<size_t>(bin) * C;
  bin_backward<T>(gfeat, H, W, C, g, ph, pw, lane, grad_scale,
                  [&](int c0, float (&gv)[Vec<T>::N]) { Vec<T>::load(go + c0, gv); });
}

// ---- per-ROI factorised kernels (round 2) -----------------------------------------------------------------------------
// A sample's bilinear weight is wy * wx and a bin's samples form a product grid, so the whole ROI pooling is a separable
// linear map:   Y[ph][pw] = 1/count * sum_py sum_px AY[py][ph] * BX[px][pw] * X[py][px],
// AY[py][ph] = total y-weight the sample rows of bin row ph give pixel row py (BX likewise; invalid samples - outside
// [-1, size] - contribute nothing, and validity is per axis, roi_align_kernel.cu bilinear_interpolate). One CTA per ROI
// builds the two small tables in shared memory once; then
//   forward : a warp per output bin gathers the (rows x cols) pixels its samples touch - (g+1)^2 loads for a g x g sample
//             grid instead of 4 g^2, no per-sample index arithmetic;
//   backward: a warp per FOOTPRINT PIXEL gathers the <= 2 x 2 bins that touch it and issues ONE vector atomic per pixel
//             and 4 channels - (P g + 1)^2 per ROI instead of P^2 (g + 1)^2 (bins share their border pixels): 2x fewer
//             for g = 2. Results equal the per-sample kernels up to fp32 summation order.
// ROIs whose footprint exceeds RA_FMAX pixels on an axis (not produced by the level assignment for boxes inside the image)
// are handled by the per-bin routines inside the same launch.
constexpr int RA_FMAX = 64;
constexpr int RA_PMAX = 14;

struct RoiTables {
  float ay[RA_FMAX][RA_PMAX];
  float bx[RA_FMAX][RA_PMAX];
  int lo[2][RA_PMAX], hi[2][RA_PMAX];      // per bin row / column: first and last touched pixel (absolute), hi < lo = none
  int bin_lo[2][RA_FMAX], bin_hi[2][RA_FMAX];   // per footprint row / column: first and last bin with a non-zero weight
  int org[2], ext[2];
};

// returns false when the footprint does not fit the tables (caller falls back); all threads of the CTA must call it
__device__ __forceinline__ bool roi_build_tables(RoiTables& t, const BinGeom& g, int P, int H, int W) {
  const int tid = threadIdx.x;
  if (tid < 2 * P) {
    const int axis = tid / P, p = tid % P;          // axis 0: y, 1: x
    const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
    const int grid = axis ? g.grid_w : g.grid_h, size = axis ? W : H;
    int lo = 0x7fffffff, hi = -1;
    for (int i = 0; i < grid; ++i) {
      const AxisSample a = axis_sample(start + p * bin + (i + 0.5f) * bin / static_cast<float>(grid), size);
      if (a.valid) {
        lo = min(lo, a.lo);
        hi = max(hi, a.hi);
      }
    }
    t.lo[axis][p] = lo;
    t.hi[axis][p] = hi;
  }
  __syncthreads();
  if (tid < 2) {
    int lo = 0x7fffffff, hi = -1;
    for (int p = 0; p < P; ++p)
      if (t.hi[tid][p] >= t.lo[tid][p]) {
        lo = min(lo, t.lo[tid][p]);
        hi = max(hi, t.hi[tid][p]);
      }
    t.org[tid] = lo;
    t.ext[tid] = hi >= lo ? hi - lo + 1 : 0;
  }
  for (int i = tid; i < RA_FMAX * RA_PMAX; i += blockDim.x) {
    (&t.ay[0][0])[i] = 0.f;
    (&t.bx[0][0])[i] = 0.f;
  }
  __syncthreads();
  if (t.ext[0] > RA_FMAX || t.ext[1] > RA_FMAX) return false;
  if (tid < 2 * P) {      // bin p only writes column p of its axis' table: no conflicts
    const int axis = tid / P, p = tid % P;
    const float start = axis ? g.start_w : g.start_h, bin = axis ? g.bin_w : g.bin_h;
    const int grid = axis ? g.grid_w : g.grid_h, size = axis ? W : H;
    float (*tab)[RA_PMAX] = axis ? t.bx : t.ay;
    const int org = t.org[axis];
    for (int i = 0; i < grid; ++i) {
      const AxisSample a = axis_sample(start + p * bin + (i + 0.5f) * bin / static_cast<float>(grid), size);
      if (a.valid) {
        tab[a.lo - org][p] += a.wlo;
        tab[a.hi - org][p] += a.whi;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * RA_FMAX; i += blockDim.x) {
    const int axis = i / RA_FMAX, r = i % RA_FMAX;
    int lo = P, hi = -1;
    if (r < t.ext[axis]) {
      const float (*tab)[RA_PMAX] = axis ? t.bx : t.ay;
      for (int p = 0; p < P; ++p)
        if (tab[r][p] != 0.f) {
          lo = min(lo, p);
          hi = p;
        }
    }
    t.bin_lo[axis][r] = lo;
    t.bin_hi[axis][r] = hi;
  }
  __syncthreads();
  return true;
}

template <typename T>
__global__ void __launch_bounds__(256)
roi_align_fwd2_kernel(Pyramid pyr, int C, const float* __restrict__ rois, const int32_t* __restrict__ levels, int K, int P,
                      T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  __shared__ RoiTables t;
  const int k = blockIdx.x, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x >> 5) * gridDim.y, warp = (threadIdx.x >> 5) * gridDim.y + blockIdx.y;   // work items are strided over the ROI's CTAs
  const float* r = rois + static_cast<size_t>(k) * 5;
  const int lvl = levels ? levels[k] : 0;
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  const T* feat = static_cast<const T*>(pyr.feat[lvl]) + static_cast<size_t>(static_cast<int>(r[0])) * H * W * C;
  const BinGeom g = bin_geometry(r, pyr.scale[lvl], P);
  T* o = out + static_cast<size_t>(k) * P * P * C;
  if (!roi_build_tables(t, g, P, H, W)) {
    for (int bin = warp; bin < P * P; bin += nwarps)
      for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
        float acc[VN];
        bin_forward<T>(feat, H, W, C, g, bin / P, bin % P, c0, acc);
        Vec<T>::store(o + static_cast<size_t>(bin) * C + c0, acc);
      }
    return;
  }
  const float inv_count = 1.f / g.count;
  const int oy = t.org[0], ox = t.org[1];
  for (int bin = warp; bin < P * P; bin += nwarps) {
    const int ph = bin / P, pw = bin % P;
    const int y0 = t.lo[0][ph], y1 = t.hi[0][ph], x0 = t.lo[1][pw], x1 = t.hi[1][pw];
    for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
      float acc[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] = 0.f;
      for (int py = y0; py <= y1; ++py) {
        const float wy = t.ay[py - oy][ph];
        if (wy == 0.f) continue;
        const T* row = feat + (static_cast<size_t>(py) * W) * C + c0;
        for (int px = x0; px <= x1; ++px) {
          const float w = wy * t.bx[px - ox][pw];
          if (w == 0.f) continue;
          float v[VN];
          Vec<T>::load(row + static_cast<size_t>(px) * C, v);
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[i] = fmaf(w, v[i], acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] *= inv_count;
      Vec<T>::store(o + static_cast<size_t>(bin) * C + c0, acc);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
roi_align_bwd2_kernel(Pyramid pyr, int C, const float* __restrict__ rois, const int32_t* __restrict__ levels, int K, int P,
                      const T* __restrict__ gout, float grad_scale) {
  constexpr int VN = Vec<T>::N;
  __shared__ RoiTables t;
  const int k = blockIdx.x, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x >> 5) * gridDim.y, warp = (threadIdx.x >> 5) * gridDim.y + blockIdx.y;   // work items are strided over the ROI's CTAs
  const float* r = rois + static_cast<size_t>(k) * 5;
  const int lvl = levels ? levels[k] : 0;
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  float* gfeat = pyr.grad[lvl] + static_cast<size_t>(static_cast<int>(r[0])) * H * W * C;
  const BinGeom g = bin_geometry(r, pyr.scale[lvl], P);
  const T* go = gout + static_cast<size_t>(k) * P * P * C;
  if (!roi_build_tables(t, g, P, H, W)) {
    for (int bin = warp; bin < P * P; bin += nwarps) {
      const T* gb = go + static_cast<size_t>(bin) * C;
      bin_backward<T>(gfeat, H, W, C, g, bin / P, bin % P, lane, grad_scale,
                      [&](int c0, float (&gv)[Vec<T>::N]) { Vec<T>::load(gb + c0, gv); });
    }
    return;
  }
  const float scale = grad_scale / g.count;
  const int oy = t.org[0], ox = t.org[1], eh = t.ext[0], ew = t.ext[1];
  for (int pix = warp; pix < eh * ew; pix += nwarps) {
    const int ry = pix / ew, rx = pix % ew;
    const int ph0 = t.bin_lo[0][ry], ph1 = t.bin_hi[0][ry], pw0 = t.bin_lo[1][rx], pw1 = t.bin_hi[1][rx];
    if (ph1 < ph0 || pw1 < pw0) continue;
    float* dst = gfeat + (static_cast<size_t>(oy + ry) * W + ox + rx) * C;
    for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
      float acc[VN];
#pragma unroll
      for (int i = 0; i < VN; ++i) acc[i] = 0.f;
      for (int ph = ph0; ph <= ph1; ++ph) {
        const float wy = t.ay[ry][ph];
        if (wy == 0.f) continue;
        for (int pw = pw0; pw <= pw1; ++pw) {
          const float w = wy * t.bx[rx][pw];
          if (w == 0.f) continue;
          float v[VN];
          Vec<T>::load(go + static_cast<size_t>(ph * P + pw) * C + c0, v);
#pragma unroll
          for (int i = 0; i < VN; ++i) acc[i] = fmaf(w, v[i], acc[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < VN; i += 4)
        atomicAdd(reinterpret_cast<float4*>(dst + c0 + i),
                  make_float4(scale * acc[i], scale * acc[i + 1], scale * acc[i + 2], scale * acc[i + 3]));
    }
  }
}

// CTAs per ROI: few ROIs (inference mask head: 100) would leave most SMs idle with one CTA each
inline unsigned roi2_splits(long long K) {
  const long long want = (4LL * u2b_num_sms() + K - 1) / K;
  return static_cast<unsigned>(want < 1 ? 1 : (want > 4 ? 4 : want));
}

// bit 0: forward, bit 1: backward use the per-ROI factorised kernel. Measured in the training step (1024 ROIs, 7x7, 256 ch,
// bf16): backward 213 -> 174 us (half the atomics); forward 86 -> 98 us (its 49 bins do not amortise the table set-up), so the
// forward stays on the warp-per-bin kernel by default.
int g_roi_align_impl = 2;

// ---- channel-major ("CHW") output layout: out (K, C, P, P) contiguous, i.e. what torch.flatten(x, 1) of the box head
// (box_head.py:99-106) wants. One CTA per ROI: the P*P bins are computed by the 8 warps into a shared-memory tile
// [bin][C + 1] (fp32), which is then written out as C*P*P contiguous elements (and read back the same way in backward),
// so the transposing copies `flatten` / `to_nhwc` on (K, 256, 7, 7) disappear (6 x 45 us per step at K = 1024).
template <typename T>
__global__ void __launch_bounds__(256)
roi_align_fwd_chw_kernel(Pyramid pyr, int C, const float* __restrict__ rois, const int32_t* __restrict__ levels,
                         int K, int P, T* __restrict__ out) {
  constexpr int VN = Vec<T>::N;
  extern __shared__ float tile[];                    // [P*P][C + 1]
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* r = rois + static_cast<size_t>(k) * 5;
  const int lvl = levels ? levels[k] : 0;
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  const T* feat = static_cast<const T*>(pyr.feat[lvl]) + static_cast<size_t>(static_cast<int>(r[0])) * H * W * C;
  const BinGeom g = bin_geometry(r, pyr.scale[lvl], P);
  const int bins = P * P, pitch = C + 1;
  for (int bl = warp; bl < bins; bl += nw) {
    for (int c0 = lane * VN; c0 < C; c0 += 32 * VN) {
      float acc[VN];
      bin_forward<T>(feat, H, W, C, g, bl / P, bl % P, c0, acc);
#pragma unroll
      for (int i = 0; i < VN; ++i) tile[bl * pitch + c0 + i] = acc[i];
    }
  }
  __syncthreads();
  T* o = out + static_cast<size_t>(k) * C * bins;
  for (int e = threadIdx.x; e < C * bins; e += blockDim.x) {       // e = c * bins + bin: contiguous in the output
    const int c = e / bins, bl = e - c * bins;
    const float v = tile[bl * pitch + c];
    if constexpr (sizeof(T) == 4) o[e] = v;
    else if constexpr (std::is_same<T, __half>::value) o[e] = __float2half(v);
    else o[e] = __float2bfloat16(v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
roi_align_bwd_chw_kernel(Pyramid pyr, int C, const float* __restrict__ rois, const int32_t* __restrict__ levels,
                         int K, int P, const T* __restrict__ gout, float grad_scale) {
  constexpr int VN = Vec<T>::N;
  extern __shared__ float tile[];                    // [P*P][C + 1]
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* r = rois + static_cast<size_t>(k) * 5;
  const int lvl = levels ? levels[k] : 0;
  const int H = pyr.H[lvl], W = pyr.W[lvl];
  float* gfeat = pyr.grad[lvl] + static_cast<size_t>(static_cast<int>(r[0])) * H * W * C;
  const BinGeom g = bin_geometry(r, pyr.scale[lvl], P);
  const int bins = P * P, pitch = C + 1;
  const T* go = gout + static_cast<size_t>(k) * C * bins;
  for (int e = threadIdx.x; e < C * bins; e += blockDim.x) {
    const int c = e / bins, bl = e - c * bins;
    float v;
    if constexpr (sizeof(T) == 4) v = go[e];
    else if constexpr (std::is_same<T, __half>::value) v = __half2float(go[e]);
    else v = __bfloat162float(go[e]);
    tile[bl * pitch + c] = v;
  }
  __syncthreads();
  for (int bl = warp; bl < bins; bl += nw) {
    const float* row = tile + bl * pitch;
    bin_backward<T>(gfeat, H, W, C, g, bl / P, bl % P, lane, grad_scale, [&](int c0, float (&gv)[VN]) {
#pragma unroll
      for (int i = 0; i < VN; ++i) gv[i] = row[c0 + i];
    });
  }
}

int fill_pyramid(Pyramid* p, int num_levels, const void* const* feats, float* const* grads,
                 const int32_t* hs, const int32_t* ws, const float* scales) {
  if (num_levels < 1 || num_levels > MAX_LEVELS) return -1;
  p->num_levels = num_levels;
  for (int i = 0; i < num_levels; ++i) {
    p->feat[i] = feats ? feats[i] : nullptr;
    p->grad[i] = grads ? grads[i] : nullptr;
    p->H[i] = hs[i];
    p->W[i] = ws[i];
    p->scale[i] = scales[i];
  }
  return 0;
}

}  // namespace

extern "C" {

int u2b_assign_levels(const float* rois5, int64_t K, int min_level, int max_level,
                      float canonical_size, int canonical_level, int32_t* levels,
                      cudaStream_t stream) {
  if (K == 0) return 0;
  U2B_CHECK_ARG(rois5 && levels && K > 0, "assign_levels: bad arguments");
  assign_levels_kernel<<<static_cast<unsigned>((K + 255) / 256), 256, 0, stream>>>(
      rois5, static_cast<int>(K), min_level, max_level, canonical_size, canonical_level, levels);
  U2B_LAUNCH_CHECK();
  return 0;
}

// dtype: 0 = fp32, 1 = fp16, 2 = bf16. feats[l]: (N, H_l, W_l, C) NHWC. out: (K, P, P, C).
int u2b_roi_align_fwd(int dtype, int num_levels, const void* const* feats, const int32_t* hs,
                      const int32_t* ws, const float* scales, int64_t C, const float* rois5,
                      const int32_t* levels, int64_t K, int P, void* out, cudaStream_t stream) {
  if (K == 0) return 0;
  Pyramid p;
  U2B_CHECK_ARG(feats && hs && ws && scales && rois5 && out, "roi_align_fwd: null pointer");
  U2B_CHECK_ARG(fill_pyramid(&p, num_levels, feats, nullptr, hs, ws, scales) == 0,
                "roi_align_fwd: 1..4 levels supported");
  U2B_CHECK_ARG(num_levels == 1 || levels, "roi_align_fwd: levels required for a pyramid");
  const int vn = dtype == 0 ? 4 : 8;
  U2B_CHECK_ARG(C > 0 && C % vn == 0, "roi_align_fwd: C=%lld must be a multiple of %d", (long long)C, vn);
  const long long bins = static_cast<long long>(K) * P * P;
  const unsigned grid = static_cast<unsigned>((bins + 7) / 8);
  if ((g_roi_align_impl & 1) && P <= RA_PMAX && (dtype == 0 || dtype == 1 || dtype == 2)) {
    const dim3 g2(static_cast<unsigned>(K), roi2_splits(K));
    if (dtype == 0) roi_align_fwd2_kernel<float><<<g2, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (float*)out);
    else if (dtype == 1) roi_align_fwd2_kernel<__half><<<g2, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (__half*)out);
    else roi_align_fwd2_kernel<__nv_bfloat16><<<g2, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (__nv_bfloat16*)out);
    U2B_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == 0)
    roi_align_fwd_kernel<float><<<grid, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (float*)out);
  else if (dtype == 1)
    roi_align_fwd_kernel<__half><<<grid, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (__half*)out);
  else if (dtype == 2)
    roi_align_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P,
                                                                 (__nv_bfloat16*)out);
  else {
    u2b_set_error("roi_align_fwd: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
  U2B_LAUNCH_CHECK();
  return 0;
}

// grad_feats[l]: (N, H_l, W_l, C) fp32, must be zero-initialised by the caller (accumulated into).
int u2b_roi_align_bwd(int dtype, int num_levels, float* const* grad_feats, const int32_t* hs,
                      const int32_t* ws, const float* scales, int64_t C, const float* rois5,
                      const int32_t* levels, int64_t K, int P, const void* grad_out, float grad_scale,
                      cudaStream_t stream) {
  if (K == 0) return 0;
  Pyramid p;
  U2B_CHECK_ARG(grad_feats && hs && ws && scales && rois5 && grad_out, "roi_align_bwd: null pointer");
  U2B_CHECK_ARG(fill_pyramid(&p, num_levels, nullptr, grad_feats, hs, ws, scales) == 0,
                "roi_align_bwd: 1..4 levels supported");
  U2B_CHECK_ARG(num_levels == 1 || levels, "roi_align_bwd: levels required for a pyramid");
  const int vn = dtype == 0 ? 4 : 8;
  U2B_CHECK_ARG(C > 0 && C % vn == 0, "roi_align_bwd: C must be a multiple of %d", vn);
  const long long bins = static_cast<long long>(K) * P * P;
  const unsigned grid = static_cast<unsigned>((bins + 7) / 8);
  if ((g_roi_align_impl & 2) && P <= RA_PMAX && (dtype == 0 || dtype == 1 || dtype == 2)) {
    const dim3 g2(static_cast<unsigned>(K), roi2_splits(K));
    if (dtype == 0) roi_align_bwd2_kernel<float><<<g2, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const float*)grad_out, grad_scale);
    else if (dtype == 1) roi_align_bwd2_kernel<__half><<<g2, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const __half*)grad_out, grad_scale);
    else roi_align_bwd2_kernel<__nv_bfloat16><<<g2, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const __nv_bfloat16*)grad_out, grad_scale);
    U2B_LAUNCH_CHECK();
    return 0;
  }
  if (dtype == 0)
    roi_align_bwd_kernel<float><<<grid, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const float*)grad_out, grad_scale);
  else if (dtype == 1)
    roi_align_bwd_kernel<__half><<<grid, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const __half*)grad_out, grad_scale);
  else if (dtype == 2)
    roi_align_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p, (int)C, rois5, levels, (int)K, P,
                                                                 (const __nv_bfloat16*)grad_out, grad_scale);
  else {
    u2b_set_error("roi_align_bwd: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
  U2B_LAUNCH_CHECK();
  return 0;
}


// bit 0: forward, bit 1: backward through the per-ROI factorised kernels (default 2); 0 = one warp per output bin for both.
// Both produce the same values up to fp32 summation order.
int u2b_roi_align_set_impl(int impl) {
  U2B_CHECK_ARG(impl >= 0 && impl <= 3, "roi_align_set_impl: 0..3");
  g_roi_align_impl = impl;
  return 0;
}

// ---- channel-major layout (round-2 draft): out / grad_out are (K, C, P, P) contiguous ----
int u2b_roi_align_chw_supported(int64_t C, int P) {
  return C > 0 && C % 8 == 0 && P > 0 && static_cast<size_t>(P) * P * (C + 1) * sizeof(float) <= 200 * 1024;
}

#define U2B_CHW_ATTR(KERNEL)                                                                                       \
  do {                                                                                                             \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      U2B_CUDA(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));             \
      attr_set = true;                                                                                             \
    }                                                                                                              \
  } while (0)

int u2b_roi_align_fwd_chw(int dtype, int num_levels, const void* const* feats, const int32_t* hs, const int32_t* ws,
                          const float* scales, int64_t C, const float* rois5, const int32_t* levels, int64_t K, int P,
                          void* out, cudaStream_t stream) {
  if (K == 0) return 0;
  Pyramid p;
  U2B_CHECK_ARG(feats && hs && ws && scales && rois5 && out, "roi_align_fwd_chw: null pointer");
  U2B_CHECK_ARG(fill_pyramid(&p, num_levels, feats, nullptr, hs, ws, scales) == 0, "roi_align_fwd_chw: 1..4 levels supported");
  U2B_CHECK_ARG(num_levels == 1 || levels, "roi_align_fwd_chw: levels required for a pyramid");
  U2B_CHECK_ARG(u2b_roi_align_chw_supported(C, P), "roi_align_fwd_chw: C=%lld P=%d not supported", (long long)C, P);
  const size_t smem = static_cast<size_t>(P) * P * (C + 1) * sizeof(float);
  const unsigned grid = static_cast<unsigned>(K);
  if (dtype == 0) {
    U2B_CHW_ATTR(roi_align_fwd_chw_kernel<float>);
    roi_align_fwd_chw_kernel<float><<<grid, 256, smem, stream>>>(p, (int)C, rois5, levels, (int)K, P, (float*)out);
  } else if (dtype == 1) {
    U2B_CHW_ATTR(roi_align_fwd_chw_kernel<__half>);
    roi_align_fwd_chw_kernel<__half><<<grid, 256, smem, stream>>>(p, (int)C, rois5, levels, (int)K, P, (__half*)out);
  } else if (dtype == 2) {
    U2B_CHW_ATTR(roi_align_fwd_chw_kernel<__nv_bfloat16>);
    roi_align_fwd_chw_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(p, (int)C, rois5, levels, (int)K, P,
                                                                       (__nv_bfloat16*)out);
  } else {
    u2b_set_error("roi_align_fwd_chw: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_roi_align_bwd_chw(int dtype, int num_levels, float* const* grad_feats, const int32_t* hs, const int32_t* ws,
                          const float* scales, int64_t C, const float* rois5, const int32_t* levels, int64_t K, int P,
                          const void* grad_out, float grad_scale, cudaStream_t stream) {
  if (K == 0) return 0;
  Pyramid p;
  U2B_CHECK_ARG(grad_feats && hs && ws && scales && rois5 && grad_out, "roi_align_bwd_chw: null pointer");
  U2B_CHECK_ARG(fill_pyramid(&p, num_levels, nullptr, grad_feats, hs, ws, scales) == 0, "roi_align_bwd_chw: 1..4 levels supported");
  U2B_CHECK_ARG(num_levels == 1 || levels, "roi_align_bwd_chw: levels required for a pyramid");
  U2B_CHECK_ARG(u2b_roi_align_chw_supported(C, P), "roi_align_bwd_chw: C=%lld P=%d not supported", (long long)C, P);
  const size_t smem = static_cast<size_t>(P) * P * (C + 1) * sizeof(float);
  const unsigned grid = static_cast<unsigned>(K);
  if (dtype == 0) {
    U2B_CHW_ATTR(roi_align_bwd_chw_kernel<float>);
    roi_align_bwd_chw_kernel<float><<<grid, 256, smem, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const float*)grad_out, grad_scale);
  } else if (dtype == 1) {
    U2B_CHW_ATTR(roi_align_bwd_chw_kernel<__half>);
    roi_align_bwd_chw_kernel<__half><<<grid, 256, smem, stream>>>(p, (int)C, rois5, levels, (int)K, P, (const __half*)grad_out, grad_scale);
  } else if (dtype == 2) {
    U2B_CHW_ATTR(roi_align_bwd_chw_kernel<__nv_bfloat16>);
    roi_align_bwd_chw_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(p, (int)C, rois5, levels, (int)K, P,
                                                                       (const __nv_bfloat16*)grad_out, grad_scale);
  } else {
    u2b_set_error("roi_align_bwd_chw: unknown dtype %d", dtype);
    return U2B_ERR_BAD_ARG;
  }
  U2B_LAUNCH_CHECK();
  return 0;
}
#undef U2B_CHW_ATTR

}  // extern "C"
