// ResNet stem convolution (7x7, stride 2, pad 3, 3 -> 64 channels) of the Panoptic-FPN backbone, forward and
// weight gradient. detectron2/modeling/backbone/resnet.py:338-362 BasicStem.conv1 (the only layer whose input has 3
// channels; the input image needs no gradient, so backward is the weight gradient alone).
//
// Why not the tcgen05 implicit-GEMM kernel of conv_tc.cu: with Cin = 3 a TMA box row is 6 bytes and K = 147 — the
// tensor-core work is negligible (9.9 GFLOP per pass at 2 x 1024^2) and both passes are bound by streaming the
// 64-channel side once: forward writes N*OH*OW*64 bf16 (67 MB), wgrad reads the same amount of dY. The library
// kernels need 0.28 ms (fprop) and 1.29 ms (wgrad, Cin=3 'indexed' path); the HBM floor is ~11 us each.
// Both kernels here are warp-level mma.sync (m16n8k16, bf16 x bf16 -> fp32) implicit GEMMs whose im2col operand is
// gathered straight from a zero-padded input window in shared memory (13 rows x 69 px x 3 ch for a 4 x 32 output
// tile); K is padded 147 -> 160 with zero weights.
//   forward:  M = pixels (one m16 tile per warp), N = 64, K = 160; weights resident in smem for the whole
//             (persistent) CTA; fp32 accumulators -> bf16 through a warp-private staging tile -> 16-byte stores.
//   wgrad:    M = 64 (co), N = 160 (patch index), K = pixels; dY tile in swizzled smem (ldmatrix.trans), each
//             persistent CTA keeps its 64 x 160 partial in registers across all its tiles and writes it once;
//             the caller sums the per-CTA partials (fixed order: deterministic).
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

constexpr int CO = 64, KS = 7, CI = 3, KREAL = KS * KS * CI, KPAD = 160;
constexpr int TH = 4, TW = 32;                     // output tile
constexpr int WIN_H = 2 * TH + 5, WIN_W = 2 * TW + 5;   // 13 x 69 input pixels
constexpr int WIN_PITCH = 208;                     // elements per window row (69*3 = 207, padded)
constexpr int WS_PITCH = 168;                      // weight row pitch in smem (conflict-free ldmatrix)
constexpr int THREADS = 256;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ptx::smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ptx::smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(const __nv_bfloat16* p0, const __nv_bfloat16* p1) {
  return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(p0)) |
         (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(p1)) << 16);
}
// window offset of patch element k = (r*7 + s)*3 + ci relative to the window position of the output pixel
__device__ __forceinline__ int patch_off(int k) {
  if (k >= KREAL) return 0;       // padded K: the weight is zero, any finite operand will do
  const int r = k / (KS * CI);
  return r * WIN_PITCH + (k - r * KS * CI);
}

// zero-padded input window of output tile (oh0, ow0) of image n -> win[WIN_H][WIN_PITCH]
__device__ __forceinline__ void load_window(__nv_bfloat16* win, const __nv_bfloat16* __restrict__ x, int n, int H,
                                            int W, int oh0, int ow0) {
  const int ih0 = 2 * oh0 - 3, iw0 = 2 * ow0 - 3;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  for (int i = threadIdx.x; i < WIN_H * WIN_W * CI; i += THREADS) {
    const int r = i / (WIN_W * CI), e = i - r * (WIN_W * CI);
    const int ih = ih0 + r, iw = iw0 + e / CI;
    __nv_bfloat16 v = zero;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W)
      v = x[(static_cast<size_t>(n) * H + ih) * W * CI + static_cast<size_t>(iw) * CI + (e % CI)];
    win[r * WIN_PITCH + e] = v;
  }
}

struct Tiles {
  int tiles_x, tiles_y, total;
  __device__ __host__ Tiles(int N, int OH, int OW)
      : tiles_x((OW + TW - 1) / TW), tiles_y((OH + TH - 1) / TH), total(N * tiles_x * tiles_y) {}
  __device__ void decode(int t, int& n, int& oh0, int& ow0) const {
    n = t / (tiles_x * tiles_y);
    const int r = t - n * tiles_x * tiles_y;
    oh0 = (r / tiles_x) * TH;
    ow0 = (r % tiles_x) * TW;
  }
};

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(THREADS)
stem_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, int N, int H, int W,
                int OH, int OW, __nv_bfloat16* __restrict__ y) {
  __shared__ __align__(16) __nv_bfloat16 ws[CO * WS_PITCH];             // [co][k], k >= 147 zero
  __shared__ __align__(16) __nv_bfloat16 win[WIN_H * WIN_PITCH];
  __shared__ __align__(16) __nv_bfloat16 stage[8][16 * (CO + 8)];       // per warp: 16 pixels x 64 channels
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < CO * WS_PITCH; i += THREADS) {
    const int co = i / WS_PITCH, k = i - co * WS_PITCH;
    ws[i] = k < KREAL ? w[co * KREAL + k] : __float2bfloat16(0.f);
  }
  const int g = lane >> 2, q = lane & 3;
  // this warp's m16 tile: row warp/2 of the output tile, 16 columns; thread rows g and g+8
  const int oh_l = warp >> 1, ow_l = (warp & 1) * 16 + g;
  const int base0 = (2 * oh_l) * WIN_PITCH + (2 * ow_l) * CI, base1 = base0 + 8 * 2 * CI;
  int koff[KPAD / 16][4];
#pragma unroll
  for (int ks = 0; ks < KPAD / 16; ++ks) {
    const int k0 = ks * 16 + q * 2;
    koff[ks][0] = patch_off(k0);
    koff[ks][1] = patch_off(k0 + 1);
    koff[ks][2] = patch_off(k0 + 8);
    koff[ks][3] = patch_off(k0 + 9);
  }
  const Tiles tiles(N, OH, OW);
  for (int t = blockIdx.x; t < tiles.total; t += gridDim.x) {
    int n, oh0, ow0;
    tiles.decode(t, n, oh0, ow0);
    __syncthreads();                       // previous tile's window fully consumed (also orders the weight load)
    load_window(win, x, n, H, W, oh0, ow0);
    __syncthreads();
    float acc[CO / 8][4];
#pragma unroll
    for (int i = 0; i < CO / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KPAD / 16; ++ks) {
      uint32_t a[4];
      a[0] = pack2(win + base0 + koff[ks][0], win + base0 + koff[ks][1]);
      a[1] = pack2(win + base1 + koff[ks][0], win + base1 + koff[ks][1]);
      a[2] = pack2(win + base0 + koff[ks][2], win + base0 + koff[ks][3]);
      a[3] = pack2(win + base1 + koff[ks][2], win + base1 + koff[ks][3]);
#pragma unroll
      for (int np = 0; np < CO / 16; ++np) {       // two n8 tiles per ldmatrix.x4
        uint32_t b[4];
        const int j = lane >> 3, i = lane & 7;
        ldmatrix_x4(b, ws + ((np * 2 + (j >> 1)) * 8 + i) * WS_PITCH + ks * 16 + (j & 1) * 8);
        mma_bf16(acc[np * 2], a, b[0], b[1]);
        mma_bf16(acc[np * 2 + 1], a, b[2], b[3]);
      }
    }
    // fp32 -> bf16 through the warp's staging tile, then 16-byte coalesced stores (128 B per pixel)
    __nv_bfloat16* st = stage[warp];
#pragma unroll
    for (int nt = 0; nt < CO / 8; ++nt) {
      *reinterpret_cast<__nv_bfloat162*>(st + g * (CO + 8) + nt * 8 + q * 2) = __floats2bfloat162_rn(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<__nv_bfloat162*>(st + (g + 8) * (CO + 8) + nt * 8 + q * 2) = __floats2bfloat162_rn(acc[nt][2], acc[nt][3]);
    }
    __syncwarp();
    const int oh = oh0 + oh_l;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int p = it * 4 + (lane >> 3), chunk = lane & 7;
      const int ow = ow0 + (warp & 1) * 16 + p;
      if (oh < OH && ow < OW)
        *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * OH + oh) * OW + ow) * CO + chunk * 8) =
            *reinterpret_cast<const uint4*>(st + p * (CO + 8) + chunk * 8);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
__global__ void __launch_bounds__(THREADS)
stem_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int N, int H, int W,
                  int OH, int OW, float* __restrict__ partials) {
  __shared__ __align__(16) __nv_bfloat16 win[WIN_H * WIN_PITCH];
  __shared__ __align__(16) __nv_bfloat16 dys[TH * TW * CO];            // [pixel][co], 16-byte chunks XOR-swizzled
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  constexpr int NT = KPAD / 8;                   // 20 n8 tiles of patch indices; warp w owns w, w+8, w+16
  int koff[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) koff[i] = patch_off((warp + 8 * i) * 8 + g);
  float acc[4][3][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[m][i][0] = acc[m][i][1] = acc[m][i][2] = acc[m][i][3] = 0.f;
  const Tiles tiles(N, OH, OW);
  for (int t = blockIdx.x; t < tiles.total; t += gridDim.x) {
    int n, oh0, ow0;
    tiles.decode(t, n, oh0, ow0);
    __syncthreads();
    // dY tile: 128 pixels x 8 chunks of 16 bytes; pixels outside the image are zero-filled (src-size 0)
    for (int i = tid; i < TH * TW * 8; i += THREADS) {
      const int p = i >> 3, chunk = i & 7;
      const int oh = oh0 + p / TW, ow = ow0 + p % TW;
      const bool ok = oh < OH && ow < OW;
      const __nv_bfloat16* src = dy + ((static_cast<size_t>(n) * OH + (ok ? oh : 0)) * OW + (ok ? ow : 0)) * CO + chunk * 8;
      const uint32_t dst = ptx::smem_u32(dys + p * CO + ((chunk ^ (p & 7)) * 8));
      const int bytes = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    load_window(win, x, n, H, W, oh0, ow0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 2
    for (int ks = 0; ks < TH * TW / 16; ++ks) {          // 16 pixels per step: row ks/2, columns (ks%2)*16 ..
      uint32_t a[4][4];
      const int j = lane >> 3, i = lane & 7;
      const int p = ks * 16 + (j >> 1) * 8 + i;
#pragma unroll
      for (int m = 0; m < 4; ++m) ldmatrix_x4_trans(a[m], dys + p * CO + (((m * 2 + (j & 1)) ^ (p & 7)) * 8));
      const int pbase = (2 * (ks >> 1)) * WIN_PITCH + (2 * ((ks & 1) * 16 + q * 2)) * CI;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        if (warp + 8 * nt >= NT) break;
        const __nv_bfloat16* b = win + pbase + koff[nt];
        const uint32_t b0 = pack2(b, b + 2 * CI), b1 = pack2(b + 16 * CI, b + 18 * CI);
#pragma unroll
        for (int m = 0; m < 4; ++m) mma_bf16(acc[m][nt], a[m], b0, b1);
      }
    }
  }
  float* out = partials + static_cast<size_t>(blockIdx.x) * CO * KPAD;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      if (warp + 8 * nt >= NT) break;
      const int k = (warp + 8 * nt) * 8 + q * 2;
      *reinterpret_cast<float2*>(out + (m * 16 + g) * KPAD + k) = make_float2(acc[m][nt][0], acc[m][nt][1]);
      *reinterpret_cast<float2*>(out + (m * 16 + g + 8) * KPAD + k) = make_float2(acc[m][nt][2], acc[m][nt][3]);
    }
}

int stem_grid(int N, int OH, int OW, int per_sm) {
  const Tiles t(N, OH, OW);
  const int want = u2b_num_sms() * per_sm;
  return t.total < want ? t.total : want;
}

}  // namespace

extern "C" {

int u2b_stem_conv_supported(int Cin, int Cout, int R, int S, int stride, int pad) {
  return Cin == CI && Cout == CO && R == KS && S == KS && stride == 2 && pad == 3;
}

int u2b_stem_conv_fwd(const void* x, int64_t N, int H, int W, const void* w, void* y, cudaStream_t stream) {
  if (N == 0) return 0;
  U2B_CHECK_ARG(x && w && y && N > 0 && H > 0 && W > 0, "stem_conv_fwd: bad arguments");
  U2B_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "stem_conv_fwd: y must be 16-byte aligned");
  const int OH = (H + 6 - KS) / 2 + 1, OW = (W + 6 - KS) / 2 + 1;
  stem_fwd_kernel<<<stem_grid((int)N, OH, OW, 4), THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(w), (int)N, H, W, OH, OW,
      static_cast<__nv_bfloat16*>(y));
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_stem_conv_wgrad_num_partials(int64_t N, int H, int W) {
  const int OH = (H + 6 - KS) / 2 + 1, OW = (W + 6 - KS) / 2 + 1;
  return stem_grid((int)N, OH, OW, 2);
}

int u2b_stem_conv_wgrad(const void* x, const void* dy, int64_t N, int H, int W, float* partials, cudaStream_t stream) {
  if (N == 0) return 0;
  U2B_CHECK_ARG(x && dy && partials && N > 0 && H > 0 && W > 0, "stem_conv_wgrad: bad arguments");
  U2B_CHECK_ARG((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "stem_conv_wgrad: dy must be 16-byte aligned");
  const int OH = (H + 6 - KS) / 2 + 1, OW = (W + 6 - KS) / 2 + 1;
  stem_wgrad_kernel<<<stem_grid((int)N, OH, OW, 2), THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), (int)N, H, W, OH, OW, partials);
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
