// ROUND-2 DRAFT (never run on hardware): weight gradient of the NHWC convolutions on tcgen05 tensor cores.
//
//   dW[co, r, s, ci] = sum over output pixels (n, oh, ow) of dY[n, oh, ow, co] * X[n, oh*stride + r - pad, ow*stride + s - pad, ci]
//
// GEMM view per filter tap (r, s): D[Cout, Cin] = dY^T [Cout x P] * X_shift [P x Cin], GEMM-K = output pixels. Both
// operands have K as the slow (row) axis of an NHWC tensor, i.e. they are MN-major UMMA operands, and the TMA boxes the
// forward kernel uses ([pixels][64 channels], 128-byte rows, SWIZZLE_128B; the tap's shifted box with OOB zero fill and
// element strides for stride 2) ARE the canonical MN-major SW128 layout (cute/atom/mma_traits_sm100.hpp:
// ((T,8,m),(8,k)):((1,T,LBO),(8T,SBO)), T = 8): 64 channels = one swizzle row, 8 pixels = one 1024-byte atom
// (SBO = 1024 B), the next 64-channel block LBO bytes further, one K=16 MMA = 16 pixels = +2048 B of start address.
//
// Work item = (Cout tile of 128, Cin tile of BN, filter tap, K split): the CTA streams its share of the pixel blocks
// (64 pixels per pipeline stage: 2 dY boxes + BN/64 X boxes) through an mbarrier ring, accumulates 128 x BN fp32 in
// TMEM and writes its partial to partials[ksplit][co][r][s][ci]; the caller sums the K-split partials (fixed order).
// Warp roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

constexpr int WG_BM = 128;           // Cout rows of the accumulator (TMEM lanes)
constexpr int WG_PB = 64;            // pixels per pipeline stage (GEMM-K per stage)
constexpr int WG_BOX_BYTES = WG_PB * 128;   // one [64 pixels][64 channels] box: 8 KB
constexpr int WG_THREADS = 256;

struct WgradParams {
  int N, H, W, Cin, Cout, R, S, stride, pad, OH, OW;
  int BW, BH, tiles_w, tiles_h, tiles_p;    // pixel blocks of BW x BH = 64 output pixels
  int co_tiles, ci_tiles, ksplit, num_work;
  float* partials;                          // (ksplit, Cout, R, S, Cin) fp32
};

template <int BN>
struct WgCfg {
  static constexpr int A_BYTES = 2 * WG_BOX_BYTES;                 // Cout 128 = 2 boxes
  static constexpr int B_BYTES = (BN / 64) * WG_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * STAGES + 2) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;             // power of two >= 32
};

// MN-major, 128B swizzle: LBO = byte distance between 64-element MN blocks, SBO = 1024 (8 K-rows of 128 B)
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor with BOTH operands MN-major (bits 15 and 16)
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(uint32_t M, uint32_t N, uint32_t ab_fmt) {
  return ptx::umma_idesc_f16(M, N, ab_fmt) | (1u << 15) | (1u << 16);
}

template <int BN, bool BF16>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                     const WgradParams p) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* T_full = empty + Cfg::STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(T_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_dy);
    ptx::prefetch_tmap(&tmap_x);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&T_full[0], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work item of this CTA (one per CTA: grid = num_work)
  int wk = blockIdx.x;
  const int ks = wk % p.ksplit; wk /= p.ksplit;
  const int tap = wk % (p.R * p.S); wk /= p.R * p.S;
  const int ci_t = wk % p.ci_tiles;
  const int co_t = wk / p.ci_tiles;
  const int r = tap / p.S, s = tap % p.S;
  const int per = (p.tiles_p + p.ksplit - 1) / p.ksplit;
  const int pb0 = ks * per, pb1 = min(p.tiles_p, pb0 + per);
  const int nblocks = max(pb1 - pb0, 0);

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int pb = pb0; pb < pb1; ++pb) {
        const int owb = pb % p.tiles_w, ohb = (pb / p.tiles_w) % p.tiles_h, n = pb / (p.tiles_w * p.tiles_h);
        const int ow0 = owb * p.BW, oh0 = ohb * p.BH;
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 2; ++j)          // dY: 128 output channels = 2 boxes of 64
          ptx::tma_load_4d(sa + j * WG_BOX_BYTES, &tmap_dy, &full[stage], co_t * WG_BM + j * 64, ow0, oh0, n);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)    // X at the tap's shifted origin (zero fill = padding)
          ptx::tma_load_4d(sa + Cfg::A_BYTES + j * WG_BOX_BYTES, &tmap_x, &full[stage], ci_t * BN + j * 64,
                           ow0 * p.stride + s - p.pad, oh0 * p.stride + r - p.pad, n);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = umma_idesc_f16_mn(WG_BM, BN, BF16 ? 1u : 0u);
      const uint32_t sbase = ptx::smem_u32(smem);
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < nblocks; ++it) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        const uint64_t a_desc = umma_desc_sw128_mn(sbase + stage * Cfg::STAGE_BYTES, WG_BOX_BYTES);
        const uint64_t b_desc = umma_desc_sw128_mn(sbase + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES, WG_BOX_BYTES);
#pragma unroll
        for (int k = 0; k < WG_PB / 16; ++k)       // 16 pixels per MMA: +2048 B = +128 in descriptor units
          ptx::umma_f16(tmem_base, a_desc + 128 * k, b_desc + 128 * k, idesc, (it | k) != 0);
        ptx::umma_commit(&empty[stage]);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      ptx::umma_commit(&T_full[0]);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int co = co_t * WG_BM + q * 32 + lane;             // TMEM lane = accumulator row = output channel
    float* orow = p.partials + ((static_cast<size_t>(ks) * p.Cout + co) * (p.R * p.S) + tap) * p.Cin + ci_t * BN;
    if (nblocks > 0) {
      ptx::mbar_wait(&T_full[0], 0);
      ptx::tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      uint32_t v[16];
      if (nblocks > 0) {
        ptx::tmem_ld16(taddr + c * 16, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;              // empty K range: the partial is zero
      }
      if (co < p.Cout) {
        float4* o = reinterpret_cast<float4*>(orow + c * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, bool BF16>
int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgCfg<BN>;
  static bool attr = false;
  if (!attr) {
    U2B_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    attr = true;
  }
  conv_wgrad_tc_kernel<BN, BF16><<<p.num_work, WG_THREADS, Cfg::SMEM_BYTES, stream>>>(tdy, tx, p);
  U2B_LAUNCH_CHECK();
  return 0;
}

int wgrad_plan(WgradParams& p, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.stride = stride; p.pad = pad;
  p.OH = (H + 2 * pad - R) / stride + 1;
  p.OW = (W + 2 * pad - S) / stride + 1;
  long long best = -1;
  for (int bw = 64; bw >= 8; bw >>= 1) {       // 64-pixel block with the least padding
    const int bh = 64 / bw;
    const long long cost = static_cast<long long>((p.OW + bw - 1) / bw) * bw * ((p.OH + bh - 1) / bh) * bh;
    if (best < 0 || cost < best) {
      best = cost;
      p.BW = bw;
      p.BH = bh;
    }
  }
  p.tiles_w = (p.OW + p.BW - 1) / p.BW;
  p.tiles_h = (p.OH + p.BH - 1) / p.BH;
  p.tiles_p = p.tiles_w * p.tiles_h * N;
  const int BN = (Cin % 256 == 0) ? 256 : (Cin % 128 == 0 ? 128 : 64);
  p.co_tiles = Cout / WG_BM;
  p.ci_tiles = Cin / BN;
  const int base = p.co_tiles * p.ci_tiles * R * S;
  int ks = (u2b_num_sms() + base - 1) / base;           // fill the machine once
  if (ks > p.tiles_p) ks = p.tiles_p;
  if (ks < 1) ks = 1;
  p.ksplit = ks;
  p.num_work = base * ks;
  return BN;
}

}  // namespace

extern "C" {

int u2b_conv2d_wgrad_supported(int Cin, int Cout, int R, int S, int stride, int pad) {
  if (Cin <= 0 || Cin % 64 != 0 || Cout <= 0 || Cout % 128 != 0) return 0;
  if (!((R == 1 && S == 1 && pad == 0) || (R == 3 && S == 3 && pad == 1))) return 0;
  return stride == 1 || stride == 2;
}

// number of K-split partials the kernel writes for this problem: partials is (ksplit, Cout, R, S, Cin) fp32
int u2b_conv2d_wgrad_ksplit(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  WgradParams p;
  wgrad_plan(p, N, H, W, Cin, Cout, R, S, stride, pad);
  return p.ksplit;
}

// dtype: 1 = fp16, 2 = bf16. x (N,H,W,Cin), dy (N,OH,OW,Cout) NHWC. dW = sum over the ksplit partials.
int u2b_conv2d_nhwc_wgrad(int dtype, const void* x, const void* dy, int N, int H, int W, int Cin, int Cout, int R, int S,
                          int stride, int pad, float* partials, cudaStream_t stream) {
  U2B_CHECK_ARG(x && dy && partials && N > 0 && H > 0 && W > 0, "conv2d_nhwc_wgrad: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "conv2d_nhwc_wgrad: dtype must be fp16(1) or bf16(2)");
  if (!u2b_conv2d_wgrad_supported(Cin, Cout, R, S, stride, pad)) {
    u2b_set_error("conv2d_nhwc_wgrad: unsupported shape Cin=%d Cout=%d k=%dx%d stride=%d pad=%d", Cin, Cout, R, S, stride, pad);
    return U2B_ERR_UNSUPPORTED;
  }
  WgradParams p;
  const int BN = wgrad_plan(p, N, H, W, Cin, Cout, R, S, stride, pad);
  p.partials = partials;
  const CUtensorMapDataType tdt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tdy, tx;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Cout * 2, (uint64_t)p.OW * Cout * 2, (uint64_t)p.OH * p.OW * Cout * 2};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    int rc = u2b_encode_tmap(&tdy, tdt, 4, dy, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {64, (uint32_t)(p.BW * stride), (uint32_t)(p.BH * stride), 1};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = u2b_encode_tmap(&tx, tdt, 4, x, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  const bool bf = dtype == 2;
  if (BN == 256) return bf ? launch_wgrad<256, true>(tdy, tx, p, stream) : launch_wgrad<256, false>(tdy, tx, p, stream);
  if (BN == 128) return bf ? launch_wgrad<128, true>(tdy, tx, p, stream) : launch_wgrad<128, false>(tdy, tx, p, stream);
  return bf ? launch_wgrad<64, true>(tdy, tx, p, stream) : launch_wgrad<64, false>(tdy, tx, p, stream);
}

}  // extern "C"
