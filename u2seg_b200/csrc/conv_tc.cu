// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC fp16/bf16, fp32 accumulate in TMEM.
// Covers the conv layers of the u2seg hot path (detectron2/layers/wrappers.py:127 F.conv2d call sites:
// ResNet bottlenecks resnet.py:149-176, FPN lateral/output fpn.py:77-88, RPN head rpn.py:116-134, mask
// head mask_head.py:242-262, sem-seg head semantic_seg.py:196-214) and, viewed as a 1x1 conv over a
// (1,1,M,K) "image", the Linear layers (box_head.py:70, fast_rcnn.py:236-239).
//
// GEMM view: M = N*OH*OW output pixels, N = Cout, K = R*S*Cin.
//   A tile (128 pixels x 64 channels) = one TMA 4-D box {64, BW*stride, BH*stride, 1} of the NHWC input at the
//   filter tap's shifted origin, element strides {1,stride,stride,1}; out-of-bounds pixels (zero padding, ragged
//   edges) are zero-filled by TMA. BW*BH = 128, so the box lands in shared memory as 128 rows x 128 B with the
//   128B swizzle = the K-major UMMA operand layout: im2col is never materialised.
//   B tile (BN x 64) = TMA 2-D box of the (Cout, R*S*Cin) packed filter.
//   D: 128 x BN fp32 accumulator in TMEM, double buffered; epilogue warps convert (+bias, +residual, ReLU)
//   and store NHWC while the next tile's MMAs run.
// Warp roles: warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator, warps 4-7 epilogue.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;  // 16384
constexpr int CONV_THREADS = 256;

struct ConvParams {
  int N, H, W, Cin, Cout, R, S, stride, pad, OH, OW;
  int BW, BH, tiles_w, tiles_h, tiles_m, tiles_n, num_tiles, kblocks_c;
  int m_groups, num_work;  // cluster scheduling: work = (m_group, n_tile), a CTA owns m_tile = m_group*CL + rank
  int relu, is_bf16;
  const float* bias;
  const void* residual;
  void* out;
};

template <int BN>
struct ConvCfg {
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
};

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if (BF16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}

// CL = cluster size (1, 2, 4). With CL > 1 the CTAs of a cluster work on CL neighbouring M tiles with the same
// filter tile sequence: each loads 1/CL of every B tile and TMA-multicasts it to all of them, so the filter is read
// from L2 once per cluster; a smem stage is reusable only when every CTA of the cluster has consumed it
// (tcgen05.commit multicast onto all `empty` barriers, which count CL arrivals).
template <int BN, bool BF16, int CL>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* T_full = empty + Cfg::STAGES;
  uint64_t* T_empty = T_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(T_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], CL);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&T_full[i], 1);
      ptx::mbar_init(&T_empty[i], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();  // peers' barriers must be initialised before any remote arrive / multicast
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int kblocks = p.R * p.S * p.kblocks_c;
  const int rank = CL > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);
  constexpr int PART_ROWS = BN / CL;

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int work = cluster_id; work < p.num_work; work += num_clusters) {
        const int tn = work % p.tiles_n;
        int tm = (work / p.tiles_n) * CL + rank;
        if (tm >= p.tiles_m) tm = p.tiles_m - 1;  // padding CTA of the last group: loads valid data, stores nothing
        const int owb = tm % p.tiles_w, ohb = (tm / p.tiles_w) % p.tiles_h, n = tm / (p.tiles_w * p.tiles_h);
        const int x_base = owb * p.BW * p.stride - p.pad, y_base = ohb * p.BH * p.stride - p.pad;
        for (int r = 0; r < p.R; ++r)
          for (int s = 0; s < p.S; ++s)
            for (int cb = 0; cb < p.kblocks_c; ++cb) {
              ptx::mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
              ptx::tma_load_4d(sa, &tmap_x, &full[stage], cb * BK, x_base + s, y_base + r, n);
              const int kcol = ((r * p.S + s) * p.kblocks_c + cb) * BK;
              if (CL == 1)
                ptx::tma_load_2d(sa + A_BYTES, &tmap_w, &full[stage], kcol, tn * BN);
              else
                ptx::tma_load_2d_mc(sa + A_BYTES + rank * (PART_ROWS * BK * 2), &tmap_w, &full[stage], kcol,
                                    tn * BN + rank * PART_ROWS, kMask);
              if (++stage == Cfg::STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_f16(BM, BN, BF16 ? 1u : 0u);
      const uint32_t sbase = ptx::smem_u32(smem);
      uint32_t stage = 0, phase = 0, acc_it = 0;
      for (int work = cluster_id; work < p.num_work; work += num_clusters, ++acc_it) {
        const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
        ptx::mbar_wait(&T_empty[buf], tphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::umma_desc_sw128(sbase + stage * Cfg::STAGE_BYTES);
          const uint64_t b_desc = ptx::umma_desc_sw128(sbase + stage * Cfg::STAGE_BYTES + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            ptx::umma_f16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          if (CL == 1)
            ptx::umma_commit(&empty[stage]);
          else
            ptx::umma_commit_mc(&empty[stage], kMask);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::umma_commit(&T_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int bh = row / p.BW, bw = row % p.BW;
    uint32_t acc_it = 0;
    for (int work = cluster_id; work < p.num_work; work += num_clusters, ++acc_it) {
      const int tn = work % p.tiles_n, tm_raw = (work / p.tiles_n) * CL + rank;
      const int tm = tm_raw < p.tiles_m ? tm_raw : p.tiles_m - 1;
      const int owb = tm % p.tiles_w, ohb = (tm / p.tiles_w) % p.tiles_h, n = tm / (p.tiles_w * p.tiles_h);
      const int oh = ohb * p.BH + bh, ow = owb * p.BW + bw;
      const bool valid = oh < p.OH && ow < p.OW && tm_raw < p.tiles_m;
      const size_t pix = (static_cast<size_t>(n) * p.OH + oh) * p.OW + ow;
      uint16_t* orow = static_cast<uint16_t*>(p.out) + pix * p.Cout + tn * BN;
      const uint16_t* rrow =
          p.residual ? static_cast<const uint16_t*>(p.residual) + pix * p.Cout + tn * BN : nullptr;
      const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
      ptx::mbar_wait(&T_full[buf], tphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c * 16, v);
        ptx::tmem_ld_wait();
        if (valid) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + tn * BN + c * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b = b4[j];
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (rrow) {
            const uint4* r4 = reinterpret_cast<const uint4*>(rrow + c * 16);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint4 r = r4[j];
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 x = unpack2<BF16>(rr[t]);
                f[8 * j + 2 * t] += x.x;
                f[8 * j + 2 * t + 1] += x.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          uint4 o0, o1;
          o0.x = pack2<BF16>(f[0], f[1]); o0.y = pack2<BF16>(f[2], f[3]);
          o0.z = pack2<BF16>(f[4], f[5]); o0.w = pack2<BF16>(f[6], f[7]);
          o1.x = pack2<BF16>(f[8], f[9]); o1.y = pack2<BF16>(f[10], f[11]);
          o1.z = pack2<BF16>(f[12], f[13]); o1.w = pack2<BF16>(f[14], f[15]);
          uint4* o = reinterpret_cast<uint4*>(orow + c * 16);
          o[0] = o0;
          o[1] = o1;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&T_empty[buf]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();  // no CTA may exit while a peer can still multicast into / arrive on its smem
  if (warp == 2) ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, bool BF16, int CL>
int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr = false;
  if (!attr) {
    U2B_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, BF16, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    attr = true;
  }
  int clusters = u2b_num_sms() / CL;
  if (clusters > p.num_work) clusters = p.num_work;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(CONV_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  U2B_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, BF16, CL>, tx, tw, p));
  return 0;
}

int g_conv_cluster = 2;  // default cluster size (u2b_conv2d_set_cluster)

}  // namespace

extern "C" {

// 1 if (shape) is handled by the tcgen05 kernel, else 0 (caller keeps the library convolution).
int u2b_conv2d_supported(int Cin, int Cout, int R, int S, int stride, int pad) {
  if (Cin <= 0 || Cin % 64 != 0 || Cout <= 0 || Cout % 64 != 0) return 0;
  if (!((R == 1 && S == 1 && pad == 0) || (R == 3 && S == 3 && pad == 1))) return 0;
  if (stride != 1 && stride != 2) return 0;
  return 1;
}

// dtype: 1 = fp16, 2 = bf16. x: (N,H,W,Cin) NHWC. w: (Cout,R,S,Cin) ("OHWI"). out: (N,OH,OW,Cout) NHWC,
// OH = (H + 2*pad - R)/stride + 1. bias: Cout fp32 or NULL. residual: same shape/dtype as out or NULL.
int u2b_conv2d_nhwc_fwd(int dtype, const void* x, int N, int H, int W, int Cin, const void* w, int Cout,
                        int R, int S, int stride, int pad, const float* bias, const void* residual, int relu,
                        void* out, cudaStream_t stream) {
  U2B_CHECK_ARG(x && w && out && N > 0 && H > 0 && W > 0, "conv2d_nhwc_fwd: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "conv2d_nhwc_fwd: dtype must be fp16(1) or bf16(2)");
  if (!u2b_conv2d_supported(Cin, Cout, R, S, stride, pad)) {
    u2b_set_error("conv2d_nhwc_fwd: unsupported shape Cin=%d Cout=%d k=%dx%d stride=%d pad=%d", Cin, Cout, R, S,
                  stride, pad);
    return U2B_ERR_UNSUPPORTED;
  }
  ConvParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.stride = stride; p.pad = pad;
  p.OH = (H + 2 * pad - R) / stride + 1;
  p.OW = (W + 2 * pad - S) / stride + 1;
  // pick the 128-pixel tile shape with the least padding
  long long best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    const int bh = 128 / bw;
    const long long cost = static_cast<long long>((p.OW + bw - 1) / bw) * bw * ((p.OH + bh - 1) / bh) * bh;
    if (best < 0 || cost < best) {
      best = cost;
      p.BW = bw;
      p.BH = bh;
    }
  }
  const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
  p.tiles_w = (p.OW + p.BW - 1) / p.BW;
  p.tiles_h = (p.OH + p.BH - 1) / p.BH;
  p.tiles_m = p.tiles_w * p.tiles_h * N;
  p.tiles_n = Cout / BN;
  p.num_tiles = p.tiles_m * p.tiles_n;
  p.kblocks_c = Cin / BK;
  int CL = g_conv_cluster;
  if (p.tiles_m < CL) CL = 1;
  p.m_groups = (p.tiles_m + CL - 1) / CL;
  p.num_work = p.m_groups * p.tiles_n;
  p.relu = relu; p.is_bf16 = dtype == 2;
  p.bias = bias; p.residual = residual; p.out = out;
  const CUtensorMapDataType tdt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {BK, (uint32_t)(p.BW * stride), (uint32_t)(p.BH * stride), 1};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = u2b_encode_tmap(&tx, tdt, 4, x, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)R * S * Cin, (uint64_t)Cout};
    uint64_t strides[1] = {(uint64_t)R * S * Cin * 2};
    uint32_t box[2] = {BK, (uint32_t)(BN / CL)};
    int rc = u2b_encode_tmap(&tw, tdt, 2, w, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  const bool bf = dtype == 2;
#define U2B_DISPATCH(BN_)                                                                              \
  if (BN == BN_) {                                                                                     \
    if (CL == 4) return bf ? launch_conv<BN_, true, 4>(tx, tw, p, stream) : launch_conv<BN_, false, 4>(tx, tw, p, stream); \
    if (CL == 2) return bf ? launch_conv<BN_, true, 2>(tx, tw, p, stream) : launch_conv<BN_, false, 2>(tx, tw, p, stream); \
    return bf ? launch_conv<BN_, true, 1>(tx, tw, p, stream) : launch_conv<BN_, false, 1>(tx, tw, p, stream);             \
  }
  U2B_DISPATCH(256)
  U2B_DISPATCH(128)
  U2B_DISPATCH(64)
#undef U2B_DISPATCH
  return U2B_ERR_UNSUPPORTED;
}

// cluster size used by the convolution kernel: 1 (no multicast), 2 or 4
int u2b_conv2d_set_cluster(int cluster) {
  U2B_CHECK_ARG(cluster == 1 || cluster == 2 || cluster == 4, "conv2d_set_cluster: 1, 2 or 4");
  g_conv_cluster = cluster;
  return 0;
}

}  // extern "C"
