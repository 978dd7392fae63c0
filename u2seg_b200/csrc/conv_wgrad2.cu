// Weight gradient of the NHWC convolutions / Linear layers on tcgen05 tensor cores, 2-CTA (cta_group::2) version.
// Reference call sites: the backward of every F.conv2d at detectron2/layers/wrappers.py:127 and of nn.Linear at
// roi_heads/box_head.py:70 (autograd's convolution_backward / addmm backward in the reference).
//
//   dW[co, r, s, ci] = sum over output pixels (n, oh, ow) of dY[n, oh, ow, co] * X[n, oh*stride + r - pad, ow*stride + s - pad, ci]
//
// Per filter tap (r, s) this is a GEMM whose K dimension is the PIXEL axis: D[Ca, Cb] = A^T [Ca x P] * B [P x Cb] with
// (A, B) = (dY, X shifted by the tap) - or (X, dY) when only Cin is a multiple of 256 ("swapped": D = dW^T). Both
// operands have K as the slow axis of an NHWC tensor, i.e. they are MN-major UMMA operands, and the TMA boxes of the
// forward kernel ([pixels][64 channels], 128-byte rows, SWIZZLE_128B, the tap's shifted box with out-of-bounds zero
// fill and element strides for stride 2) ARE the canonical MN-major SW128 layout: 64 channels = one swizzle row,
// 8 pixels = one 1024-byte atom (SBO), the next 64-channel block LBO bytes further, one K=16 MMA = 16 pixels.
//
// A cluster of two CTAs (one SM pair) owns a 256 (Ca) x BN (Cb) tile of one tap and one slice of the pixel range
// (split-K): each CTA stages its own 128 A channels and HALF of the B channels per 64-pixel stage (32 KB at BN = 256,
// the same bytes per MMA as the forward kernel), the leader issues tcgen05.mma.cta_group::2 (M = 256, N = BN, K = 16),
// and each CTA's epilogue writes its 128 x BN fp32 block of partials[ksplit][tap][Ca][Cb]; a second kernel sums the
// K-split partials in a fixed order and writes dW (OHWI) - deterministic, no atomics. Without a K split (enough tiles
// to fill the machine, e.g. the 12544 -> 1024 Linear) the epilogue writes dW itself. Persistent: a pair loops over its
// work items, so TMEM allocation / barrier setup / descriptor prefetch are paid once per launch.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (leader), 2 = TMEM allocator, 4-7 = epilogue.
#include <cuda_bf16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

constexpr int PB = 64;                 // pixels per pipeline stage (GEMM-K per stage)
constexpr int BOX_BYTES = PB * 128;    // one [64 pixels][64 channels] box: 8 KB
constexpr int A_BYTES = 2 * BOX_BYTES; // 128 A channels per CTA
constexpr int W2_THREADS = 256;

struct Wg2Params {
  int N, H, W, OH, OW, R, S, stride, pad;
  int Ca, Cb, a_is_dy;
  int BW, BH, tiles_w, tiles_h, tiles_p;  // pixel blocks of BW x BH = 64 output pixels
  int m_tiles, n_tiles, ksplit, num_work;
  float* partials;  // (ksplit, R*S, Ca, Cb) fp32; with ksplit == 1 and direct_out: unused
  void* dw;         // ksplit == 1: the epilogue writes dW (Cout,R,S,Cin) itself, no reduction pass
  int direct_out;   // 0 = partials, 1 = fp32 dW, 2 = fp16, 3 = bf16
};

template <int BN>
struct Wg2Cfg {
  static constexpr int B_BYTES = (BN / 128) * BOX_BYTES;  // this CTA's half of the B channels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 6 : 8;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * STAGES + 2) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = BN;  // 128 or 256: powers of two >= 32
};

__device__ __forceinline__ uint32_t w2_mapa(uint32_t cta_addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(cta_addr), "r"(rank));
  return out;
}
__device__ __forceinline__ bool w2_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(ptx::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void w2_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (w2_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!w2_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > U2B_MBAR_TIMEOUT_CYCLES) {
      printf("u2b wgrad2: mbarrier timeout block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x,
             ptx::smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void w2_tma_load_4d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                               int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void w2_umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void w2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// MN-major, 128B swizzle: LBO = byte distance between 64-element MN blocks, SBO = 1024 (8 K-rows of 128 B)
__device__ __forceinline__ uint64_t w2_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, bool BF16>
__global__ void __launch_bounds__(W2_THREADS, 1)
conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const Wg2Params p) {
  using Cfg = Wg2Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* T_full = empty + Cfg::STAGES;
  uint64_t* T_empty = T_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(T_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  ptx_free::pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    ptx::mbar_init(&T_full[0], 1);
    ptx::mbar_init(&T_empty[0], 8);  // 4 epilogue warps x 2 CTAs (the leader's copy is the one waited on)
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(tmem_ptr)),
                 "r"(static_cast<uint32_t>(Cfg::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  ptx_free::pdl_wait();
  const uint32_t tmem_base = *tmem_ptr;

  // persistent: cluster c takes work items c, c + #clusters, ...; work = (m tile, n tile, tap, k split)
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int per = (p.tiles_p + p.ksplit - 1) / p.ksplit;
  const int a_mul = p.a_is_dy ? 1 : p.stride, b_mul = p.a_is_dy ? p.stride : 1;
  struct Work {
    int ks, tap, n_t, m_t, r, s, pb0, pb1;
  };
  auto decode = [&](int wk) {
    Work w;
    w.ks = wk % p.ksplit; wk /= p.ksplit;
    w.tap = wk % (p.R * p.S); wk /= p.R * p.S;
    w.n_t = wk % p.n_tiles;
    w.m_t = wk / p.n_tiles;
    w.r = w.tap / p.S; w.s = w.tap % p.S;
    w.pb0 = w.ks * per;
    w.pb1 = min(p.tiles_p, w.pb0 + per);
    if (w.pb1 < w.pb0) w.pb1 = w.pb0;
    return w;
  };

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int wk = cluster_id; wk < p.num_work; wk += num_clusters) {
        const Work w = decode(wk);
        const int a_c0 = w.m_t * 256 + static_cast<int>(rank) * 128;
        const int b_c0 = w.n_t * BN + static_cast<int>(rank) * (BN / 2);
        const int a_dx = p.a_is_dy ? 0 : w.s - p.pad, a_dy = p.a_is_dy ? 0 : w.r - p.pad;
        const int b_dx = p.a_is_dy ? w.s - p.pad : 0, b_dy = p.a_is_dy ? w.r - p.pad : 0;
        for (int pb = w.pb0; pb < w.pb1; ++pb) {
          const int owb = pb % p.tiles_w, ohb = (pb / p.tiles_w) % p.tiles_h, n = pb / (p.tiles_w * p.tiles_h);
          const int ow0 = owb * p.BW, oh0 = ohb * p.BH;
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          const uint32_t full_leader = w2_mapa(ptx::smem_u32(&full[stage]), 0);
          if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            w2_tma_load_4d(sa + j * BOX_BYTES, &tmap_a, full_leader, a_c0 + j * 64, ow0 * a_mul + a_dx, oh0 * a_mul + a_dy, n);
#pragma unroll
          for (int j = 0; j < BN / 128; ++j)
            w2_tma_load_4d(sa + A_BYTES + j * BOX_BYTES, &tmap_b, full_leader, b_c0 + j * 64, ow0 * b_mul + b_dx,
                           oh0 * b_mul + b_dy, n);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_f16(256, BN, BF16 ? 1u : 0u) | (1u << 15) | (1u << 16);  // A, B MN-major
      const uint32_t sbase = ptx::smem_u32(smem);
      uint32_t stage = 0, phase = 0, it_w = 0;
      for (int wk = cluster_id; wk < p.num_work; wk += num_clusters, ++it_w) {
        const Work w = decode(wk);
        const int nblocks = w.pb1 - w.pb0;
        w2_wait_cluster(&T_empty[0], (it_w & 1) ^ 1);  // both epilogues have drained the accumulator of the previous work
        ptx::tc_fence_after();
        for (int it = 0; it < nblocks; ++it) {
          w2_wait_cluster(&full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = w2_desc_mn(sbase + stage * Cfg::STAGE_BYTES, BOX_BYTES);
          const uint64_t b_desc = w2_desc_mn(sbase + stage * Cfg::STAGE_BYTES + A_BYTES, BOX_BYTES);
#pragma unroll
          for (int k = 0; k < PB / 16; ++k)  // 16 pixels per MMA: +2048 B = +128 in descriptor units
            w2_umma(tmem_base, a_desc + 128 * k, b_desc + 128 * k, idesc, (it | k) != 0);
          w2_commit(&empty[stage]);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        w2_commit(&T_full[0]);  // with nblocks == 0 nothing is pending: arrives at once
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_empty_leader = w2_mapa(ptx::smem_u32(&T_empty[0]), 0);
    uint32_t it_w = 0;
    for (int wk = cluster_id; wk < p.num_work; wk += num_clusters, ++it_w) {
      const Work w = decode(wk);
      const int nblocks = w.pb1 - w.pb0;
      const int row = w.m_t * 256 + static_cast<int>(rank) * 128 + q * 32 + lane;  // accumulator row = A-side channel
      ptx::mbar_wait(&T_full[0], it_w & 1);
      ptx::tc_fence_after();
      float* orow = p.partials + ((static_cast<size_t>(w.ks) * (p.R * p.S) + w.tap) * p.Ca + row) * p.Cb + w.n_t * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        if (nblocks > 0) {
          ptx::tmem_ld16(taddr + c * 32, *reinterpret_cast<uint32_t(*)[16]>(v));
          ptx::tmem_ld16(taddr + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(v + 16));
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;  // empty K range: the partial is zero
        }
        if (!p.direct_out) {
          float4* o = reinterpret_cast<float4*>(orow + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
        } else if (p.a_is_dy) {
          // dW[co = row][tap][ci = n_t*BN + c*32 + j]: 32 consecutive elements of one filter row
          const size_t off = (static_cast<size_t>(row) * (p.R * p.S) + w.tap) * p.Cb + w.n_t * BN + c * 32;
          if (p.direct_out == 1) {
            float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.dw) + off);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
          } else {
            uint4* o = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.dw) + off);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float a = __uint_as_float(v[8 * j + 2 * t]), b = __uint_as_float(v[8 * j + 2 * t + 1]);
                if (p.direct_out == 3) {
                  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                  pk[t] = *reinterpret_cast<uint32_t*>(&h);
                } else {
                  __half2 h = __floats2half2_rn(a, b);
                  pk[t] = *reinterpret_cast<uint32_t*>(&h);
                }
              }
              o[j] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        } else {
          // swapped: row = ci, columns = co: dW[co][tap][ci] is a strided scatter (small tensors only take this path)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int co = w.n_t * BN + c * 32 + j;
            const size_t off = (static_cast<size_t>(co) * (p.R * p.S) + w.tap) * p.Ca + row;
            const float a = __uint_as_float(v[j]);
            if (p.direct_out == 1) static_cast<float*>(p.dw)[off] = a;
            else if (p.direct_out == 3) static_cast<__nv_bfloat16*>(p.dw)[off] = __float2bfloat16_rn(a);
            else static_cast<__half*>(p.dw)[off] = __float2half_rn(a);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(t_empty_leader) : "memory");
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(Cfg::TMEM_COLS))
                 : "memory");
}

// dW[co][tap][ci] = sum over k splits of partials[ks][tap][a][b], (a, b) = (co, ci) or, swapped, (ci, co)
template <typename OutT>
__global__ void __launch_bounds__(256)
wgrad2_reduce_kernel(const float* __restrict__ partials, int ksplit, int taps, int Ca, int Cb, int a_is_dy,
                     OutT* __restrict__ dw, long long total) {
  ptx_free::pdl_prologue();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes dW in OHWI order: co, tap, ci
  const int Cout = a_is_dy ? Ca : Cb, Cin = a_is_dy ? Cb : Ca;
  const int ci = static_cast<int>(i % Cin);
  const int tap = static_cast<int>((i / Cin) % taps);
  const int co = static_cast<int>(i / (static_cast<long long>(Cin) * taps));
  const size_t a = a_is_dy ? co : ci, b = a_is_dy ? ci : co;
  const size_t plane = static_cast<size_t>(taps) * Ca * Cb;
  const float* src = partials + (static_cast<size_t>(tap) * Ca + a) * Cb + b;
  float acc = 0.f;
  for (int k = 0; k < ksplit; ++k) acc += src[k * plane];
  dw[i] = static_cast<OutT>(acc);
}

template <int BN, bool BF16>
int launch_wgrad2(const CUtensorMap& ta, const CUtensorMap& tb, const Wg2Params& p, cudaStream_t stream) {
  using Cfg = Wg2Cfg<BN>;
  static bool attr = false;
  if (!attr) {
    U2B_CUDA(cudaFuncSetAttribute(conv_wgrad2_kernel<BN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  int clusters = u2b_persistent_sms() / 2;
  if (clusters > p.num_work) clusters = p.num_work;
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(W2_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = u2b_pdl_enabled() ? 2 : 1;
  U2B_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad2_kernel<BN, BF16>, ta, tb, p));
  return 0;
}

// returns BN (0 = unsupported)
int wgrad2_plan(Wg2Params& p, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  p.N = N; p.H = H; p.W = W; p.R = R; p.S = S; p.stride = stride; p.pad = pad;
  p.OH = (H + 2 * pad - R) / stride + 1;
  p.OW = (W + 2 * pad - S) / stride + 1;
  if (Cout % 256 == 0 && Cin % 128 == 0) {
    p.a_is_dy = 1; p.Ca = Cout; p.Cb = Cin;
  } else if (Cin % 256 == 0 && Cout % 128 == 0) {
    p.a_is_dy = 0; p.Ca = Cin; p.Cb = Cout;
  } else {
    return 0;
  }
  long long best = -1;
  for (int bw = 64; bw >= 8; bw >>= 1) {  // 64-pixel block with the least padding
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
  const int BN = (p.Cb % 256 == 0) ? 256 : 128;
  p.m_tiles = p.Ca / 256;
  p.n_tiles = p.Cb / BN;
  const int base = p.m_tiles * p.n_tiles * R * S;
  const int pairs = u2b_persistent_sms() / 2;
  // one wave of work items over the SM pairs when the tile count allows it (a second, mostly empty wave would double
  // the time), at least 8 pixel blocks (512 pixels) per split; more tiles than pairs: no split, the kernel is persistent
  int ks = base >= pairs ? 1 : pairs / base;
  const int max_ks = (p.tiles_p + 7) / 8;
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  p.ksplit = ks;
  p.num_work = base * ks;
  return BN;
}

}  // namespace

extern "C" {

int u2b_conv_wgrad2_supported(int Cin, int Cout, int R, int S, int stride, int pad) {
  if (Cin <= 0 || Cout <= 0) return 0;
  if (!((Cout % 256 == 0 && Cin % 128 == 0) || (Cin % 256 == 0 && Cout % 128 == 0))) return 0;
  if (R == 2 && S == 2) return pad == 0 && stride == 2;  // ConvTranspose2d(k=2, s=2) seen as the conv it is the gradient of
  if (!((R == 1 && S == 1 && pad == 0) || (R == 3 && S == 3 && pad == 1))) return 0;
  return stride == 1 || stride == 2;
}

// fp32 elements of the partial buffer the kernel needs for this problem
int64_t u2b_conv_wgrad2_workspace_floats(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  Wg2Params p;
  if (!wgrad2_plan(p, N, H, W, Cin, Cout, R, S, stride, pad)) return 0;
  return static_cast<int64_t>(p.ksplit) * R * S * Cin * Cout;
}

// dtype: 1 = fp16, 2 = bf16. x (N,H,W,Cin), dy (N,OH,OW,Cout) NHWC. dw: (Cout,R,S,Cin) OHWI, out_dtype 0 = fp32,
// 1 = fp16, 2 = bf16. workspace: u2b_conv_wgrad2_workspace_floats(...) floats.
int u2b_conv_wgrad2(int dtype, const void* x, const void* dy, int N, int H, int W, int Cin, int Cout, int R, int S,
                    int stride, int pad, float* workspace, int out_dtype, void* dw, cudaStream_t stream) {
  U2B_CHECK_ARG(x && dy && workspace && dw && N > 0 && H > 0 && W > 0, "conv_wgrad2: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "conv_wgrad2: dtype must be fp16(1) or bf16(2)");
  U2B_CHECK_ARG(out_dtype >= 0 && out_dtype <= 2, "conv_wgrad2: out_dtype must be 0, 1 or 2");
  if (!u2b_conv_wgrad2_supported(Cin, Cout, R, S, stride, pad)) {
    u2b_set_error("conv_wgrad2: unsupported shape Cin=%d Cout=%d k=%dx%d stride=%d pad=%d", Cin, Cout, R, S, stride, pad);
    return U2B_ERR_UNSUPPORTED;
  }
  Wg2Params p;
  const int BN = wgrad2_plan(p, N, H, W, Cin, Cout, R, S, stride, pad);
  p.partials = workspace;
  p.dw = dw;
  p.direct_out = p.ksplit == 1 ? (out_dtype == 0 ? 1 : (out_dtype == 1 ? 2 : 3)) : 0;
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
  const CUtensorMap& ta = p.a_is_dy ? tdy : tx;
  const CUtensorMap& tb = p.a_is_dy ? tx : tdy;
  const bool bf = dtype == 2;
  int rc;
  if (BN == 256) rc = bf ? launch_wgrad2<256, true>(ta, tb, p, stream) : launch_wgrad2<256, false>(ta, tb, p, stream);
  else rc = bf ? launch_wgrad2<128, true>(ta, tb, p, stream) : launch_wgrad2<128, false>(ta, tb, p, stream);
  if (rc) return rc;
  if (p.direct_out) return 0;  // no K split: the epilogue wrote dW
  const long long total = static_cast<long long>(Cout) * R * S * Cin;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (out_dtype == 0)
    u2b_launch_pdl(wgrad2_reduce_kernel<float>, dim3(grid), dim3(256), 0, stream, workspace, p.ksplit, R * S, p.Ca, p.Cb, p.a_is_dy,
                                                          static_cast<float*>(dw), total);
  else if (out_dtype == 1)
    u2b_launch_pdl(wgrad2_reduce_kernel<__half>, dim3(grid), dim3(256), 0, stream, workspace, p.ksplit, R * S, p.Ca, p.Cb, p.a_is_dy,
                                                           static_cast<__half*>(dw), total);
  else
    u2b_launch_pdl(wgrad2_reduce_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, workspace, p.ksplit, R * S, p.Ca, p.Cb, p.a_is_dy,
                                                                  static_cast<__nv_bfloat16*>(dw), total);
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
