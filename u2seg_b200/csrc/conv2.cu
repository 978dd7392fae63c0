// 2-CTA (tcgen05.mma.cta_group::2) implicit-GEMM convolution / GEMM on sm_100a: NHWC fp16/bf16 operands, fp32
// accumulation in TMEM, bf16/fp16 NHWC output. Second generation of csrc/conv_tc.cu (which stays as the cta_group::1
// variant); call sites on the u2seg hot path: detectron2/layers/wrappers.py:127 (every F.conv2d of
// backbone/resnet.py:149-176, backbone/fpn.py:77-88, proposal_generator/rpn.py:116-134, roi_heads/mask_head.py:242-262,
// meta_arch/semantic_seg.py:196-214) and, viewed as a 1x1 convolution over a (1,1,M,K) image, the nn.Linear layers
// (roi_heads/box_head.py:70,94-97, roi_heads/fast_rcnn.py:236-239).
//
// GEMM view: M = N*OH*OW output pixels, N = Cout, K = R*S*Cin. A thread-block CLUSTER OF TWO CTAs (one SM pair)
// computes a 256 x BN output tile:
//   A: each CTA TMA-loads its own 128-pixel box {64 ch, BW*stride, BH*stride, 1} of the NHWC input at the filter tap's
//      shifted origin (zero padding / ragged edges = TMA out-of-bounds fill; im2col is never materialised);
//   B: each CTA loads HALF of the (BN x 64) filter tile (BN/2 rows); the pair's UMMA reads both halves, so every SM
//      stages and reads half the B bytes of the 1-CTA kernel (the shared-memory port is what capped conv_tc);
//   D: 256 x BN fp32 accumulator, rows 0-127 in the leader's TMEM, rows 128-255 in the peer's, double buffered;
//      ONE thread of the leader CTA issues tcgen05.mma.cta_group::2 (M=256, N=BN, K=16) for the pair.
// Pipelines: smem ring full/empty (TMA of both CTAs -> leader's `full`; tcgen05.commit multicast -> both `empty`),
// TMEM full/empty (commit multicast -> both epilogues; both epilogues -> leader's `T_empty`).
// Epilogue (8 warps per CTA: two per TMEM lane quadrant, each pair of quadrant-mates alternates over the 64-column
// chunks): tcgen05.ld of 32 columns per wait, + bias, ReLU, round to bf16, stage 128 x 64 chunks in swizzled shared
// memory and write them with TMA stores (coalesced 128-byte rows, ragged tiles clipped by the TMA unit). Optionally emits
// per-channel sum / sum-of-squares of the ROUNDED output tile (read back from the staged chunk, conflict-free; one partial
// row per 128-pixel tile, deterministic): the statistics pass of the SyncBN that follows almost every conv of the
// backbone (layers/batch_norm.py:187) costs no extra read of the activation.
// Input gradient: dX = conv(dY, rot180(W)^T) is the same kernel reading the UNTRANSPOSED (Cout,R,S,Cin) filter as an
// MN-major B operand (GEMM-K = Cout is the slow axis of the filter rows), taps flipped by index arithmetic.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA only), 2 = TMEM allocator, 4-11 = epilogue.
#include <cuda_bf16.h>

#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

constexpr int BM = 128;  // rows per CTA; the pair computes 256
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int STG_BYTES = BM * 64 * 2;  // epilogue staging chunk: 128 rows x 64 channels
constexpr int C2_THREADS = 384;
constexpr int EPI_THREADS = 256;  // warps 4..11

struct Conv2Params {
  int N, H, W, Cin, Cout, R, S, stride, pad, OH, OW;
  int BW, BH, tiles_w, tiles_h, tiles_m, tiles_n, kblocks_c;
  int lbw;       // log2(BW): tile rows -> (bh, bw) by shift / mask
  int num_work;  // (m pair, n tile)
  int b_mn;      // 0: filter (Cout_gemm, R*S*K) K-major rows (forward); 1: dgrad, filter (K, R, S, Cout_gemm) read MN-major, taps flipped
  int relu;
  const float* bias;
  float* stats;  // (tiles_m, 2*Cout) fp32 partial sums [sum | sum of squares] or NULL
  int stages;    // operand ring depth actually used (<= Cfg2::STAGES)
  int nstg;      // epilogue staging chunks per epilogue half (1..3): shared memory not used by the ring
};

template <int BN>
struct Cfg2 {
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the filter tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 5 : (BN == 128 ? 7 : 8);
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int STAT_OFF = STG_OFF + 2 * STG_BYTES;  // one staging chunk per epilogue half
  static constexpr int BAR_OFF = STAT_OFF + 4 * 2 * BN * 4;
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * STAGES + 4) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// ---- PTX used only by the 2-CTA kernel ----
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(cta_addr), "r"(rank));
  return out;
}
// relaxed: the only data this arrive publishes are completed TMEM reads, ordered by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync; a release.cluster arrive compiles to MEMBAR.ALL + ERRBAR and stalls every
// epilogue warp for hundreds of cycles per tile (ncu source page, round 2)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > U2B_MBAR_TIMEOUT_CYCLES) {
      printf("u2b conv2: mbarrier timeout block %d thread %d bar@%u parity %u\n", blockIdx.x, threadIdx.x,
             ptx::smem_u32(bar), parity);
      __trap();
    }
  }
}
// TMA loads whose completion bytes land on the LEADER CTA's mbarrier (cluster address), executed by both CTAs
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(ptx::smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {  // whole warp, both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the mbarrier at this CTA-relative offset in both CTAs of the pair once all prior MMAs retire
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          ptx::smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

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

template <int BN, bool BF16>
__global__ void __launch_bounds__(C2_THREADS, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
             const __grid_constant__ CUtensorMap tmap_y, const Conv2Params p) {
  using Cfg = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stg = smem + p.stages * Cfg::STAGE_BYTES;   // staging chunks follow the ring actually used
  float* sstat = reinterpret_cast<float*>(smem + Cfg::STAT_OFF);  // [4 quadrants][2*BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* T_full = empty + Cfg::STAGES;
  uint64_t* T_empty = T_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(T_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  ptx_free::pdl_launch_dependents();   // the next kernel of the stream may set itself up while this one runs (common.cuh)
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);   // leader's producer (expect_tx covers both CTAs' bytes)
      ptx::mbar_init(&empty[i], 1);  // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&T_full[i], 1);
      ptx::mbar_init(&T_empty[i], 16);  // 8 epilogue warps x 2 CTAs (only the leader's copy is used)
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // the peer's barriers exist before any remote arrive / complete_tx
  ptx::tc_fence_after();
  ptx_free::pdl_wait();     // barriers, TMEM and descriptor prefetch above overlap the predecessor's tail; its data is read below
  const uint32_t tmem_base = *tmem_ptr;
  const int kblocks = p.R * p.S * p.kblocks_c;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int work = cluster_id; work < p.num_work; work += num_clusters) {
        const int tn = work % p.tiles_n;
        int tm = (work / p.tiles_n) * 2 + static_cast<int>(rank);
        if (tm >= p.tiles_m) tm = p.tiles_m - 1;  // padding CTA of an odd tile count: valid loads, no stores
        const int owb = tm % p.tiles_w, ohb = (tm / p.tiles_w) % p.tiles_h, n = tm / (p.tiles_w * p.tiles_h);
        const int x_base = owb * p.BW * p.stride - p.pad, y_base = ohb * p.BH * p.stride - p.pad;
        for (int r = 0; r < p.R; ++r)
          for (int s = 0; s < p.S; ++s)
            for (int cb = 0; cb < p.kblocks_c; ++cb) {
              ptx::mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              const uint32_t full_leader = mapa_rank(ptx::smem_u32(&full[stage]), 0);
              if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
              tma_load_4d_2sm(sa, &tmap_x, full_leader, cb * BK, x_base + s, y_base + r, n);
              if (!p.b_mn) {
                const int kcol = ((r * p.S + s) * p.kblocks_c + cb) * BK;
                tma_load_2d_2sm(sa + A_BYTES, &tmap_w, full_leader, kcol, tn * BN + static_cast<int>(rank) * (BN / 2));
              } else {
                // filter rows = GEMM-K channels (cb*64 .. +63), columns = flipped tap x this CTA's BN/2 output channels:
                // boxes of [64 K rows][64 N columns] = the MN-major SW128 operand layout (next 64-column block 8 KB on)
                const int tap = (p.R - 1 - r) * p.S + (p.S - 1 - s);
                const int col0 = tap * p.Cout + tn * BN + static_cast<int>(rank) * (BN / 2);
#pragma unroll
                for (int j = 0; j < BN / 128; ++j)
                  tma_load_2d_2sm(sa + A_BYTES + j * (64 * 128), &tmap_w, full_leader, col0 + j * 64, cb * BK);
              }
              if (++stage == static_cast<uint32_t>(p.stages)) {
                stage = 0;
                phase ^= 1;
              }
            }
      }
    }
  } else if (warp == 1) {
    if (leader && ptx::elect_one()) {
      const uint32_t idesc = ptx::umma_idesc_f16(2 * BM, BN, BF16 ? 1u : 0u) | (p.b_mn ? (1u << 16) : 0u);  // bit 16: B MN-major
      const uint32_t b_step = p.b_mn ? 128u : 2u;   // K = 16: 16 rows of 128 B (MN-major) or 32 B along the row (K-major)
      const uint32_t sbase = ptx::smem_u32(smem);
      uint32_t stage = 0, phase = 0, acc_it = 0;
      for (int work = cluster_id; work < p.num_work; work += num_clusters, ++acc_it) {
        const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
        mbar_wait_cluster(&T_empty[buf], tphase ^ 1);  // both CTAs' epilogues have drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_cluster(&full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::umma_desc_sw128(sbase + stage * Cfg::STAGE_BYTES);
          uint64_t b_desc = ptx::umma_desc_sw128(sbase + stage * Cfg::STAGE_BYTES + A_BYTES);
          if (p.b_mn) b_desc |= static_cast<uint64_t>((64 * 128) >> 4) << 16;   // LBO: next 64-column block
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_2sm(tmem_d, a_desc + 2 * k, b_desc + b_step * k, idesc, (kb | k) != 0);
          umma_commit_2sm(&empty[stage]);
          if (++stage == static_cast<uint32_t>(p.stages)) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&T_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;          // the two warps of a quadrant alternate over the 64-column chunks
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 128;     // 0..255
    const int half_tid = epi_tid & 127;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    // p.nstg staging chunks per half, used round-robin: the TMA store of chunk i drains while chunks i+1.. are produced
    // (one chunk per half made every 16 KB store a ~2 us round trip on the memory-bound 1x1 layers)
    uint32_t sc = 0;
    const uint32_t t_empty_leader = mapa_rank(ptx::smem_u32(&T_empty[0]), 0);
    uint32_t acc_it = 0;
    for (int work = cluster_id; work < p.num_work; work += num_clusters, ++acc_it) {
      const int tn = work % p.tiles_n, tm_raw = (work / p.tiles_n) * 2 + static_cast<int>(rank);
      const bool store_tile = tm_raw < p.tiles_m;
      const int tm = store_tile ? tm_raw : p.tiles_m - 1;
      const int owb = tm % p.tiles_w, ohb = (tm / p.tiles_w) % p.tiles_h, n = tm / (p.tiles_w * p.tiles_h);
      const bool tile_full = (ohb + 1) * p.BH <= p.OH && (owb + 1) * p.BW <= p.OW;
      const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
      ptx::mbar_wait(&T_full[buf], tphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = half; c < BN / 64; c += 2) {
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 64 + h * 32, v);
          ptx::tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + tn * BN + c * 64 + h * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[h * 16 + j] = pack2<BF16>(f[2 * j], f[2 * j + 1]);
        }
        uint8_t* sbuf = stg + (half * p.nstg + static_cast<int>(sc % static_cast<uint32_t>(p.nstg))) * STG_BYTES;
        const uint32_t srow = ptx::smem_u32(sbuf) + static_cast<uint32_t>(row) * 128u;
        ++sc;
        // the store that last used this chunk (nstg chunks ago) must have finished reading it before it is overwritten
        if (half_tid == 0) {
          if (p.nstg == 1) bulk_wait_group_read<0>();
          else if (p.nstg == 2) bulk_wait_group_read<1>();
          else bulk_wait_group_read<2>();
        }
        ptx::named_bar_sync(1 + half, 128);
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((j ^ sw) << 4)), "r"(pk[4 * j]),
                       "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                       : "memory");
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1 + half, 128);
        if (half_tid == 0) {
          if (store_tile) tma_store_4d(&tmap_y, sbuf, tn * BN + c * 64, owb * p.BW, ohb * p.BH, n);
          bulk_commit_group();     // also for the padding tile (an empty group): wait_group counts groups, one per chunk
        }
        if (p.stats) {
          // per-channel sum / sum of squares of the staged (rounded) chunk: this warp takes rows q*32..q*32+31, lane l
          // the channel pair (2l, 2l+1); for a fixed row the 32 lanes touch 32 distinct banks (16-byte slot
          // (l>>2)^(row&7), word l&3). Rows outside the image (ragged / padding tiles) are skipped.
          float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
          const uint32_t base = ptx::smem_u32(sbuf) + static_cast<uint32_t>(lane & 3) * 4u;
          const uint32_t slot = static_cast<uint32_t>(lane >> 2);
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const int r2 = q * 32 + rr;
            bool ok = store_tile;
            if (!tile_full) ok = ok && (ohb * p.BH + (r2 >> p.lbw)) < p.OH && (owb * p.BW + (r2 & (p.BW - 1))) < p.OW;
            uint32_t u;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(base + static_cast<uint32_t>(r2) * 128u + ((slot ^ (r2 & 7)) << 4)));
            const float2 x = unpack2<BF16>(u);
            if (ok) {
              s0 += x.x; s1 += x.y;
              q0 = fmaf(x.x, x.x, q0); q1 = fmaf(x.y, x.y, q1);
            }
          }
          float* st = sstat + q * 2 * BN + c * 64 + 2 * lane;
          *reinterpret_cast<float2*>(st) = make_float2(s0, s1);
          *reinterpret_cast<float2*>(st + BN) = make_float2(q0, q1);
        }
      }
      // this warp's share of the accumulator has been read: hand the TMEM buffer back to the MMA issuer
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(t_empty_leader + buf * 8);
      if (p.stats) {
        ptx::named_bar_sync(3, EPI_THREADS);
        if (store_tile) {
          float* dst = p.stats + static_cast<size_t>(tm) * 2 * p.Cout;
          for (int i = epi_tid; i < 2 * BN; i += EPI_THREADS) {
            const float sum = (sstat[i] + sstat[2 * BN + i]) + (sstat[4 * BN + i] + sstat[6 * BN + i]);
            dst[(i < BN ? 0 : p.Cout - BN) + tn * BN + i] = sum;
          }
        }
        ptx::named_bar_sync(3, EPI_THREADS);
      }
    }
    if (half_tid == 0) bulk_wait_group_all();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // no CTA may exit while its peer can still read its smem / arrive on its barriers
  if (warp == 2) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
}

// 1: trade ring stages for epilogue staging chunks on short-K layers; 0 (default): full ring + one chunk per half. Measured in
// the training step (A/B, twice each): 15.60 ms off vs 15.65 ms on - the memory-bound 1x1 layers are not limited by the number
// of TMA stores in flight, so the experiment stays behind u2b_conv2_set_staging / U2B_CONV2_STAGING=1.
int g_conv2_staging = 0;

// Experiment (see g_conv2_staging): short reductions (1x1 convolutions with K <= 256: 1-4 k-blocks per tile) move 64 KB of
// output per CTA and tile against 16-64 KB of operands and do not need a deep operand ring; the shared memory can go to
// staging chunks instead so that several TMA stores are in flight per epilogue half.
template <int BN>
void conv2_pick_pipeline(Conv2Params& p) {
  using Cfg = Cfg2<BN>;
  const int kblocks = p.R * p.S * p.kblocks_c;
  p.stages = Cfg::STAGES;
  p.nstg = 1;
  if (!g_conv2_staging) return;
  if (kblocks <= 4) {
    p.stages = 3;
    p.nstg = 3;
  } else if (kblocks <= 12) {
    p.stages = 4;
    p.nstg = 2;
  }
  // staging chunks live between the ring actually used and the fixed statistics / barrier area
  while (p.nstg > 1 && p.stages * Cfg::STAGE_BYTES + 2 * p.nstg * STG_BYTES > Cfg::STAT_OFF) --p.nstg;
}

template <int BN, bool BF16>
int launch_conv2(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const Conv2Params& p_in,
                 cudaStream_t stream) {
  using Cfg = Cfg2<BN>;
  Conv2Params p = p_in;
  conv2_pick_pipeline<BN>(p);
  static bool attr = false;
  if (!attr) {
    U2B_CUDA(cudaFuncSetAttribute(conv2_kernel<BN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    attr = true;
  }
  int clusters = u2b_persistent_sms() / 2;
  if (clusters > p.num_work) clusters = p.num_work;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(C2_THREADS);
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
  U2B_CUDA(cudaLaunchKernelEx(&cfg, conv2_kernel<BN, BF16>, tx, tw, ty, p));
  return 0;
}

int g_conv2_bn = 0;  // 0 = automatic tile width; 64 / 128 / 256 force it (u2b_conv2_set_tile_n, benchmarking)

void conv2_geometry(Conv2Params& p, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.stride = stride; p.pad = pad;
  p.OH = (H + 2 * pad - R) / stride + 1;
  p.OW = (W + 2 * pad - S) / stride + 1;
  long long best = -1;  // the 128-pixel tile shape with the least padding
  for (int bw = 128; bw >= 8; bw >>= 1) {
    const int bh = 128 / bw;
    const long long cost = static_cast<long long>((p.OW + bw - 1) / bw) * bw * ((p.OH + bh - 1) / bh) * bh;
    if (best < 0 || cost < best) {
      best = cost;
      p.BW = bw;
      p.BH = bh;
    }
  }
  p.lbw = 0;
  while ((1 << p.lbw) < p.BW) ++p.lbw;
  p.tiles_w = (p.OW + p.BW - 1) / p.BW;
  p.tiles_h = (p.OH + p.BH - 1) / p.BH;
  p.tiles_m = p.tiles_w * p.tiles_h * N;
  p.kblocks_c = Cin / BK;
}

int conv2_pick_bn(const Conv2Params& p, int min_bn) {
  if (g_conv2_bn >= min_bn && p.Cout % g_conv2_bn == 0) return g_conv2_bn;
  const int pairs = (p.tiles_m + 1) / 2, want = u2b_persistent_sms() / 2;
  // widest tile that still gives every SM pair a work item; otherwise the narrowest (most parallelism)
  for (int bn = 256; bn >= min_bn; bn >>= 1)
    if (p.Cout % bn == 0 && pairs * (p.Cout / bn) >= want) return bn;
  for (int bn = min_bn; bn <= 256; bn <<= 1)
    if (p.Cout % bn == 0) return bn;
  return 0;
}

// x: (N,H,W,Kc) NHWC activations whose Kc channels are the GEMM-K axis; out: (N,OH,OW,Nc). b_mn = 0: w is (Nc, R*S*Kc)
// K-major rows (forward). b_mn = 1: w is the forward filter (Kc, R*S*Nc) read MN-major with flipped taps (dgrad).
int conv2_run(int dtype, const void* x, int N, int H, int W, int Kc, const void* w, int Nc, int R, int S, int stride,
              int pad, int b_mn, const float* bias, int relu, void* out, float* stats, cudaStream_t stream) {
  Conv2Params p;
  conv2_geometry(p, N, H, W, Kc, Nc, R, S, stride, pad);
  const int BN = conv2_pick_bn(p, b_mn ? 128 : 64);
  if (BN == 0) {
    u2b_set_error("conv2: no tile width for %d output channels (mode %d)", Nc, b_mn);
    return U2B_ERR_UNSUPPORTED;
  }
  p.tiles_n = Nc / BN;
  p.num_work = ((p.tiles_m + 1) / 2) * p.tiles_n;
  p.b_mn = b_mn;
  p.relu = relu;
  p.bias = bias;
  p.stats = stats;
  const CUtensorMapDataType tdt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tx, tw, ty;
  {
    uint64_t dims[4] = {(uint64_t)Kc, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Kc * 2, (uint64_t)W * Kc * 2, (uint64_t)H * W * Kc * 2};
    uint32_t box[4] = {BK, (uint32_t)(p.BW * stride), (uint32_t)(p.BH * stride), 1};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = u2b_encode_tmap(&tx, tdt, 4, x, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  if (!b_mn) {
    uint64_t dims[2] = {(uint64_t)R * S * Kc, (uint64_t)Nc};
    uint64_t strides[1] = {(uint64_t)R * S * Kc * 2};
    uint32_t box[2] = {BK, (uint32_t)(BN / 2)};
    int rc = u2b_encode_tmap(&tw, tdt, 2, w, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  } else {
    uint64_t dims[2] = {(uint64_t)R * S * Nc, (uint64_t)Kc};
    uint64_t strides[1] = {(uint64_t)R * S * Nc * 2};
    uint32_t box[2] = {64, 64};
    int rc = u2b_encode_tmap(&tw, tdt, 2, w, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Nc, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Nc * 2, (uint64_t)p.OW * Nc * 2, (uint64_t)p.OH * p.OW * Nc * 2};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    int rc = u2b_encode_tmap(&ty, tdt, 4, out, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  const bool bf = dtype == 2;
  if (BN == 256) return bf ? launch_conv2<256, true>(tx, tw, ty, p, stream) : launch_conv2<256, false>(tx, tw, ty, p, stream);
  if (BN == 128) return bf ? launch_conv2<128, true>(tx, tw, ty, p, stream) : launch_conv2<128, false>(tx, tw, ty, p, stream);
  if (BN == 64) return bf ? launch_conv2<64, true>(tx, tw, ty, p, stream) : launch_conv2<64, false>(tx, tw, ty, p, stream);
  return U2B_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" {

// 1 if (shape) is handled by the 2-CTA tcgen05 kernel
int u2b_conv2_supported(int Cin, int Cout, int R, int S, int stride, int pad) {
  if (Cin <= 0 || Cin % 64 != 0 || Cout <= 0 || Cout % 64 != 0) return 0;
  if (R == 2 && S == 2) return pad == 0 && stride == 2;  // the input gradient of ConvTranspose2d(k=2, s=2) (mask_head.py:256)
  if (!((R == 1 && S == 1 && pad == 0) || (R == 3 && S == 3 && pad == 1))) return 0;
  if (stride != 1 && stride != 2) return 0;
  return 1;
}

// rows of the BN-statistics partial buffer (one per 128-pixel output tile) for this problem
int64_t u2b_conv2_stats_rows(int N, int H, int W, int R, int S, int stride, int pad) {
  Conv2Params p;
  conv2_geometry(p, N, H, W, 64, 64, R, S, stride, pad);
  return p.tiles_m;
}

int u2b_conv2_set_staging(int on) {
  g_conv2_staging = on ? 1 : 0;
  return 0;
}

int u2b_conv2_set_tile_n(int bn) {
  U2B_CHECK_ARG(bn == 0 || bn == 64 || bn == 128 || bn == 256, "conv2_set_tile_n: 0 (auto), 64, 128 or 256");
  g_conv2_bn = bn;
  return 0;
}

// dtype: 1 = fp16, 2 = bf16. x: (N,H,W,Cin) NHWC. w: (Cout,R,S,Cin). out: (N,OH,OW,Cout) NHWC. bias: Cout fp32 or NULL.
// stats: NULL, or (u2b_conv2_stats_rows, 2*Cout) fp32: row t = [sum_c | sumsq_c] of the rounded outputs of tile t
// (every row and column is written; feed it to u2b_bn_finalize with S = rows).
int u2b_conv2_nhwc_fwd(int dtype, const void* x, int N, int H, int W, int Cin, const void* w, int Cout, int R, int S,
                       int stride, int pad, const float* bias, int relu, void* out, float* stats,
                       cudaStream_t stream) {
  U2B_CHECK_ARG(x && w && out && N > 0 && H > 0 && W > 0, "conv2_nhwc_fwd: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "conv2_nhwc_fwd: dtype must be fp16(1) or bf16(2)");
  if (!u2b_conv2_supported(Cin, Cout, R, S, stride, pad)) {
    u2b_set_error("conv2_nhwc_fwd: unsupported shape Cin=%d Cout=%d k=%dx%d stride=%d pad=%d", Cin, Cout, R, S, stride,
                  pad);
    return U2B_ERR_UNSUPPORTED;
  }
  return conv2_run(dtype, x, N, H, W, Cin, w, Cout, R, S, stride, pad, 0, bias, relu, out, stats, stream);
}

// 1 if the input gradient of this (stride-1) convolution runs on the kernel with the untransposed filter
int u2b_conv2_dgrad_supported(int Cin, int Cout, int R, int S, int stride, int pad) {
  if (stride != 1 || Cin <= 0 || Cin % 128 != 0 || Cout <= 0 || Cout % 64 != 0) return 0;
  return (R == 1 && S == 1 && pad == 0) || (R == 3 && S == 3 && pad == 1);
}

// dX = conv(dY, rot180(W)^T) for a stride-1 'same' convolution. dy: (N,H,W,Cout) NHWC; w: the FORWARD filter
// (Cout,R,S,Cin), read in place as an MN-major operand; dx: (N,H,W,Cin) NHWC.
int u2b_conv2_nhwc_dgrad(int dtype, const void* dy, int N, int H, int W, int Cout, const void* w, int Cin, int R, int S,
                         int pad, void* dx, cudaStream_t stream) {
  U2B_CHECK_ARG(dy && w && dx && N > 0 && H > 0 && W > 0, "conv2_nhwc_dgrad: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "conv2_nhwc_dgrad: dtype must be fp16(1) or bf16(2)");
  if (!u2b_conv2_dgrad_supported(Cin, Cout, R, S, 1, pad)) {
    u2b_set_error("conv2_nhwc_dgrad: unsupported shape Cin=%d Cout=%d k=%dx%d pad=%d", Cin, Cout, R, S, pad);
    return U2B_ERR_UNSUPPORTED;
  }
  return conv2_run(dtype, dy, N, H, W, Cout, w, Cin, R, S, 1, R - 1 - pad, 1, nullptr, 0, dx, nullptr, stream);
}

// ConvTranspose2d(kernel 2, stride 2, pad 0) forward (roi_heads/mask_head.py:256 `deconv`), NHWC:
//   y[n, 2h+i, 2w+j, co] = [relu](bias[co] + sum_ci x[n,h,w,ci] * Wt[ci,co,i,j])
// as four 1x1 GEMMs on the 2-CTA kernel, one per output phase (i, j): the filter slice Wt[:, :, i, j] is read in place from
// the channels_last weight (physical (Cin,2,2,Cout)) as an MN-major operand, and the TMA-store epilogue writes each
// phase's rows straight into its interleaved positions of y (tensor map with doubled pixel / row strides).
// (Its input gradient is u2b_conv2_nhwc_fwd with the same weight as a 2x2 / stride-2 filter, its weight gradient
// u2b_conv_wgrad2 of that convolution.)
int u2b_deconv2x2_supported(int Cin, int Cout) { return Cin > 0 && Cin % 64 == 0 && Cout > 0 && Cout % 128 == 0; }

int u2b_deconv2x2_nhwc_fwd(int dtype, const void* x, int N, int H, int W, int Cin, const void* w, int Cout,
                           const float* bias, int relu, void* y, cudaStream_t stream) {
  U2B_CHECK_ARG(x && w && y && N > 0 && H > 0 && W > 0, "deconv2x2_nhwc_fwd: bad arguments");
  U2B_CHECK_ARG(dtype == 1 || dtype == 2, "deconv2x2_nhwc_fwd: dtype must be fp16(1) or bf16(2)");
  if (!u2b_deconv2x2_supported(Cin, Cout)) {
    u2b_set_error("deconv2x2_nhwc_fwd: unsupported channels Cin=%d Cout=%d", Cin, Cout);
    return U2B_ERR_UNSUPPORTED;
  }
  Conv2Params p;
  conv2_geometry(p, N, H, W, Cin, Cout, 1, 1, 1, 0);
  const int BN = conv2_pick_bn(p, 128);
  if (BN == 0) return U2B_ERR_UNSUPPORTED;
  p.tiles_n = Cout / BN;
  p.num_work = ((p.tiles_m + 1) / 2) * p.tiles_n;
  p.b_mn = 1;
  p.relu = relu;
  p.bias = bias;
  p.stats = nullptr;
  const CUtensorMapDataType tdt = dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tx;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {BK, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    int rc = u2b_encode_tmap(&tx, tdt, 4, x, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  const bool bf = dtype == 2;
  for (int ph = 0; ph < 4; ++ph) {
    const int i = ph >> 1, j = ph & 1;
    CUtensorMap tw, ty;
    {
      // rows = Cin (GEMM-K), columns = Cout of phase (i, j): physical (Cin, 2, 2, Cout) -> row pitch 4*Cout elements
      const uint16_t* wp = static_cast<const uint16_t*>(w) + static_cast<size_t>(ph) * Cout;
      uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)Cin};
      uint64_t strides[1] = {(uint64_t)4 * Cout * 2};
      uint32_t box[2] = {64, 64};
      int rc = u2b_encode_tmap(&tw, tdt, 2, wp, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc) return rc;
    }
    {
      uint16_t* yp = static_cast<uint16_t*>(y) + (static_cast<size_t>(i) * 2 * W + j) * Cout;
      uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
      uint64_t strides[3] = {(uint64_t)2 * Cout * 2, (uint64_t)2 * (2 * W) * Cout * 2, (uint64_t)(2 * H) * (2 * W) * Cout * 2};
      uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
      int rc = u2b_encode_tmap(&ty, tdt, 4, yp, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc) return rc;
    }
    int rc;
    if (BN == 256) rc = bf ? launch_conv2<256, true>(tx, tw, ty, p, stream) : launch_conv2<256, false>(tx, tw, ty, p, stream);
    else rc = bf ? launch_conv2<128, true>(tx, tw, ty, p, stream) : launch_conv2<128, false>(tx, tw, ty, p, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
