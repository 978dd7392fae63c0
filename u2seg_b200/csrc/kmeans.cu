// k-means (Lloyd) hot path of u2seg/Instance_Clustering, B200-native.
// Reference semantics: u2seg/Instance_Clustering/shared/utils/nn_utils.py:304-379
//   E-step  (nn_utils.py:353-355): cl = argmin_j sum_d (x_i - c_j)^2   (first minimum)
//   M-step  (nn_utils.py:359-364): c = scatter_add(x by cl) / bincount(cl)   (NaN when empty)
//
// E-step design (tensor-pipe bound, 2*N*K*D flop):
//   argmin_j |x-c_j|^2 = argmin_j (|c_j|^2 - 2 x.c_j). The N x K dot products run on tcgen05
//   (fp16 operands, fp32 accumulate in TMEM): a 128-row X tile stays resident in shared memory
//   (TMA, 128B swizzle), centroid tiles of 160 rows stream through a 4-stage TMA ring, two
//   128x160 fp32 accumulators alternate in TMEM so the epilogue (|c|^2 - 2 acc, running top-3)
//   overlaps the next tile's MMAs. Centroids are rounded to fp16 for the tensor pipe, so every
//   row whose best/second gap is inside the rigorous rounding bound 2^-9*|x|max*|c|max is
//   re-decided exactly in fp32 by kmeans_refine_kernel (reference formula, fp32 centroids).
//   A NaN centroid (empty cluster) is never selected, as in the KeOps reduction the reference uses.
#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

constexpr int BM = 128;          // X rows per tile (UMMA M)
constexpr int NT = 160;          // centroids per accumulator tile (UMMA N)
constexpr int BK = 64;           // fp16 elements per 128B swizzle row
constexpr int MAXKB = 6;         // D <= 384 resident
constexpr int BSTAGES = 4;
constexpr int A_KB_BYTES = BM * BK * 2;     // 16384
constexpr int B_STAGE_BYTES = NT * BK * 2;  // 20480
constexpr int EPI_WARPS = 8;
constexpr int ASSIGN_THREADS = 128 + EPI_WARPS * 32;  // 384
constexpr int TMEM_COLS = 512;
constexpr int HALF_N = NT / 2;  // 80 columns per epilogue warp

struct AssignSmem {
  // offsets into dynamic smem (after 1024B alignment)
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + MAXKB * A_KB_BYTES;          // 98304
  static constexpr int MERGE_OFF = B_OFF + BSTAGES * B_STAGE_BYTES;  // 180224
  static constexpr int MERGE_BYTES = 2 * BM * 5 * 4;                 // 5120
  static constexpr int BAR_OFF = MERGE_OFF + MERGE_BYTES;
  static constexpr int NBARS = 2 * MAXKB + 2 * BSTAGES + 4;  // 24
  static constexpr int TMEMPTR_OFF = BAR_OFF + NBARS * 8;
  static constexpr int CNORM_OFF = (TMEMPTR_OFF + 16 + 15) / 16 * 16;
  static int bytes(int kpad) { return CNORM_OFF + kpad * 4 + 1024; }
};

struct Top3 {
  float v1, v2, v3;
  int i1, i2;
  __device__ __forceinline__ void init() {
    v1 = v2 = v3 = __int_as_float(0x7f800000);
    i1 = 0;
    i2 = 0;
  }
  // strict '<' keeps the earliest index among equal values within one thread's ascending scan;
  // NaN compares false and is never inserted.
  __device__ __forceinline__ void push(float d, int j) {
    bool lt1 = d < v1, lt2 = d < v2, lt3 = d < v3;
    v3 = lt2 ? v2 : (lt3 ? d : v3);
    i2 = lt1 ? i1 : (lt2 ? j : i2);
    v2 = lt1 ? v1 : (lt2 ? d : v2);
    i1 = lt1 ? j : i1;
    v1 = lt1 ? d : v1;
  }
};

// CL = thread-block cluster size: the CL CTAs of a cluster work on CL neighbouring row tiles and walk the same
// centroid-tile sequence; each loads NT/CL rows of every centroid tile and TMA-multicasts them to the whole cluster,
// so the (K x D) centroid matrix is read from L2 once per cluster instead of once per CTA.
template <int CL>
__global__ void __launch_bounds__(ASSIGN_THREADS, 1)
kmeans_assign_kernel(const __grid_constant__ CUtensorMap tmap_x,
                     const __grid_constant__ CUtensorMap tmap_c, const float* __restrict__ cnorm,
                     const float* __restrict__ xmax, const float* __restrict__ cmax2,
                     int32_t* __restrict__ labels, int4* __restrict__ amb,
                     int* __restrict__ amb_count, int amb_capacity, int N, int kpad, int kblocks,
                     int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + AssignSmem::A_OFF;
  uint8_t* sB = smem + AssignSmem::B_OFF;
  float* sMerge = reinterpret_cast<float*>(smem + AssignSmem::MERGE_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AssignSmem::BAR_OFF);
  uint64_t* A_full = bars;
  uint64_t* A_empty = bars + MAXKB;
  uint64_t* B_full = bars + 2 * MAXKB;
  uint64_t* B_empty = B_full + BSTAGES;
  uint64_t* T_full = B_empty + BSTAGES;
  uint64_t* T_empty = T_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + AssignSmem::TMEMPTR_OFF);
  float* sCnorm = reinterpret_cast<float*>(smem + AssignSmem::CNORM_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ntiles_n = kpad / NT;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < MAXKB; ++i) {
      ptx::mbar_init(&A_full[i], 1);
      ptx::mbar_init(&A_empty[i], 1);
    }
    for (int i = 0; i < BSTAGES; ++i) {
      ptx::mbar_init(&B_full[i], 1);
      ptx::mbar_init(&B_empty[i], CL);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&T_full[i], 1);
      ptx::mbar_init(&T_empty[i], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kpad; i += blockDim.x) sCnorm[i] = cnorm[i];
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int rank = CL > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  const int num_groups = (num_tiles + CL - 1) / CL;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);
  constexpr int PART_ROWS = NT / CL;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (ptx::elect_one()) {
      uint32_t bstage = 0, bphase = 0;
      int it = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters, ++it) {
        int tile = grp * CL + rank;
        if (tile >= num_tiles) tile = num_tiles - 1;  // padding CTA of the last group (results discarded)
        for (int n = 0; n < ntiles_n; ++n) {
          for (int kb = 0; kb < kblocks; ++kb) {
            if (n == 0) {
              ptx::mbar_wait(&A_empty[kb], (it & 1) ^ 1);
              ptx::mbar_arrive_expect_tx(&A_full[kb], A_KB_BYTES);
              ptx::tma_load_2d(sA + kb * A_KB_BYTES, &tmap_x, &A_full[kb], kb * BK, tile * BM);
            }
            ptx::mbar_wait(&B_empty[bstage], bphase ^ 1);
            ptx::mbar_arrive_expect_tx(&B_full[bstage], B_STAGE_BYTES);
            if (CL == 1)
              ptx::tma_load_2d(sB + bstage * B_STAGE_BYTES, &tmap_c, &B_full[bstage], kb * BK, n * NT);
            else
              ptx::tma_load_2d_mc(sB + bstage * B_STAGE_BYTES + rank * (PART_ROWS * BK * 2), &tmap_c,
                                  &B_full[bstage], kb * BK, n * NT + rank * PART_ROWS, kMask);
            if (++bstage == BSTAGES) {
              bstage = 0;
              bphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(BM, NT, /*fp16*/ 0);
      const uint32_t a_addr = ptx::smem_u32(sA);
      const uint32_t b_addr = ptx::smem_u32(sB);
      uint32_t bstage = 0, bphase = 0, acc_it = 0;
      int it = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters, ++it) {
        for (int n = 0; n < ntiles_n; ++n, ++acc_it) {
          const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
          ptx::mbar_wait(&T_empty[buf], tphase ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * NT;
          for (int kb = 0; kb < kblocks; ++kb) {
            if (n == 0) ptx::mbar_wait(&A_full[kb], it & 1);
            ptx::mbar_wait(&B_full[bstage], bphase);
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::umma_desc_sw128(a_addr + kb * A_KB_BYTES);
            const uint64_t b_desc = ptx::umma_desc_sw128(b_addr + bstage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              ptx::umma_f16(tmem_d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            if (CL == 1)
              ptx::umma_commit(&B_empty[bstage]);
            else
              ptx::umma_commit_mc(&B_empty[bstage], kMask);
            if (n == ntiles_n - 1) ptx::umma_commit(&A_empty[kb]);
            if (++bstage == BSTAGES) {
              bstage = 0;
              bphase ^= 1;
            }
          }
          ptx::umma_commit(&T_full[buf]);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps, (lane quarter q) x (column half) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row_in_tile = q * 32 + lane;
    const float xm = *xmax;
    const float cm2 = *cmax2;
    // rigorous bound: |(cn - 2 x.c16) - (cn - 2 x.c)| <= 2^-10 |x||c|; order certain if gap >= 2^-9|x||c|
    const float margin = 1.25f * 0.001953125f * xm * sqrtf(cm2) + 1e-30f;
    uint32_t acc_it = 0;
    int it = 0;
    for (int grp = cluster_id; grp < num_groups; grp += num_clusters, ++it) {
      const int tile = grp * CL + rank;   // >= num_tiles for the padding CTA: row >= N below, nothing is written
      Top3 t;
      t.init();
      for (int n = 0; n < ntiles_n; ++n, ++acc_it) {
        const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
        ptx::mbar_wait(&T_full[buf], tphase);
        ptx::tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * NT + half * HALF_N;
        uint32_t r[HALF_N / 16][16];
#pragma unroll
        for (int c = 0; c < HALF_N / 16; ++c) ptx::tmem_ld16(taddr + c * 16, r[c]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&T_empty[buf]);  // accumulator is in registers now
        const int jbase = n * NT + half * HALF_N;
#pragma unroll
        for (int c = 0; c < HALF_N / 16; ++c) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            // |c|^2 for 4 consecutive centroids in one 16-byte shared-memory load (broadcast to the warp)
            const float4 cn = *reinterpret_cast<const float4*>(sCnorm + jbase + c * 16 + j4 * 4);
            const float cnv[4] = {cn.x, cn.y, cn.z, cn.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int col = jbase + c * 16 + j4 * 4 + jj;
              const float d = fmaf(-2.0f, __uint_as_float(r[c][j4 * 4 + jj]), cnv[jj]);
              t.push(d, col);
            }
          }
        }
      }
      // merge the two column halves through smem (double-buffered by tile parity)
      float* mb = sMerge + (it & 1) * (BM * 5);
      if (half == 1) {
        mb[row_in_tile * 5 + 0] = t.v1;
        mb[row_in_tile * 5 + 1] = t.v2;
        mb[row_in_tile * 5 + 2] = t.v3;
        mb[row_in_tile * 5 + 3] = __int_as_float(t.i1);
        mb[row_in_tile * 5 + 4] = __int_as_float(t.i2);
      }
      ptx::named_bar_sync(1, EPI_WARPS * 32);
      if (half == 0) {
        const float o1 = mb[row_in_tile * 5 + 0], o2 = mb[row_in_tile * 5 + 1],
                    o3 = mb[row_in_tile * 5 + 2];
        const int oi1 = __float_as_int(mb[row_in_tile * 5 + 3]),
                  oi2 = __float_as_int(mb[row_in_tile * 5 + 4]);
        t.push(o1, oi1);
        t.push(o2, oi2);
        t.push(o3, -1);  // only its value matters (can only land in v3)
        const long long row = static_cast<long long>(tile) * BM + row_in_tile;
        if (row < N) {
          labels[row] = t.i1;
          if (t.v2 - t.v1 <= margin) {
            const int slot = atomicAdd(amb_count, 1);
            if (slot < amb_capacity)
              amb[slot] = make_int4(static_cast<int>(row), t.i1, t.i2, (t.v3 - t.v1 <= margin) ? 1 : 0);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

int g_kmeans_cluster = 2;
int g_kmeans_mstep_sort = 1;   // 1: sort-by-label M-step (u2b_kmeans_set_mstep), 0: shared-memory accumulators

template <int CL>
int launch_assign(const CUtensorMap& tx, const CUtensorMap& tc, const float* cnorm, const float* xmax,
                  const float* cmax2, int32_t* labels, int4* amb, int* amb_count, int N, int kpad, int kblocks,
                  int num_tiles, int smem, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    U2B_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    attr_set = true;
  }
  const int groups = (num_tiles + CL - 1) / CL;
  int clusters = u2b_num_sms() / CL;
  if (clusters > groups) clusters = groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(ASSIGN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  U2B_CUDA(cudaLaunchKernelEx(&cfg, kmeans_assign_kernel<CL>, tx, tc, cnorm, xmax, cmax2, labels, amb, amb_count, N,
                              N, kpad, kblocks, num_tiles));
  return 0;
}

// Exact fp32 decision (reference formula sum_d (x-c)^2) for rows the fp16 tensor pass could not
// decide. entry = {row, i1, i2, full}: full==0 -> only {i1,i2} can be the argmin (third-best was outside the
// bound): one warp per entry. full==1 -> every centroid is scanned: one CTA per entry, warps stride over the
// centroids, lanes over D (coalesced fp32 rows), (distance, index) pairs min-reduced through shared memory.
__device__ __forceinline__ float warp_dist(const float* __restrict__ xs, const float* __restrict__ cr, int D,
                                           int lane) {
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float df = xs[d] - cr[d];
    s = fmaf(df, df, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

__global__ void __launch_bounds__(256)
kmeans_refine_kernel(const __half* __restrict__ x, const float* __restrict__ c32,
                     const int4* __restrict__ amb, const int* __restrict__ amb_count,
                     int amb_capacity, int32_t* __restrict__ labels, int D, int K) {
  extern __shared__ float rs[];   // [8 warps][D] x rows (pair pass) / [D] x row + reduction scratch (full pass)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt = *amb_count;
  if (cnt > amb_capacity) cnt = amb_capacity;
  // ---- pass 1: two-candidate entries, one warp each ----
  float* xs = rs + warp * D;
  for (int e = blockIdx.x * 8 + warp; e < cnt; e += gridDim.x * 8) {
    const int4 ent = amb[e];
    if (ent.w) continue;
    const __half* xr = x + static_cast<size_t>(ent.x) * D;
    for (int d = lane; d < D; d += 32) xs[d] = __half2float(xr[d]);
    __syncwarp();
    const float d1 = warp_dist(xs, c32 + static_cast<size_t>(ent.y) * D, D, lane);
    const float d2 = warp_dist(xs, c32 + static_cast<size_t>(ent.z) * D, D, lane);
    // first-minimum semantics: on an exact tie the lower index wins; NaN never wins
    const bool take2 = (d2 < d1) || (d2 == d1 && ent.z < ent.y) || (d1 != d1 && d2 == d2);
    if (lane == 0) labels[ent.x] = take2 ? ent.z : ent.y;
    __syncwarp();
  }
  __syncthreads();
  // ---- pass 2: full scans, one CTA each; every lane owns one centroid (float4 reads of its fp32 row, x broadcast
  //      from shared memory), sequential fp32 accumulation over D as a plain sum of squares ----
  __shared__ float s_best[8];
  __shared__ int s_idx[8];
  for (int e = blockIdx.x; e < cnt; e += gridDim.x) {
    const int4 ent = amb[e];
    if (!ent.w) continue;     // uniform across the CTA
    const __half* xr = x + static_cast<size_t>(ent.x) * D;
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) rs[d] = __half2float(xr[d]);
    __syncthreads();
    float bv = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float4* cr = reinterpret_cast<const float4*>(c32 + static_cast<size_t>(k) * D);
      float acc = 0.f;
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 c = cr[d4];
        const float a0 = rs[4 * d4] - c.x, a1 = rs[4 * d4 + 1] - c.y, a2 = rs[4 * d4 + 2] - c.z, a3 = rs[4 * d4 + 3] - c.w;
        acc = fmaf(a0, a0, acc);
        acc = fmaf(a1, a1, acc);
        acc = fmaf(a2, a2, acc);
        acc = fmaf(a3, a3, acc);
      }
      if (acc < bv) {   // ascending k per thread: first minimum kept; NaN compares false
        bv = acc;
        bi = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_best[warp] = bv;
      s_idx[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = s_best[0];
      int i = s_idx[0];
      for (int w = 1; w < 8; ++w)
        if (s_best[w] < v || (s_best[w] == v && s_idx[w] < i)) {
          v = s_best[w];
          i = s_idx[w];
        }
      labels[ent.x] = (i == 0x7fffffff) ? ent.y : i;   // every distance NaN/inf: keep the tensor pass's label
    }
  }
}

// fp32 centroids -> fp16 tensor operand (padded to kpad rows), |c|^2 in fp32, max finite |c|^2.
__global__ void kmeans_prepare_kernel(const float* __restrict__ c32, __half* __restrict__ c16,
                                      float* __restrict__ cnorm, float* __restrict__ cmax2, int K,
                                      int kpad, int D) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= kpad) return;
  __half* o = c16 + static_cast<size_t>(row) * D;
  if (row >= K) {
    for (int d = lane; d < D; d += 32) o[d] = __float2half(0.f);
    if (lane == 0) cnorm[row] = __int_as_float(0x7f800000);
    return;
  }
  const float* cr = c32 + static_cast<size_t>(row) * D;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = cr[d];
    o[d] = __float2half(v);
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    cnorm[row] = s;
    if (s == s && s < __int_as_float(0x7f800000))
      atomicMax(reinterpret_cast<unsigned int*>(cmax2), __float_as_uint(s));
  }
}

__global__ void kmeans_xnorm_kernel(const __half* __restrict__ x, float* __restrict__ xmax,
                                    long long N, int D) {
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  float m = 0.f;
  for (long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
       row < N; row += warps) {
    const __half2* xr = reinterpret_cast<const __half2*>(x + row * D);
    float s = 0.f;
    for (int d = lane; d < D / 2; d += 32) {
      const float2 v = __half22float2(xr[d]);
      s = fmaf(v.x, v.x, s);
      s = fmaf(v.y, v.y, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    m = fmaxf(m, s);
  }
  if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(xmax), __float_as_uint(sqrtf(m)));
}

// ---- M-step: per-CTA shared-memory accumulators over a (row chunk) x (W-column slice) ----
// grid = (slices, R). smem: acc[K][W] fp32 (+ counts[K] for slice 0). Column pair (2l, 2l+1) of
// lane l is stored at (l, 32+l) so shared atomics are bank-conflict free.
template <int W>
__global__ void __launch_bounds__(1024, 1)
kmeans_accum_kernel(const __half* __restrict__ x, const int32_t* __restrict__ labels,
                    float* __restrict__ partial, long long N, int D, int K) {
  extern __shared__ float sacc[];
  int* scnt = reinterpret_cast<int*>(sacc + static_cast<size_t>(K) * W);
  const int slice = blockIdx.x, r = blockIdx.y, R = gridDim.y;
  for (int i = threadIdx.x; i < K * W; i += blockDim.x) sacc[i] = 0.f;
  if (slice == 0)
    for (int i = threadIdx.x; i < K; i += blockDim.x) scnt[i] = 0;
  __syncthreads();
  const long long rows_per = (N + R - 1) / R;
  const long long r0 = r * rows_per;
  const long long r1 = (r0 + rows_per < N) ? (r0 + rows_per) : N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int col0 = slice * W;
  constexpr int UNROLL = 8;
  for (long long base = r0 + static_cast<long long>(warp) * UNROLL; base < r1;
       base += static_cast<long long>(nwarps) * UNROLL) {
    int lab[UNROLL];
    float2 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long row = base + u;
      lab[u] = -1;
      v[u] = make_float2(0.f, 0.f);
      if (row < r1) {
        lab[u] = labels[row];
        if (W == 64) {
          v[u] = __half22float2(
              *reinterpret_cast<const __half2*>(x + row * D + col0 + 2 * lane));
        } else {
          v[u].x = __half2float(x[row * D + col0 + lane]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (lab[u] >= 0 && lab[u] < K) {
        float* a = sacc + static_cast<size_t>(lab[u]) * W;
        atomicAdd(a + lane, v[u].x);
        if (W == 64) atomicAdd(a + 32 + lane, v[u].y);
        if (slice == 0 && lane == 0) atomicAdd(scnt + lab[u], 1);
      }
    }
  }
  __syncthreads();
  // partial layout: [R][K][D+1]
  float* out = partial + static_cast<size_t>(r) * K * (D + 1);
  for (int i = threadIdx.x; i < K * W; i += blockDim.x) {
    const int k = i / W, p = i % W;
    int col;
    if (W == 64)
      col = (p < 32) ? 2 * p : 2 * (p - 32) + 1;
    else
      col = p;
    out[static_cast<size_t>(k) * (D + 1) + col0 + col] = sacc[i];
  }
  if (slice == 0)
    for (int i = threadIdx.x; i < K; i += blockDim.x)
      out[static_cast<size_t>(i) * (D + 1) + D] = static_cast<float>(scnt[i]);
}

// ---- M-step, second generation: rows grouped by label (counting sort), then segment sums without shared atomics ----
// fp32 atomicAdd on shared memory is a compare-and-swap loop (ATOMS.CAST.SPIN): the accumulator kernel above spends its
// time there (28 % of the HBM roofline). Here the labels are counting-sorted into a row-index list (native integer
// shared atomics only: ATOMS.POPC.INC / ATOMS.ADD), and every warp then walks 64 consecutive list entries - rows of at
// most a few different labels, nondecreasing - summing whole 768-byte rows into registers (coalesced 8-byte loads, no
// atomics) and flushing one partial per label run with global fp32 reductions (RED.ADD.F32, ~20k per iteration).
// X is read exactly once, at gather granularity of a full row.
constexpr int SORT_THREADS = 1024;

__global__ void __launch_bounds__(SORT_THREADS)
kmeans_label_count_kernel(const int32_t* __restrict__ labels, long long N, int K, int* __restrict__ count) {
  extern __shared__ int sh[];
  for (int i = threadIdx.x; i < K; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const long long per = (N + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(N, r0 + per);
  for (long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    const int l = labels[i];
    if (l >= 0 && l < K) atomicAdd(&sh[l], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x)
    if (sh[i]) atomicAdd(&count[i], sh[i]);
}

// start[k] = exclusive prefix sum of count[]; cursor[k] = start[k]; sums[k][D] = count[k] (the count column)
__global__ void __launch_bounds__(1024)
kmeans_label_scan_kernel(const int* __restrict__ count, int K, int D, int* __restrict__ start, int* __restrict__ cursor,
                         float* __restrict__ sums) {
  __shared__ int sh[1024];
  int carry = 0;
  for (int base = 0; base < K; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < K ? count[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      const int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = sh[threadIdx.x];
    if (i < K) {
      start[i] = carry + incl - v;
      cursor[i] = carry + incl - v;
      sums[static_cast<size_t>(i) * (D + 1) + D] = static_cast<float>(v);
    }
    const int total = sh[1023];
    __syncthreads();
    carry += total;
  }
}

__global__ void __launch_bounds__(SORT_THREADS)
kmeans_label_scatter_kernel(const int32_t* __restrict__ labels, long long N, int K, int* __restrict__ cursor,
                            int32_t* __restrict__ sorted_rows, int32_t* __restrict__ sorted_labels) {
  extern __shared__ int sh[];   // [K] counts of this chunk, then reused as this chunk's write cursors
  for (int i = threadIdx.x; i < K; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const long long per = (N + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per, r1 = min(N, r0 + per);
  for (long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    const int l = labels[i];
    if (l >= 0 && l < K) atomicAdd(&sh[l], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const int c = sh[i];
    sh[i] = c ? atomicAdd(&cursor[i], c) : 0;   // reserve this chunk's range of label i's segment
  }
  __syncthreads();
  for (long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    const int l = labels[i];
    if (l >= 0 && l < K) {
      const int pos = atomicAdd(&sh[l], 1);
      sorted_rows[pos] = static_cast<int32_t>(i);
      sorted_labels[pos] = l;
    }
  }
}

// one warp per run of SEG consecutive entries of the sorted list; NP = D / 128 eight-byte pieces per lane
template <int NP>
__global__ void __launch_bounds__(256)
kmeans_gather_sum_kernel(const __half* __restrict__ x, const int32_t* __restrict__ sorted_rows,
                         const int32_t* __restrict__ sorted_labels, const int* __restrict__ start,
                         const int* __restrict__ count, int K, int D, float* __restrict__ sums) {
  constexpr int SEG = 64;
  const int lane = threadIdx.x & 31;
  const long long M = static_cast<long long>(start[K - 1]) + count[K - 1];   // rows with a label in [0, K)
  const long long w = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long e0 = w * SEG, e1 = min(M, e0 + SEG);
  if (e0 >= M) return;
  float acc[NP][4];
#pragma unroll
  for (int j = 0; j < NP; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
  int cur = sorted_labels[e0];
  auto flush = [&](int label) {
    float* dst = sums + static_cast<size_t>(label) * (D + 1);
#pragma unroll
    for (int j = 0; j < NP; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(dst + j * 128 + lane * 4 + k, acc[j][k]);
        acc[j][k] = 0.f;
      }
  };
  for (long long e = e0; e < e1; e += 4) {
    // 4 rows in flight per step: labels / row ids first (uniform across the warp), then the loads, then the adds
    int lab[4];
    long long row[4];
    uint2 v[4][NP];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ee = e + u < e1 ? e + u : e1 - 1;
      lab[u] = sorted_labels[ee];
      row[u] = sorted_rows[ee];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < NP; ++j)
        v[u][j] = *reinterpret_cast<const uint2*>(x + row[u] * D + j * 128 + lane * 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (e + u >= e1) break;
      if (lab[u] != cur) {
        flush(cur);
        cur = lab[u];
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v[u][j].x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v[u][j].y));
        acc[j][0] += a.x; acc[j][1] += a.y; acc[j][2] += b.x; acc[j][3] += b.y;
      }
    }
  }
  flush(cur);
}

__global__ void kmeans_reduce_partials_kernel(const float* __restrict__ partial,
                                              float* __restrict__ sums, int R, long long KD1) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= KD1) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += partial[static_cast<size_t>(r) * KD1 + i];
  sums[i] = s;
}

__global__ void kmeans_finalize_kernel(const float* __restrict__ sums, float* __restrict__ c32,
                                       int K, int D) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(K) * D) return;
  const int k = static_cast<int>(i / D), d = static_cast<int>(i % D);
  // in-place mean; 0/0 -> NaN for an empty cluster, as nn_utils.py:364 (`c /= Ncl`)
  c32[i] = sums[static_cast<size_t>(k) * (D + 1) + d] / sums[static_cast<size_t>(k) * (D + 1) + D];
}

inline int kpad_of(int64_t K) { return static_cast<int>(ceil_div64(K, NT) * NT); }

struct AccumPlan {
  int W, slices, R;
  size_t smem;
};
inline bool accum_plan(int64_t D, int64_t K, AccumPlan* p) {
  for (int W : {64, 32}) {
    if (D % W) continue;
    size_t smem = static_cast<size_t>(K) * W * 4 + static_cast<size_t>(K) * 4;
    if (smem > 220 * 1024) continue;
    p->W = W;
    p->slices = static_cast<int>(D / W);
    int R = u2b_num_sms() / p->slices;
    p->R = R < 1 ? 1 : R;
    p->smem = smem;
    return true;
  }
  return false;
}

}  // namespace

extern "C" {

int64_t u2b_kmeans_kpad(int64_t K) { return kpad_of(K); }

size_t u2b_kmeans_workspace_bytes(int64_t N, int64_t D, int64_t K) {
  AccumPlan p;
  size_t accum = 0;
  if (accum_plan(D, K, &p)) accum = static_cast<size_t>(p.R) * K * (D + 1) * 4;
  size_t amb = static_cast<size_t>(N) * 16 + 256;
  return (accum > amb ? accum : amb) + 1024;
}

int u2b_kmeans_xnorm_max(const void* x16, int64_t N, int64_t D, float* xmax, cudaStream_t stream) {
  U2B_CHECK_ARG(x16 && xmax && N > 0 && D > 0 && D % 2 == 0, "kmeans_xnorm_max: bad arguments");
  U2B_CUDA(cudaMemsetAsync(xmax, 0, sizeof(float), stream));
  int blocks = u2b_num_sms() * 4;
  kmeans_xnorm_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __half*>(x16), xmax, N,
                                                  static_cast<int>(D));
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_kmeans_prepare(const float* c32, int64_t K, int64_t D, void* c16, float* cnorm,
                       float* cmax2, cudaStream_t stream) {
  U2B_CHECK_ARG(c32 && c16 && cnorm && cmax2 && K > 0 && D > 0, "kmeans_prepare: bad arguments");
  const int kpad = kpad_of(K);
  U2B_CUDA(cudaMemsetAsync(cmax2, 0, sizeof(float), stream));
  const int wpb = 8;
  kmeans_prepare_kernel<<<(kpad + wpb - 1) / wpb, wpb * 32, 0, stream>>>(
      c32, static_cast<__half*>(c16), cnorm, cmax2, static_cast<int>(K), kpad, static_cast<int>(D));
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_kmeans_assign(const void* x16, int64_t N, int64_t D, int64_t K, const void* c16,
                      const float* c32, const float* cnorm, const float* xmax, const float* cmax2,
                      int32_t* labels, int32_t* amb_count_out, void* workspace,
                      size_t workspace_bytes, cudaStream_t stream) {
  U2B_CHECK_ARG(x16 && c16 && c32 && cnorm && xmax && cmax2 && labels && workspace,
                "kmeans_assign: null pointer");
  U2B_CHECK_ARG(N > 0 && N < (1LL << 31) && K > 0, "kmeans_assign: bad N=%lld K=%lld",
                (long long)N, (long long)K);
  if (D % BK != 0 || D / BK > MAXKB) {
    u2b_set_error("kmeans_assign: D=%lld unsupported (need D %% 64 == 0 and D <= %d)",
                  (long long)D, MAXKB * BK);
    return U2B_ERR_UNSUPPORTED;
  }
  const int kpad = kpad_of(K);
  if (kpad > 8000) {
    u2b_set_error("kmeans_assign: K=%lld too large for the resident |c|^2 table", (long long)K);
    return U2B_ERR_UNSUPPORTED;
  }
  U2B_CHECK_ARG(workspace_bytes >= static_cast<size_t>(N) * 16 + 256,
                "kmeans_assign: workspace too small");
  int* amb_count = static_cast<int*>(workspace);
  int4* amb = reinterpret_cast<int4*>(static_cast<uint8_t*>(workspace) + 256);
  U2B_CUDA(cudaMemsetAsync(amb_count, 0, sizeof(int), stream));

  const int num_tiles = static_cast<int>(ceil_div64(N, BM));
  int CL = g_kmeans_cluster;
  if (num_tiles < CL) CL = 1;
  CUtensorMap tx, tc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(D), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(D) * 2};
    uint32_t box[2] = {BK, BM};
    int rc = u2b_encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x16, dims, strides, box,
                             nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(D), static_cast<uint64_t>(kpad)};
    uint64_t strides[1] = {static_cast<uint64_t>(D) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(NT / CL)};
    int rc = u2b_encode_tmap(&tc, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c16, dims, strides, box,
                             nullptr, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  const int smem = AssignSmem::bytes(kpad);
  U2B_CHECK_ARG(smem <= 227 * 1024, "kmeans_assign: shared memory budget exceeded");
  int rc;
  if (CL == 4)
    rc = launch_assign<4>(tx, tc, cnorm, xmax, cmax2, labels, amb, amb_count, (int)N, kpad, (int)(D / BK), num_tiles, smem, stream);
  else if (CL == 2)
    rc = launch_assign<2>(tx, tc, cnorm, xmax, cmax2, labels, amb, amb_count, (int)N, kpad, (int)(D / BK), num_tiles, smem, stream);
  else
    rc = launch_assign<1>(tx, tc, cnorm, xmax, cmax2, labels, amb, amb_count, (int)N, kpad, (int)(D / BK), num_tiles, smem, stream);
  if (rc) return rc;
  kmeans_refine_kernel<<<u2b_num_sms() * 2, 256, static_cast<size_t>(8) * D * sizeof(float), stream>>>(
      static_cast<const __half*>(x16), c32, amb, amb_count, static_cast<int>(N), labels,
      static_cast<int>(D), static_cast<int>(K));
  U2B_LAUNCH_CHECK();
  if (amb_count_out)
    U2B_CUDA(cudaMemcpyAsync(amb_count_out, amb_count, sizeof(int), cudaMemcpyDeviceToDevice,
                             stream));
  return 0;
}

int u2b_kmeans_set_cluster(int cluster) {
  U2B_CHECK_ARG(cluster == 1 || cluster == 2 || cluster == 4, "kmeans_set_cluster: 1, 2 or 4");
  g_kmeans_cluster = cluster;
  return 0;
}

int u2b_kmeans_set_mstep(int sort_by_label) {
  g_kmeans_mstep_sort = sort_by_label ? 1 : 0;
  return 0;
}

int u2b_kmeans_accumulate(const void* x16, const int32_t* labels, int64_t N, int64_t D, int64_t K,
                          float* sums, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream) {
  U2B_CHECK_ARG(x16 && labels && sums && workspace && N > 0, "kmeans_accumulate: bad arguments");
  AccumPlan p;
  if (!accum_plan(D, K, &p)) {
    u2b_set_error("kmeans_accumulate: K=%lld D=%lld does not fit the shared-memory accumulators",
                  (long long)K, (long long)D);
    return U2B_ERR_UNSUPPORTED;
  }
  if (g_kmeans_mstep_sort && D % 128 == 0 && D <= 512 && N < (1LL << 31) && K <= 54000 &&
      workspace_bytes >= static_cast<size_t>(N) * 8 + static_cast<size_t>(K) * 12 + 64) {
    // sort-by-label M-step: workspace = sorted_rows[N] | sorted_labels[N] | count[K] | start[K] | cursor[K]
    int32_t* sorted_rows = static_cast<int32_t*>(workspace);
    int32_t* sorted_labels = sorted_rows + N;
    int* count = reinterpret_cast<int*>(sorted_labels + N);
    int* start = count + K;
    int* cursor = start + K;
    U2B_CUDA(cudaMemsetAsync(count, 0, static_cast<size_t>(K) * 4, stream));
    U2B_CUDA(cudaMemsetAsync(sums, 0, static_cast<size_t>(K) * (D + 1) * 4, stream));
    const int chunks = u2b_num_sms() * 2;
    const size_t sh = static_cast<size_t>(K) * 4;
    kmeans_label_count_kernel<<<chunks, SORT_THREADS, sh, stream>>>(labels, N, static_cast<int>(K), count);
    U2B_LAUNCH_CHECK();
    kmeans_label_scan_kernel<<<1, 1024, 0, stream>>>(count, static_cast<int>(K), static_cast<int>(D), start, cursor, sums);
    U2B_LAUNCH_CHECK();
    kmeans_label_scatter_kernel<<<chunks, SORT_THREADS, sh, stream>>>(labels, N, static_cast<int>(K), cursor, sorted_rows,
                                                                     sorted_labels);
    U2B_LAUNCH_CHECK();
    // the list holds the rows with a label in [0, K) (all of them after u2b_kmeans_assign); its length is read on the device
    const long long warps = (N + 63) / 64;
    const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
    const __half* xh = static_cast<const __half*>(x16);
    if (D == 128) kmeans_gather_sum_kernel<1><<<grid, 256, 0, stream>>>(xh, sorted_rows, sorted_labels, start, count, (int)K, (int)D, sums);
    else if (D == 256) kmeans_gather_sum_kernel<2><<<grid, 256, 0, stream>>>(xh, sorted_rows, sorted_labels, start, count, (int)K, (int)D, sums);
    else if (D == 384) kmeans_gather_sum_kernel<3><<<grid, 256, 0, stream>>>(xh, sorted_rows, sorted_labels, start, count, (int)K, (int)D, sums);
    else kmeans_gather_sum_kernel<4><<<grid, 256, 0, stream>>>(xh, sorted_rows, sorted_labels, start, count, (int)K, (int)D, sums);
    U2B_LAUNCH_CHECK();
    return 0;
  }
  const size_t need = static_cast<size_t>(p.R) * K * (D + 1) * 4;
  U2B_CHECK_ARG(workspace_bytes >= need, "kmeans_accumulate: workspace too small");
  float* partial = static_cast<float*>(workspace);
  dim3 grid(p.slices, p.R);
  if (p.W == 64) {
    static bool attr = false;
    if (!attr) {
      U2B_CUDA(cudaFuncSetAttribute(kmeans_accum_kernel<64>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr = true;
    }
    kmeans_accum_kernel<64><<<grid, 1024, p.smem, stream>>>(static_cast<const __half*>(x16), labels,
                                                            partial, N, static_cast<int>(D),
                                                            static_cast<int>(K));
  } else {
    static bool attr = false;
    if (!attr) {
      U2B_CUDA(cudaFuncSetAttribute(kmeans_accum_kernel<32>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr = true;
    }
    kmeans_accum_kernel<32><<<grid, 1024, p.smem, stream>>>(static_cast<const __half*>(x16), labels,
                                                            partial, N, static_cast<int>(D),
                                                            static_cast<int>(K));
  }
  U2B_LAUNCH_CHECK();
  const long long KD1 = static_cast<long long>(K) * (D + 1);
  kmeans_reduce_partials_kernel<<<static_cast<unsigned>((KD1 + 255) / 256), 256, 0, stream>>>(
      partial, sums, p.R, KD1);
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_kmeans_finalize(const float* sums, int64_t K, int64_t D, float* c32, cudaStream_t stream) {
  U2B_CHECK_ARG(sums && c32 && K > 0 && D > 0, "kmeans_finalize: bad arguments");
  const long long n = static_cast<long long>(K) * D;
  kmeans_finalize_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      sums, c32, static_cast<int>(K), static_cast<int>(D));
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
