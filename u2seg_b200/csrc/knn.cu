// Exact k-nearest-neighbour search (squared L2) over dense embeddings, B200-native.
// Reference: u2seg/Instance_Clustering/shared/utils/nn_utils.py:203-224 kNN() - KeOps
//   D_ij = ((x_test_i - x_train_j)^2).sum(-1);  d_knn, ind_knn = D_ij.Kmin_argKmin(K, dim=1)
// (the K smallest distances of every test row, ascending, and their train indices) and :227-299 partitioned_kNN(), the
// density-peak selection's neighbour search (K = 20) - 2*N^2*D flop, the most expensive step of the clustering pipeline.
//
// Design (same shape as the k-means E-step, csrc/kmeans.cu):
//   1. candidate pass on tcgen05: |x-y|^2 = |x|^2 + |y|^2 - 2 x.y; the N1 x N2 dot products run as fp16 UMMAs with fp32
//      accumulation in TMEM - a 128-row query tile resident in shared memory, train tiles of 160 rows streamed through a
//      TMA ring (multicast across a cluster), two accumulators so the epilogue overlaps the next tile's MMAs. The
//      epilogue forms |y|^2 - 2 acc and keeps, per (row, column half), the KC = 24 smallest values and their indices in
//      sorted shared-memory lists (insertion only when a value beats the list's worst: ~K ln(N2/K) times per row);
//   2. exact pass: one warp per query takes the K-th smallest candidate value, recomputes sum_d (x-y)^2 in fp32 (the
//      reference's formula) for the candidates within 2 eps of it, selects the K smallest (ties: lower index), and
//      certifies the result: the fp16 rounding of the candidate pass perturbs a value by at most eps - computed on the host
//      from the MEASURED rounding-error norms of the operands, |x.y - x~.y~| <= |x| |y - y~| + |x - x~| |y~| - so if the
//      exact K-th distance + eps is below every list's worst kept value (in distance space, minus eps) no discarded train
//      row can belong to the answer. Rows that cannot be certified are flagged; the host (u2seg_b200/clustering.py) gives
//      them an fp32 pass over all train rows (128 candidates, bound 4 D 2^-24 |x||y|) and, for exact ties beyond that, an
//      exhaustive search, so the result is always the exact fp32 answer.
#include <cuda_fp16.h>

#include "../../include/u2b200.h"
#include "common.cuh"

namespace {

constexpr int BM = 128;   // query rows per tile (UMMA M)
constexpr int NT = 160;   // train rows per accumulator tile (UMMA N)
constexpr int BK = 64;
constexpr int MAXKB = 6;  // D <= 384
constexpr int BSTAGES = 4;
constexpr int A_KB_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = NT * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int KNN_THREADS = 128 + EPI_WARPS * 32;  // 384
constexpr int TMEM_COLS = 512;
constexpr int HALF_N = NT / 2;
constexpr int KC = 24;    // candidates kept per (row, column half)
constexpr int EPI_T = EPI_WARPS * 32;

struct KnnSmem {
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_OFF + MAXKB * A_KB_BYTES;            // 98304
  static constexpr int VAL_OFF = B_OFF + BSTAGES * B_STAGE_BYTES;     // 180224
  static constexpr int IDX_OFF = VAL_OFF + KC * EPI_T * 4;            // + 24576
  static constexpr int BAR_OFF = IDX_OFF + KC * EPI_T * 4;            // 229376
  static constexpr int NBARS = 2 * MAXKB + 2 * BSTAGES + 4;
  static constexpr int TMEMPTR_OFF = BAR_OFF + NBARS * 8;
  static constexpr int BYTES = TMEMPTR_OFF + 16 + 1024;
  static_assert(BYTES <= 232448, "shared memory budget");
};

// Sorted insertion of (d, col) into one thread's list (entries KC apart by LIST_PITCH bytes, ascending; an equal value
// goes behind the earlier column). Returns the list's new worst value. Shared-space addresses; deliberately not inlined.
constexpr int LIST_PITCH = EPI_T * 4;
__device__ __noinline__ float knn_list_insert(uint32_t val0, uint32_t idx0, float d, int col) {
  int k = KC - 1;
#pragma unroll 1
  while (k > 0) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(val0 + (k - 1) * LIST_PITCH));
    if (!(v > d)) break;
    int i;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(i) : "r"(idx0 + (k - 1) * LIST_PITCH));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(val0 + k * LIST_PITCH), "f"(v) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(idx0 + k * LIST_PITCH), "r"(i) : "memory");
    --k;
  }
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(val0 + k * LIST_PITCH), "f"(d) : "memory");
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(idx0 + k * LIST_PITCH), "r"(col) : "memory");
  float worst;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(worst) : "r"(val0 + (KC - 1) * LIST_PITCH));
  return worst;
}

template <int CL>
__global__ void __launch_bounds__(KNN_THREADS, 1)
knn_candidates_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                      const float* __restrict__ ynorm, int32_t* __restrict__ cand_idx, float* __restrict__ cand_val,
                      float* __restrict__ cand_thr, int N1, int ntiles_n, int kblocks, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem + KnnSmem::A_OFF;
  uint8_t* sB = smem + KnnSmem::B_OFF;
  float* sVal = reinterpret_cast<float*>(smem + KnnSmem::VAL_OFF);   // [KC][EPI_T], ascending along KC
  int* sIdx = reinterpret_cast<int*>(smem + KnnSmem::IDX_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KnnSmem::BAR_OFF);
  uint64_t* A_full = bars;
  uint64_t* A_empty = bars + MAXKB;
  uint64_t* B_full = bars + 2 * MAXKB;
  uint64_t* B_empty = B_full + BSTAGES;
  uint64_t* T_full = B_empty + BSTAGES;
  uint64_t* T_empty = T_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + KnnSmem::TMEMPTR_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_y);
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
    if (ptx::elect_one()) {
      uint32_t bstage = 0, bphase = 0;
      int it = 0;
      for (int grp = cluster_id; grp < num_groups; grp += num_clusters, ++it) {
        int tile = grp * CL + rank;
        if (tile >= num_tiles) tile = num_tiles - 1;
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
              ptx::tma_load_2d(sB + bstage * B_STAGE_BYTES, &tmap_y, &B_full[bstage], kb * BK, n * NT);
            else
              ptx::tma_load_2d_mc(sB + bstage * B_STAGE_BYTES + rank * (PART_ROWS * BK * 2), &tmap_y, &B_full[bstage],
                                  kb * BK, n * NT + rank * PART_ROWS, kMask);
            if (++bstage == BSTAGES) {
              bstage = 0;
              bphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(BM, NT, /*fp16*/ 0);
      const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB);
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
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int te = threadIdx.x - 128;            // 0..255: this thread's list column
    const int row_in_tile = q * 32 + lane;
    const uint32_t list_val = ptx::smem_u32(sVal + te), list_idx = ptx::smem_u32(sIdx + te);
    uint32_t acc_it = 0;
    for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
      const int tile = grp * CL + rank;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        sVal[k * EPI_T + te] = __int_as_float(0x7f800000);
        sIdx[k * EPI_T + te] = -1;
      }
      float thr = __int_as_float(0x7f800000);
      for (int n = 0; n < ntiles_n; ++n, ++acc_it) {
        const uint32_t buf = acc_it & 1, tphase = (acc_it >> 1) & 1;
        ptx::mbar_wait(&T_full[buf], tphase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * NT + half * HALF_N;
        uint32_t r[HALF_N / 16][16];
#pragma unroll
        for (int c = 0; c < HALF_N / 16; ++c) ptx::tmem_ld16(taddr + c * 16, r[c]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&T_empty[buf]);
        const int jbase = n * NT + half * HALF_N;
        // Hot path: 4 FFMA + 3 FMNMX + 1 compare per 4 columns, one (rarely taken) branch. The insertion is a real call:
        // inlined and unrolled at all 80 sites it made 240 KB of code whose branch targets missed the instruction
        // cache on every element (23 k cycles per tile instead of ~2 k).
#pragma unroll
        for (int c = 0; c < HALF_N / 16; ++c) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 yn = __ldg(reinterpret_cast<const float4*>(ynorm + jbase + c * 16 + j4 * 4));
            const float d0 = fmaf(-2.0f, __uint_as_float(r[c][j4 * 4 + 0]), yn.x);
            const float d1 = fmaf(-2.0f, __uint_as_float(r[c][j4 * 4 + 1]), yn.y);
            const float d2 = fmaf(-2.0f, __uint_as_float(r[c][j4 * 4 + 2]), yn.z);
            const float d3 = fmaf(-2.0f, __uint_as_float(r[c][j4 * 4 + 3]), yn.w);
            if (fminf(fminf(d0, d1), fminf(d2, d3)) < thr) {
              const int col = jbase + c * 16 + j4 * 4;
              if (d0 < thr) thr = knn_list_insert(list_val, list_idx, d0, col);
              if (d1 < thr) thr = knn_list_insert(list_val, list_idx, d1, col + 1);
              if (d2 < thr) thr = knn_list_insert(list_val, list_idx, d2, col + 2);
              if (d3 < thr) thr = knn_list_insert(list_val, list_idx, d3, col + 3);
            }
          }
        }
      }
      const long long row = static_cast<long long>(tile) * BM + row_in_tile;
      if (tile < num_tiles && row < N1) {
        int32_t* ci = cand_idx + row * (2 * KC) + half * KC;
        float* cv = cand_val + row * (2 * KC) + half * KC;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          ci[k] = sIdx[k * EPI_T + te];
          cv[k] = sVal[k * EPI_T + te];
        }
        cand_thr[row * 2 + half] = thr;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// One warp per query: exact fp32 distances of the 2*KC candidates, K smallest (ascending; ties: lower index), certificate.
template <int NP>   // D = 128 * NP
__global__ void __launch_bounds__(256)
knn_refine_kernel(const float* __restrict__ x, const float* __restrict__ y, const int32_t* __restrict__ cand_idx,
                  const float* __restrict__ cand_val, const float* __restrict__ cand_thr,
                  const float* __restrict__ xnorm, long long N1, int K, float eps,
                  float* __restrict__ d_out, int64_t* __restrict__ i_out, int32_t* __restrict__ flagged,
                  int* __restrict__ n_flagged) {
  constexpr int D = 128 * NP;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N1) return;
  float4 xv[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) xv[j] = *reinterpret_cast<const float4*>(x + row * D + j * 128 + lane * 4);
  const int32_t* ci = cand_idx + row * (2 * KC);
  // Only candidates that can still belong to the answer are re-evaluated exactly: with |approx - true| <= eps for every
  // pair, a candidate whose approximate value exceeds the K-th smallest approximate value by more than 2 eps is farther
  // than the true K-th neighbour. (Typically ~K + a few of the 48 survive: the exact pass is a random 1.5 KB gather per
  // surviving candidate and would otherwise dominate the search.)
  float va = lane < 2 * KC ? cand_val[row * (2 * KC) + lane] : __int_as_float(0x7f800000);
  float vb = lane + 32 < 2 * KC ? cand_val[row * (2 * KC) + lane + 32] : __int_as_float(0x7f800000);
  const float keep_a = va, keep_b = vb;
  float vK = __int_as_float(0x7f800000);
  for (int k = 0; k < K; ++k) {   // K-th smallest approximate value: K rounds of warp-min with removal
    const float m = fminf(va, vb);
    float best = m;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, off));
    const unsigned who = __ballot_sync(0xffffffffu, m == best);
    if (lane == __ffs(who) - 1) {   // one holder retires one entry
      if (va == best) va = __int_as_float(0x7f800000);
      else vb = __int_as_float(0x7f800000);
    }
    vK = best;
  }
  const float tau = vK + 2.f * eps;
  float myd[2] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000)};
  int myi[2] = {0x7fffffff, 0x7fffffff};
#pragma unroll 4
  for (int c = 0; c < 2 * KC; ++c) {
    const float vc = __shfl_sync(0xffffffffu, (c < 32) ? keep_a : keep_b, c & 31);
    const int j = (vc <= tau) ? ci[c] : -1;
    float d = __int_as_float(0x7f800000);
    if (j >= 0) {   // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const float4 yv = *reinterpret_cast<const float4*>(y + static_cast<long long>(j) * D + p * 128 + lane * 4);
        const float a = xv[p].x - yv.x, b = xv[p].y - yv.y, cc = xv[p].z - yv.z, dd = xv[p].w - yv.w;
        s = fmaf(a, a, s); s = fmaf(b, b, s); s = fmaf(cc, cc, s); s = fmaf(dd, dd, s);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      d = s;
    }
    if (lane == (c & 31)) {
      myd[c >> 5] = d;
      myi[c >> 5] = j >= 0 ? j : 0x7fffffff;
    }
  }
  // K rounds of warp arg-min over (distance bits, index): distances are >= 0, so their bit patterns order like floats
  float dK = 0.f;
  for (int k = 0; k < K; ++k) {
    const int s0 = (myd[0] < myd[1] || (myd[0] == myd[1] && myi[0] <= myi[1])) ? 0 : 1;
    unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(myd[s0])) << 32) | static_cast<unsigned int>(myi[s0]);
    unsigned long long best = key;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
      best = o < best ? o : best;
    }
    if (best == key && myi[s0] != 0x7fffffff) {   // the winner retires its entry (keys are unique: indices differ)
      myd[s0] = __int_as_float(0x7f800000);
      myi[s0] = 0x7fffffff;
    }
    const float bd = __uint_as_float(static_cast<unsigned int>(best >> 32));
    if (lane == 0) {
      d_out[row * K + k] = bd;
      i_out[row * K + k] = static_cast<int64_t>(static_cast<unsigned int>(best & 0xffffffffu));
    }
    dK = bd;
  }
  if (lane == 0) {
    // discarded rows of half h have candidate-pass value >= thr_h, i.e. true distance >= thr_h + |x|^2 - eps
    const float xn = xnorm[row];
    const float bound = fminf(cand_thr[row * 2], cand_thr[row * 2 + 1]) + xn - eps;
    if (!(dK + eps < bound)) flagged[atomicAdd(n_flagged, 1)] = static_cast<int32_t>(row);
  }
}

int g_knn_cluster = 2;

template <int CL>
int launch_knn(const CUtensorMap& tx, const CUtensorMap& ty, const float* ynorm, int32_t* cand_idx, float* cand_val,
               float* cand_thr, int N1, int ntiles_n, int kblocks, int num_tiles, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    U2B_CUDA(cudaFuncSetAttribute(knn_candidates_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, KnnSmem::BYTES));
    attr = true;
  }
  const int groups = (num_tiles + CL - 1) / CL;
  int clusters = u2b_num_sms() / CL;
  if (clusters > groups) clusters = groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(KNN_THREADS);
  cfg.dynamicSmemBytes = KnnSmem::BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  U2B_CUDA(cudaLaunchKernelEx(&cfg, knn_candidates_kernel<CL>, tx, ty, ynorm, cand_idx, cand_val, cand_thr, N1, ntiles_n,
                              kblocks, num_tiles));
  return 0;
}

}  // namespace

extern "C" {

int u2b_knn_candidates_per_row(void) { return 2 * KC; }
int64_t u2b_knn_npad(int64_t N2) { return ceil_div64(N2, NT) * NT; }

// Candidate pass. x16 (N1, D) / y16 (npad(N2), D) fp16 row-major (rows beyond N2 zero), ynorm (npad(N2)) fp32 = |y|^2 of
// the fp32 rows, +inf beyond N2 (u2b_kmeans_prepare produces both). D % 64 == 0, D <= 384.
// cand_idx (N1, 48) int32 (-1 = empty slot), cand_val (N1, 48) fp32 = their values in |y|^2 - 2 x.y space (+inf = empty),
// cand_thr (N1, 2) fp32 = worst kept value of each of the two lists.
int u2b_knn_candidates(const void* x16, int64_t N1, const void* y16, const float* ynorm, int64_t N2, int64_t D,
                       int32_t* cand_idx, float* cand_val, float* cand_thr, cudaStream_t stream) {
  U2B_CHECK_ARG(x16 && y16 && ynorm && cand_idx && cand_val && cand_thr && N1 > 0 && N2 > 0, "knn_candidates: bad arguments");
  U2B_CHECK_ARG(N1 < (1LL << 31) && N2 < (1LL << 31), "knn_candidates: more than 2^31 rows");
  if (D % BK != 0 || D / BK > MAXKB) {
    u2b_set_error("knn_candidates: D=%lld unsupported (need D %% 64 == 0 and D <= %d)", (long long)D, MAXKB * BK);
    return U2B_ERR_UNSUPPORTED;
  }
  const int64_t npad = u2b_knn_npad(N2);
  const int num_tiles = static_cast<int>(ceil_div64(N1, BM));
  CUtensorMap tx, ty;
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)N1};
    uint64_t strides[1] = {(uint64_t)D * 2};
    uint32_t box[2] = {BK, BM};
    int rc = u2b_encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x16, dims, strides, box, nullptr,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  int CL = g_knn_cluster;
  if (num_tiles < CL) CL = 1;
  {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)npad};
    uint64_t strides[1] = {(uint64_t)D * 2};
    uint32_t box[2] = {BK, (uint32_t)(NT / CL)};
    int rc = u2b_encode_tmap(&ty, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y16, dims, strides, box, nullptr,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) return rc;
  }
  const int ntn = static_cast<int>(npad / NT), kb = static_cast<int>(D / BK);
  if (CL == 4) return launch_knn<4>(tx, ty, ynorm, cand_idx, cand_val, cand_thr, (int)N1, ntn, kb, num_tiles, stream);
  if (CL == 2) return launch_knn<2>(tx, ty, ynorm, cand_idx, cand_val, cand_thr, (int)N1, ntn, kb, num_tiles, stream);
  return launch_knn<1>(tx, ty, ynorm, cand_idx, cand_val, cand_thr, (int)N1, ntn, kb, num_tiles, stream);
}

// Exact pass. x (N1, D), y (N2, D) fp32; xnorm (N1) fp32; eps = rounding bound of the candidate pass in distance units.
// d_out (N1, K) fp32 ascending, i_out (N1, K) int64; flagged (N1) int32 + n_flagged (device int, zeroed by the caller)
// list the rows whose result could not be certified. D in {128, 256, 384}; K <= 2*KC.
int u2b_knn_refine(const float* x, const float* y, const int32_t* cand_idx, const float* cand_val, const float* cand_thr,
                   const float* xnorm, int64_t N1, int64_t D, int K, float eps, float* d_out, int64_t* i_out,
                   int32_t* flagged, int32_t* n_flagged, cudaStream_t stream) {
  U2B_CHECK_ARG(x && y && cand_idx && cand_val && cand_thr && xnorm && d_out && i_out && flagged && n_flagged && N1 > 0,
                "knn_refine: bad arguments");
  U2B_CHECK_ARG(K > 0 && K <= 2 * KC, "knn_refine: K=%d outside 1..%d", K, 2 * KC);
  const unsigned grid = static_cast<unsigned>((N1 + 7) / 8);
  if (D == 128) knn_refine_kernel<1><<<grid, 256, 0, stream>>>(x, y, cand_idx, cand_val, cand_thr, xnorm, N1, K, eps, d_out, i_out, flagged, n_flagged);
  else if (D == 256) knn_refine_kernel<2><<<grid, 256, 0, stream>>>(x, y, cand_idx, cand_val, cand_thr, xnorm, N1, K, eps, d_out, i_out, flagged, n_flagged);
  else if (D == 384) knn_refine_kernel<3><<<grid, 256, 0, stream>>>(x, y, cand_idx, cand_val, cand_thr, xnorm, N1, K, eps, d_out, i_out, flagged, n_flagged);
  else {
    u2b_set_error("knn_refine: D=%lld unsupported (128, 256 or 384)", (long long)D);
    return U2B_ERR_UNSUPPORTED;
  }
  U2B_LAUNCH_CHECK();
  return 0;
}

int u2b_knn_set_cluster(int cluster) {
  U2B_CHECK_ARG(cluster == 1 || cluster == 2 || cluster == 4, "knn_set_cluster: 1, 2 or 4");
  g_knn_cluster = cluster;
  return 0;
}

}  // extern "C"
