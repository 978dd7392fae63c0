// Box kernels of the detector path: fused pairwise-IoU + Matcher, and batched NMS with the
// sequential scan on the device (no host round trip).
//  * pairwise_iou  detectron2/structures/boxes.py:310-358
//  * Matcher       detectron2/modeling/matcher.py:62-127 (incl. set_low_quality_matches_)
//    The reference materialises the G x A IoU matrix (G*1 MB at A=261,888 plus ~6 temporaries);
//    here each thread owns one prediction, loops over the G ground-truth boxes held in shared
//    memory, and only the (A,) results are written. Low-quality matches need the per-GT maximum
//    over all predictions: pass 1 folds it with atomicMax on the fp32 bit pattern (IoU >= 0).
//  * batched_nms   detectron2/layers/nms.py:9-21 -> torchvision nms (class-by-class semantics)
#include "common.cuh"
#include "../../include/u2b200.h"

namespace {

constexpr int MAX_GT = 1024;

using ptx_free::iou_ref;

__global__ void __launch_bounds__(256)
iou_match_kernel(const float4* __restrict__ gt, int G, const uint8_t* __restrict__ gt_valid,
                 const float4* __restrict__ pred, int A, int64_t* __restrict__ matches,
                 float* __restrict__ matched_vals, unsigned int* __restrict__ gt_max_bits) {
  __shared__ float4 sgt[MAX_GT];
  __shared__ float sarea[MAX_GT];
  __shared__ unsigned int smax[MAX_GT];
  __shared__ uint8_t sval[MAX_GT];
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const float4 g = gt[i];
    sgt[i] = g;
    sarea[i] = (g.z - g.x) * (g.w - g.y);
    smax[i] = 0u;
    sval[i] = gt_valid ? gt_valid[i] : 1;
  }
  __syncthreads();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < A) {
    const float4 p = pred[a];
    const float parea = __fmul_rn(p.z - p.x, p.w - p.y);   // rounded on its own, as boxes.py:area()
    float best = -1.f;
    int bi = 0;
    for (int g = 0; g < G; ++g) {
      if (!sval[g]) continue;  // padded GT slot (fixed-capacity buffers): not a candidate
      const float v = iou_ref(sgt[g], sarea[g], p, parea);
      if (v > best) {  // first maximum, as torch.max(dim=0)
        best = v;
        bi = g;
      }
      if (gt_max_bits && v > 0.f) atomicMax(&smax[g], __float_as_uint(v));
    }
    matches[a] = bi;
    matched_vals[a] = best < 0.f ? 0.f : best;  // no valid GT at all: quality 0 (matcher.py:80-88 labels it background)
  }
  if (gt_max_bits) {
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x)
      if (smax[i]) atomicMax(&gt_max_bits[i], smax[i]);
  }
}

// labels from thresholds (matcher.py:96-101) + low-quality promotion (matcher.py:116-127)
__global__ void __launch_bounds__(256)
match_label_kernel(const float4* __restrict__ gt, int G, const uint8_t* __restrict__ gt_valid,
                   const float4* __restrict__ pred, int A, const float* __restrict__ matched_vals, const float* __restrict__ thresholds,
                   const int* __restrict__ labels, int nthr,
                   const unsigned int* __restrict__ gt_max_bits, int8_t* __restrict__ out_labels) {
  __shared__ float4 sgt[MAX_GT];
  __shared__ float sarea[MAX_GT];
  __shared__ float smax[MAX_GT];
  if (gt_max_bits) {
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
      const float4 g = gt[i];
      sgt[i] = g;
      sarea[i] = (g.z - g.x) * (g.w - g.y);
      // padded GT slots never promote anything: an unreachable maximum
      smax[i] = (gt_valid && !gt_valid[i]) ? -1.f : __uint_as_float(gt_max_bits[i]);
    }
    __syncthreads();
  }
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= A) return;
  const float v = matched_vals[a];
  int lab = 1;
  // thresholds = [-inf, t0, .., +inf] (nthr entries), labels nthr-1 entries; later intervals overwrite
  for (int i = 0; i + 1 < nthr; ++i)
    if (v >= thresholds[i] && v < thresholds[i + 1]) lab = labels[i];
  if (gt_max_bits) {
    const float4 p = pred[a];
    const float parea = __fmul_rn(p.z - p.x, p.w - p.y);   // rounded on its own, as boxes.py:area()
    for (int g = 0; g < G; ++g)
      if (iou_ref(sgt[g], sarea[g], p, parea) == smax[g]) {
        lab = 1;
        break;
      }
  }
  out_labels[a] = static_cast<int8_t>(lab);
}

// ---------------- NMS ----------------
// torchvision nms_kernel.cu devIoU: (inter / (Sa + Sb - inter)) > threshold
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  const float inter = width * height;
  const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
  return (inter / (sa + sb - inter)) > thr;
}

// boxes/cats already gathered in descending-score order. mask[i][cb] bit j: box (cb*64+j) is
// suppressed by box i (same category, IoU > thr, j after i).
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, const int64_t* __restrict__ cats, int n, float thr,
                unsigned long long* __restrict__ mask, int col_blocks) {
  const int row_start = blockIdx.y, col_start = blockIdx.x;
  if (row_start > col_start) return;
  const int row_size = min(n - row_start * 64, 64), col_size = min(n - col_start * 64, 64);
  __shared__ float4 sb[64];
  __shared__ int64_t sc[64];
  if (threadIdx.x < col_size) {
    sb[threadIdx.x] = boxes[col_start * 64 + threadIdx.x];
    sc[threadIdx.x] = cats ? cats[col_start * 64 + threadIdx.x] : 0;
  }
  __syncthreads();
  if (threadIdx.x < row_size) {
    const int cur = row_start * 64 + threadIdx.x;
    const float4 b = boxes[cur];
    const int64_t c = cats ? cats[cur] : 0;
    unsigned long long t = 0;
    const int start = (row_start == col_start) ? threadIdx.x + 1 : 0;
    for (int i = start; i < col_size; ++i)
      if (sc[i] == c && iou_gt(b, sb[i], thr)) t |= 1ULL << i;
    mask[static_cast<size_t>(cur) * col_blocks + col_start] = t;
  }
}

// Sequential part of greedy NMS, on the device. One CTA walks the sorted boxes in blocks of 64. The 64 mask rows of
// a block (only the columns >= the block's own, rows padded to an even number of words) are staged in shared memory
// by the TMA engine — one cp.async.bulk per row issued by a single thread, completion on an mbarrier, double
// buffered — so the load of block b+1 hides behind the work on block b at a cost of 64 instructions per block, and
// nothing on the serial chain touches global memory: thread 0 resolves the 64 intra-block decisions from the
// diagonal words (registers), then two threads per later column OR the kept rows' words into `removed[]`.
// Stops as soon as max_keep boxes are kept (proposal_utils.py:122 `keep[:post_nms_topk]`).
// developer instrumentation: cycles thread 0 spends per phase of the scan (u2b_debug_nms_profile); off by default
__device__ unsigned long long g_nms_prof[8];
__device__ int g_nms_prof_on = 0;

__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned long long* __restrict__ mask, const int64_t* __restrict__ order,
                const uint8_t* __restrict__ valid, int n, int col_blocks, int pitch, int max_keep,
                int64_t* __restrict__ keep, int* __restrict__ num_keep) {
  extern __shared__ __align__(16) unsigned long long nms_smem[];
  unsigned long long* tile0 = nms_smem;                                        // [2][64][pitch]
  const size_t tile_words = static_cast<size_t>(64) * pitch;
  unsigned long long* removed = nms_smem + 2 * tile_words;                     // [col_blocks]
  __shared__ __align__(8) uint64_t full[2];
  __shared__ unsigned long long s_kept;
  __shared__ int s_nk, s_stop;
  const int t = threadIdx.x;
  const int issuer = 32;            // warp 1 issues the copies; thread 0 (warp 0) runs the serial chain

  auto stage = [&](int b) {         // called by the issuing thread only
    const int rows = min(64, n - b * 64);
    const int c0 = b & ~1;          // 16-byte aligned first column
    const uint32_t row_bytes = static_cast<uint32_t>(pitch - c0) * 8u;
    unsigned long long* dst = tile0 + (b & 1) * tile_words;
    ptx::fence_proxy_async();
    ptx::mbar_arrive_expect_tx(&full[b & 1], row_bytes * rows);
    for (int i = 0; i < rows; ++i) {
      const unsigned long long* src = mask + (static_cast<size_t>(b) * 64 + i) * pitch + c0;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              ptx::smem_u32(dst + static_cast<size_t>(i) * pitch + c0)),
          "l"(src), "r"(row_bytes), "r"(ptx::smem_u32(&full[b & 1]))
          : "memory");
    }
  };

  if (t == issuer) {
    ptx::mbar_init(&full[0], 1);
    ptx::mbar_init(&full[1], 1);
    ptx::fence_barrier_init();
  }
  for (int i = t; i < col_blocks; i += blockDim.x) {
    unsigned long long r = 0ULL;
    if (valid)  // boxes flagged invalid (fixed-capacity buffers) start out removed
      for (int j = 0; j < 64 && i * 64 + j < n; ++j)
        if (!valid[order[i * 64 + j]]) r |= 1ULL << j;
    removed[i] = r;
  }
  if (t == 0) { s_nk = 0; s_stop = 0; }
  __syncthreads();
  if (t == issuer) stage(0);
  const bool prof = g_nms_prof_on != 0 && t == 0;
  long long pc[6] = {0, 0, 0, 0, 0, 0}, tp = prof ? clock64() : 0;
#define U2B_NMS_TICK(k)              \
  if (prof) {                        \
    const long long now = clock64(); \
    pc[k] += now - tp;               \
    tp = now;                        \
  }
  for (int b = 0; b < col_blocks; ++b) {
    const int rows = min(64, n - b * 64);
    const unsigned long long* tile = tile0 + (b & 1) * tile_words;
    if (t == issuer && b + 1 < col_blocks) stage(b + 1);    // buffer (b+1)&1 was released by the sync ending block b-1
    ptx::mbar_wait(&full[b & 1], (b >> 1) & 1);
    U2B_NMS_TICK(0)
    const int base = s_nk;
    if (t == 0) {
      // the 64 dependent decisions of this block. Only 32-bit halves sit on the dependent chain (3 ALU ops per
      // step: extract bit, make mask, OR-in the masked row): rows 0..31 are decided by the low half of `removed`,
      // rows 32..63 by the high half (a row's word has no bits at or below its own index). The max_keep cut is taken
      // afterwards: earlier decisions never depend on later ones.
      const unsigned long long rem0 = removed[b] | (rows < 64 ? ~0ULL << rows : 0ULL);   // rows past n: removed
      unsigned int rlo = static_cast<unsigned int>(rem0), rhi = static_cast<unsigned int>(rem0 >> 32);
      unsigned int klo = 0u, khi = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned long long d = tile[static_cast<size_t>(i) * pitch + b];
        const unsigned int m = ((rlo >> i) & 1u) - 1u;          // all ones when box i is still alive -> kept
        rlo |= static_cast<unsigned int>(d) & m;
        rhi |= static_cast<unsigned int>(d >> 32) & m;
        klo |= m & (1u << i);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned int dh = static_cast<unsigned int>(tile[static_cast<size_t>(32 + i) * pitch + b] >> 32);
        const unsigned int m = ((rhi >> i) & 1u) - 1u;
        rhi |= dh & m;
        khi |= m & (1u << i);
      }
      unsigned long long kw = (static_cast<unsigned long long>(khi) << 32) | klo;
      const int room = max_keep - base;
      if (__popcll(kw) >= room) {                // reached the cap inside this block: keep the first `room` only
        while (__popcll(kw) > room) kw &= ~(1ULL << (63 - __clzll(static_cast<long long>(kw))));
        s_stop = 1;
      }
      s_kept = kw;
    }
    U2B_NMS_TICK(1)
    __syncthreads();
    U2B_NMS_TICK(2)
    const unsigned long long kw = s_kept;
    if (!s_stop) {
      // two threads per later column: warps 0-3 OR the kept rows 0..31, warps 4-7 rows 32..63. All 32 row words are
      // loaded first (independent, back to back) and masked with the kept bits afterwards: no load latency is
      // exposed per row and no divergence.
      const int half = t >> 7;
      const unsigned int bits = static_cast<unsigned int>(kw >> (32 * half));
      const unsigned long long* trow = tile + static_cast<size_t>(32 * half) * pitch;
      for (int j = b + 1 + (t & 127); j < col_blocks; j += 128) {
        unsigned long long v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = trow[static_cast<size_t>(i) * pitch + j];
        unsigned long long acc = 0ULL;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc |= ((bits >> i) & 1u) ? v[i] : 0ULL;
        if (acc) atomicOr(&removed[j], acc);
      }
    }
    if (t < rows && ((kw >> t) & 1ULL))   // kept boxes of this block written in parallel, in order
      keep[base + __popcll(kw & ((1ULL << t) - 1ULL))] = order[b * 64 + t];
    if (s_stop) break;
    if (t == 0) s_nk = base + __popcll(kw);
    U2B_NMS_TICK(3)
    __syncthreads();  // removed[] complete, s_nk published, all reads of tile b done
    U2B_NMS_TICK(4)
  }
#undef U2B_NMS_TICK
  __syncthreads();
  if (t == 0) *num_keep = s_stop ? s_nk + __popcll(s_kept) : s_nk;
  if (prof)
    for (int i = 0; i < 6; ++i) g_nms_prof[i] += static_cast<unsigned long long>(pc[i]);
}

__global__ void gather_sorted_kernel(const float4* __restrict__ boxes, const int64_t* __restrict__ cats,
                                     const int64_t* __restrict__ order, int n,
                                     float4* __restrict__ sboxes, int64_t* __restrict__ scats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t o = order[i];
  sboxes[i] = boxes[o];
  if (cats) scats[i] = cats[o];
}

}  // namespace

extern "C" {

// developer tool: enable != 0 switches the scan's phase counters on; out6 (host, nullable) receives and clears the
// cycles accumulated so far: [wait tile, serial chain, barrier 1, keep + OR phase, barrier 2, unused]
int u2b_debug_nms_profile(int enable, uint64_t* out6) {
  unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  U2B_CUDA(cudaDeviceSynchronize());
  if (out6) {
    U2B_CUDA(cudaMemcpyFromSymbol(h, g_nms_prof, sizeof(h)));
    for (int i = 0; i < 6; ++i) out6[i] = h[i];
  }
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  U2B_CUDA(cudaMemcpyToSymbol(g_nms_prof, z, sizeof(z)));
  U2B_CUDA(cudaMemcpyToSymbol(g_nms_prof_on, &enable, sizeof(int)));
  return 0;
}

size_t u2b_nms_workspace_bytes(int64_t n) {
  const size_t pitch = ((static_cast<size_t>((n + 63) / 64) + 1) / 2) * 2;      // mask row pitch: even number of words
  return static_cast<size_t>(n) * 16 + static_cast<size_t>(n) * 8 + 16 + static_cast<size_t>(n) * pitch * 8 + 256;
}

// boxes (n,4) fp32 xyxy, cats (n) int64 or NULL, order (n) int64 = indices sorted by score descending
// (stable). keep (n) int64, num_keep device int. Semantics: torchvision batched_nms, class by class.
int u2b_batched_nms(const float* boxes, const int64_t* cats, const int64_t* order, const uint8_t* valid, int64_t n,
                    float iou_threshold, int64_t max_keep, int64_t* keep, int32_t* num_keep, void* workspace,
                    size_t workspace_bytes, cudaStream_t stream) {
  U2B_CHECK_ARG(num_keep, "batched_nms: num_keep is NULL");
  if (n == 0) {
    U2B_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    return 0;
  }
  U2B_CHECK_ARG(boxes && order && keep && workspace && n > 0, "batched_nms: bad arguments");
  U2B_CHECK_ARG(workspace_bytes >= u2b_nms_workspace_bytes(n), "batched_nms: workspace too small");
  const int cb = static_cast<int>((n + 63) / 64);
  const int pitch = (cb + 1) / 2 * 2;
  const size_t smem = (static_cast<size_t>(pitch) * 128 + cb) * 8;   // 2 tiles of 64 x pitch words + removed[cb]
  U2B_CHECK_ARG(smem <= 220 * 1024, "batched_nms: n=%lld too large (<= %d boxes)", (long long)n, 216 * 64);
  U2B_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "batched_nms: workspace must be 16-byte aligned");
  uint8_t* w = static_cast<uint8_t*>(workspace);
  float4* sboxes = reinterpret_cast<float4*>(w);
  int64_t* scats = reinterpret_cast<int64_t*>(w + static_cast<size_t>(n) * 16);
  unsigned long long* mask =
      reinterpret_cast<unsigned long long*>(w + (static_cast<size_t>(n) * 24 + 15) / 16 * 16);    // 16-byte aligned rows
  gather_sorted_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(boxes), cats, order, (int)n, sboxes, scats);
  U2B_LAUNCH_CHECK();
  // lower-triangular blocks are never read by the scan (j starts at i>>6), no memset needed
  dim3 grid(cb, cb);
  nms_mask_kernel<<<grid, 64, 0, stream>>>(sboxes, cats ? scats : nullptr, (int)n, iou_threshold, mask, pitch);
  U2B_LAUNCH_CHECK();
  static bool scan_attr = false;
  if (!scan_attr) {
    U2B_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    scan_attr = true;
  }
  const int mk = (max_keep < 0 || max_keep > n) ? (int)n : (int)max_keep;
  if (mk == 0) {
    U2B_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int32_t), stream));
    return 0;
  }
  nms_scan_kernel<<<1, 256, smem, stream>>>(mask, order, valid, (int)n, cb, pitch, mk, keep, num_keep);
  U2B_LAUNCH_CHECK();
  return 0;
}

// gt (G,4), pred (A,4) fp32. thresholds: nthr floats on the device [-inf, t.., +inf]; labels: nthr-1
// ints on the device. matches int64 (A), matched_vals fp32 (A), out_labels int8 (A).
// gt_max_scratch: G uint32 device scratch, required iff allow_low_quality.
int u2b_iou_match(const float* gt, int64_t G, const uint8_t* gt_valid, const float* pred, int64_t A,
                  const float* thresholds, const int32_t* labels, int nthr, int allow_low_quality,
                  int64_t* matches, float* matched_vals, int8_t* out_labels, uint32_t* gt_max_scratch,
                  cudaStream_t stream) {
  if (A == 0) return 0;
  U2B_CHECK_ARG(pred && thresholds && labels && matches && matched_vals && out_labels && nthr >= 2,
                "iou_match: bad arguments");
  U2B_CHECK_ARG(G > 0 && G <= MAX_GT, "iou_match: G=%lld outside 1..%d (empty GT is handled by the caller)",
                (long long)G, MAX_GT);
  U2B_CHECK_ARG(!allow_low_quality || gt_max_scratch, "iou_match: low-quality matching needs scratch");
  const unsigned grid = static_cast<unsigned>((A + 255) / 256);
  if (allow_low_quality) U2B_CUDA(cudaMemsetAsync(gt_max_scratch, 0, sizeof(uint32_t) * G, stream));
  iou_match_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(gt), (int)G, gt_valid,
                                             reinterpret_cast<const float4*>(pred), (int)A, matches,
                                             matched_vals, allow_low_quality ? gt_max_scratch : nullptr);
  U2B_LAUNCH_CHECK();
  match_label_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(gt), (int)G, gt_valid,
                                               reinterpret_cast<const float4*>(pred), (int)A, matched_vals,
                                               thresholds, labels, nthr,
                                               allow_low_quality ? gt_max_scratch : nullptr, out_labels);
  U2B_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
