/* libu2b200.so — C ABI of the B200-native U2Seg hot paths (sm_100a only).
 *
 * The reference (u2seg/U2Seg, a Detectron2 fork) has no native boundary on these paths: every
 * device op is reached through torch / torchvision / pykeops Python calls. Each entry point below
 * names the reference call site (file:line under the reference root) it replaces; the Python
 * host code in u2seg_b200/ mirrors those call sites' signatures and calls these functions with
 * raw device pointers (tensor.data_ptr()), shapes and the caller's CUDA stream.
 *
 * Conventions
 *  - return 0 on success; >0 = cudaError_t; <0 = library error (U2B_ERR_*). u2b_last_error()
 *    returns a thread-local message. No C++ exception crosses the boundary.
 *  - the caller owns every buffer (inputs, outputs, workspace); the library never allocates,
 *    frees or retains device pointers. All work is enqueued on `stream`; no hidden sync.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef U2B200_H_
#define U2B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U2B200_VERSION 100

typedef struct CUstream_st* u2b_stream_t; /* == cudaStream_t */

const char* u2b_last_error(void);
int u2b_version(void);
int u2b_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * k-means (Lloyd) — u2seg/Instance_Clustering/shared/utils/nn_utils.py:304-379 (KMeans)
 * x16: (N, D) fp16 row-major embeddings, D % 64 == 0, D <= 384. Centroids c32: (K, D) fp32.
 * ------------------------------------------------------------------------------------------ */

/* rows of the padded fp16 centroid operand (multiple of the 160-wide accumulator tile) */
int64_t u2b_kmeans_kpad(int64_t K);
size_t u2b_kmeans_workspace_bytes(int64_t N, int64_t D, int64_t K);

/* max_i |x_i| -> *xmax (device float). Once per data set; feeds the refinement bound. */
int u2b_kmeans_xnorm_max(const void* x16, int64_t N, int64_t D, float* xmax, u2b_stream_t stream);

/* c32 -> c16 (kpad x D fp16, zero rows beyond K), cnorm (kpad fp32, +inf beyond K),
 * *cmax2 = max finite |c|^2 (device float). Run before every E-step. */
int u2b_kmeans_prepare(const float* c32, int64_t K, int64_t D, void* c16, float* cnorm,
                       float* cmax2, u2b_stream_t stream);

/* E-step, nn_utils.py:353-355: labels[i] = argmin_j sum_d (x_i - c_j)^2 (first minimum; a NaN
 * centroid is never selected). tcgen05 fp16 pass + exact fp32 refinement of undecidable rows.
 * amb_count_out (device int, may be NULL) receives the number of refined rows. */
int u2b_kmeans_assign(const void* x16, int64_t N, int64_t D, int64_t K, const void* c16,
                      const float* c32, const float* cnorm, const float* xmax, const float* cmax2,
                      int32_t* labels, int32_t* amb_count_out, void* workspace,
                      size_t workspace_bytes, u2b_stream_t stream);

/* M-step part 1, nn_utils.py:359-363: sums[k, 0:D] = sum of rows with label k, sums[k, D] =
 * count (fp32). (K, D+1) fp32 — the buffer a row-sharded job all-reduces across ranks. */
int u2b_kmeans_accumulate(const void* x16, const int32_t* labels, int64_t N, int64_t D, int64_t K,
                          float* sums, void* workspace, size_t workspace_bytes,
                          u2b_stream_t stream);

/* M-step part 2, nn_utils.py:364: c32[k] = sums[k, 0:D] / sums[k, D] (NaN when empty). */
int u2b_kmeans_finalize(const float* sums, int64_t K, int64_t D, float* c32, u2b_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* U2B200_H_ */
