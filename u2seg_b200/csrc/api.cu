// libu2b200.so: library-wide entry points (version, error string) and host utilities.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "../../include/u2b200.h"

static thread_local char g_err[512] = "";

void u2b_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int u2b_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Persistent kernels (conv2, conv_wgrad2) assign tiles statically to one cluster per SM pair. When another kernel holds SMs
// for the whole duration (NCCL's all-reduce CTAs during the overlapped backward pass), clusters that cannot be co-resident
// run as a second wave and double the kernel's time; a budget below the SM count leaves room for them.
static int g_sm_budget = 0;
int u2b_persistent_sms() {
  const int n = u2b_num_sms();
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}
extern "C" int u2b_set_sm_budget(int sms) {
  if (sms < 0 || (sms > 0 && sms < 2)) {
    u2b_set_error("set_sm_budget: %d (0 = all SMs, otherwise >= 2)", sms);
    return U2B_ERR_BAD_ARG;
  }
  g_sm_budget = sms & ~1;
  return 0;
}

static int g_pdl = 1;
int u2b_pdl_enabled() { return g_pdl; }
extern "C" int u2b_set_pdl(int on) {
  g_pdl = on ? 1 : 0;
  return 0;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

int u2b_encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* gaddr,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    const uint32_t* elem_strides, CUtensorMapSwizzle swizzle,
                    CUtensorMapFloatOOBfill oob) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      u2b_set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return U2B_ERR_DRIVER;
    }
    fn = (PFN_tmapEncodeTiled)p;
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(gaddr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, oob);
  if (r != CUDA_SUCCESS) {
    u2b_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dim0 %llu box0 %u)",
                  (int)r, rank, (unsigned long long)dims[0], box[0]);
    return U2B_ERR_DRIVER;
  }
  return 0;
}

extern "C" {
const char* u2b_last_error(void) { return g_err; }
int u2b_version(void) { return U2B200_VERSION; }
int u2b_sm_count(void) { return u2b_num_sms(); }
}
