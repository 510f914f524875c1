// otter_b200 — host-side helpers: error channel, launch counter, TMA descriptor encoding.
#include "otb_host.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace otb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked: the driver entry point is resolved through the runtime at first use, so
// the library loads (and exports its symbols) on a machine without a GPU driver.
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(OTB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(OTB_ERR_INVALID, "TMA base not 16B aligned");
  if ((ld * 2) % 16 != 0) return set_error(OTB_ERR_INVALID, "TMA row pitch %llu elements not a multiple of 8",
                                            (unsigned long long)ld);
  if (box_cols * 2 != 128 || box_rows > 256) return set_error(OTB_ERR_INVALID, "bad TMA box");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(OTB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return OTB_OK;
}

int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(OTB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(OTB_ERR_INVALID, "TMA base not 16B aligned");
  if ((ld * 4) % 16 != 0) return set_error(OTB_ERR_INVALID, "TMA row pitch %llu fp32 elements not a multiple of 4",
                                            (unsigned long long)ld);
  if (box_cols * 4 != 128 || box_rows > 256) return set_error(OTB_ERR_INVALID, "bad TMA box");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(OTB_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return OTB_OK;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("OTB_PDL"); return (e && e[0] == '1'); }();
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace otb

extern "C" {
const char* otb_last_error(void) { return otb::g_err; }
int otb_version(void) { return 1; }
int otb_compiled_arch(void) { return 100; }
long long otb_launch_count(void) { return otb::g_launches.load(); }
int otb_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(otb_gemm_epilogue);
    case 1: return (int)sizeof(otb_attn_desc);
    case 2: return (int)sizeof(otb_attn_grads);
    case 3: return (int)sizeof(otb_lm_attn_desc);
    case 4: return (int)sizeof(otb_lm_attn_grads);
    default: return -1;
  }
}
}
