// otter_b200 — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../include/otter_b200.h"

namespace otb {

// thread-local error string behind otb_last_error()
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define OTB_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return otb::set_error(OTB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define OTB_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return otb::set_error(OTB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                                    \
  } while (0)

// Encode a 2D bf16 row-major tensor [rows][cols] (row pitch ld elements) as a TMA tensor map with a
// SWIZZLE_128B box of box_cols x box_rows (box_cols * 2 bytes must be 128).
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);

// Same for an fp32 tensor (box_cols * 4 bytes must be 128): output maps of the TMA-store / reduce-add epilogue.
int make_tmap_f32_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols);

int sm_count();   // of the CURRENT device (cached per device)
// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (kernel, device); thread-safe
cudaError_t ensure_dyn_smem(const void* func, int bytes);
template <typename F>
inline cudaError_t ensure_dyn_smem(F* kern, int bytes) { return ensure_dyn_smem(reinterpret_cast<const void*>(kern), bytes); }
bool pdl_enabled();   // OTB_PDL=1 enables programmatic dependent launch (measured neutral under graph replay; off by default)

// Launch with the programmatic-stream-serialization attribute (all otter_b200 kernels call pdl_wait()).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace otb
