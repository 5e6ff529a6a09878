// Programmatic dependent launch (PDL) helpers.
//
// A kernel launched through launch_pdl() may be scheduled while its stream predecessor is still running (as soon as
// every block of the predecessor has executed pdl_trigger(), or exited).  It MUST call pdl_wait() before it touches
// memory the predecessor writes: pdl_wait() returns once the predecessor grid has completed and flushed.  Everything a
// kernel does before pdl_wait() (barrier init, TMEM allocation, tensor-map prefetch, loading weights that were packed
// long before) overlaps the predecessor's tail; inside a captured CUDA graph the attribute becomes a programmatic edge.
// Only kernels that contain pdl_wait() are ever launched with the attribute.
#pragma once
#include <cuda_runtime.h>

#include <stdlib.h>

#include <utility>

namespace srg {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  // SRG_PDL=0 (developer switch): plain stream-order launches everywhere, so that a kernel timeline shows each kernel's own
  // duration instead of a duration that starts while the predecessor is still running
  // SRG_PDL=2: only grids of at most 8 blocks (the single-block finalize kernels) start early: a big grid that is parked at
  // griddepcontrol.wait holds registers that another graph branch's ready kernel could use
  static int pdl_on = -1;
  if (pdl_on < 0) { const char* e = getenv("SRG_PDL"); pdl_on = e ? atoi(e) : 1; }
  cfg.numAttrs = (pdl_on == 1 || (pdl_on == 2 && grid.x * grid.y * grid.z <= 8)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// A/B switches (read once): SRG_PDL_CONV=0 launches the SM-exclusive convolution kernels WITHOUT the attribute (a CTA that
// is resident early holds ~190 KB of shared memory while it waits for its predecessor, which keeps another graph branch's
// ready convolution off that SM); SRG_PDL_EW_LATE=1 makes the elementwise producers of a convolution operand trigger their
// dependents at the END of their loop instead of at the start (the convolution still pre-pays its launch latency).
inline bool pdl_conv() { static int v = -1; if (v < 0) { const char* e = getenv("SRG_PDL_CONV"); v = (e && e[0] == '0') ? 0 : 1; } return v != 0; }
inline int pdl_ew_late() { static int v = -1; if (v < 0) { const char* e = getenv("SRG_PDL_EW_LATE"); v = (e && e[0] == '1') ? 1 : 0; } return v; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_opt_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args&&... args) {
  if (pdl) return launch_pdl(kernel, grid, block, smem, st, std::forward<Args>(args)...);
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace srg
