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
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace srg
