// Developer microbenchmark: issue rate of tcgen05.mma (bf16, M=128, K=16, operands in shared memory) for N = 64/128/256,
// and with the A operand re-used from the same shared-memory tile vs streamed over different tiles.
#include <stdio.h>
#include "../super_resolution-image-reconstructer-multi_generator_gan_b200/csrc/ptx.cuh"
using namespace srg;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters, int a_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t hi = (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
    const uint32_t a0 = smem_u32(smem) >> 4;                 // A tiles: 16 KB each (128 rows x 64 ch)
    const uint32_t b0 = smem_u32(smem + 96 * 1024) >> 4;     // B tile: N rows x 64 ch (<= 32 KB)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t ad = hi | uint64_t(a0 + uint32_t(it % a_tiles) * 1024u);
      const uint64_t bd = hi | uint64_t(b0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N>
void run(int a_tiles) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 2000;
  cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mma_rate_kernel<N><<<148, 128, 170 * 1024>>>(d, iters, a_tiles);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (iters * 4);
  printf("N=%3d a_tiles=%d: %s  %.1f cycles per M128xN%dxK16 MMA (floor %d) -> %.0f%% of the floor rate\n", N, a_tiles,
         cudaGetErrorString(e), per, N, 128 * N / 256, 100.0 * (128.0 * N / 256) / per);
  cudaFree(d);
}

int main() {
  run<64>(1); run<128>(1); run<192>(1); run<192>(6); run<256>(1); run<96>(1); run<160>(1); run<224>(1);
  return 0;
}
