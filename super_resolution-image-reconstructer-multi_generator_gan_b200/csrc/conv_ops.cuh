// Per-operator convolution launchers behind srg_conv2d_* (internal C++ interface; see conv_ops.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace srg {

// bf16 elements of a packed k x k filter (fprop and dgrad copies have the same size)
size_t conv2d_packed_elems(int cout, int cin, int k);
// fp32 OIHW [cout][cin][k][k] -> packed bf16 operand; dgrad != 0: the transposed, spatially flipped copy dgrad convolves with
int launch_conv2d_pack(const float* w_oihw, int cout, int cin, int k, int dgrad, void* packed, cudaStream_t st);
// out[N,H,W,n_out] = act(conv_k(x[N,H,W,k_ch]) + bias) (+ residual | zeroed where mask_src <= 0); k in {1, 3}, pad k/2
int launch_conv2d(const void* x, int N, int H, int W, int k_ch, const void* packed, int n_out, int k, const float* bias, int act,
                  float slope, const void* residual, const void* mask_src, void* out, cudaStream_t st);
size_t conv2d_wgrad_workspace_bytes(int N, int H, int W, int cin, int cout);
// 3x3: dw_oihw[cout][cin][3][3] = sum_p x[p + tap] * dy[p]; dbias[cout] (optional) = sum_p dy[p]
int launch_conv2d_wgrad(const void* x, const void* dy, int N, int H, int W, int cin, int cout, void* workspace,
                        size_t workspace_bytes, float* dw_oihw, float* dbias, cudaStream_t st);

}  // namespace srg
