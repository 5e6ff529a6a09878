// Helpers of the VGG19 perceptual loss between the convolutions (internal C++ interface; see vgg_ops.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace srg {

constexpr int kL1FeatBlocks = 1184;
// NCHW fp32 3-channel image -> [N][H][W][64] bf16 im2col of the 3x3 / pad 1 window: channel (kh*3 + kw)*3 + c (27 used)
int launch_unfold3(const float* src, int N, int H, int W, void* dst, cudaStream_t st);
// adjoint of launch_unfold3: gradient of the unfolded tensor -> NCHW fp32 image gradient (times scale)
int launch_fold3(const void* d_unf, int N, int H, int W, float scale, float* dimg, cudaStream_t st);
// nn.MaxPool2d(2, 2) on NHWC bf16 (C % 8 == 0): out [N][H/2][W/2][C]
int launch_maxpool2_forward(const void* x, int N, int H, int W, int C, void* out, cudaStream_t st);
// dx [N][H][W][C] = dy routed to the first maximum of every window (+ add, optional, same shape as dx)
int launch_maxpool2_backward(const void* x, const void* dy, const void* add, int N, int H, int W, int C, void* dx, cudaStream_t st);
// out[0] (+)= weight * mean|a - b|; grad (optional, bf16) = weight * grad_scale * sign(a - b) / n, zero where relu_mask && a <= 0
size_t l1_feat_scratch_bytes();
int launch_l1_feat(const void* a, const void* b, long long n, float weight, int accumulate, float grad_scale, int relu_mask,
                   void* grad, void* scratch, float* out, cudaStream_t st);

}  // namespace srg
