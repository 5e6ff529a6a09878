// Pillow-exact antialiased resize of 8-bit RGB batches (+ ToTensor + additive noise); see resample.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace srg {

// filter: 0 = bilinear (support 1), 1 = bicubic (a = -0.5, support 2).  ksize = taps per output element, or -1
int resize_plan_ksize(int in_size, int out_size, int filter);
// HOST arrays: bounds [out_size][2] = (first input index, tap count), coeffs [out_size][ksize] fixed point (22 bits)
int resize_plan(int in_size, int out_size, int filter, int* bounds, int* coeffs);
// src uint8 [N][H][W][3] -> out_u8 [N][out_h][out_w][3] and / or out_f32 [N][3][out_h][out_w] = v / 255 (+ noise * sigma[n]);
// plans are DEVICE copies of resize_plan's arrays; tmp: uint8 [N][H][out_w][3] (needed when out_w != W)
int launch_resize_u8(const uint8_t* src, int N, int H, int W, int out_h, int out_w, const int* bounds_w, const int* coeffs_w,
                     int ksize_w, const int* bounds_h, const int* coeffs_h, int ksize_h, uint8_t* tmp, uint8_t* out_u8,
                     float* out_f32, const float* noise, const float* sigma, cudaStream_t st);

}  // namespace srg
