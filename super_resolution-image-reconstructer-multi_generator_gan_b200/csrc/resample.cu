// On-GPU image-size transforms either side of the hot path (reference: src/transformers.py:73-82 through
// torchvision.transforms.Resize on PIL images = Pillow's ImagingResample, src/libImaging/Resample.c): antialiased bilinear
// / bicubic resize of 8-bit RGB batches, bit-exact with Pillow (separable two-pass convolution, per-pixel windows scaled by
// the reduction factor, double-precision coefficients normalised and converted to 22-bit fixed point, int32 accumulation
// from 1 << 21, clip to uint8 after EACH pass), with ToTensor (/255, NCHW fp32) and the additive Gaussian degradation of
// downward_img_quality fused into the second pass.  Integer work: the parity bar is bit-exact.
//
// HBM-bound byte work: one thread per output pixel (3 channels), taps read through L1 (every input byte is used by
// ~ksize / scale outputs of neighbouring threads); the intermediate of the horizontal pass is uint8 like Pillow's.
#include "resample.cuh"

#include <math.h>

#include "conv_gemm.cuh"

namespace srg {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}
double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: dst[n][y][xx][c] = clip8(half + sum_t src[n][y][xmin + t][c] * k[xx][t])
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ src, int N, int H, int W, int Wo,
                                                         const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize,
                                                         uint8_t* __restrict__ dst) {
  const long long total = (long long)N * H * Wo;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int xx = int(i % Wo);
    const long long row = i / Wo;
    const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
    const int* k = coeffs + (long long)xx * ksize;
    const uint8_t* p = src + (row * W + xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < cnt; ++t) {
      const int kv = k[t];
      s0 += int(p[3 * t]) * kv; s1 += int(p[3 * t + 1]) * kv; s2 += int(p[3 * t + 2]) * kv;
    }
    uint8_t* o = dst + i * 3;
    o[0] = uint8_t(clip8(s0)); o[1] = uint8_t(clip8(s1)); o[2] = uint8_t(clip8(s2));
  }
}

// vertical pass + output conversion: uint8 NHWC and / or fp32 NCHW (v / 255 [+ noise * sigma[n]], two roundings like torch)
__global__ void __launch_bounds__(256) resample_v_kernel(const uint8_t* __restrict__ src, int N, int H, int Wo, int Ho,
                                                         const int* __restrict__ bounds, const int* __restrict__ coeffs, int ksize,
                                                         uint8_t* __restrict__ out_u8, float* __restrict__ out_f32,
                                                         const float* __restrict__ noise, const float* __restrict__ sigma) {
  const long long total = (long long)N * Ho * Wo;
  const long long plane = (long long)Ho * Wo;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int x = int(i % Wo);
    const int yy = int((i / Wo) % Ho);
    const int n = int(i / plane);
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const int* k = coeffs + (long long)yy * ksize;
    const uint8_t* p = src + (((long long)n * H + ymin) * Wo + x) * 3;
    const long long pitch = (long long)Wo * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int t = 0; t < cnt; ++t) {
      const int kv = k[t];
      const uint8_t* q = p + t * pitch;
      s0 += int(q[0]) * kv; s1 += int(q[1]) * kv; s2 += int(q[2]) * kv;
    }
    const int v[3] = {clip8(s0), clip8(s1), clip8(s2)};
    if (out_u8 != nullptr) {
      uint8_t* o = out_u8 + i * 3;
      o[0] = uint8_t(v[0]); o[1] = uint8_t(v[1]); o[2] = uint8_t(v[2]);
    }
    if (out_f32 != nullptr) {
      const float sg = (noise != nullptr && sigma != nullptr) ? sigma[n] : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long o = ((long long)n * 3 + c) * plane + (long long)yy * Wo + x;
        float f = __fdiv_rn(float(v[c]), 255.f);
        if (noise != nullptr && sigma != nullptr) f = __fadd_rn(f, __fmul_rn(noise[o], sg));
        out_f32[o] = f;
      }
    }
  }
}

int grid_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : int(b);
}

}  // namespace

int resize_plan_ksize(int in_size, int out_size, int filter) {
  if (in_size < 1 || out_size < 1 || (filter != 0 && filter != 1)) return -1;
  double filterscale = double(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = (filter == 0 ? 1.0 : 2.0) * filterscale;
  return int(ceil(support)) * 2 + 1;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box
int resize_plan(int in_size, int out_size, int filter, int* bounds, int* coeffs) {
  const int ksize = resize_plan_ksize(in_size, out_size, filter);
  if (ksize < 0) { set_error("resize_plan: bad sizes or filter"); return -90; }
  double (*fn)(double) = filter == 0 ? bilinear_filter : bicubic_filter;
  const double scale = double(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = (filter == 0 ? 1.0 : 2.0) * filterscale;
  const double ss = 1.0 / filterscale;
  double* k = new double[ksize];
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < ksize; ++x) k[x] = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double w = fn((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < ksize; ++x)
      coeffs[size_t(xx) * ksize + x] = k[x] < 0 ? int(-0.5 + k[x] * (1 << kPrecisionBits)) : int(0.5 + k[x] * (1 << kPrecisionBits));
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  delete[] k;
  return 0;
}

int launch_resize_u8(const uint8_t* src, int N, int H, int W, int out_h, int out_w, const int* bounds_w, const int* coeffs_w,
                     int ksize_w, const int* bounds_h, const int* coeffs_h, int ksize_h, uint8_t* tmp, uint8_t* out_u8,
                     float* out_f32, const float* noise, const float* sigma, cudaStream_t st) {
  if (N < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1) { set_error("resize_u8: empty image"); return -91; }
  if (out_u8 == nullptr && out_f32 == nullptr) { set_error("resize_u8: no output"); return -92; }
  if (bounds_h == nullptr || coeffs_h == nullptr || ksize_h < 1) { set_error("resize_u8: the vertical plan is required"); return -93; }
  const uint8_t* mid = src;
  if (out_w != W) {
    if (tmp == nullptr || bounds_w == nullptr || coeffs_w == nullptr || ksize_w < 1) {
      set_error("resize_u8: the width changes: horizontal plan and intermediate buffer required"); return -94;
    }
    resample_h_kernel<<<grid_for((long long)N * H * out_w), 256, 0, st>>>(src, N, H, W, out_w, bounds_w, coeffs_w, ksize_w, tmp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("resample_h launch: %s", cudaGetErrorString(e)); return int(e); }
    count_launch();
    mid = tmp;
  }
  // the vertical pass also converts the output; with an unchanged height its plan is the identity (one tap of weight 1),
  // which reproduces Pillow skipping the pass
  resample_v_kernel<<<grid_for((long long)N * out_h * out_w), 256, 0, st>>>(mid, N, H, out_w, out_h, bounds_h, coeffs_h, ksize_h,
                                                                           out_u8, out_f32, noise, sigma);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("resample_v launch: %s", cudaGetErrorString(e)); return int(e); }
  count_launch();
  return 0;
}

}  // namespace srg
