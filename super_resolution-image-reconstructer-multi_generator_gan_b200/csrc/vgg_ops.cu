// HBM-bound helpers of the VGG19 perceptual loss (reference: src/models.py:123-151 VGGFeatureExtractor over
// torchvision's vgg19.features, src/utils.py:154-166 perceptal_loss): the 3x3 convolutions run through conv_ops.cu; this file
// holds what sits between them -- the 3-channel first layer's im2col / col2im, MaxPool2d(2, 2) forward / backward and
// the L1 feature loss with its (ReLU-masked) gradient.  Activations are NHWC bf16 with C % 8 == 0, images NCHW fp32.
#include "vgg_ops.cuh"

#include <cuda_bf16.h>

#include "conv_gemm.cuh"

namespace srg {

namespace {

#define VGG_LAUNCH_CHECK(name)                                                    \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      set_error("%s launch: %s", name, cudaGetErrorString(e_));                   \
      return int(e_);                                                             \
    }                                                                             \
    count_launch();                                                               \
  } while (0)

__device__ __forceinline__ float blo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bhi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t bpack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  f[0] = blo(r.x); f[1] = bhi(r.x); f[2] = blo(r.y); f[3] = bhi(r.y);
  f[4] = blo(r.z); f[5] = bhi(r.z); f[6] = blo(r.w); f[7] = bhi(r.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(bpack(f[0], f[1]), bpack(f[2], f[3]), bpack(f[4], f[5]), bpack(f[6], f[7]));
}

int grid_for(long long n, int per_block = 256, int cap = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return int(b);
}

// dst[n][h][w][ch], ch = (r*3 + s)*3 + c  <-  src[n][c][h + r - 1][w + s - 1]  (0 outside; ch 27..63 = 0)
__global__ void __launch_bounds__(256) unfold3_kernel(const float* __restrict__ src, int N, int H, int W, uint4* __restrict__ dst) {
  const long long total = (long long)N * H * W * 8;
  const long long plane = (long long)H * W;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int j = int(i & 7);
    const long long pix = i >> 3;
    const int w = int(pix % W);
    const int h = int((pix / W) % H);
    const int n = int(pix / plane);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = j * 8 + e;
      float v = 0.f;
      if (ch < 27) {
        const int r = ch / 9, s = (ch % 9) / 3, c = ch % 3;
        const int hh = h + r - 1, ww = w + s - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = src[((long long)n * 3 + c) * plane + (long long)hh * W + ww];
      }
      f[e] = v;
    }
    dst[i] = pack8(f);
  }
}

// dimg[n][c][h][w] = scale * sum_{r,s} d_unf[n][h - r + 1][w - s + 1][(r*3 + s)*3 + c]
__global__ void __launch_bounds__(256) fold3_kernel(const __nv_bfloat16* __restrict__ d_unf, int N, int H, int W, float scale,
                                                    float* __restrict__ dimg) {
  const long long plane = (long long)H * W;
  const long long total = (long long)N * 3 * plane;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int w = int(i % W);
    const int h = int((i / W) % H);
    const int c = int((i / plane) % 3);
    const int n = int(i / (3 * plane));
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int hh = h - r + 1, ww = w - s + 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
          acc += __bfloat162float(d_unf[(((long long)n * H + hh) * W + ww) * 64 + (r * 3 + s) * 3 + c]);
      }
    dimg[i] = acc * scale;
  }
}

// MaxPool2d(2, 2): out[n][ho][wo][c] = max of the 2x2 window (floor semantics for odd sizes)
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const uint4* __restrict__ x, int N, int H, int W, int C8,
                                                           uint4* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C8;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int j = int(i % C8);
    const long long pix = i / C8;
    const int wo = int(pix % Wo);
    const int ho = int((pix / Wo) % Ho);
    const int n = int(pix / ((long long)Ho * Wo));
    const long long base = (((long long)n * H + 2 * ho) * W + 2 * wo) * C8 + j;
    float a[8], b[8];
    unpack8(x[base], a);
    unpack8(x[base + C8], b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
    unpack8(x[base + (long long)W * C8], b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
    unpack8(x[base + (long long)W * C8 + C8], b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = fmaxf(a[e], b[e]);
    out[i] = pack8(a);
  }
}

// dx = (gradient routed to the FIRST maximum of each window, torch's tie rule) [+ add]; pixels no window covers get 0 [+ add]
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                           const uint4* __restrict__ add, int N, int H, int W, int C8,
                                                           uint4* __restrict__ dx) {
  const int Hb = (H + 1) / 2, Wb = (W + 1) / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Hb * Wb * C8;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int j = int(i % C8);
    const long long blk = i / C8;
    const int wb = int(blk % Wb);
    const int hb = int((blk / Wb) % Hb);
    const int n = int(blk / ((long long)Hb * Wb));
    const int h0 = 2 * hb, w0 = 2 * wb;
    const bool full = hb < Ho && wb < Wo;
    float v[4][8], g[8], o[4][8];
    long long idx[4];
    bool ok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int hh = h0 + (q >> 1), ww = w0 + (q & 1);
      ok[q] = hh < H && ww < W;
      idx[q] = (((long long)n * H + hh) * W + ww) * C8 + j;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[q][e] = 0.f;
    }
    if (full) {
#pragma unroll
      for (int q = 0; q < 4; ++q) unpack8(x[idx[q]], v[q]);
      unpack8(dy[(((long long)n * Ho + hb) * Wo + wb) * C8 + j], g);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int best = 0;
        float m = v[0][e];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (v[q][e] > m) { m = v[q][e]; best = q; }
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q][e] = (q == best) ? g[e] : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!ok[q]) continue;
      if (add != nullptr) {
        float t[8];
        unpack8(add[idx[q]], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[q][e] += t[e];
      }
      dx[idx[q]] = pack8(o[q]);
    }
  }
}

// partial[block] = sum |a - b| over the block's elements; grad (optional) = sign(a - b) * gscale, zero where a <= 0
__global__ void __launch_bounds__(256) l1_feat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, long long n_vec,
                                                      float gscale, int relu_mask, uint4* __restrict__ grad,
                                                      double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n_vec; i += gridDim.x * 256ll) {
    float x[8], y[8], g[8];
    unpack8(a[i], x);
    unpack8(b[i], y);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = x[e] - y[e];
      s += fabsf(d);
      float t = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
      if (relu_mask && !(x[e] > 0.f)) t = 0.f;
      g[e] = t;
    }
    acc += double(s);
    if (grad != nullptr) grad[i] = pack8(g);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void l1_feat_final_kernel(const double* __restrict__ partial, int blocks, double inv_n, float weight, int accumulate,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < blocks; ++i) t += partial[i];
  const float v = float(t * inv_n) * weight;
  out[0] = accumulate ? out[0] + v : v;
}

}  // namespace

int launch_unfold3(const float* src, int N, int H, int W, void* dst, cudaStream_t st) {
  if (N < 1 || H < 1 || W < 1) { set_error("unfold3: empty image"); return -80; }
  unfold3_kernel<<<grid_for((long long)N * H * W * 8), 256, 0, st>>>(src, N, H, W, reinterpret_cast<uint4*>(dst));
  VGG_LAUNCH_CHECK("unfold3");
  return 0;
}
int launch_fold3(const void* d_unf, int N, int H, int W, float scale, float* dimg, cudaStream_t st) {
  if (N < 1 || H < 1 || W < 1) { set_error("fold3: empty image"); return -80; }
  fold3_kernel<<<grid_for((long long)N * 3 * H * W), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_unf), N, H, W, scale, dimg);
  VGG_LAUNCH_CHECK("fold3");
  return 0;
}
int launch_maxpool2_forward(const void* x, int N, int H, int W, int C, void* out, cudaStream_t st) {
  if (N < 1 || H < 2 || W < 2 || C < 8 || C % 8 != 0) { set_error("maxpool2: needs H, W >= 2 and C %% 8 == 0"); return -81; }
  maxpool2_fwd_kernel<<<grid_for((long long)N * (H / 2) * (W / 2) * (C / 8)), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(x), N, H, W, C / 8, reinterpret_cast<uint4*>(out));
  VGG_LAUNCH_CHECK("maxpool2_forward");
  return 0;
}
int launch_maxpool2_backward(const void* x, const void* dy, const void* add, int N, int H, int W, int C, void* dx, cudaStream_t st) {
  if (N < 1 || H < 2 || W < 2 || C < 8 || C % 8 != 0) { set_error("maxpool2: needs H, W >= 2 and C %% 8 == 0"); return -81; }
  maxpool2_bwd_kernel<<<grid_for((long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8)), 256, 0, st>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(add), N, H, W, C / 8,
      reinterpret_cast<uint4*>(dx));
  VGG_LAUNCH_CHECK("maxpool2_backward");
  return 0;
}
size_t l1_feat_scratch_bytes() { return size_t(kL1FeatBlocks) * sizeof(double); }
int launch_l1_feat(const void* a, const void* b, long long n, float weight, int accumulate, float grad_scale, int relu_mask,
                   void* grad, void* scratch, float* out, cudaStream_t st) {
  if (n < 8 || n % 8 != 0) { set_error("l1_feat: element count must be a positive multiple of 8"); return -82; }
  const long long n_vec = n / 8;
  int blocks = grid_for(n_vec, 256, kL1FeatBlocks);
  l1_feat_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), n_vec,
                                         grad_scale * weight / float(n), relu_mask, reinterpret_cast<uint4*>(grad),
                                         reinterpret_cast<double*>(scratch));
  VGG_LAUNCH_CHECK("l1_feat");
  l1_feat_final_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const double*>(scratch), blocks, 1.0 / double(n), weight, accumulate, out);
  VGG_LAUNCH_CHECK("l1_feat_final");
  return 0;
}

}  // namespace srg
