// Per-operator convolution entry points (conv2d_{fprop,dgrad,wgrad} + workspace queries): nn.Conv2d(Cin, Cout,
// k, padding=k/2) forward, input gradient and weight / bias gradient for k in {1, 3} on NHWC bf16 activations whose channel
// counts are multiples of 64 -- the same tcgen05 implicit-GEMM kernels the generator / discriminator engines launch
// (conv_gemm.cu, wgrad_gemm.cu), reachable for topologies other than the two fixed ones (e.g. the VGG19 feature extractor
// of the perceptual loss, src/models.py:123-151).  Weights are fp32 OIHW on the caller's side; srg_conv2d_pack_weights
// re-packs them into the k-block-major bf16 operand layout once per weight update.
#include "conv_ops.cuh"

#include <string.h>

#include "conv_gemm.cuh"
#include "elementwise.cuh"

namespace srg {

namespace {

#define OPS_LAUNCH_CHECK(name)                                                    \
  do {                                                                            \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      set_error("%s launch: %s", name, cudaGetErrorString(e_));                   \
      return int(e_);                                                             \
    }                                                                             \
    count_launch();                                                               \
  } while (0)

// dst[(((c*k + s)*k + r) * n_out + n) * 64 + kk]: k-block (c, s = kw, r = kh), GEMM column n, 64 K-channels
//   fprop: = w[n][c*64 + kk][r][s]                       (n_out = cout, K = cin)
//   dgrad: = w[c*64 + kk][n][k-1-r][k-1-s]               (n_out = cin,  K = cout): the transposed, flipped filter
__global__ void __launch_bounds__(256) conv2d_pack_kernel(const float* __restrict__ w, int cout, int cin, int k, int dgrad,
                                                          __nv_bfloat16* __restrict__ dst, long long total) {
  const int n_out = dgrad ? cin : cout;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int kk = int(i & 63);
    long long t = i >> 6;
    const int n = int(t % n_out); t /= n_out;
    const int r = int(t % k); t /= k;
    const int s = int(t % k);
    const int c = int(t / k);
    const int kc = c * 64 + kk;
    float v;
    if (!dgrad) v = w[((size_t(n) * cin + kc) * k + r) * k + s];
    else v = w[((size_t(kc) * cin + n) * k + (k - 1 - r)) * k + (k - 1 - s)];
    dst[i] = __float2bfloat16(v);
  }
}

// partial element j of one split, layout [n_blocks][kw 3][(2-kh)*64 + co][ci 64] -> OIHW element of the full filter
__global__ void __launch_bounds__(256) wgrad3_reduce_oihw_kernel(const float* __restrict__ partials, int per_split, int splits,
                                                                 int cin, int ci0, int co0, float* __restrict__ dw) {
  for (int j = blockIdx.x * 256 + threadIdx.x; j < per_split; j += gridDim.x * 256) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partials[size_t(s) * per_split + j];
    const int ci = j & 63;
    const int row = (j >> 6) % 192;
    const int kh = 2 - row / 64, co = row & 63;
    const int kw = (j / (64 * 192)) % 3;
    const int nb = j / (64 * 192 * 3);
    dw[((size_t(co0 + nb * 64 + co) * cin + ci0 + ci) * 3 + kh) * 3 + kw] = acc;
  }
}

// per-channel sums of a [P][C] bf16 tensor (C % 64 == 0): partials [C/64][rows][64], then a fixed-order sum over rows
__global__ void __launch_bounds__(256) chan_sum_c_kernel(const uint4* __restrict__ a, long long pixels, int C,
                                                         float* __restrict__ partials) {
  __shared__ float red[32][65];
  const int cg = threadIdx.x & 7, lane_p = threadIdx.x >> 3, chunk = blockIdx.y;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  const int vec_per_pixel = C / 8;
  for (long long p = blockIdx.x * 32ll + lane_p; p < pixels; p += gridDim.x * 32ll) {
    const uint4 r = a[p * vec_per_pixel + chunk * 8 + cg];
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[2 * e] += __uint_as_float(w[e] << 16);
      s[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[lane_p][cg * 8 + e] = s[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float acc = 0.f;
    for (int l = 0; l < 32; ++l) acc += red[l][threadIdx.x];
    partials[(size_t(chunk) * gridDim.x + blockIdx.x) * 64 + threadIdx.x] = acc;
  }
}
__global__ void chan_sum_c_final_kernel(const float* __restrict__ partials, int rows, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int chunk = c >> 6, cc = c & 63;
  double acc = 0.0;
  for (int r = 0; r < rows; ++r) acc += double(partials[(size_t(chunk) * rows + r) * 64 + cc]);
  out[c] = float(acc);
}

InView view_of(const void* base, int H, int W, int C, int c0, int channels) {
  InView v;
  v.ptr = reinterpret_cast<const uint16_t*>(base) + c0;
  v.stride_w = C; v.stride_h = int64_t(W) * C; v.stride_n = int64_t(H) * W * C; v.channels = channels;
  return v;
}

int check_geometry(const char* who, int N, int H, int W, int cin, int cout, int k) {
  if (N < 1 || H < 1 || W < 1) { set_error("%s: empty tensor", who); return -70; }
  if (cin < 64 || cout < 64 || cin % 64 != 0 || cout % 64 != 0) {
    set_error("%s: channel counts must be multiples of 64 (got %d -> %d)", who, cin, cout); return -71;
  }
  if (k != 1 && k != 3) { set_error("%s: kernel size 1 or 3 (got %d)", who, k); return -72; }
  return 0;
}

}  // namespace

size_t conv2d_packed_elems(int cout, int cin, int k) { return size_t(cout) * cin * k * k; }

int launch_conv2d_pack(const float* w_oihw, int cout, int cin, int k, int dgrad, void* packed, cudaStream_t st) {
  int rc = check_geometry("conv2d_pack_weights", 1, 1, 1, cin, cout, k);
  if (rc) return rc;
  const long long total = (long long)conv2d_packed_elems(cout, cin, k);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv2d_pack_kernel<<<int(blocks), 256, 0, st>>>(w_oihw, cout, cin, k, dgrad, reinterpret_cast<__nv_bfloat16*>(packed), total);
  OPS_LAUNCH_CHECK("conv2d_pack");
  return 0;
}

// out[N,H,W,n_out] = act(conv_k(x[N,H,W,k_ch], packed) + bias) [+ residual] [masked by mask_src > 0]
int launch_conv2d(const void* x, int N, int H, int W, int k_ch, const void* packed, int n_out, int k, const float* bias, int act,
                  float slope, const void* residual, const void* mask_src, void* out, cudaStream_t st) {
  int rc = check_geometry("conv2d", N, H, W, k_ch, n_out, k);
  if (rc) return rc;
  ConvGemmArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.H = H; a.W = W; a.TH = 16; a.TW = 8;
  if (k == 3) {
    a.n_strips = 3; a.n_taps = 3; a.strip_rows = 18; a.strip_dh = -1;
    for (int s = 0; s < 3; ++s) a.strip_dw[s] = s - 1;
    for (int r = 0; r < 3; ++r) a.tap_row[r] = r;
  } else {
    a.n_strips = 1; a.n_taps = 1; a.strip_rows = 16; a.strip_dh = 0; a.strip_dw[0] = 0; a.tap_row[0] = 0;
  }
  a.n_views = 1; a.views[0] = view_of(x, H, W, k_ch, 0, k_ch); a.in_H = H; a.in_W = W;
  a.weights = packed; a.cout_total = n_out; a.block_n = 64;
  a.bias = bias; a.act = act; a.slope = slope; a.residual = residual; a.mask_src = mask_src;
  a.out = out; a.out_mode = OUT_NHWC;
  a.exclusive = 1;
  // variant 0: launch_conv_gemm picks the row-interleaved kernel for 3x3 / 64 input channels / <= 256 output channels and
  // the generic strip kernel (weights streamed when they do not fit shared memory) for everything else
  return launch_conv_gemm(a, st);
}

size_t conv2d_wgrad_workspace_bytes(int N, int H, int W, int cin, int cout) {
  if (N < 1 || H < 1 || W < 1 || cin % 64 != 0 || cout % 64 != 0) return 0;
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.N = N; a.H = H; a.W = W; a.TH = 16; a.TW = 8; a.in_H = H; a.in_W = W;
  a.n_blocks = cout >= 256 ? 4 : cout / 64;
  a.n_strips = 3; a.n_taps = 3; a.strip_rows = 18; a.strip_dh = -1;
  int splits = 0;
  const size_t part = size_t(wgrad3x3_partials_floats(a, &splits)) * 4;
  const size_t bias = size_t(cout / 64) * kRedBlocksMax * 64 * 4;
  return part + bias + 1024;
}

int launch_conv2d_wgrad(const void* x, const void* dy, int N, int H, int W, int cin, int cout, void* workspace,
                        size_t workspace_bytes, float* dw_oihw, float* dbias, cudaStream_t st) {
  int rc = check_geometry("conv2d_wgrad", N, H, W, cin, cout, 3);
  if (rc) return rc;
  if (workspace_bytes < conv2d_wgrad_workspace_bytes(N, H, W, cin, cout)) { set_error("conv2d_wgrad: workspace too small"); return -73; }
  float* partials = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  for (int co0 = 0; co0 < cout; co0 += 256) {
    const int nb = (cout - co0 >= 256 ? 256 : cout - co0) / 64;
    for (int ci0 = 0; ci0 < cin; ci0 += 64) {
      WgradArgs a;
      memset(&a, 0, sizeof(a));
      a.N = N; a.H = H; a.W = W; a.TH = 16; a.TW = 8;
      a.x = view_of(x, H, W, cin, ci0, 64); a.in_H = H; a.in_W = W;
      a.dy_views = 1; a.dy[0] = view_of(dy, H, W, cout, co0, nb * 64);
      a.n_blocks = nb;
      a.n_strips = 3; a.n_taps = 3; a.strip_rows = 18; a.strip_dh = -1;
      for (int s = 0; s < 3; ++s) a.strip_dw[s] = s - 1;
      for (int r = 0; r < 3; ++r) a.tap_row[r] = r;
      a.partials = partials;
      int splits = 0;
      wgrad3x3_partials_floats(a, &splits);
      rc = launch_wgrad3x3(a, st);
      if (rc) return rc;
      const int per_split = nb * 3 * 192 * 64;
      int blocks = (per_split + 255) / 256;
      if (blocks > 148 * 4) blocks = 148 * 4;
      wgrad3_reduce_oihw_kernel<<<blocks, 256, 0, st>>>(partials, per_split, splits, cin, ci0, co0, dw_oihw);
      OPS_LAUNCH_CHECK("wgrad3_reduce_oihw");
    }
  }
  if (dbias != nullptr) {
    const long long P = (long long)N * H * W;
    int rows = int((P + 255) / 256);
    if (rows > kRedBlocksMax) rows = kRedBlocksMax;
    if (rows < 1) rows = 1;
    float* bpart = partials + (conv2d_wgrad_workspace_bytes(N, H, W, cin, cout) - 1024) / 4 - size_t(cout / 64) * kRedBlocksMax * 64;
    chan_sum_c_kernel<<<dim3(rows, cout / 64), 256, 0, st>>>(reinterpret_cast<const uint4*>(dy), P, cout, bpart);
    OPS_LAUNCH_CHECK("chan_sum_c");
    chan_sum_c_final_kernel<<<(cout + 127) / 128, 128, 0, st>>>(bpart, rows, cout, dbias);
    OPS_LAUNCH_CHECK("chan_sum_c_final");
  }
  return 0;
}

}  // namespace srg
