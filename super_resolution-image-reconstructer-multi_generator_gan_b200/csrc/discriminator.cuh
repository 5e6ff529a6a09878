// Discriminator engine (reference: src/models.py:90-120): 4 x [strided conv, MaxPool(3,2), InstanceNorm, LeakyReLU]
// (last stage: Sigmoid instead of LeakyReLU).  Internal C++ interface; C ABI wrappers live in api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "generator.cuh"

namespace srg {

struct DiscStage {
  int Cin, Cout;        // reference channel counts (3->64, 64->128, 128->256, 256->512)
  int Hin, Win;         // spatial size of the stage input (image for stage 0, previous pooled map otherwise)
  int Ho, Wo;           // conv output
  int Hp, Wp;           // pooled output
  int Hs, Ws, Cs;       // operand tensor of the conv: stage 0: unfolded image [Ho+3][Wo][64]; else s2d [Ho+1][Wo+1][4*Cin]
};

struct DiscriminatorEngine {
  int N, H, W;
  DiscStage st[4];
  std::vector<ParamInfo> params;
  std::vector<TensorInfo> tensors;
  int64_t param_elems = 0;
  int64_t packed_elems = 0;
  size_t workspace_bytes_train = 0, workspace_bytes_eval = 0;
  int* d_pack_idx = nullptr;
  float* master = nullptr;
  float* grads = nullptr;
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  bool ws_training = false;
  long long launches = 0;
  virtual ~DiscriminatorEngine();
};

// returns nullptr (and sets the error string) when the geometry is invalid for the reference network
DiscriminatorEngine* discriminator_create(int N, int H, int W);
int discriminator_bind(DiscriminatorEngine* d, float* master, float* grads, void* ws, size_t ws_bytes, int training);
int discriminator_pack(DiscriminatorEngine* d, cudaStream_t st);
// x: fp32 NCHW [N,3,H,W]  ->  out: fp32 NCHW [N,512,Hp3,Wp3] (sigmoid map)
int discriminator_forward(DiscriminatorEngine* d, const float* x_nchw, float* out_nchw, int keep_for_backward,
                          cudaStream_t st);
// dout: gradient w.r.t. the output map.  param_grads != 0: writes every parameter gradient into the bound flat
// `grads`.  dx (may be null): gradient w.r.t. the input image, fp32 NCHW.
int discriminator_backward(DiscriminatorEngine* d, const float* dout_nchw, int param_grads, float* dx_nchw,
                           cudaStream_t st);

}  // namespace srg
