// Launch interface of the generic CUDA-core MLP forward / backward kernels (mlp_generic.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nndt {
namespace mlpg {

constexpr int kMaxLayers = 8;
enum Act : int { kNone = 0, kRelu = 1, kTanh = 2, kSigmoid = 3 };

struct Args {
  const void* x;         // [M, dims[0]]
  const void* params;    // flat parameter vector: W_l [dims[l+1], dims[l]] row-major at w_off[l], b_l at b_off[l]
  int M, nl, dtype64;
  int dims[kMaxLayers + 1];
  int act[kMaxLayers];
  int w_off[kMaxLayers], b_off[kMaxLayers];
  void* acts;            // [M, act_stride] post-activation output of every layer (column offset act_off[l])
  int act_stride;
  int act_off[kMaxLayers];
  const void* gout;      // backward: dL/d(last output) [M, dims[nl]]
  void* gparams;         // backward: flat gradient, accumulated with atomicAdd (zeroed by the caller)
  void* gx;              // backward: dL/dx [M, dims[0]] or nullptr
};

bool supported(const Args& a);
cudaError_t launch_forward(const Args& a, cudaStream_t st);
cudaError_t launch_backward(const Args& a, cudaStream_t st);

}  // namespace mlpg
}  // namespace nndt
