// Launch interface of the fused MNIST conv-net kernels (mnist.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nndt {
namespace mnist {

struct Args {
  // parameters: row l of theta [L, n_pad]; slot offsets from FlatLayout
  const float* theta;
  int n_pad, L;
  int off_wc, off_bc, off_w1, off_b1, off_w2, off_b2;
  // data (uint8 pixels + (mean, 1/std), or float32 already normalised), int64 labels
  const void* x;
  const int64_t* y;
  int x_is_u8;
  float mean, inv_std;
  // sampling: direct==1 -> rows [l*batch + t]; else stateless permutation of shard l
  int direct, batch, seed, node0;
  const int* direct_bs;   // [L] valid rows per node when direct (nullptr: all `batch`)
  const int* shard_off;   // [L]
  const int* shard_len;   // [L]
  const int* calls;       // [L] draw counter per node (device, advanced by the update kernel)
  // training outputs
  float* grad_part;       // [L, S, n_pad]
  float* loss_part;       // [L, S]
  // evaluation
  int n_val;
  float* val_loss;              // [L, n_val]
  unsigned char* val_correct;   // [L, n_val]
};

cudaError_t launch_train(const Args& a, int spb, int S, cudaStream_t st);
cudaError_t launch_eval(const Args& a, int ctas_per_node, cudaStream_t st);
cudaError_t launch_batch_indices(int m, int B, int call, int seed, int node, int* out, int* out_size, cudaStream_t st);

}  // namespace mnist
}  // namespace nndt
