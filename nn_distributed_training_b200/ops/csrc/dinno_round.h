// One-launch-per-round DiNNO on the MNIST conv net (dinno_round.cu).
#pragma once
#include "consensus.h"
#include "mnist.h"

namespace nndt {
namespace round {

constexpr int kMaxSteps = 8;

struct RoundArgs {
  mnist::Args m;                        // forward/backward geometry (x / y / direct_bs are taken per step below)
  consensus::DinnoArgs<float> d;        // consensus state; d.step is ignored (all d.pits steps run in the launch)
  const void* x_step[kMaxSteps];
  const int64_t* y_step[kMaxSteps];
  const int* bs_step[kMaxSteps];
  long long* prof;                      // optional [L*S, 64] %globaltimer stamps of the phases (debug / profiling)
};

// grid = (S slices, L nodes), one thread-block cluster of S CTAs per node (S <= 8)
cudaError_t launch_dinno_round(const RoundArgs& a, int S, cudaStream_t st);
int max_active_clusters(int S);

}  // namespace round
}  // namespace nndt
