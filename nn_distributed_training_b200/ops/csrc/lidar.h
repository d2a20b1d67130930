#pragma once
#include <cuda_runtime.h>

namespace nndt {
namespace lidar {

struct Args {
  const double* tx; const double* ty; const double* coef;   // bicubic spline: knots and (ntx-4) x (nty-4) coefficients
  int ntx, nty;
  const double* poses;   // [n_poses, 2]
  int n_poses, num_beams, beam_samps, collision_samps, fine_samps;
  double beam_len, samp_df;
  double* out;           // [n_poses, num_beams * beam_samps, 3]
};

cudaError_t launch_scan(const Args& a, cudaStream_t st);
cudaError_t launch_density(const Args& a, const double* xy, int n, double* out, cudaStream_t st);

}  // namespace lidar
}  // namespace nndt
