// Fused neighbor-exchange + mixing + optimizer-update kernels for DiNNO / DSGD / DSGT.
//
// Reference call sites replaced (all Python loops over nodes x parameter tensors):
//   optimizers/dinno.py:103-125 + :74-91  -> dinno_update   (exchange, dual ascent, prox-grad, Adam/SGD/AdamW)
//   optimizers/dsgd.py:37-46 / :55-58      -> dsgd_mix / dsgd_step
//   optimizers/dsgt.py:58-75 / :87-103     -> dsgt_mix / dsgt_track
//
// Every kernel is a single pass over the node's 16-byte vectorised parameter row: neighbor
// rows are pulled straight from the (local or NVLink-peer) published buffers named by the
// pointer table, combined with the Metropolis row in registers, the gradient partials of the
// forward/backward kernel are summed on the fly, the optimizer update is applied and the new
// row is published — no [d_i, n] stack, no cdist, no separate reduce or elementwise launch.
//
// Cross-GPU protocol (pull model): published rows are double buffered by round parity.  A
// rank announces "round k published" by writing k into its slot of every peer's flag array
// (st.release.sys over NVLink after __threadfence_system()); consumers spin with
// ld.acquire.sys on their *local* flag array only for the ranks that own a neighbor.  Since a
// node publishes k+1 only after finishing its round-k reads, two buffers suffice.
#include <mutex>
#include <unordered_map>

#include "consensus_device.cuh"

namespace nndt {
namespace consensus {

// S_local[par][chan] = sum over this rank's nodes of the published rows of round k
template <typename T>
__global__ void __launch_bounds__(THREADS) local_sum_kernel(const Common<T> c) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Vec<T>::N;
  const RoundInfo<T> ri = round_info(c);
  for (int ch = 0; ch < c.C; ++ch) {
    for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
      double s[N];
#pragma unroll
      for (int u = 0; u < N; ++u) s[u] = 0.0;
      for (int l = 0; l < c.L; ++l) {
        const Pack<T> q = ldv(pub_row(c, ri.par, ch, l) + i);
#pragma unroll
        for (int u = 0; u < N; ++u) s[u] += (double)q.v[u];
      }
      double* dst = c.sum_local + (size_t)(ri.par * c.C + ch) * c.n_pad + i;
#pragma unroll
      for (int u = 0; u < N; u += 2) *reinterpret_cast<double2*>(dst + u) = make_double2(s[u], s[u + 1]);
    }
  }
  // last block: tell every peer that this rank's partial sum of round k is ready
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(c.done_ctr, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) *c.done_ctr = 0;
    if (c.world > 1) {
      __threadfence_system();
      if ((int)threadIdx.x < c.world && (int)threadIdx.x != c.rank)
        st_release_sys(reinterpret_cast<int*>(c.peer_sum_flag[threadIdx.x]), ri.k + 1);
    }
  }
}


// Announce "every round below *round_ctr is published" to the peers.  Runs after the round's last kernel (kernel
// boundary = happens-before for all of its stores), so one thread's system fence + release stores are cumulative over
// the whole round; launched on a forked graph branch, nothing local waits for it.  Reading the *current* counter is
// always truthful: it only advances once the corresponding rows are written.
template <typename T>
__global__ void publish_round_kernel(const Common<T> c) {
  const int k = *reinterpret_cast<volatile int*>(c.round_ctr);
  announce_round(c, k);
}

// ------------------------------------------------------------------ DiNNO ----
template <typename T, int U>
__global__ void __launch_bounds__(THREADS) dinno_update_kernel(const DinnoArgs<T> a) {
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = node_of_block(c);
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  const DinnoCoef<T> cf = dinno_coef(a, ri.k, a.step, deg);
  const T rho = cf.rho;
  const bool first = a.step == 0, last = a.step == a.pits - 1;
  tl_stamp(c, ri.k, a.step, 0);
  if (first) begin_round(c, ri.gid, l, ri.k);
  tl_stamp(c, ri.k, a.step, 1);

  const bool fresh = first && !a.persistent;  // Adam moments restart every round (reference Q3)

  const size_t row = (size_t)l * c.n_pad;
  const T* thk_row = pub_row(c, ri.par, 0, l);
  // The grid covers the row exactly once (one vector per thread), so all state that the preceding
  // forward/backward kernel does not write is fetched BEFORE the programmatic-dependency wait and
  // overlaps that kernel's tail; only the gradient partials are read after it.
  bool waited = false;
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    Pack<T> th = ldv(c.theta + row + i);
    Pack<T> thk, dl, du;
    if (first) {
      thk = th;  // live row == theta^k at round start
#pragma unroll
      for (int u = 0; u < N; ++u) dl.v[u] = (T)0;
      if (c.sum_mode) {   // delta = S_all - N theta_i   (every other node is a neighbor)
        const DPack<N> sall = network_sum(c, ri.par, 0, i);
#pragma unroll
        for (int u = 0; u < N; ++u) dl.v[u] = (T)(sall.v[u] - (double)c.n_total * (double)thk.v[u]);
      } else {
        // up to four neighbor rows in flight per thread: over NVLink a load is ~2 us, issued one by one they add up
        for (int e0 = 0; e0 < deg; e0 += 4) {
          Pack<T> q[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (e0 + j < deg) q[j] = ldv(nbr_row(c, ri.gid, l, e0 + j, ri.par, 0) + i);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (e0 + j < deg) {
#pragma unroll
              for (int u = 0; u < N; ++u) dl.v[u] += q[j].v[u] - thk.v[u];
            }
        }
      }
      du = ldv(a.dual + row + i);
#pragma unroll
      for (int u = 0; u < N; ++u) du.v[u] -= rho * dl.v[u];
      stv(a.delta + row + i, dl);
      stv(a.dual + row + i, du);
    } else {
      thk = ldv(thk_row + i);
      dl = ldv(a.delta + row + i);
      du = ldv(a.dual + row + i);
    }
    Pack<T> m, v;
    if (a.opt != kSGD) {
      if (fresh) {
#pragma unroll
        for (int u = 0; u < N; ++u) { m.v[u] = (T)0; v.v[u] = (T)0; }
      } else {
        m = ldv(a.m + row + i);
        v = ldv(a.v + row + i);
      }
    }
    if (!waited) { tl_stamp(c, ri.k, a.step, 2); pdl_wait(); pdl_launch_dependents(); waited = true; tl_stamp(c, ri.k, a.step, 3); }
    const Pack<T> gl = sum_partials<U>(c, l, i);
    dinno_apply(cf, th, thk, dl, du, m, v, gl);
    if (a.opt != kSGD) {
      stv(a.m + row + i, m);
      stv(a.v + row + i, v);
    }
    stv(c.theta + row + i, th);
    if (last) stv(pub_row(c, ri.par ^ 1, 0, l) + i, th);
  }
  if (!waited) { pdl_wait(); pdl_launch_dependents(); }
  tl_stamp(c, ri.k, a.step, 4);
  step_bookkeeping(c, l);
  if (last) { tag_published(c, l, ri.k); finish_round(c, ri.k); }
}

// ------------------------------------------------------------------- DSGD ----
template <typename T>
__global__ void __launch_bounds__(THREADS) dsgd_mix_kernel(const Common<T> c) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Vec<T>::N;
  const int l = node_of_block(c);
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  begin_round(c, ri.gid, l, ri.k);
  const T ws = c.self_w[ri.gid * c.L + l];
  const T* w = c.nbr_w + (size_t)(ri.gid * c.L + l) * c.dmax;
  const size_t row = (size_t)l * c.n_pad;
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    if (c.sum_mode) {     // W = 11^T / N: the mixed row is the network mean
      const DPack<N> sall = network_sum(c, ri.par, 0, i);
      Pack<T> th;
#pragma unroll
      for (int u = 0; u < N; ++u) th.v[u] = (T)(sall.v[u] / (double)c.n_total);
      stv(c.theta + row + i, th);
      continue;
    }
    Pack<T> th = ldv(c.theta + row + i);
#pragma unroll
    for (int u = 0; u < N; ++u) th.v[u] *= ws;
    for (int e0 = 0; e0 < deg; e0 += 4) {
      Pack<T> q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e0 + j < deg) q[j] = ldv(nbr_row(c, ri.gid, l, e0 + j, ri.par, 0) + i);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e0 + j < deg) {
          const T we = w[e0 + j];
#pragma unroll
          for (int u = 0; u < N; ++u) th.v[u] += we * q[j].v[u];
        }
    }
    stv(c.theta + row + i, th);
  }
}

template <typename T, int U>
__global__ void __launch_bounds__(THREADS) dsgd_step_kernel(const Common<T> c) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Vec<T>::N;
  const int l = node_of_block(c);
  const RoundInfo<T> ri = round_info(c);
  const T alpha = c.alpha[ri.k];
  const size_t row = (size_t)l * c.n_pad;
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    Pack<T> th = ldv(c.theta + row + i);
    const Pack<T> g = sum_partials<U>(c, l, i);
#pragma unroll
    for (int u = 0; u < N; ++u) th.v[u] -= alpha * g.v[u];
    stv(c.theta + row + i, th);
    stv(pub_row(c, ri.par ^ 1, 0, l) + i, th);
  }
  step_bookkeeping(c, l);
  tag_published(c, l, ri.k);
  finish_round(c, ri.k);
}

// ------------------------------------------------------------------- DSGT ----
// channel 0 of the published buffer is theta, channel 1 the gradient tracker y.
template <typename T, int U>
__global__ void __launch_bounds__(THREADS) dsgt_init_kernel(const DsgtArgs<T> a) {
  pdl_wait();
  pdl_launch_dependents();
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = node_of_block(c);
  const size_t row = (size_t)l * c.n_pad;
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    const Pack<T> g = sum_partials<U>(c, l, i);
    stv(a.g_old + row + i, g);
    stv(pub_row(c, 0, 1, l) + i, g);
  }
  step_bookkeeping(c, l);
}

template <typename T>
__global__ void __launch_bounds__(THREADS) dsgt_mix_kernel(const DsgtArgs<T> a) {
  pdl_wait();
  pdl_launch_dependents();
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = node_of_block(c);
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  begin_round(c, ri.gid, l, ri.k);
  const T alpha = c.alpha[ri.k];
  const T ws = c.self_w[ri.gid * c.L + l];
  const T* w = c.nbr_w + (size_t)(ri.gid * c.L + l) * c.dmax;
  const size_t row = (size_t)l * c.n_pad;
  const T* ys = pub_row(c, ri.par, 1, l);
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    if (c.sum_mode) {
      const DPack<N> st = network_sum(c, ri.par, 0, i), sy = network_sum(c, ri.par, 1, i);
      Pack<T> th;
#pragma unroll
      for (int u = 0; u < N; ++u) th.v[u] = (T)((st.v[u] - (double)alpha * sy.v[u]) / (double)c.n_total);
      stv(c.theta + row + i, th);
      continue;
    }
    Pack<T> th = ldv(c.theta + row + i);
    const Pack<T> y = ldv(ys + i);
#pragma unroll
    for (int u = 0; u < N; ++u) th.v[u] = ws * (th.v[u] - alpha * y.v[u]);
    for (int e0 = 0; e0 < deg; e0 += 2) {
      Pack<T> qt[2], qy[2];
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (e0 + j < deg) {
          qt[j] = ldv(nbr_row(c, ri.gid, l, e0 + j, ri.par, 0) + i);
          qy[j] = ldv(nbr_row(c, ri.gid, l, e0 + j, ri.par, 1) + i);
        }
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (e0 + j < deg) {
          const T we = w[e0 + j];
#pragma unroll
          for (int u = 0; u < N; ++u) th.v[u] += we * (qt[j].v[u] - alpha * qy[j].v[u]);
        }
    }
    stv(c.theta + row + i, th);
  }
}

template <typename T, int U>
__global__ void __launch_bounds__(THREADS) dsgt_track_kernel(const DsgtArgs<T> a) {
  pdl_wait();
  pdl_launch_dependents();
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = node_of_block(c);
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  const T ws = c.self_w[ri.gid * c.L + l];
  const T* w = c.nbr_w + (size_t)(ri.gid * c.L + l) * c.dmax;
  const size_t row = (size_t)l * c.n_pad;
  const T* ys = pub_row(c, ri.par, 1, l);
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    Pack<T> y;
    if (c.sum_mode) {
      const DPack<N> sy = network_sum(c, ri.par, 1, i);
#pragma unroll
      for (int u = 0; u < N; ++u) y.v[u] = (T)(sy.v[u] / (double)c.n_total);
    } else {
      y = ldv(ys + i);
#pragma unroll
      for (int u = 0; u < N; ++u) y.v[u] *= ws;
      for (int e0 = 0; e0 < deg; e0 += 4) {
        Pack<T> q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e0 + j < deg) q[j] = ldv(nbr_row(c, ri.gid, l, e0 + j, ri.par, 1) + i);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (e0 + j < deg) {
            const T we = w[e0 + j];
#pragma unroll
            for (int u = 0; u < N; ++u) y.v[u] += we * q[j].v[u];
          }
      }
    }
    const Pack<T> gn = sum_partials<U>(c, l, i);
    const Pack<T> go = ldv(a.g_old + row + i);
#pragma unroll
    for (int u = 0; u < N; ++u) y.v[u] += gn.v[u] - go.v[u];
    stv(a.g_old + row + i, gn);
    stv(pub_row(c, ri.par ^ 1, 1, l) + i, y);
    stv(pub_row(c, ri.par ^ 1, 0, l) + i, ldv(c.theta + row + i));
  }
  step_bookkeeping(c, l);
  tag_published(c, l, ri.k);
  finish_round(c, ri.k);
}

// ------------------------------------------------------------ consensus metric ----
NNDT_DEVINL double block_sum(double v) {
  __shared__ double red[THREADS / 32];
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < THREADS / 32; ++w) t += red[w];
  return t;
}
template <typename T>
__global__ void __launch_bounds__(THREADS) inv_norm_kernel(const int64_t* rows, int n_pad, double* inv_norm) {
  const T* r = reinterpret_cast<const T*>(rows[blockIdx.x]);
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_pad; i += THREADS) { const double v = (double)r[i]; acc += v * v; }
  acc = block_sum(acc);
  if (threadIdx.x == 0) inv_norm[blockIdx.x] = 1.0 / fmax(sqrt(acc), 1e-12);   // F.normalize eps
}
// block (j, l): j < N -> pair distance to node j;  j == N -> distance to the mean of the normalised rows
template <typename T>
__global__ void __launch_bounds__(THREADS) consensus_metric_kernel(const int64_t* rows, int N, int n_pad, int local0,
                                                                   const double* inv_norm, double* out_pair, double* out_mean) {
  const int j = blockIdx.x, l = blockIdx.y, gi = local0 + l;
  const T* ri = reinterpret_cast<const T*>(rows[gi]);
  const double ni = inv_norm[gi];
  double acc = 0.0;
  if (j < N) {
    const T* rj = reinterpret_cast<const T*>(rows[j]);
    const double nj = inv_norm[j];
    for (int e = threadIdx.x; e < n_pad; e += THREADS) {
      const double d = (double)ri[e] * ni - (double)rj[e] * nj;
      acc += d * d;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) out_pair[(size_t)l * N + j] = sqrt(acc);
  } else {
    for (int e = threadIdx.x; e < n_pad; e += THREADS) {
      double mu = 0.0;
      for (int q = 0; q < N; ++q) mu += (double)reinterpret_cast<const T*>(rows[q])[e] * inv_norm[q];
      const double d = (double)ri[e] * ni - mu / (double)N;
      acc += d * d;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) out_mean[l] = sqrt(acc);
  }
}
template <typename T>
cudaError_t launch_consensus_metric(const int64_t* rows, int N, int n_pad, int local0, int L, double* inv_norm,
                                    double* out_pair, double* out_mean, cudaStream_t st) {
  inv_norm_kernel<T><<<N, THREADS, 0, st>>>(rows, n_pad, inv_norm);
  consensus_metric_kernel<T><<<dim3(N + 1, L), THREADS, 0, st>>>(rows, N, n_pad, local0, inv_norm, out_pair, out_mean);
  return cudaGetLastError();
}

// ------------------------------------------------------------ rank barrier ----
__global__ void rank_barrier_kernel(int* slots, const int64_t* peer_slot, int world, int rank, int epoch,
                                    const volatile int* gate, int* err) {
  if (gate != nullptr && threadIdx.x == 0) {
    const long long t0 = clock64();
    while (*gate == 0) {
      if (clock64() - t0 > kSpinLimit) { if (err) *err = 1; break; }
    }
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r < world && r != rank) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<int*>(peer_slot[r]), epoch);
    const long long t0 = clock64();
    while (ld_acquire_sys(slots + r) < epoch) {
      if (clock64() - t0 > kSpinLimit) { if (err) *err = 1; break; }
    }
  }
}
cudaError_t launch_rank_barrier(int* slots, const int64_t* peer_slot, int world, int rank, int epoch,
                                const volatile int* gate, int* err, cudaStream_t st) {
  rank_barrier_kernel<<<1, 64, 0, st>>>(slots, peer_slot, world, rank, epoch, gate, err);
  return cudaGetLastError();
}
__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}
cudaError_t launch_spin(long long cycles, cudaStream_t st) {
  spin_kernel<<<1, 1, 0, st>>>(cycles);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- launchers ----
// One wave: the grid of an update kernel is capped at the number of CTAs that can be resident at once (the kernels
// loop over the row with a grid stride).  A second wave costs a full CTA start-up + the dependent header loads
// (round counter -> schedules -> topology -> data, ~4 L2 round trips), more than a second loop iteration does.
template <typename K>
static int resident_ctas(K kernel) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;       // per kernel instantiation (one device type per process)
  std::lock_guard<std::mutex> lock(mu);
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
  return cache[key] = sms * occ;
}
template <typename T, typename K>
static dim3 grid_for(const Common<T>& c, K kernel) {
  const int per_block = THREADS * Vec<T>::N;
  int gx = (c.n_pad + per_block - 1) / per_block;
  const int slots = resident_ctas(kernel);
  if (gx * c.L > slots) {
    const int iters = (gx * c.L + slots - 1) / slots;       // loop iterations per thread that make it fit
    gx = (gx + iters - 1) / iters;
  }
  return dim3(gx > 0 ? gx : 1, c.L);
}

template <typename T> cudaError_t launch_local_sum(const Common<T>& c, cudaStream_t st) {
  const int per_block = THREADS * Vec<T>::N;
  return launch_pdl(local_sum_kernel<T>, dim3((c.n_pad + per_block - 1) / per_block), dim3(THREADS), 0, st, c);
}
template <typename T> cudaError_t launch_publish_round(const Common<T>& c, cudaStream_t st) {
  publish_round_kernel<T><<<1, 32, 0, st>>>(c);
  return cudaGetLastError();
}
// kernels that sum gradient partials: the 4-deep variant when the producer writes <= 4 partial rows per node
#define NNDT_BY_S(S, KERNEL, ARG, C)                                                              \
  ((S) <= 4 ? launch_pdl(KERNEL<T, 4>, grid_for(C, KERNEL<T, 4>), dim3(THREADS), 0, st, ARG)     \
            : launch_pdl(KERNEL<T, 16>, grid_for(C, KERNEL<T, 16>), dim3(THREADS), 0, st, ARG))
template <typename T> cudaError_t launch_dinno_update(const DinnoArgs<T>& a, cudaStream_t st) {
  return NNDT_BY_S(a.c.S, dinno_update_kernel, a, a.c);
}
template <typename T> cudaError_t launch_dsgd_mix(const Common<T>& c, cudaStream_t st) {
  return launch_pdl(dsgd_mix_kernel<T>, grid_for(c, dsgd_mix_kernel<T>), dim3(THREADS), 0, st, c);
}
template <typename T> cudaError_t launch_dsgd_step(const Common<T>& c, cudaStream_t st) {
  return NNDT_BY_S(c.S, dsgd_step_kernel, c, c);
}
template <typename T> cudaError_t launch_dsgt_init(const DsgtArgs<T>& a, cudaStream_t st) {
  return NNDT_BY_S(a.c.S, dsgt_init_kernel, a, a.c);
}
template <typename T> cudaError_t launch_dsgt_mix(const DsgtArgs<T>& a, cudaStream_t st) {
  return launch_pdl(dsgt_mix_kernel<T>, grid_for(a.c, dsgt_mix_kernel<T>), dim3(THREADS), 0, st, a);
}
template <typename T> cudaError_t launch_dsgt_track(const DsgtArgs<T>& a, cudaStream_t st) {
  return NNDT_BY_S(a.c.S, dsgt_track_kernel, a, a.c);
}
#undef NNDT_BY_S

#define NNDT_INST(T)                                                                  \
  template cudaError_t launch_local_sum<T>(const Common<T>&, cudaStream_t);           \
  template cudaError_t launch_publish_round<T>(const Common<T>&, cudaStream_t);       \
  template cudaError_t launch_consensus_metric<T>(const int64_t*, int, int, int, int, double*, double*, double*, cudaStream_t); \
  template cudaError_t launch_dinno_update<T>(const DinnoArgs<T>&, cudaStream_t);     \
  template cudaError_t launch_dsgd_mix<T>(const Common<T>&, cudaStream_t);            \
  template cudaError_t launch_dsgd_step<T>(const Common<T>&, cudaStream_t);           \
  template cudaError_t launch_dsgt_init<T>(const DsgtArgs<T>&, cudaStream_t);         \
  template cudaError_t launch_dsgt_mix<T>(const DsgtArgs<T>&, cudaStream_t);          \
  template cudaError_t launch_dsgt_track<T>(const DsgtArgs<T>&, cudaStream_t);
NNDT_INST(float)
NNDT_INST(double)

}  // namespace consensus
}  // namespace nndt