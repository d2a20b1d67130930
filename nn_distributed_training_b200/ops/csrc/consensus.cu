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
#include "common.cuh"
#include "consensus.h"

namespace nndt {
namespace consensus {

constexpr int THREADS = 256;
constexpr long long kSpinLimit = 20000000000LL;  // ~10 s at 2 GHz, then flag an error and go on

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

template <typename T> struct Pack { T v[Vec<T>::N]; };

template <typename T>
NNDT_DEVINL Pack<T> ldv(const T* p) {
  Pack<T> r;
  *reinterpret_cast<typename Vec<T>::type*>(r.v) = *reinterpret_cast<const typename Vec<T>::type*>(p);
  return r;
}
template <typename T>
NNDT_DEVINL void stv(T* p, const Pack<T>& r) {
  *reinterpret_cast<typename Vec<T>::type*>(p) = *reinterpret_cast<const typename Vec<T>::type*>(r.v);
}

template <typename T>
struct RoundInfo { int k, par, gid; };

template <typename T>
NNDT_DEVINL RoundInfo<T> round_info(const Common<T>& c) {
  RoundInfo<T> r;
  r.k = *c.round_ctr;
  r.par = r.k & 1;
  r.gid = c.graph_id[r.k];
  return r;
}

// wait until every rank owning a neighbor of local node l has published round k
template <typename T>
NNDT_DEVINL void wait_neighbors(const Common<T>& c, int gid, int l, int k) {
  if (c.world > 1) {
    const int d = c.deg[gid * c.L + l];
    if ((int)threadIdx.x < d) {
      const int r = c.nbr_rank[(gid * c.L + l) * c.dmax + threadIdx.x];
      if (r >= 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(c.flags + r) < k) {
          if (clock64() - t0 > kSpinLimit) { *c.err = 1; break; }
        }
      }
    }
    __syncthreads();
  }
}

// last block of the launch: advance the round counter and announce the new round to peers
template <typename T>
NNDT_DEVINL void finish_round(const Common<T>& c, int k) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    is_last = (atomicAdd(c.done_ctr, 1u) == total - 1);
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) {
      *c.done_ctr = 0;
      *c.round_ctr = k + 1;
    }
    if (c.world > 1) {
      __threadfence_system();
      if ((int)threadIdx.x < c.world && (int)threadIdx.x != c.rank)
        st_release_sys(reinterpret_cast<int*>(c.peer_flag[threadIdx.x]), k + 1);
    }
  }
}

template <typename T>
NNDT_DEVINL const T* nbr_row(const Common<T>& c, int gid, int l, int e, int par, int chan) {
  return reinterpret_cast<const T*>(c.nbr_ptr[(((size_t)(gid * c.L + l) * c.dmax + e) * 2 + par) * c.C + chan]);
}
template <typename T>
NNDT_DEVINL T* pub_row(const Common<T>& c, int par, int chan, int l) {
  return c.pub + ((size_t)(par * c.C + chan) * c.pub_L + l) * c.n_pad;
}

// ---- complete-graph mode -----------------------------------------------------------------------
// network-wide sum of channel `chan` at element i for parity `par`
template <int N> struct DPack { double v[N]; };
template <typename T>
NNDT_DEVINL DPack<Vec<T>::N> network_sum(const Common<T>& c, int par, int chan, int i) {
  constexpr int N = Vec<T>::N;
  const size_t off = (size_t)(par * c.C + chan) * c.n_pad + i;
  DPack<N> r;
  if (c.sum_mc != nullptr) {
#pragma unroll
    for (int u = 0; u < N; ++u)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(r.v[u]) : "l"(c.sum_mc + off + u) : "memory");
  } else {
#pragma unroll
    for (int u = 0; u < N; u += 2) {
      const double2 q = *reinterpret_cast<const double2*>(c.sum_local + off + u);
      r.v[u] = q.x; r.v[u + 1] = q.y;
    }
  }
  return r;
}
// every rank's partial sum of round k must be in place before the in-switch reduction reads it
template <typename T>
NNDT_DEVINL void wait_all_sums(const Common<T>& c, int k) {
  if (c.world > 1) {
    if ((int)threadIdx.x < c.world && (int)threadIdx.x != c.rank) {
      const long long t0 = clock64();
      while (ld_acquire_sys(c.sum_flags + threadIdx.x) < k + 1) {
        if (clock64() - t0 > kSpinLimit) { *c.err = 1; break; }
      }
    }
    __syncthreads();
  }
}
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

template <typename T>
NNDT_DEVINL Pack<T> sum_partials(const Common<T>& c, int l, int i) {
  const T* gp = c.grad_part + (size_t)l * c.S * c.n_pad + i;
  Pack<T> g = ldv(gp);
  for (int s = 1; s < c.S; ++s) {
    const Pack<T> q = ldv(gp + (size_t)s * c.n_pad);
#pragma unroll
    for (int u = 0; u < Vec<T>::N; ++u) g.v[u] += q.v[u];
  }
  return g;
}

// step bookkeeping done by one thread per node in the kernel that consumes a gradient: advance the sampler's
// draw counter and fold the step's training loss into the moving average (problems/dist_online_dense_problem.py:129-137)
template <typename T>
NNDT_DEVINL void step_bookkeeping(const Common<T>& c, int l) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (c.calls != nullptr) c.calls[l] += 1;
    if (c.tloss != nullptr) {
      float loss = 0.f;
      for (int s = 0; s < c.loss_S; ++s) loss += c.loss_part[l * c.loss_S + s];
      const float t = c.tloss[l];
      c.tloss[l] = t != 0.f ? (1.f - c.tdecay) * t + c.tdecay * loss : loss;
    }
  }
}

// ------------------------------------------------------------------ DiNNO ----
template <typename T>
__global__ void __launch_bounds__(THREADS) dinno_update_kernel(const DinnoArgs<T> a) {
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = blockIdx.y;
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  const T rho = c.rho[ri.k], lr = c.lr[ri.k];
  const bool first = a.step == 0, last = a.step == a.pits - 1;
  if (first) { if (c.sum_mode) wait_all_sums(c, ri.k); else wait_neighbors(c, ri.gid, l, ri.k); }

  const T b1 = (T)0.9, b2 = (T)0.999, eps = (T)1e-8, wd = (T)1e-2;
  const int t = a.persistent ? ri.k * a.pits + a.step + 1 : a.step + 1;
  const T bc1 = (T)1 - pow((T)0.9, (T)t);
  const T bc2s = sqrt((T)1 - pow((T)0.999, (T)t));
  const T step_size = lr / bc1;
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
        for (int e = 0; e < deg; ++e) {
          const Pack<T> q = ldv(nbr_row(c, ri.gid, l, e, ri.par, 0) + i);
#pragma unroll
          for (int u = 0; u < N; ++u) dl.v[u] += q.v[u] - thk.v[u];
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
    if (!waited) { pdl_wait(); pdl_launch_dependents(); waited = true; }
    const Pack<T> gl = sum_partials(c, l, i);
    Pack<T> g;
#pragma unroll
    for (int u = 0; u < N; ++u)
      g.v[u] = gl.v[u] + du.v[u] + (T)2 * rho * (T)deg * (th.v[u] - thk.v[u]) - rho * dl.v[u];
    if (a.opt == kSGD) {
#pragma unroll
      for (int u = 0; u < N; ++u) th.v[u] -= lr * g.v[u];
    } else {
#pragma unroll
      for (int u = 0; u < N; ++u) {
        if (a.opt == kAdamW) th.v[u] *= ((T)1 - lr * wd);
        m.v[u] = b1 * m.v[u] + ((T)1 - b1) * g.v[u];
        v.v[u] = b2 * v.v[u] + ((T)1 - b2) * g.v[u] * g.v[u];
        th.v[u] -= step_size * m.v[u] / (sqrt(v.v[u]) / bc2s + eps);
      }
      stv(a.m + row + i, m);
      stv(a.v + row + i, v);
    }
    stv(c.theta + row + i, th);
    if (last) stv(pub_row(c, ri.par ^ 1, 0, l) + i, th);
  }
  if (!waited) { pdl_wait(); pdl_launch_dependents(); }
  step_bookkeeping(c, l);
  if (last) finish_round(c, ri.k);
}

// ------------------------------------------------------------------- DSGD ----
template <typename T>
__global__ void __launch_bounds__(THREADS) dsgd_mix_kernel(const Common<T> c) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Vec<T>::N;
  const int l = blockIdx.y;
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  if (c.sum_mode) wait_all_sums(c, ri.k); else wait_neighbors(c, ri.gid, l, ri.k);
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
    for (int e = 0; e < deg; ++e) {
      const Pack<T> q = ldv(nbr_row(c, ri.gid, l, e, ri.par, 0) + i);
      const T we = w[e];
#pragma unroll
      for (int u = 0; u < N; ++u) th.v[u] += we * q.v[u];
    }
    stv(c.theta + row + i, th);
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS) dsgd_step_kernel(const Common<T> c) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Vec<T>::N;
  const int l = blockIdx.y;
  const RoundInfo<T> ri = round_info(c);
  const T alpha = c.alpha[ri.k];
  const size_t row = (size_t)l * c.n_pad;
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    Pack<T> th = ldv(c.theta + row + i);
    const Pack<T> g = sum_partials(c, l, i);
#pragma unroll
    for (int u = 0; u < N; ++u) th.v[u] -= alpha * g.v[u];
    stv(c.theta + row + i, th);
    stv(pub_row(c, ri.par ^ 1, 0, l) + i, th);
  }
  step_bookkeeping(c, l);
  finish_round(c, ri.k);
}

// ------------------------------------------------------------------- DSGT ----
// channel 0 of the published buffer is theta, channel 1 the gradient tracker y.
template <typename T>
__global__ void __launch_bounds__(THREADS) dsgt_init_kernel(const DsgtArgs<T> a) {
  pdl_wait();
  pdl_launch_dependents();
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = blockIdx.y;
  const size_t row = (size_t)l * c.n_pad;
  for (int i = (blockIdx.x * THREADS + threadIdx.x) * N; i < c.n_pad; i += gridDim.x * THREADS * N) {
    const Pack<T> g = sum_partials(c, l, i);
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
  const int l = blockIdx.y;
  const RoundInfo<T> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  if (c.sum_mode) wait_all_sums(c, ri.k); else wait_neighbors(c, ri.gid, l, ri.k);
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
    for (int e = 0; e < deg; ++e) {
      const Pack<T> qt = ldv(nbr_row(c, ri.gid, l, e, ri.par, 0) + i);
      const Pack<T> qy = ldv(nbr_row(c, ri.gid, l, e, ri.par, 1) + i);
      const T we = w[e];
#pragma unroll
      for (int u = 0; u < N; ++u) th.v[u] += we * (qt.v[u] - alpha * qy.v[u]);
    }
    stv(c.theta + row + i, th);
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS) dsgt_track_kernel(const DsgtArgs<T> a) {
  pdl_wait();
  pdl_launch_dependents();
  const Common<T>& c = a.c;
  constexpr int N = Vec<T>::N;
  const int l = blockIdx.y;
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
      for (int e = 0; e < deg; ++e) {
        const Pack<T> qy = ldv(nbr_row(c, ri.gid, l, e, ri.par, 1) + i);
        const T we = w[e];
#pragma unroll
        for (int u = 0; u < N; ++u) y.v[u] += we * qy.v[u];
      }
    }
    const Pack<T> gn = sum_partials(c, l, i);
    const Pack<T> go = ldv(a.g_old + row + i);
#pragma unroll
    for (int u = 0; u < N; ++u) y.v[u] += gn.v[u] - go.v[u];
    stv(a.g_old + row + i, gn);
    stv(pub_row(c, ri.par ^ 1, 1, l) + i, y);
    stv(pub_row(c, ri.par ^ 1, 0, l) + i, ldv(c.theta + row + i));
  }
  step_bookkeeping(c, l);
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

// ---------------------------------------------------------------- launchers ----
template <typename T>
static dim3 grid_for(const Common<T>& c) {
  const int per_block = THREADS * Vec<T>::N;
  int gx = (c.n_pad + per_block - 1) / per_block;
  return dim3(gx, c.L);
}

template <typename T> cudaError_t launch_local_sum(const Common<T>& c, cudaStream_t st) {
  const int per_block = THREADS * Vec<T>::N;
  return launch_pdl(local_sum_kernel<T>, dim3((c.n_pad + per_block - 1) / per_block), dim3(THREADS), 0, st, c);
}
template <typename T> cudaError_t launch_dinno_update(const DinnoArgs<T>& a, cudaStream_t st) {
  return launch_pdl(dinno_update_kernel<T>, grid_for(a.c), dim3(THREADS), 0, st, a);
}
template <typename T> cudaError_t launch_dsgd_mix(const Common<T>& c, cudaStream_t st) {
  return launch_pdl(dsgd_mix_kernel<T>, grid_for(c), dim3(THREADS), 0, st, c);
}
template <typename T> cudaError_t launch_dsgd_step(const Common<T>& c, cudaStream_t st) {
  return launch_pdl(dsgd_step_kernel<T>, grid_for(c), dim3(THREADS), 0, st, c);
}
template <typename T> cudaError_t launch_dsgt_init(const DsgtArgs<T>& a, cudaStream_t st) {
  return launch_pdl(dsgt_init_kernel<T>, grid_for(a.c), dim3(THREADS), 0, st, a);
}
template <typename T> cudaError_t launch_dsgt_mix(const DsgtArgs<T>& a, cudaStream_t st) {
  return launch_pdl(dsgt_mix_kernel<T>, grid_for(a.c), dim3(THREADS), 0, st, a);
}
template <typename T> cudaError_t launch_dsgt_track(const DsgtArgs<T>& a, cudaStream_t st) {
  return launch_pdl(dsgt_track_kernel<T>, grid_for(a.c), dim3(THREADS), 0, st, a);
}

#define NNDT_INST(T)                                                                  \
  template cudaError_t launch_local_sum<T>(const Common<T>&, cudaStream_t);           \
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
