// Device-side helpers of the consensus kernels (shared by consensus.cu and dinno_round.cu).
#pragma once
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

// local node handled by this CTA row: nodes whose neighbors live on other GPUs are launched first, so their NVLink
// pulls run in the shadow of the preceding forward/backward kernel (the update grid only becomes fully resident when
// that kernel drains)
template <typename T>
NNDT_DEVINL int node_of_block(const Common<T>& c) {
  return c.node_order != nullptr ? c.node_order[blockIdx.y] : (int)blockIdx.y;
}

// debug timeline: stamp `which` (0..7) of update launch (round k, primal step) by the first and the last block of the grid
template <typename T>
NNDT_DEVINL void tl_stamp(const Common<T>& c, int k, int step, int which) {
  if (c.timeline == nullptr || threadIdx.x != 0) return;
  const bool first = blockIdx.x == 0 && blockIdx.y == 0, last = blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1;
  if (!first && !last) return;
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  c.timeline[(size_t)((k * 4 + step) & 4095) * 16 + (first ? 0 : 8) + which] = t;
}

// spin until rank r has published round k (push: peers store into our local slot; pull: poll r's own counter over NVLink)
template <typename T>
NNDT_DEVINL void wait_rank(const Common<T>& c, int r, int k) {
  const int* f = c.flag_pull ? reinterpret_cast<const int*>(c.peer_pub[r]) : c.flags + r;
  if (ld_acquire_sys(f) >= k) return;
  if (*reinterpret_cast<volatile int*>(c.err) != 0) return;      // a peer already timed out: do not stack 10 s spins
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < k) {
    if (clock64() - t0 > kSpinLimit) { *c.err = 1; break; }
  }
}

// Wait until every rank owning a neighbor of local node l has published round k; with the sequence check enabled,
// also verify that the row about to be read is tagged with round k.
// Buffer reuse on time-varying graphs: this node overwrites pub[(k+1)&1] at the end of round k, the buffer its
// round-(k-1) neighbors read during round k-1.  A rank publishes round k only after its round-(k-1) reads, so also
// waiting for "round k published" from the ranks of the round-(k-1) neighbors (a no-op on static graphs: same set)
// closes the write-after-read window without a separate "consumed" counter.
template <typename T>
NNDT_DEVINL void wait_neighbors(const Common<T>& c, int gid, int l, int k) {
  const bool check = c.nbr_seq != nullptr;
  if (c.world > 1 || check) {
    const int d = c.deg[gid * c.L + l];
    if ((int)threadIdx.x < d) {
      const int r = c.world > 1 ? c.nbr_rank[(gid * c.L + l) * c.dmax + threadIdx.x] : -1;
      if (r >= 0) wait_rank(c, r, k);
      if (check) {
        const int* tag = reinterpret_cast<const int*>(c.nbr_seq[((size_t)(gid * c.L + l) * c.dmax + threadIdx.x) * 2 + (k & 1)]);
        if (ld_acquire_sys(tag) != k) *c.err = 2;
      }
    }
    if (c.world > 1 && k > 0) {
      const int gp = c.graph_id[k - 1];
      if (gp != gid) {
        const int t = (int)threadIdx.x - 32;                  // a different warp than the current-graph waiters
        if (t >= 0 && t < c.deg[gp * c.L + l]) {
          const int r = c.nbr_rank[(gp * c.L + l) * c.dmax + t];
          if (r >= 0) wait_rank(c, r, k);
        }
      }
    }
    __syncthreads();
  }
}

// tag the rows published for round k + 1 (one thread per node, before the launch's arrival counter)
template <typename T>
NNDT_DEVINL void tag_published(const Common<T>& c, int l, int k) {
  if (c.pub_seq != nullptr && blockIdx.x == 0 && threadIdx.x == 0) c.pub_seq[((k + 1) & 1) * c.pub_L + l] = k + 1;
}

// tell the peers that every round below `kn` is published (called by threads 0..world-1 of ONE block after the rows are
// ordered before this point at gpu scope).  push: one remote store per rank that ever owns a neighbor; pull: a single
// release of this rank's own counter — nothing crosses NVLink on the producer's critical path.
template <typename T>
NNDT_DEVINL void announce_round(const Common<T>& c, int kn) {
  if (c.flag_pull) {
    if (threadIdx.x == 0) { __threadfence_system(); st_release_sys(c.flags + c.rank, kn); }
  } else {
    __threadfence_system();
    if ((int)threadIdx.x < c.world && (int)threadIdx.x != c.rank && ((c.notify_mask >> threadIdx.x) & 1ull))
      st_release_sys(reinterpret_cast<int*>(c.peer_flag[threadIdx.x]), kn);
  }
}

// First consensus kernel of round k (the one that reads neighbor rows).  flags_in_kernel == 2: block (0, 0) announces
// "round k published" HERE instead of the last block of round k - 1's final kernel: that kernel wrote the rows, it is
// complete and flushed by the time any block of this one runs (every kernel passes griddepcontrol.wait before it lets its
// dependents launch), and the system fence + NVLink flag stores (3-4 us when they sit at the end of a kernel the next
// forward/backward waits for) now overlap the forward/backward kernel this launch runs under.
template <typename T>
NNDT_DEVINL void begin_round(const Common<T>& c, int gid, int l, int k) {
  if (c.world > 1 && c.flags_in_kernel == 2 && blockIdx.x == 0 && blockIdx.y == 0) announce_round(c, k);
  if (c.sum_mode) wait_all_sums(c, k); else wait_neighbors(c, gid, l, k);
}

// last block of the launch: advance the round counter and announce the new round to peers
template <typename T>
NNDT_DEVINL void finish_round(const Common<T>& c, int k) {
  // flags_in_kernel == 0: a separate publish_round_kernel on a forked graph branch announces the round to the
  // peers (system fence + remote flag stores off the local critical path); this kernel only advances the counter.
  const bool announce = c.world > 1 && c.flags_in_kernel == 1;
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    // release this block's published rows before arriving (the CTA barrier makes the fence cumulative over every
    // thread's stores).  Otherwise the kernel boundary is the only consumer-side ordering needed.
    if (announce) __threadfence();
    const unsigned total = gridDim.x * gridDim.y;
    is_last = (atomicAdd(c.done_ctr, 1u) == total - 1);
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) {
      *c.done_ctr = 0;
      *c.round_ctr = k + 1;
    }
    if (announce) announce_round(c, k + 1);
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

template <int U, typename T>
NNDT_DEVINL Pack<T> sum_partials(const Common<T>& c, int l, int i) {
  // all (up to U) partial loads are issued before the first add: one L2 round trip instead of S dependent ones;
  // the summation order stays s = 0, 1, 2, ...  (U = 4 for the cluster kernels' <= 4 partial rows per node: the
  // 16-deep variant costs 48 more registers and halves the occupancy of the update kernels)
  constexpr int N = Vec<T>::N;
  const T* gp = c.grad_part + (size_t)l * c.S * c.n_pad + i;
  Pack<T> q[U];
#pragma unroll
  for (int s = 0; s < U; ++s)
    if (s < c.S) q[s] = ldv(gp + (size_t)s * c.n_pad);
  Pack<T> g = q[0];
#pragma unroll
  for (int s = 1; s < U; ++s)
    if (s < c.S) {
#pragma unroll
      for (int u = 0; u < N; ++u) g.v[u] += q[s].v[u];
    }
  for (int s = U; s < c.S; ++s) {
    const Pack<T> r = ldv(gp + (size_t)s * c.n_pad);
#pragma unroll
    for (int u = 0; u < N; ++u) g.v[u] += r.v[u];
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

// ---- DiNNO arithmetic on one vector (optimizers/dinno.py:74-91): augmented-Lagrangian gradient + optimizer step ----
template <typename T>
struct DinnoCoef {
  T rho, lr, step_size, bc2s;
  int deg, opt;
};
template <typename T>
NNDT_DEVINL DinnoCoef<T> dinno_coef(const DinnoArgs<T>& a, int k, int step, int deg) {
  DinnoCoef<T> q;
  q.rho = a.c.rho[k]; q.lr = a.c.lr[k];
  const int t = a.persistent ? k * a.pits + step + 1 : step + 1;
  const T bc1 = (T)1 - pow((T)0.9, (T)t);
  q.bc2s = sqrt((T)1 - pow((T)0.999, (T)t));
  q.step_size = q.lr / bc1;
  q.deg = deg; q.opt = a.opt;
  return q;
}
template <typename T>
NNDT_DEVINL void dinno_apply(const DinnoCoef<T>& q, Pack<T>& th, const Pack<T>& thk, const Pack<T>& dl, const Pack<T>& du,
                             Pack<T>& m, Pack<T>& v, const Pack<T>& gl) {
  constexpr int N = Vec<T>::N;
  const T b1 = (T)0.9, b2 = (T)0.999, eps = (T)1e-8, wd = (T)1e-2;
  Pack<T> g;
#pragma unroll
  for (int u = 0; u < N; ++u)
    g.v[u] = gl.v[u] + du.v[u] + (T)2 * q.rho * (T)q.deg * (th.v[u] - thk.v[u]) - q.rho * dl.v[u];
  if (q.opt == kSGD) {
#pragma unroll
    for (int u = 0; u < N; ++u) th.v[u] -= q.lr * g.v[u];
  } else {
#pragma unroll
    for (int u = 0; u < N; ++u) {
      if (q.opt == kAdamW) th.v[u] *= ((T)1 - q.lr * wd);
      m.v[u] = b1 * m.v[u] + ((T)1 - b1) * g.v[u];
      v.v[u] = b2 * v.v[u] + ((T)1 - b2) * g.v[u] * g.v[u];
      th.v[u] -= q.step_size * m.v[u] / (sqrt(v.v[u]) / q.bc2s + eps);
    }
  }
}

}  // namespace consensus
}  // namespace nndt
