// Launch interface of the fused consensus kernels (consensus.cu).
//
// One launch covers every graph node hosted by this GPU.  Neighbor rows are read through a
// device pointer table, so a neighbor may be another row of the local arena (virtual nodes on
// one GPU) or a row of a peer GPU's symmetric-memory arena mapped over NVLink — the kernel is
// identical, the graph only changes which pointers are in the table.  All per-round scalars
// (rho_k, lr_k, alpha_k, graph id) are indexed by a device-side round counter so a captured
// CUDA graph can be replayed for any number of rounds with no host involvement.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nndt {
namespace consensus {

enum Opt : int { kSGD = 0, kAdam = 1, kAdamW = 2 };

template <typename T>
struct Common {
  // geometry
  int L, n_pad, S;            // local nodes, padded row length, gradient partials per node
  // live state (rows of the local arena)
  T* theta;                   // [L, n_pad]
  const T* grad_part;         // [L, S, n_pad]
  T* pub;                     // [2 parity, C chan, pub_L, n_pad] published rows (symmetric memory)
  int C, pub_L;
  // topology tables for G graphs
  const int64_t* nbr_ptr;     // [G, L, dmax, 2, C] device addresses of neighbor rows per parity/chan
  const T* nbr_w;             // [G, L, dmax]
  const T* self_w;            // [G, L]
  const int* deg;             // [G, L]
  const int* nbr_rank;        // [G, L, dmax] owning rank of each neighbor, -1 when local
  int dmax;
  // device-side schedules
  int* round_ctr;             // [1]
  const T* rho; const T* lr; const T* alpha;   // [oits]
  const int* graph_id;        // [oits]
  // sampler bookkeeping
  int* calls;                 // [L] or nullptr
  // optional device-side moving average of the training loss (online density metric, tloss_decay)
  const float* loss_part;     // [L, loss_S] per-CTA loss partials of the forward/backward kernel
  float* tloss;               // [L] tracker, nullptr = off
  float tdecay; int loss_S;
  // cross-GPU sync (nullptr / 0 when single GPU)
  int* flags;                 // [world] local slots written by peers (round published)
  const int64_t* peer_flag;   // [world] address of *their* slot for this rank
  int world, rank;
  // flag transport: push (default) = the producer stores k+1 into the reader's local slot over NVLink and readers spin on
  // local memory; pull = the producer only releases its own counter (slot `rank` of its own array, no remote store on its
  // critical path) and readers poll that slot over NVLink through `peer_pub`.
  int flag_pull;
  const int64_t* peer_pub;    // [world] address of rank r's own counter (pull mode)
  unsigned long long notify_mask;  // ranks that ever own a neighbor of a local node: the only ones told about a new round
  const int* node_order;      // [L] launch order of the local nodes (nodes with remote neighbors first), nullptr = identity
  long long* timeline;        // debug (NNDT_TIMELINE=1): [4096][16] %globaltimer stamps of the update kernels, nullptr = off
  unsigned int* done_ctr;     // [1] last-block detection
  int* err;                   // [1] 1 = spin timeout, 2 = sequence check failed
  // optional debug build of the protocol: every published row carries the round it belongs to and every
  // neighbor read verifies it (nullptr = off)
  int* pub_seq;               // [2 parity, pub_L] local tags, written with the published rows
  const int64_t* nbr_seq;     // [G, L, dmax, 2] device addresses of the neighbors' tags per parity
  int flags_in_kernel;        // who tells the peers that a round is published: 2 = the first consensus kernel of the round that
                              // reads it (default), 1 = the last kernel of the round that wrote it, 0 = publish_round_kernel
  // complete-graph ("sum") mode: Metropolis weights are uniform 1/N, so every aggregate is a function of
  // S = sum over ALL nodes.  Each rank reduces its local rows into `sum_local` and the consumers fetch the
  // network-wide sum either with one NVLS in-switch reduction (multimem.ld_reduce over `sum_mc`) or, on a
  // single GPU, straight from `sum_local`.
  int sum_mode;               // 0 = pointer-table neighbors, 1 = complete graph via sums
  int n_total;                // N (all nodes of the graph)
  double* sum_local;          // [2 parity, C, n_pad] this rank's partial sums, always fp64: S - N theta_i cancels
  const double* sum_mc;       //   catastrophically in fp32 near consensus.  sum_mc = multicast mapping (or nullptr)
  int* sum_flags;             // [world] "partial sum of round k ready" flags written by peers
  const int64_t* peer_sum_flag;  // [world]
};

template <typename T>
struct DinnoArgs {
  Common<T> c;
  T* dual; T* delta; T* m; T* v;   // [L, n_pad]
  int step, pits, opt, persistent;
};

template <typename T>
struct DsgtArgs {
  Common<T> c;
  T* g_old;                        // [L, n_pad]
};

template <typename T> cudaError_t launch_dinno_update(const DinnoArgs<T>& a, cudaStream_t st);
template <typename T> cudaError_t launch_dsgd_mix(const Common<T>& c, cudaStream_t st);
template <typename T> cudaError_t launch_dsgd_step(const Common<T>& c, cudaStream_t st);
template <typename T> cudaError_t launch_dsgt_init(const DsgtArgs<T>& a, cudaStream_t st);
template <typename T> cudaError_t launch_dsgt_mix(const DsgtArgs<T>& a, cudaStream_t st);
template <typename T> cudaError_t launch_dsgt_track(const DsgtArgs<T>& a, cudaStream_t st);
template <typename T> cudaError_t launch_local_sum(const Common<T>& c, cudaStream_t st);
template <typename T> cudaError_t launch_publish_round(const Common<T>& c, cudaStream_t st);

// All-rank barrier on the device (bench start alignment, metric quiescence): every rank stores `epoch` into its slot of
// every peer's array and spins until all of its own slots reached it; with `gate` != nullptr the kernel first spins on that
// (pinned host) word until the host sets it, so everything enqueued behind it starts at the host's command.
cudaError_t launch_rank_barrier(int* slots, const int64_t* peer_slot, int world, int rank, int epoch,
                                const volatile int* gate, int* err, cudaStream_t st);
// Busy-wait `cycles` SM clocks (tests: a deliberately delayed rank)
cudaError_t launch_spin(long long cycles, cudaStream_t st);

// K6 consensus metric (problems/dist_mnist_problem.py:155-169): distances between L2-normalised parameter rows.
// rows[j] is the device address of node j's current row (local, or a peer GPU's published row over NVLink).
// out_pair [L, N] = |th_i/|th_i| - th_j/|th_j||,  out_mean [L] = |th_i/|th_i| - mean_j th_j/|th_j||   (fp64 accumulation)
template <typename T>
cudaError_t launch_consensus_metric(const int64_t* rows, int N, int n_pad, int local0, int L, double* inv_norm,
                                    double* out_pair, double* out_mean, cudaStream_t st);

}  // namespace consensus
}  // namespace nndt
