// A whole DiNNO communication round of the MNIST problem in ONE launch.
//
// Per-step kernels (mnist.cu + consensus.cu) need 2 x primal_iterations launches per round because the optimizer
// step needs the gradient of the *whole* minibatch, which is spread over the S CTAs that share a node.  Here the
// S CTAs of a node form a thread-block cluster: the hardware co-schedules them, `barrier.cluster` replaces the
// kernel boundary, and between the barriers every CTA applies the consensus update to its 1/S slice of the
// node's parameter row (neighbor pull, dual ascent, augmented-Lagrangian gradient, Adam) and publishes it.
// Nodes only meet through the published rows and the round flags, exactly as with the per-step kernels, so a
// round is:  [fwd/bwd -> cluster barrier -> update slice -> cluster barrier] x primal_iterations  in one grid.
// Reference semantics: optimizers/dinno.py:74-125 (primal_update / train loop).
#include "mnist_device.cuh"
#include "consensus_device.cuh"
#include "dinno_round.h"

namespace nndt {
namespace round {

constexpr int kNT = 768, kSPB = 8;

NNDT_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// order this thread's generic-proxy writes (global theta, smem scratch) before later async-proxy (TMA) accesses
NNDT_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

using consensus::Pack;
NNDT_DEVINL void stamp(long long* prof, int cta, int idx, int tid) {
  if (prof != nullptr && tid == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prof[cta * 64 + idx] = t;
  }
}
NNDT_DEVINL Pack<float> ldcg4(const float* p) {   // L2 load: data produced by other CTAs of this launch
  Pack<float> r;
  *reinterpret_cast<float4*>(r.v) = __ldcg(reinterpret_cast<const float4*>(p));
  return r;
}

template <int SPB, int NT>
__global__ void __launch_bounds__(NT, 1) dinno_round_kernel(const RoundArgs ra) {
  using namespace consensus;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  mnist::Smem<SPB, NT>& sm = *reinterpret_cast<mnist::Smem<SPB, NT>*>(smem_raw);
  const int tid = threadIdx.x, slice = blockIdx.x, S = gridDim.x, l = blockIdx.y;
  const Common<float>& c = ra.d.c;
  long long* prof = ra.prof;
  const int cta = l * S + slice;
  stamp(prof, cta, 0, tid);
  pdl_wait();                 // the previous round's launch (or whatever preceded) is complete and visible
  pdl_launch_dependents();
  stamp(prof, cta, 1, tid);

  const RoundInfo<float> ri = round_info(c);
  const int deg = c.deg[ri.gid * c.L + l];
  const size_t row = (size_t)l * c.n_pad;
  float* th_row = c.theta + row;
  const float* thk_row = pub_row(c, ri.par, 0, l);
  const int calls0 = ra.m.calls != nullptr ? ra.m.calls[l] : 0;
  const int nvec = c.n_pad >> 2, per = (nvec + S - 1) / S;
  const int v0 = slice * per, v1 = min(nvec, v0 + per);
  const int pits = ra.d.pits;

  for (int p = 0; p < pits; ++p) {
    // ---- forward/backward of this CTA's batch slice at the node's current parameters -------------------
    mnist::Args a = ra.m;
    a.x = ra.x_step[p]; a.y = ra.y_step[p]; a.direct_bs = ra.bs_step[p];
    mnist::stage_params<SPB, NT>(sm, a, th_row, tid, p == 0);
    const mnist::BatchGeom bg = mnist::batch_geom<true>(a, l, calls0 + p);
    mnist::process_chunk<SPB, NT, true>(sm, a, l, slice, S, 0, bg, (uint32_t)(p & 1), tid,
                                        prof != nullptr && p < 3 ? prof + cta * 64 + 16 + 16 * p : nullptr);
    stamp(prof, cta, 2 + 4 * p, tid);
    cluster_sync();           // all S gradient partials of node l are written
    stamp(prof, cta, 3 + 4 * p, tid);

    // ---- consensus update of this CTA's slice of the row ----------------------------------------------------
    const bool first = p == 0, last = p == pits - 1;
    if (first) begin_round(c, ri.gid, l, ri.k);
    const DinnoCoef<float> cf = dinno_coef(ra.d, ri.k, p, deg);
    const bool fresh = first && !ra.d.persistent;
    for (int v = v0 + tid; v < v1; v += NT) {
      const int i = v << 2;
      Pack<float> th = ldv(th_row + i);
      Pack<float> thk, dl, du;
      if (first) {
        thk = th;
#pragma unroll
        for (int u = 0; u < 4; ++u) dl.v[u] = 0.f;
        if (c.sum_mode) {
          const DPack<4> sall = network_sum(c, ri.par, 0, i);
#pragma unroll
          for (int u = 0; u < 4; ++u) dl.v[u] = (float)(sall.v[u] - (double)c.n_total * (double)thk.v[u]);
        } else {
          for (int e = 0; e < deg; ++e) {
            const Pack<float> q = ldcg4(nbr_row(c, ri.gid, l, e, ri.par, 0) + i);
#pragma unroll
            for (int u = 0; u < 4; ++u) dl.v[u] += q.v[u] - thk.v[u];
          }
        }
        du = ldv(ra.d.dual + row + i);
#pragma unroll
        for (int u = 0; u < 4; ++u) du.v[u] -= cf.rho * dl.v[u];
        stv(ra.d.delta + row + i, dl);
        stv(ra.d.dual + row + i, du);
      } else {
        thk = ldv(thk_row + i);
        dl = ldv(ra.d.delta + row + i);
        du = ldv(ra.d.dual + row + i);
      }
      Pack<float> m, mv;
      if (ra.d.opt != kSGD) {
        if (fresh) {
#pragma unroll
          for (int u = 0; u < 4; ++u) { m.v[u] = 0.f; mv.v[u] = 0.f; }
        } else {
          m = ldv(ra.d.m + row + i);
          mv = ldv(ra.d.v + row + i);
        }
      }
      const float* gp = c.grad_part + (size_t)l * S * c.n_pad + i;
      Pack<float> gq[8];
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < S) gq[s] = ldcg4(gp + (size_t)s * c.n_pad);
      Pack<float> gl = gq[0];
#pragma unroll
      for (int s = 1; s < 8; ++s)
        if (s < S) {
#pragma unroll
          for (int u = 0; u < 4; ++u) gl.v[u] += gq[s].v[u];
        }
      dinno_apply(cf, th, thk, dl, du, m, mv, gl);
      if (ra.d.opt != kSGD) {
        stv(ra.d.m + row + i, m);
        stv(ra.d.v + row + i, mv);
      }
      stv(th_row + i, th);
      if (last) stv(pub_row(c, ri.par ^ 1, 0, l) + i, th);
    }
    if (slice == 0 && tid == 0 && c.tloss != nullptr) {   // moving average of the training loss
      float loss = 0.f;
      for (int s = 0; s < c.loss_S; ++s) loss += __ldcg(c.loss_part + l * c.loss_S + s);
      const float t = c.tloss[l];
      c.tloss[l] = t != 0.f ? (1.f - c.tdecay) * t + c.tdecay * loss : loss;
    }
    stamp(prof, cta, 4 + 4 * p, tid);
    if (!last) {
      fence_proxy_async_all();   // new theta (generic stores) and smem scratch vs. the next step's TMA staging
      cluster_sync();
      stamp(prof, cta, 5 + 4 * p, tid);
    }
  }
  if (slice == 0 && tid == 0 && ra.m.calls != nullptr) ra.m.calls[l] = calls0 + pits;
  tag_published(c, l, ri.k);      // blockIdx.x == 0 is slice 0
  finish_round(c, ri.k);
  stamp(prof, cta, 63, tid);
}

static cudaError_t prepare_once() {
  static cudaError_t st = cudaFuncSetAttribute(dinno_round_kernel<kSPB, kNT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(mnist::Smem<kSPB, kNT>));
  return st;
}

static void fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int S, int L, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(S, L); cfg.blockDim = dim3(kNT);
  cfg.dynamicSmemBytes = sizeof(mnist::Smem<kSPB, kNT>); cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
}

cudaError_t launch_dinno_round(const RoundArgs& a, int S, cudaStream_t st) {
  if (S < 1 || S > 8 || a.d.pits < 1 || a.d.pits > kMaxSteps) return cudaErrorInvalidValue;
  cudaError_t e = prepare_once();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
  fill_cfg(cfg, attr, S, a.m.L, st);
  return cudaLaunchKernelEx(&cfg, dinno_round_kernel<kSPB, kNT>, a);
}

int max_active_clusters(int S) {
  if (prepare_once() != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[2];
  fill_cfg(cfg, attr, S, 1, nullptr);
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, dinno_round_kernel<kSPB, kNT>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

}  // namespace round
}  // namespace nndt
