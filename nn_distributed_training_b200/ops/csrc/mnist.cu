// Fused forward+backward (and forward-only evaluation) of the MNIST conv classifier
//   conv 1->3 5x5 -> ReLU -> maxpool2 -> fc 432->64 -> ReLU -> fc 64->10 -> log-softmax -> NLL
// for every graph node hosted by this GPU in ONE launch (reference op chain:
// models/mnist_conv_nn.py:16-25 driven by problems/dist_mnist_problem.py:96-98, where it is
// ~30 eager ATen/cuDNN launches per node and step).
//
// Work decomposition: grid = (S batch slices, L nodes); a CTA owns SPB samples of one node.
// The 110 KB fc1 weight matrix is staged once per CTA into shared memory with cp.async
// (row stride padded to 436 floats -> conflict-free 128-bit reads along j), overlapped with
// the conv+ReLU+pool phase which runs out of an even/odd column-split image tile
// (conflict-free stride-2 accesses).  Minibatch rows are gathered in-kernel through the
// stateless Feistel sampler, so there is no host work, no index tensor and no H2D copy per
// step.  Each CTA writes its slice's partial gradient row; the consensus update kernel that
// follows sums the S partials while applying the optimizer step (no separate reduce launch).
//
// Batch 64 x 28k parameters is far below a tcgen05 tile's break-even (M=64 of a 128-row MMA,
// K=432; the whole fc1 GEMM is 1.8 MFLOP) — see DESIGN.md §MNIST for the arithmetic; the
// tensor-core path of this framework is ops/csrc/mlp_tc.cu (Fourier / ReLU MLPs).
#include "common.cuh"
#include "sampler.cuh"
#include "mnist.h"

namespace nndt {
namespace mnist {

constexpr int F = 3, KS = 5, HW = 28, PHW = 12, NPOOL = 144;
constexpr int FC1_IN = 432, HID = 64, NCLS = 10;
constexpr int W1_STRIDE = 436;          // padded fc1 row stride in smem (floats)
constexpr int XROW = 22;                // padded row stride of the even/odd column planes
constexpr int XPLANE = HW * XROW;       // 616 floats per plane
constexpr int CGROUP = 256;             // threads cooperating on one conv channel in the backward

// NT threads per CTA: 64 hidden units x (NT/64) k-slices must tile the 108 float4 of an fc1 row.
template <int NT> struct Geo {
  static constexpr int KSLICES = NT / 64;
  static constexpr int K4S = (FC1_IN / 4) / KSLICES;
  static_assert((FC1_IN / 4) % KSLICES == 0, "NT/64 must divide 108");
  static_assert(NT >= 3 * CGROUP && NT >= FC1_IN, "need >= 768 threads");
};

template <int SPB, int NT>
struct Smem {
  float w1[HID * W1_STRIDE];            // fc1 weights; later scratch for dW1 transposition / conv-grad reduce
  float xe[SPB * XPLANE];
  float xo[SPB * XPLANE];
  float a1[SPB * FC1_IN];
  float da1[SPB * FC1_IN];
  float hpart[Geo<NT>::KSLICES * SPB * HID];
  float h[SPB * HID];
  float dh[SPB * HID];
  float dhT[HID * SPB];
  float w2[NCLS * HID];
  float z[SPB * 16];
  float dz[SPB * 16];
  float wc[F * KS * KS + 4];
  float red[80];
  float b1[HID];
  float b2[16];
  int sidx[SPB];
  int label[SPB];
  float valid[SPB];
  unsigned char arg[SPB * FC1_IN];
  alignas(8) uint64_t w1_bar;           // TMA transaction barrier of the fc1 weight staging
};

template <int SPB, int NT>
__device__ __forceinline__ void load_images(Smem<SPB, NT>& sm, const Args& a, int tid) {
  // 196 groups of 4 pixels per sample; issue every global load of this thread before the
  // first conversion so the latencies overlap
  constexpr int PER = (SPB * 196 + NT - 1) / NT;
  uint32_t raw_u8[PER];
  float4 raw_f[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int o = tid + i * NT;
    raw_u8[i] = 0; raw_f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o < SPB * 196) {
      const int s = o / 196, q = o - s * 196;
      if (sm.valid[s] != 0.f) {
        const size_t base = (size_t)sm.sidx[s] * 784 + 4 * q;
        if (a.x_is_u8) raw_u8[i] = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(a.x) + base);
        else raw_f[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + base);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int o = tid + i * NT;
    if (o < SPB * 196) {
      const int s = o / 196, q = o - s * 196;
      const int row = (4 * q) / HW, col = (4 * q) - row * HW;
      float v0, v1, v2, v3;
      if (a.x_is_u8) {
        const bool ok = sm.valid[s] != 0.f;
        const uint32_t p = raw_u8[i];
        v0 = ok ? ((p & 0xff) * (1.f / 255.f) - a.mean) * a.inv_std : 0.f;
        v1 = ok ? (((p >> 8) & 0xff) * (1.f / 255.f) - a.mean) * a.inv_std : 0.f;
        v2 = ok ? (((p >> 16) & 0xff) * (1.f / 255.f) - a.mean) * a.inv_std : 0.f;
        v3 = ok ? ((p >> 24) * (1.f / 255.f) - a.mean) * a.inv_std : 0.f;
      } else {
        v0 = raw_f[i].x; v1 = raw_f[i].y; v2 = raw_f[i].z; v3 = raw_f[i].w;
      }
      float* e = sm.xe + s * XPLANE + row * XROW + (col >> 1);
      float* d = sm.xo + s * XPLANE + row * XROW + (col >> 1);
      e[0] = v0; d[0] = v1; e[1] = v2; d[1] = v3;
    }
  }
}

// pixel (r, c) of sample s
template <int SPB, int NT>
__device__ __forceinline__ float px(const Smem<SPB, NT>& sm, int s, int r, int c) {
  return ((c & 1) ? sm.xo : sm.xe)[s * XPLANE + r * XROW + (c >> 1)];
}

template <int SPB, int NT>
__device__ __forceinline__ void conv_relu_pool(Smem<SPB, NT>& sm, int tid) {
  for (int it = tid; it < SPB * NPOOL; it += NT) {
    const int s = it / NPOOL, p = it - s * NPOOL;
    const int py = p / PHW, pxx = p - py * PHW;
    float patch[6][6];
    const float* e = sm.xe + s * XPLANE + (2 * py) * XROW + pxx;
    const float* o = sm.xo + s * XPLANE + (2 * py) * XROW + pxx;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        patch[r][2 * c] = e[r * XROW + c];
        patch[r][2 * c + 1] = o[r * XROW + c];
      }
    }
#pragma unroll
    for (int c = 0; c < F; ++c) {
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float w = sm.wc[c * 25 + ky * 5 + kx];
          acc[0][0] = fmaf(w, patch[ky][kx], acc[0][0]);
          acc[0][1] = fmaf(w, patch[ky][kx + 1], acc[0][1]);
          acc[1][0] = fmaf(w, patch[ky + 1][kx], acc[1][0]);
          acc[1][1] = fmaf(w, patch[ky + 1][kx + 1], acc[1][1]);
        }
      }
      // first maximum wins, like ATen's max_pool2d
      float m = acc[0][0]; int ai = 0;
      if (acc[0][1] > m) { m = acc[0][1]; ai = 1; }
      if (acc[1][0] > m) { m = acc[1][0]; ai = 2; }
      if (acc[1][1] > m) { m = acc[1][1]; ai = 3; }
      m += sm.wc[75 + c];
      sm.a1[s * FC1_IN + c * NPOOL + p] = fmaxf(m, 0.f);
      sm.arg[s * FC1_IN + c * NPOOL + p] = (unsigned char)ai;
    }
  }
}

template <int SPB, int NT>
__device__ __forceinline__ void fc1_forward(Smem<SPB, NT>& sm, int tid) {
  constexpr int K4S = Geo<NT>::K4S;
  const int j = tid & 63, ks = tid >> 6;
  float acc[SPB];
#pragma unroll
  for (int s = 0; s < SPB; ++s) acc[s] = 0.f;
  const float4* wrow = reinterpret_cast<const float4*>(sm.w1 + j * W1_STRIDE) + ks * K4S;
  const float4* arow = reinterpret_cast<const float4*>(sm.a1) + ks * K4S;
#pragma unroll
  for (int i = 0; i < K4S; ++i) {
    const float4 w = wrow[i];
#pragma unroll
    for (int s = 0; s < SPB; ++s) {
      const float4 x = arow[s * (FC1_IN / 4) + i];
      acc[s] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[s]))));
    }
  }
#pragma unroll
  for (int s = 0; s < SPB; ++s) sm.hpart[(ks * SPB + s) * HID + j] = acc[s];
}

template <int SPB, int NT, bool TRAIN>
__global__ void __launch_bounds__(NT, 1) mnist_kernel(const Args a) {
  constexpr int KSLICES = Geo<NT>::KSLICES, K4S = Geo<NT>::K4S;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<SPB, NT>& sm = *reinterpret_cast<Smem<SPB, NT>*>(smem_raw);
  const int tid = threadIdx.x;
  const int l = blockIdx.y;
  const float* th = a.theta + (size_t)l * a.n_pad;
  // everything below reads parameters / draw counters written by the preceding consensus kernel
  pdl_wait();
  pdl_launch_dependents();

  // ---- stage fc1 weights with the TMA engine (one bulk copy per 1728-byte row into the padded smem rows,
  //      completion tracked by an mbarrier transaction count) and the small tensors with plain loads -------
  {
    const float* w1g = th + a.off_w1;
    if (tid == 0) {
      mbarrier_init(&sm.w1_bar, 1);
      mbarrier_expect_tx(&sm.w1_bar, HID * FC1_IN * 4);
      for (int j = 0; j < HID; ++j) tma_bulk_g2s(sm.w1 + j * W1_STRIDE, w1g + j * FC1_IN, FC1_IN * 4, &sm.w1_bar);
    }
    if (tid < 75) sm.wc[tid] = th[a.off_wc + tid];
    if (tid < 3) sm.wc[75 + tid] = th[a.off_bc + tid];
    if (tid < HID) sm.b1[tid] = th[a.off_b1 + tid];
    if (tid < NCLS) sm.b2[tid] = th[a.off_b2 + tid];
    for (int o = tid; o < NCLS * HID; o += NT) sm.w2[o] = th[a.off_w2 + o];
  }

  // ---- batch geometry ---------------------------------------------------------------------
  uint32_t bs = 0, start = 0, key = 0, m = 0;
  int shard_off = 0;
  if (TRAIN) {
    if (a.direct) {
      bs = a.direct_bs != nullptr ? (uint32_t)a.direct_bs[l] : (uint32_t)a.batch;
    } else {
      m = (uint32_t)a.shard_len[l];
      shard_off = a.shard_off[l];
      const BatchLoc loc = locate_batch((uint32_t)a.calls[l], m, (uint32_t)a.batch);
      bs = loc.size; start = loc.start;
      key = mix_key((uint32_t)a.seed, (uint32_t)(a.node0 + l), loc.epoch);
    }
  }
  const float inv_bs = TRAIN ? 1.f / (float)(bs ? bs : 1) : 1.f;

  const int n_chunks = TRAIN ? 1 : (a.n_val + SPB - 1) / SPB;
  for (int chunk = TRAIN ? 0 : blockIdx.x; chunk < n_chunks; chunk += TRAIN ? 1 : gridDim.x) {
    // ---- which samples ---------------------------------------------------------------------
    if (tid < SPB) {
      int idx = 0; float ok = 0.f;
      if (TRAIN) {
        const uint32_t t = blockIdx.x * SPB + tid;
        if (t < bs) {
          ok = 1.f;
          idx = a.direct ? (int)(l * a.batch + t) : shard_off + (int)feistel_permute(start + t, m, key);
        }
      } else {
        const int t = chunk * SPB + tid;
        if (t < a.n_val) { ok = 1.f; idx = t; }
      }
      sm.sidx[tid] = idx;
      sm.valid[tid] = ok;
      sm.label[tid] = ok != 0.f ? (int)a.y[idx] : 0;
    }
    __syncthreads();
    load_images<SPB, NT>(sm, a, tid);
    __syncthreads();
    conv_relu_pool<SPB, NT>(sm, tid);
    mbarrier_wait_parity(&sm.w1_bar, 0);   // fc1 weights have landed (no-op after the first chunk)
    __syncthreads();

    // ---- fc1 -------------------------------------------------------------------------------
    fc1_forward<SPB, NT>(sm, tid);
    __syncthreads();
    for (int o = tid; o < SPB * HID; o += NT) {
      const int s = o >> 6, j = o & 63;
      float v = sm.b1[j];
#pragma unroll
      for (int ks = 0; ks < KSLICES; ++ks) v += sm.hpart[(ks * SPB + s) * HID + j];
      sm.h[o] = fmaxf(v, 0.f);
    }
    __syncthreads();

    // ---- fc2 + log-softmax + NLL -----------------------------------------------------------
    if (tid < SPB * NCLS) {
      const int s = tid / NCLS, c = tid - s * NCLS;
      float v = sm.b2[c];
#pragma unroll 8
      for (int j = 0; j < HID; ++j) v = fmaf(sm.h[s * HID + j], sm.w2[c * HID + j], v);
      sm.z[s * 16 + c] = v;
    }
    __syncthreads();
    if (tid < SPB) {
      const int s = tid;
      float mx = sm.z[s * 16];
      int am = 0;
#pragma unroll
      for (int c = 1; c < NCLS; ++c) if (sm.z[s * 16 + c] > mx) { mx = sm.z[s * 16 + c]; am = c; }
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < NCLS; ++c) se += __expf(sm.z[s * 16 + c] - mx);
      const float lse = mx + __logf(se);
      const int y = sm.label[s];
      const float ok = sm.valid[s];
      const float loss = ok * (lse - sm.z[s * 16 + y]);
      if (TRAIN) {
#pragma unroll
        for (int c = 0; c < NCLS; ++c)
          sm.dz[s * 16 + c] = ok * inv_bs * (__expf(sm.z[s * 16 + c] - lse) - (c == y ? 1.f : 0.f));
        sm.red[s] = loss;
      } else if (ok != 0.f) {
        const size_t o = (size_t)l * a.n_val + sm.sidx[s];
        a.val_loss[o] = loss;
        a.val_correct[o] = (unsigned char)(am == y);
      }
    }
    __syncthreads();
    if (!TRAIN) continue;

    float* gp = a.grad_part + ((size_t)l * gridDim.x + blockIdx.x) * a.n_pad;
    if (tid == 0) {
      float tot = 0.f;
#pragma unroll
      for (int s = 0; s < SPB; ++s) tot += sm.red[s];
      a.loss_part[l * gridDim.x + blockIdx.x] = tot * inv_bs;
    }
    // ---- fc2 grads, dh ----------------------------------------------------------------------
    for (int o = tid; o < NCLS * HID; o += NT) {
      const int c = o >> 6, j = o & 63;
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < SPB; ++s) v = fmaf(sm.dz[s * 16 + c], sm.h[s * HID + j], v);
      gp[a.off_w2 + o] = v;
    }
    if (tid < NCLS) {
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < SPB; ++s) v += sm.dz[s * 16 + tid];
      gp[a.off_b2 + tid] = v;
    }
    for (int o = tid; o < SPB * HID; o += NT) {
      const int s = o >> 6, j = o & 63;
      float v = 0.f;
#pragma unroll
      for (int c = 0; c < NCLS; ++c) v = fmaf(sm.dz[s * 16 + c], sm.w2[c * HID + j], v);
      v = sm.h[o] > 0.f ? v : 0.f;
      sm.dh[o] = v;
      sm.dhT[j * SPB + s] = v;
    }
    __syncthreads();
    if (tid < HID) {
      float v = 0.f;
#pragma unroll
      for (int s = 0; s < SPB; ++s) v += sm.dh[s * HID + tid];
      gp[a.off_b1 + tid] = v;
    }
    // ---- da1 = dh . W1 (masked by ReLU): one fc1 input k per thread ---------------------------
    if (tid < FC1_IN) {
      float acc[SPB];
#pragma unroll
      for (int s = 0; s < SPB; ++s) acc[s] = 0.f;
#pragma unroll 8
      for (int j = 0; j < HID; ++j) {
        const float w = sm.w1[j * W1_STRIDE + tid];
#pragma unroll
        for (int s = 0; s < SPB; ++s) acc[s] = fmaf(sm.dhT[j * SPB + s], w, acc[s]);
      }
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        const int k = s * FC1_IN + tid;
        sm.da1[k] = sm.a1[k] > 0.f ? acc[s] : 0.f;
      }
    }
    // ---- dW1[j][k] = sum_s dh[s][j] a1[s][k]: register tile per (j, k-slice) -------------------
    float4 dw[K4S];
    {
      const int j = tid & 63, ks = tid >> 6;
#pragma unroll
      for (int i = 0; i < K4S; ++i) dw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* arow = reinterpret_cast<const float4*>(sm.a1) + ks * K4S;
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        const float d = sm.dh[s * HID + j];
#pragma unroll
        for (int i = 0; i < K4S; ++i) {
          const float4 x = arow[s * (FC1_IN / 4) + i];
          dw[i].x = fmaf(d, x.x, dw[i].x);
          dw[i].y = fmaf(d, x.y, dw[i].y);
          dw[i].z = fmaf(d, x.z, dw[i].z);
          dw[i].w = fmaf(d, x.w, dw[i].w);
        }
      }
    }
    __syncthreads();   // every read of the staged W1 is done: its smem becomes scratch
    {
      // transpose through smem so the global stores are fully coalesced
      const int j = tid & 63, ks = tid >> 6;
      float4* srow = reinterpret_cast<float4*>(sm.w1 + j * W1_STRIDE) + ks * K4S;
#pragma unroll
      for (int i = 0; i < K4S; ++i) srow[i] = dw[i];
    }
    __syncthreads();
    {
      float4* out = reinterpret_cast<float4*>(gp + a.off_w1);
      for (int o = tid; o < HID * (FC1_IN / 4); o += NT) {
        const int j = o / (FC1_IN / 4), k4 = o - j * (FC1_IN / 4);
        out[o] = *reinterpret_cast<const float4*>(sm.w1 + j * W1_STRIDE + 4 * k4);
      }
    }
    // ---- conv grads: each pooled cell routes da1 to its argmax conv position ------------------
    // 3 groups of 256 threads, one per channel; partial sums are transposed through smem
    // (scratch = the dead W1 region) and reduced by warps — no 26x5 shuffle trees.
    float cacc[26];
#pragma unroll
    for (int i = 0; i < 26; ++i) cacc[i] = 0.f;
    const int cg = tid / CGROUP, ct = tid - cg * CGROUP;
    if (cg < F) {
      for (int it = ct; it < SPB * NPOOL; it += CGROUP) {
        const int s = it / NPOOL, p = it - s * NPOOL;
        const float g = sm.da1[s * FC1_IN + cg * NPOOL + p];
        if (g != 0.f) {
          const int ai = sm.arg[s * FC1_IN + cg * NPOOL + p];
          const int py = p / PHW, pxx = p - py * PHW;
          const int r0 = 2 * py + (ai >> 1), c0 = 2 * pxx + (ai & 1);
#pragma unroll
          for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
              cacc[ky * 5 + kx] = fmaf(g, px<SPB, NT>(sm, s, r0 + ky, c0 + kx), cacc[ky * 5 + kx]);
          cacc[25] += g;
        }
      }
    }
    __syncthreads();   // dW1 copy-out finished reading the scratch
    float* scratch = sm.w1;   // [78][CGROUP]
    if (cg < F) {
#pragma unroll
      for (int i = 0; i < 26; ++i) scratch[(cg * 26 + i) * CGROUP + ct] = cacc[i];
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int o = warp; o < 78; o += NT / 32) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < CGROUP / 32; ++q) v += scratch[o * CGROUP + lane + 32 * q];
        v = warp_sum(v);
        if (lane == 0) {
          const int c = o / 26, i = o - c * 26;
          gp[(i < 25) ? a.off_wc + c * 25 + i : a.off_bc + c] = v;
        }
      }
    }
  }
}

constexpr int kNT = 768;

template <int SPB, bool TRAIN>
static cudaError_t prepare_once() {
  // opt in to >48 KB dynamic shared memory once per process (not a stream op: legal under capture)
  static cudaError_t st = cudaFuncSetAttribute(mnist_kernel<SPB, kNT, TRAIN>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(Smem<SPB, kNT>));
  return st;
}

template <int SPB>
static cudaError_t launch_t(const Args& a, int S, bool train, int eval_ctas, cudaStream_t st) {
  const size_t smem = sizeof(Smem<SPB, kNT>);
  cudaError_t e = train ? prepare_once<SPB, true>() : prepare_once<SPB, false>();
  if (e != cudaSuccess) return e;
  if (train) {
    return launch_pdl(mnist_kernel<SPB, kNT, true>, dim3(S, a.L), dim3(kNT), smem, st, a);
  } else {
    mnist_kernel<SPB, kNT, false><<<dim3(eval_ctas, a.L), kNT, smem, st>>>(a);
  }
  return cudaGetLastError();
}

// ---- device-initiated H2D staging of one round's minibatches (one warp per row) -------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherArgs a) {
  // few, long-lived blocks: the staging copy must leave most SMs (and their register files) to the
  // training kernels it overlaps with; each warp streams several rows with all loads of a row in flight
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int rows = a.P * a.L * a.batch;
  const int r = *a.stage_round;
  const int n16 = a.row_bytes >> 4;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += nwarps) {
    const int t = row % a.batch, l = (row / a.batch) % a.L, p = row / (a.batch * a.L);
    const uint32_t m = (uint32_t)a.shard_len[l];
    const BatchLoc loc = locate_batch((uint32_t)(a.calls0[l] + r * a.P + p), m, (uint32_t)a.batch);
    if (t == 0 && lane == 0) a.bs_stage[p * a.L + l] = (int)loc.size;
    if ((uint32_t)t < loc.size) {
      const uint32_t key = mix_key((uint32_t)a.seed, (uint32_t)(a.node0 + l), loc.epoch);
      const size_t src = (size_t)a.shard_off[l] + feistel_permute(loc.start + t, m, key);
      const uint4* s4 = reinterpret_cast<const uint4*>(a.x_host + src * a.row_bytes);
      uint4* d4 = reinterpret_cast<uint4*>(a.x_stage + (size_t)row * a.row_bytes);
      uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
      if (lane < n16) v0 = s4[lane];
      if (lane + 32 < n16) v1 = s4[lane + 32];
      if (lane < n16) d4[lane] = v0;
      if (lane + 32 < n16) d4[lane + 32] = v1;
      for (int i = lane + 64; i < n16; i += 32) d4[i] = s4[i];
      if (lane == 0) a.y_stage[row] = a.y_host[src];
    }
  }
  // last block advances the staged-round counter
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(a.done_ctr, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x == 0) { *a.done_ctr = 0; *a.stage_round = r + 1; }
}
cudaError_t launch_gather(const GatherArgs& a, cudaStream_t st) {
  const int rows = a.P * a.L * a.batch;
  int blocks = (rows * 32 + 255) / 256;
  if (blocks > 24) blocks = 24;
  gather_rows_kernel<<<blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

// debug/test helper: the sampler's row indices for one draw (mirrors data/sampler.py)
__global__ void batch_indices_kernel(int m, int B, int call, int seed, int node, int* out, int* out_size) {
  const BatchLoc loc = locate_batch((uint32_t)call, (uint32_t)m, (uint32_t)B);
  const uint32_t key = mix_key((uint32_t)seed, (uint32_t)node, loc.epoch);
  for (uint32_t t = threadIdx.x; t < loc.size; t += blockDim.x)
    out[t] = (int)feistel_permute(loc.start + t, (uint32_t)m, key);
  if (threadIdx.x == 0) *out_size = (int)loc.size;
}
cudaError_t launch_batch_indices(int m, int B, int call, int seed, int node, int* out, int* out_size, cudaStream_t st) {
  batch_indices_kernel<<<1, 256, 0, st>>>(m, B, call, seed, node, out, out_size);
  return cudaGetLastError();
}

cudaError_t launch_train(const Args& a, int spb, int S, cudaStream_t st) {
  if (spb == 8) return launch_t<8>(a, S, true, 0, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_eval(const Args& a, int ctas_per_node, cudaStream_t st) {
  return launch_t<8>(a, 0, false, ctas_per_node, st);
}

}  // namespace mnist
}  // namespace nndt
