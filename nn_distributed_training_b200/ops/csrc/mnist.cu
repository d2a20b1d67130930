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
constexpr int THREADS = 256;
constexpr int KSLICES = 4, K4_PER_SLICE = FC1_IN / 4 / KSLICES;  // 27 float4 per k-slice

template <int SPB>
struct Smem {
  float w1[HID * W1_STRIDE];
  float xe[SPB * XPLANE];
  float xo[SPB * XPLANE];
  float a1[SPB * FC1_IN];
  float da1[SPB * FC1_IN];
  float hpart[KSLICES * SPB * HID];
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
};

template <int SPB>
__device__ __forceinline__ void load_images(Smem<SPB>& sm, const Args& a, int tid) {
  // 196 groups of 4 pixels per sample; columns 4q..4q+3 of one row never straddle rows (28 % 4 == 0)
  for (int o = tid; o < SPB * 196; o += THREADS) {
    const int s = o / 196, q = o - s * 196;
    const int row = (4 * q) / HW, col = (4 * q) - row * HW;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    if (sm.valid[s] != 0.f) {
      const size_t base = (size_t)sm.sidx[s] * 784 + 4 * q;
      if (a.x_is_u8) {
        const uchar4 p = *reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(a.x) + base);
        v0 = (p.x * (1.f / 255.f) - a.mean) * a.inv_std;
        v1 = (p.y * (1.f / 255.f) - a.mean) * a.inv_std;
        v2 = (p.z * (1.f / 255.f) - a.mean) * a.inv_std;
        v3 = (p.w * (1.f / 255.f) - a.mean) * a.inv_std;
      } else {
        const float4 p = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + base);
        v0 = p.x; v1 = p.y; v2 = p.z; v3 = p.w;
      }
    }
    float* e = sm.xe + s * XPLANE + row * XROW + (col >> 1);
    float* d = sm.xo + s * XPLANE + row * XROW + (col >> 1);
    e[0] = v0; d[0] = v1; e[1] = v2; d[1] = v3;
  }
}

// pixel (r, c) of sample s
template <int SPB>
__device__ __forceinline__ float px(const Smem<SPB>& sm, int s, int r, int c) {
  return ((c & 1) ? sm.xo : sm.xe)[s * XPLANE + r * XROW + (c >> 1)];
}

template <int SPB>
__device__ __forceinline__ void conv_relu_pool(Smem<SPB>& sm, int tid) {
  for (int it = tid; it < SPB * NPOOL; it += THREADS) {
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

template <int SPB>
__device__ __forceinline__ void fc1_forward(Smem<SPB>& sm, int tid) {
  const int j = tid & 63, ks = tid >> 6;
  float acc[SPB];
#pragma unroll
  for (int s = 0; s < SPB; ++s) acc[s] = 0.f;
  const float4* wrow = reinterpret_cast<const float4*>(sm.w1 + j * W1_STRIDE) + ks * K4_PER_SLICE;
  const float4* arow = reinterpret_cast<const float4*>(sm.a1) + ks * K4_PER_SLICE;
#pragma unroll 3
  for (int i = 0; i < K4_PER_SLICE; ++i) {
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

template <int SPB, bool TRAIN>
__global__ void __launch_bounds__(THREADS, 1) mnist_kernel(const Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem<SPB>& sm = *reinterpret_cast<Smem<SPB>*>(smem_raw);
  const int tid = threadIdx.x;
  const int l = blockIdx.y;
  const float* th = a.theta + (size_t)l * a.n_pad;

  // ---- stage fc1 weights (async) and the small tensors ----------------------------------
  {
    const float* w1g = th + a.off_w1;
    for (int o = tid; o < HID * (FC1_IN / 4); o += THREADS) {
      const int j = o / (FC1_IN / 4), k4 = o - j * (FC1_IN / 4);
      cp_async16(sm.w1 + j * W1_STRIDE + 4 * k4, w1g + j * FC1_IN + 4 * k4);
    }
    cp_async_commit();
    if (tid < 75) sm.wc[tid] = th[a.off_wc + tid];
    if (tid < 3) sm.wc[75 + tid] = th[a.off_bc + tid];
    if (tid < HID) sm.b1[tid] = th[a.off_b1 + tid];
    if (tid < NCLS) sm.b2[tid] = th[a.off_b2 + tid];
    for (int o = tid; o < NCLS * HID; o += THREADS) sm.w2[o] = th[a.off_w2 + o];
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
    load_images<SPB>(sm, a, tid);
    __syncthreads();
    conv_relu_pool<SPB>(sm, tid);
    cp_async_wait<0>();
    __syncthreads();

    // ---- fc1 -------------------------------------------------------------------------------
    fc1_forward<SPB>(sm, tid);
    __syncthreads();
    for (int o = tid; o < SPB * HID; o += THREADS) {
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
    for (int o = tid; o < NCLS * HID; o += THREADS) {
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
    for (int o = tid; o < SPB * HID; o += THREADS) {
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
    // ---- da1 = dh . W1 (masked by ReLU) ------------------------------------------------------
    if (tid < FC1_IN / 2) {
      float acc0[SPB], acc1[SPB];
#pragma unroll
      for (int s = 0; s < SPB; ++s) { acc0[s] = 0.f; acc1[s] = 0.f; }
#pragma unroll 4
      for (int j = 0; j < HID; ++j) {
        const float2 w = *reinterpret_cast<const float2*>(sm.w1 + j * W1_STRIDE + 2 * tid);
#pragma unroll
        for (int s = 0; s < SPB; ++s) {
          const float d = sm.dhT[j * SPB + s];
          acc0[s] = fmaf(d, w.x, acc0[s]);
          acc1[s] = fmaf(d, w.y, acc1[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        const int k = s * FC1_IN + 2 * tid;
        sm.da1[k] = sm.a1[k] > 0.f ? acc0[s] : 0.f;
        sm.da1[k + 1] = sm.a1[k + 1] > 0.f ? acc1[s] : 0.f;
      }
    }
    // ---- dW1[j][k] = sum_s dh[s][j] a1[s][k]  (register tile of 108 per thread) ---------------
    {
      const int j = tid & 63, ks = tid >> 6;
      float4 acc[K4_PER_SLICE];
#pragma unroll
      for (int i = 0; i < K4_PER_SLICE; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* arow = reinterpret_cast<const float4*>(sm.a1) + ks * K4_PER_SLICE;
#pragma unroll
      for (int s = 0; s < SPB; ++s) {
        const float d = sm.dh[s * HID + j];
#pragma unroll
        for (int i = 0; i < K4_PER_SLICE; ++i) {
          const float4 x = arow[s * (FC1_IN / 4) + i];
          acc[i].x = fmaf(d, x.x, acc[i].x);
          acc[i].y = fmaf(d, x.y, acc[i].y);
          acc[i].z = fmaf(d, x.z, acc[i].z);
          acc[i].w = fmaf(d, x.w, acc[i].w);
        }
      }
      float4* out = reinterpret_cast<float4*>(gp + a.off_w1 + j * FC1_IN) + ks * K4_PER_SLICE;
#pragma unroll
      for (int i = 0; i < K4_PER_SLICE; ++i) out[i] = acc[i];
    }
    __syncthreads();
    // ---- conv grads: each pooled cell routes da1 to its argmax conv position ------------------
    if (tid < 80) sm.red[tid] = 0.f;
    __syncthreads();
    for (int c = 0; c < F; ++c) {
      float acc[26];
#pragma unroll
      for (int i = 0; i < 26; ++i) acc[i] = 0.f;
      for (int it = tid; it < SPB * NPOOL; it += THREADS) {
        const int s = it / NPOOL, p = it - s * NPOOL;
        const float g = sm.da1[s * FC1_IN + c * NPOOL + p];
        if (g != 0.f) {
          const int ai = sm.arg[s * FC1_IN + c * NPOOL + p];
          const int py = p / PHW, pxx = p - py * PHW;
          const int r0 = 2 * py + (ai >> 1), c0 = 2 * pxx + (ai & 1);
#pragma unroll
          for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx)
              acc[ky * 5 + kx] = fmaf(g, px<SPB>(sm, s, r0 + ky, c0 + kx), acc[ky * 5 + kx]);
          acc[25] += g;
        }
      }
#pragma unroll
      for (int i = 0; i < 26; ++i) {
        const float v = warp_sum(acc[i]);
        if ((tid & 31) == 0) atomicAdd(&sm.red[i < 25 ? c * 25 + i : 75 + c], v);
      }
    }
    __syncthreads();
    if (tid < 75) gp[a.off_wc + tid] = sm.red[tid];
    if (tid < 3) gp[a.off_bc + tid] = sm.red[75 + tid];
  }
}

template <int SPB, bool TRAIN>
static cudaError_t prepare_once() {
  // opt in to >48 KB dynamic shared memory once per process (not a stream op: legal under capture)
  static cudaError_t st = cudaFuncSetAttribute(mnist_kernel<SPB, TRAIN>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(Smem<SPB>));
  return st;
}

template <int SPB>
static cudaError_t launch_t(const Args& a, int S, bool train, int eval_ctas, cudaStream_t st) {
  const size_t smem = sizeof(Smem<SPB>);
  cudaError_t e = train ? prepare_once<SPB, true>() : prepare_once<SPB, false>();
  if (e != cudaSuccess) return e;
  if (train) {
    mnist_kernel<SPB, true><<<dim3(S, a.L), THREADS, smem, st>>>(a);
  } else {
    mnist_kernel<SPB, false><<<dim3(eval_ctas, a.L), THREADS, smem, st>>>(a);
  }
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
  if (spb == 4) return launch_t<4>(a, S, true, 0, st);
  return cudaErrorInvalidValue;
}

cudaError_t launch_eval(const Args& a, int ctas_per_node, cudaStream_t st) {
  return launch_t<8>(a, 0, false, ctas_per_node, st);
}

}  // namespace mnist
}  // namespace nndt
