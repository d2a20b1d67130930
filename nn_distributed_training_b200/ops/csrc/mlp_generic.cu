// Generic fully-connected network forward / backward on CUDA cores for ANY layer widths <= 256 and
// ReLU / Tanh / Sigmoid / identity activations, fp32 or fp64 — the kernels behind FFReLUNet / FFTanhNet / FFSigmoidNet
// (reference: models/relu_nn.py:4-116 takes any `shape`) and the RL actors / critics [12,64,64,64,5] / [12,64,64,64,1]
// (reference: RL/dist_rl/model.py:6-45), whose shapes the tcgen05 kernel of mlp_tc.cu (d_in <= 4, 64-wide trunk, scalar
// output, loss fused) does not cover.  Module-level contract (ops/mlp_generic.py: torch.autograd.Function):
//   forward : x [M, d0], flat parameters  ->  every layer's post-activation output, saved in `acts` [M, sum d_l]
//   backward: dL/dout [M, d_last]          ->  dL/dparams (atomicAdd into a zeroed flat vector) and optionally dL/dx
// One CTA owns 32 batch rows and walks the layers with the activations in shared memory; weights stream from L2 in
// transposed 32-wide K chunks; a thread accumulates RG rows x 1 output column in registers.
#include "common.cuh"
#include "mlp_generic.h"

namespace nndt {
namespace mlpg {

namespace {

constexpr int NT = 256, ROWS = 32, KCH = 32, MAXW = 256, LDW = MAXW + 1;

template <typename T> NNDT_DEVINL T act_fwd(T z, int a) {
  if (a == kRelu) return z > (T)0 ? z : (T)0;
  if (a == kTanh) return tanh(z);
  if (a == kSigmoid) return (T)1 / ((T)1 + exp(-z));
  return z;
}
// derivative expressed through the OUTPUT of the activation
template <typename T> NNDT_DEVINL T act_bwd(T y, int a) {
  if (a == kRelu) return y > (T)0 ? (T)1 : (T)0;
  if (a == kTanh) return (T)1 - y * y;
  if (a == kSigmoid) return y * ((T)1 - y);
  return (T)1;
}

template <typename T>
struct Sm {
  T* in; T* out; T* wt;
  __device__ explicit Sm(unsigned char* p) {
    in = reinterpret_cast<T*>(p);
    out = in + ROWS * LDW;
    wt = out + ROWS * LDW;
  }
  static constexpr size_t bytes() { return sizeof(T) * (2 * ROWS * LDW + KCH * LDW); }
};

// out[r][j] = act(b[j] + sum_k in[r][k] W[j][k]) for RG rows per thread; thread = (row group g, column j)
template <typename T, int RG>
NNDT_DEVINL void layer_fwd(const T* in, T* out, T* wt, const T* W, const T* b, int din, int dout, int dpad, int act, int tid) {
  const int j = tid % dpad, g = tid / dpad;
  const bool live = j < dout && g * RG < ROWS;
  T acc[RG];
#pragma unroll
  for (int r = 0; r < RG; ++r) acc[r] = 0;
  for (int k0 = 0; k0 < din; k0 += KCH) {
    const int kc = min(KCH, din - k0);
    __syncthreads();
    for (int o = tid; o < dout * KCH; o += NT) {          // W[jj][k0 + kk] -> wt[kk][jj] (coalesced along k)
      const int jj = o / KCH, kk = o - jj * KCH;
      wt[kk * LDW + jj] = kk < kc ? __ldg(W + (size_t)jj * din + k0 + kk) : (T)0;
    }
    __syncthreads();
    if (live) {
      for (int kk = 0; kk < kc; ++kk) {
        const T w = wt[kk * LDW + j];
#pragma unroll
        for (int r = 0; r < RG; ++r) acc[r] += in[(g * RG + r) * LDW + k0 + kk] * w;
      }
    }
  }
  if (live) {
    const T bj = __ldg(b + j);
#pragma unroll
    for (int r = 0; r < RG; ++r) out[(g * RG + r) * LDW + j] = act_fwd(acc[r] + bj, act);
  }
}

template <typename T>
NNDT_DEVINL void layer_fwd_dispatch(const T* in, T* out, T* wt, const T* W, const T* b, int din, int dout, int act, int tid) {
  int dpad = 8;
  while (dpad < dout) dpad <<= 1;
  switch (dpad) {
    case 256: layer_fwd<T, 32>(in, out, wt, W, b, din, dout, dpad, act, tid); break;
    case 128: layer_fwd<T, 16>(in, out, wt, W, b, din, dout, dpad, act, tid); break;
    case 64: layer_fwd<T, 8>(in, out, wt, W, b, din, dout, dpad, act, tid); break;
    case 32: layer_fwd<T, 4>(in, out, wt, W, b, din, dout, dpad, act, tid); break;
    case 16: layer_fwd<T, 2>(in, out, wt, W, b, din, dout, dpad, act, tid); break;
    default: layer_fwd<T, 1>(in, out, wt, W, b, din, dout, dpad, act, tid); break;
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) mlp_generic_forward_kernel(const Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm<T> sm(smem_raw);
  const int tid = threadIdx.x, row0 = blockIdx.x * ROWS;
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* P = reinterpret_cast<const T*>(a.params);
  T* acts = reinterpret_cast<T*>(a.acts);
  T* cur = sm.in; T* nxt = sm.out;
  for (int o = tid; o < ROWS * a.dims[0]; o += NT) {
    const int r = o / a.dims[0], k = o - r * a.dims[0];
    cur[r * LDW + k] = row0 + r < a.M ? x[(size_t)(row0 + r) * a.dims[0] + k] : (T)0;
  }
  for (int l = 0; l < a.nl; ++l) {
    const int din = a.dims[l], dout = a.dims[l + 1];
    layer_fwd_dispatch<T>(cur, nxt, sm.wt, P + a.w_off[l], P + a.b_off[l], din, dout, a.act[l], tid);
    __syncthreads();
    for (int o = tid; o < ROWS * dout; o += NT) {          // save the post-activation output for the backward pass
      const int r = o / dout, j = o - r * dout;
      if (row0 + r < a.M) acts[(size_t)(row0 + r) * a.act_stride + a.act_off[l] + j] = nxt[r * LDW + j];
    }
    T* t = cur; cur = nxt; nxt = t;
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) mlp_generic_backward_kernel(const Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Sm<T> sm(smem_raw);
  T* dz = sm.in; T* ain = sm.out; T* dnext = sm.wt;       // dnext: ROWS x LDW fits in the KCH x LDW weight stage (KCH == ROWS)
  const int tid = threadIdx.x, row0 = blockIdx.x * ROWS;
  const T* x = reinterpret_cast<const T*>(a.x);
  const T* P = reinterpret_cast<const T*>(a.params);
  const T* acts = reinterpret_cast<const T*>(a.acts);
  const T* gout = reinterpret_cast<const T*>(a.gout);
  T* G = reinterpret_cast<T*>(a.gparams);
  const int L = a.nl;
  {
    const int d = a.dims[L];
    for (int o = tid; o < ROWS * d; o += NT) {             // dz = dL/dout * act'(out)
      const int r = o / d, j = o - r * d;
      T v = 0;
      if (row0 + r < a.M) {
        const T y = acts[(size_t)(row0 + r) * a.act_stride + a.act_off[L - 1] + j];
        v = gout[(size_t)(row0 + r) * d + j] * act_bwd(y, a.act[L - 1]);
      }
      dz[r * LDW + j] = v;
    }
  }
  for (int l = L - 1; l >= 0; --l) {
    const int din = a.dims[l], dout = a.dims[l + 1];
    for (int o = tid; o < ROWS * din; o += NT) {           // the layer's input: previous activation (or x)
      const int r = o / din, k = o - r * din;
      T v = 0;
      if (row0 + r < a.M)
        v = l == 0 ? x[(size_t)(row0 + r) * din + k] : acts[(size_t)(row0 + r) * a.act_stride + a.act_off[l - 1] + k];
      ain[r * LDW + k] = v;
    }
    __syncthreads();
    // dW[j][k] += sum_r dz[r][j] ain[r][k]   (k fastest: coalesced atomics)
    T* gW = G + a.w_off[l];
    for (int o = tid; o < dout * din; o += NT) {
      const int j = o / din, k = o - j * din;
      T v = 0;
#pragma unroll 8
      for (int r = 0; r < ROWS; ++r) v += dz[r * LDW + j] * ain[r * LDW + k];
      atomicAdd(gW + o, v);
    }
    for (int j = tid; j < dout; j += NT) {
      T v = 0;
#pragma unroll 8
      for (int r = 0; r < ROWS; ++r) v += dz[r * LDW + j];
      atomicAdd(G + a.b_off[l] + j, v);
    }
    if (l == 0 && a.gx == nullptr) break;
    // d(input)[r][k] = sum_j dz[r][j] W[j][k]
    const T* W = P + a.w_off[l];
    for (int o = tid; o < ROWS * din; o += NT) {
      const int r = o / din, k = o - r * din;
      T v = 0;
      for (int j = 0; j < dout; ++j) v += dz[r * LDW + j] * __ldg(W + (size_t)j * din + k);
      dnext[r * LDW + k] = l == 0 ? v : v * act_bwd(ain[r * LDW + k], a.act[l - 1]);
    }
    __syncthreads();
    if (l == 0) {
      T* gx = reinterpret_cast<T*>(a.gx);
      for (int o = tid; o < ROWS * din; o += NT) {
        const int r = o / din, k = o - r * din;
        if (row0 + r < a.M) gx[(size_t)(row0 + r) * din + k] = dnext[r * LDW + k];
      }
    } else {
      for (int o = tid; o < ROWS * din; o += NT) { const int r = o / din, k = o - r * din; dz[r * LDW + k] = dnext[r * LDW + k]; }
      __syncthreads();
    }
  }
}

template <typename T, typename K>
cudaError_t launch(K kern, const Args& a, cudaStream_t st) {
  const size_t smem = Sm<T>::bytes();
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(a.M + ROWS - 1) / ROWS, NT, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

bool supported(const Args& a) {
  if (a.nl < 1 || a.nl > kMaxLayers || a.M < 1) return false;
  for (int l = 0; l <= a.nl; ++l)
    if (a.dims[l] < 1 || a.dims[l] > MAXW) return false;
  return true;
}

cudaError_t launch_forward(const Args& a, cudaStream_t st) {
  if (!supported(a)) return cudaErrorInvalidValue;
  return a.dtype64 ? launch<double>(mlp_generic_forward_kernel<double>, a, st) : launch<float>(mlp_generic_forward_kernel<float>, a, st);
}

cudaError_t launch_backward(const Args& a, cudaStream_t st) {
  if (!supported(a)) return cudaErrorInvalidValue;
  return a.dtype64 ? launch<double>(mlp_generic_backward_kernel<double>, a, st) : launch<float>(mlp_generic_backward_kernel<float>, a, st);
}

}  // namespace mlpg
}  // namespace nndt
