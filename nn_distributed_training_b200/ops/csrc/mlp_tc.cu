// tcgen05 / TMEM implementation of the implicit-density MLPs (FourierNet and ReLU nets).
//
// One CTA processes tiles of 128 batch rows.  Per tile the hidden GEMMs run on the 5th-gen
// tensor cores: activations live in shared memory as bf16 128B-swizzled slabs that are written
// by the previous layer's epilogue, weights are staged once per CTA as bf16 slabs, accumulators
// are fp32 in TMEM (M = 128: TMEM lane == batch row) and are read back with tcgen05.ld for the
// fused bias+activation epilogue.  The K=2 Fourier/SIREN input layer (sin on the SFU) and the
// 64->1 output layer (+sigmoid, +loss) stay on CUDA cores inside the same kernel.
// Reference op chain: models/fourier_nn.py:33-35,48-57 driven by
// problems/dist_online_dense_problem.py:117-127 (5 eager GEMMs + elementwise kernels per pass).
#include "common.cuh"
#include "mlp.h"
#include "umma.cuh"

namespace nndt {
namespace mlp {

using namespace umma;

constexpr int NT = 256;         // 8 warps: two warpgroups share the 128 TMEM lanes (32 columns each per pass)
constexpr int TILE = 128;
constexpr int HID = 64;
constexpr int MAX_DIN = 4;
constexpr int ACT_SLAB = TILE * kSlabRowBytes;   // 16 KB: 128 rows x 64 bf16
constexpr int W_SLAB = HID * kSlabRowBytes;      // 8 KB: 64 rows x 64 bf16

template <int H1>
struct FwdSmem {
  alignas(1024) uint8_t w1[(H1 / 64) * W_SLAB];
  alignas(1024) uint8_t w2[W_SLAB];
  alignas(1024) uint8_t w3[W_SLAB];
  alignas(1024) uint8_t h1[(H1 / 64) * ACT_SLAB];
  alignas(1024) uint8_t ha[ACT_SLAB];
  alignas(1024) uint8_t hb[ACT_SLAB];
  float w0[H1 * MAX_DIN];
  float b0[H1];
  float b1[HID], b2[HID], b3[HID], w4[HID];
  float b4;
  float xs[TILE * MAX_DIN];
  float part[TILE];
  alignas(8) uint64_t bar;
  uint32_t tmem_base;
};

// stage a [64 x K] fp32 weight matrix (row-major, K multiple of 64) as K/64 bf16 swizzled slabs
NNDT_DEVINL void stage_weight(uint8_t* dst, const float* w, int K, int tid) {
  const int chunks = HID * (K / 8);            // 16-byte chunks
  for (int o = tid; o < chunks; o += NT) {
    const int n = o / (K / 8), c = o - n * (K / 8);
    const float4 lo = *reinterpret_cast<const float4*>(w + (size_t)n * K + 8 * c);
    const float4 hi = *reinterpret_cast<const float4*>(w + (size_t)n * K + 8 * c + 4);
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const int slab = c >> 3, cc = c & 7;
    *reinterpret_cast<uint4*>(dst + slab * W_SLAB + swz_chunk_off(n, cc)) = pack_bf16x8(v);
  }
}

template <int H1, class S>
NNDT_DEVINL void stage_all_weights(S& sm, const Args& a, const float* th, int tid) {
  stage_weight(sm.w1, th + a.off[2], H1, tid);
  stage_weight(sm.w2, th + a.off[4], HID, tid);
  stage_weight(sm.w3, th + a.off[6], HID, tid);
  for (int o = tid; o < H1 * a.d_in; o += NT) sm.w0[o] = th[a.off[0] + o];
  for (int o = tid; o < H1; o += NT) sm.b0[o] = th[a.off[1] + o];
  if (tid < HID) {
    sm.b1[tid] = th[a.off[3] + tid];
    sm.b2[tid] = th[a.off[5] + tid];
    sm.b3[tid] = th[a.off[7] + tid];
    sm.w4[tid] = th[a.off[8] + tid];
  }
  if (tid == 0) sm.b4 = th[a.off[9]];
}

// first layer on CUDA cores: h1[r][f] = relu(sin(scale * z)) or relu(z), z = x[r] . W0[f] + b0[f]
template <int H1, class S>
NNDT_DEVINL void first_layer(S& sm, const Args& a, int tid) {
  const int r = tid & (TILE - 1);
  const int half = tid >> 7;                       // 2 thread groups split the features
  float x[MAX_DIN];
#pragma unroll
  for (int d = 0; d < MAX_DIN; ++d) x[d] = d < a.d_in ? sm.xs[r * MAX_DIN + d] : 0.f;
  constexpr int CH = H1 / 8 / 2;                   // 16-byte chunks per thread
  for (int c = 0; c < CH; ++c) {
    const int chunk = half * CH + c;               // global chunk index along the H1 features
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int f = chunk * 8 + e;
      float z = sm.b0[f];
#pragma unroll
      for (int d = 0; d < MAX_DIN; ++d) if (d < a.d_in) z = fmaf(x[d], sm.w0[f * a.d_in + d], z);
      if (a.first_act == kFirstSinRelu) z = __sinf(a.scale * z);
      v[e] = fmaxf(z, 0.f);
    }
    const int slab = chunk >> 3, cc = chunk & 7;
    *reinterpret_cast<uint4*>(sm.h1 + slab * ACT_SLAB + swz_chunk_off(r, cc)) = pack_bf16x8(v);
  }
}

// D[128 x 64] (+)= A[128 x K] . B[64 x K]^T, both K-major; issued by one thread
NNDT_DEVINL void gemm_kmajor(uint32_t tmem_d, const uint8_t* a_slabs, const uint8_t* b_slabs, int k_slabs,
                             bool accumulate_first) {
  constexpr uint32_t idesc = make_idesc(128, 64, false, false);
  const uint32_t a0 = smem_u32(a_slabs), b0 = smem_u32(b_slabs);
  for (int s = 0; s < k_slabs; ++s)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      mma_bf16(tmem_d, desc_kmajor(a0 + s * ACT_SLAB, k), desc_kmajor(b0 + s * W_SLAB, k), idesc,
               accumulate_first || (s | k) != 0);
}

// epilogue of a 64-wide hidden layer: h = relu(acc + bias) -> bf16 slab; returns the fp32 values
NNDT_DEVINL void hidden_epilogue(uint32_t tmem_d, const float* bias, uint8_t* dst, int warp, int lane, float (&v)[32]) {
  const int row = (warp & 3) * 32 + lane;
  const int colhalf = warp >> 2;
  tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + colhalf * 32, v);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bias[colhalf * 32 + i], 0.f);
  if (dst != nullptr) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(dst + swz_chunk_off(row, colhalf * 4 + c)) = pack_bf16x8(v + 8 * c);
  }
}

template <int H1>
__global__ void __launch_bounds__(NT, 1) mlp_forward_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  FwdSmem<H1>& sm = *reinterpret_cast<FwdSmem<H1>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = blockIdx.y;
  const float* th = a.theta + (size_t)l * a.n_pad;

  if (warp == 0) tmem_alloc(&sm.tmem_base, 64);
  if (tid == 32) { mbar_init(&sm.bar, 1); mbar_init_fence(); }
  stage_all_weights<H1>(sm, a, th, tid);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = sm.tmem_base;
  uint32_t phase = 0;

  const int ntiles = (a.n_rows + TILE - 1) / TILE;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * TILE;
    for (int o = tid; o < TILE * a.d_in; o += NT) {
      const int r = o / a.d_in, d = o - r * a.d_in;
      sm.xs[r * MAX_DIN + d] = (row0 + r < a.n_rows) ? a.x[(size_t)(row0 + r) * a.d_in + d] : 0.f;
    }
    __syncthreads();
    first_layer<H1>(sm, a, tid);
    fence_async_smem();
    __syncthreads();

    float v[32];
    // ---- layer 2: [128 x H1] . W1^T ------------------------------------------------------------
    if (tid == 0) { fence_after_sync(); gemm_kmajor(tmem_d, sm.h1, sm.w1, H1 / 64, false); commit(&sm.bar); }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    hidden_epilogue(tmem_d, sm.b1, sm.ha, warp, lane, v);
    fence_before_sync(); fence_async_smem();
    __syncthreads();
    // ---- layer 3 --------------------------------------------------------------------------------
    if (tid == 0) { fence_after_sync(); gemm_kmajor(tmem_d, sm.ha, sm.w2, 1, false); commit(&sm.bar); }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    hidden_epilogue(tmem_d, sm.b2, sm.hb, warp, lane, v);
    fence_before_sync(); fence_async_smem();
    __syncthreads();
    // ---- layer 4 + output layer -------------------------------------------------------------------
    if (tid == 0) { fence_after_sync(); gemm_kmajor(tmem_d, sm.hb, sm.w3, 1, false); commit(&sm.bar); }
    mbar_wait(&sm.bar, phase); phase ^= 1;
    fence_after_sync();
    hidden_epilogue(tmem_d, sm.b3, nullptr, warp, lane, v);
    fence_before_sync();
    const int colhalf = warp >> 2, row = (warp & 3) * 32 + lane;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) dot = fmaf(v[i], sm.w4[colhalf * 32 + i], dot);
    if (colhalf == 1) sm.part[row] = dot;
    __syncthreads();
    if (colhalf == 0 && row0 + row < a.n_rows) {
      float z = dot + sm.part[row] + sm.b4;
      if (a.last_act == kLastSigmoid) z = 1.f / (1.f + __expf(-z));
      a.out[(size_t)l * a.n_rows + row0 + row] = z;
    }
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

template <int H1>
static cudaError_t launch_forward_t(const Args& a, int ctas, cudaStream_t st) {
  const int smem = (int)sizeof(FwdSmem<H1>) + 1024;
  static cudaError_t attr = cudaFuncSetAttribute(mlp_forward_kernel<H1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (attr != cudaSuccess) return attr;
  mlp_forward_kernel<H1><<<dim3(ctas, a.L), NT, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_forward(const Args& a, int ctas_per_node, cudaStream_t st) {
  if (a.d_in > MAX_DIN) return cudaErrorInvalidValue;
  switch (a.h1) {
    case 64: return launch_forward_t<64>(a, ctas_per_node, st);
    case 128: return launch_forward_t<128>(a, ctas_per_node, st);
    case 256: return launch_forward_t<256>(a, ctas_per_node, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_train(const Args& a, int ctas, cudaStream_t st) {
  (void)a; (void)ctas; (void)st;
  return cudaErrorNotSupported;
}

}  // namespace mlp
}  // namespace nndt
