// MNIST conv-net forward+backward with the fc1-sized contractions on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, W1 slices fetched by tensor-map TMA) — the training kernel of the
// paper's MNISTConvNet(3, 5, 64) at batch <= 64 (reference op chain: models/mnist_conv_nn.py:16-25, driven per node
// and step by problems/dist_mnist_problem.py:96-98).
//
// Decomposition: the work of a node is cut along K (the 432 fc1 inputs), not along the batch.  A node is a cluster
// of 6 CTAs; CTA c owns pooled rows {2c, 2c+1} of all three channels = 72 fc1 inputs, for ALL 64 samples:
//   * conv + ReLU + maxpool of its 8 image rows on CUDA cores into a 128B-swizzled A tile [64 samples x 3 x 32]
//     (24 real columns + 8 zero columns per channel atom);
//   * its W1 slice [64 x 3 x 32] arrives by three cp.async.bulk.tensor boxes (UTMALDG) — 24 KB per CTA instead of
//     the whole 110 KB matrix per CTA of the batch-split kernel (mnist.cu);
//   * H_c = A_c . W1_c^T  (M = 64 samples, N = 64, K = 96) is one tcgen05 accumulator; fp32 accuracy comes from the
//     3xTF32 split (x = hi + lo, both exactly representable in TF32): lo.hi + hi.lo + hi.hi accumulate in TMEM;
//   * the six partial H_c are reduced through distributed shared memory (each CTA reduces ~11 samples, runs
//     fc2 + log-softmax + NLL + their backward for them, and pushes the dH rows into every CTA's operand buffer);
//   * da1_c = dH . W1_c and dW1_c = dH^T . A_c need NO cross-CTA traffic: they reuse the same shared-memory tiles
//     through MN-major descriptors, land in TMEM, and dW1_c goes straight to its columns of the node's gradient row;
//   * conv gradients route da1 through the argmax positions; the small gradients (conv, b1, fc2) are reduced through
//     DSMEM by rank 0.  The node's gradient is ONE row (no per-slice partials for the consensus kernel to sum).
#include <cuda.h>

#include "mnist_device.cuh"
#include "umma.cuh"

namespace nndt {
namespace mnist {

namespace tc {

constexpr int NT = 768, CL = 6, CELLS = 24, KC = 72;      // threads, cluster size, pooled cells per channel, real K per CTA
constexpr int SLAB = 64 * 128;                            // one swizzle-128B slab: 64 rows x 128 B
constexpr int XP = 8 * 14;                                // floats per sample and column-parity plane (8 image rows)
constexpr int HP_STRIDE = 65;
constexpr int PART_WC = 0, PART_BC = 75, PART_B1 = 78, PART_W2 = 142, PART_B2 = 782, PART_LOSS = 792, PART_N = 793;
constexpr uint32_t TM_D1 = 0, TM_D2 = 64, TM_D3 = 160, TM_COLS = 256;

struct Smem {
  alignas(1024) unsigned char w_hi[3 * SLAB];   // TMA destination (raw fp32), rewritten in place as the TF32 "hi" part
  unsigned char w_lo[3 * SLAB];
  unsigned char a_hi[3 * SLAB];
  unsigned char a_lo[3 * SLAB];
  unsigned char dh_hi[2 * SLAB];
  unsigned char dh_lo[2 * SLAB];
  float xe[64 * XP];
  float xo[64 * XP];
  float hpart[64 * KC];                          // partial H [64][65] (read by the peers), later da1 [64][72]
  float h_loc[11 * 64];
  float dh_loc[11 * 64];
  float part[800];                               // this CTA's share of the fc2 / b1 gradients + loss (read by the peers)
  float cpart[CL * 80];                          // rank 0: the six CTAs' conv-gradient shares (written by the peers)
  float w2[NCLS * HID];
  float b1[HID];
  float b2[16];
  float wc[80];
  float z[11 * 16];
  float dz[11 * 16];
  float red[16];
  int sidx[64];
  int label[64];
  float valid[64];
  unsigned char arg[64 * KC];
  alignas(8) uint64_t bar_w, bar_m1, bar_m2, bar_m3;
  uint32_t tmem_base;
};
static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "shared memory budget");
static_assert(64 * HP_STRIDE <= 64 * KC, "hpart fits its region");

NNDT_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
NNDT_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
NNDT_DEVINL void cluster_sync() { cluster_arrive(); cluster_wait(); }
NNDT_DEVINL uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
NNDT_DEVINL uint32_t map_to(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(umma::smem_u32(p)), "r"(rank));
  return r;
}
NNDT_DEVINL float ld_dsmem(uint32_t a) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
NNDT_DEVINL void st_dsmem4(uint32_t a, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
NNDT_DEVINL void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

NNDT_DEVINL float tf32_hi(float x) { uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x)); return __uint_as_float(h); }
NNDT_DEVINL void split(float x, float& hi, float& lo) { hi = tf32_hi(x); lo = tf32_hi(x - hi); }

// byte offset of element (row, col) in a [64 rows x 32 floats] 128B-swizzled slab
NNDT_DEVINL uint32_t swz(int row, int col) { return umma::swz_chunk_off(row, col >> 2) + (uint32_t)(col & 3) * 4u; }

// ---- descriptors (kind::tf32: one MMA consumes K = 8 floats = 32 B of a K-major row, or 8 rows of an MN-major slab) ----
NNDT_DEVINL uint64_t kdesc(uint32_t op_base, int kk) {            // K-major operand [rows][K], K step kk
  return umma::make_desc(op_base + (uint32_t)(kk >> 2) * SLAB + 32u * (uint32_t)(kk & 3), 16, 1024);
}
NNDT_DEVINL uint64_t mndesc(uint32_t op_base, int kk) {           // MN-major operand: MN along the row (atoms SLAB apart), K = rows
  return umma::make_desc(op_base + 1024u * (uint32_t)kk, SLAB, 1024);
}
NNDT_DEVINL constexpr uint32_t idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
NNDT_DEVINL void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
NNDT_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(umma::smem_u32(bar))
      : "memory");
}

// MN-major reading of 32-bit operands: the only layout the tensor core accepts is "128B swizzle with 32B atomicity"
// (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B = 1: byte-address bits [5,7) ^= bits [7,9), atoms of 128 B x 4 rows), not
// the 16B-chunk swizzle of the K-major layout.  The same tile is therefore re-swizzled IN PLACE between its K-major use
// (forward) and its MN-major use (backward): a per-row permutation of the eight 16-byte chunks.
NNDT_DEVINL uint64_t mn32desc(uint32_t op_base, int kk) {        // MN along the 128 B row (atoms SLAB apart), K = rows, 8 per MMA
  uint64_t d = 0;
  d |= (uint64_t)(((op_base + 1024u * (uint32_t)kk) & 0x3FFFF) >> 4);
  d |= (uint64_t)((SLAB >> 4) & 0x3FFF) << 16;                   // leading byte offset: next 128 B-wide atom along MN
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;                    // stride byte offset: next 4-row atom along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                                        // SWIZZLE_128B_BASE32B
  return d;
}
// physical 16 B chunk of logical chunk c in row `row`: K-major (standard 128B swizzle) / MN-major (32B-base swizzle)
NNDT_DEVINL int phys_k(int row, int c) { return c ^ (row & 7); }
NNDT_DEVINL int phys_mn(int row, int c) { return ((((c >> 1) ^ (row & 3)) << 1) | (c & 1)); }
// re-swizzle `nslab` slabs starting at `base` (hi and lo buffers are `lo_off` bytes apart) from the K-major to the
// MN-major chunk order.  One 16 B chunk per thread and iteration; the eight chunks of a row sit in eight consecutive lanes.
template <int NSLAB>
NNDT_DEVINL void reswizzle_k_to_mn(unsigned char* base, uint32_t lo_off, int tid) {
  constexpr int CH = 2 * NSLAB * 64 * 8, IT = (CH + NT - 1) / NT;
  float4 v[IT];
#pragma unroll
  for (int i = 0; i < IT; ++i) {
    const int o = tid + i * NT;
    if (o < CH) {
      const int buf = o / (NSLAB * 512), r = o - buf * (NSLAB * 512), slab = r >> 9, row = (r >> 3) & 63, p = r & 7;
      v[i] = *reinterpret_cast<const float4*>(base + buf * lo_off + slab * SLAB + row * 128 + 16 * p);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < IT; ++i) {
    const int o = tid + i * NT;
    if (o < CH) {
      const int buf = o / (NSLAB * 512), r = o - buf * (NSLAB * 512), slab = r >> 9, row = (r >> 3) & 63, p = r & 7;
      const int c = p ^ (row & 7);                               // logical chunk held at physical position p
      *reinterpret_cast<float4*>(base + buf * lo_off + slab * SLAB + row * 128 + 16 * phys_mn(row, c)) = v[i];
    }
  }
}

NNDT_DEVINL void stamp(long long* prof, int idx, int tid) {
  if (prof != nullptr && tid == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prof[idx] = t;
  }
}

// MS = samples per cluster (64 / batch splits per node): grid (CL, 64 / MS, L); the tensor-core tiles keep M = 64 rows
template <int MS>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT, 1)
mnist_tc_train_kernel(const Args a, const __grid_constant__ CUtensorMap w1_map) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment of the swizzled slabs (the dynamic window starts at the same offset in every CTA of the
  // cluster, so the peers' copies of every member sit at the same offset — what mapa needs)
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = blockIdx.z, bsplit = blockIdx.y, nsplit = gridDim.y;
  const int c = (int)cluster_rank();                       // == blockIdx.x: K slice of this CTA
  const float* th = a.theta + (size_t)l * a.n_pad;
  long long* prof = a.prof != nullptr ? a.prof + ((l * nsplit + bsplit) * CL + c) * 64 : nullptr;
  stamp(prof, 0, tid);
  auto own_lo = [](int r) { return (MS * r + CL - 1) / CL; };   // first sample whose fc2 part CTA r computes

  // ---- data half (before the programmatic-dependency wait: depends only on the dataset and the draw counter) ------
  const int call = a.calls != nullptr ? a.calls[l] : 0;
  const BatchGeom bg = batch_geom<true>(a, l, call);
  if (tid < MS) {
    int idx = 0, lab = 0; float ok = 0.f;
    const uint32_t t = (uint32_t)(bsplit * MS + tid);
    if (t < bg.bs) {
      ok = 1.f;
      idx = a.direct ? (int)(l * a.batch + t) : bg.shard_off + (int)feistel_permute(bg.start + t, bg.m, bg.key);
      lab = (int)a.y[idx];
    }
    sm.sidx[tid] = idx; sm.valid[tid] = ok; sm.label[tid] = lab;
  }
  if (tid == 64) {
    umma::mbar_init(&sm.bar_w, 1); umma::mbar_init(&sm.bar_m1, 1); umma::mbar_init(&sm.bar_m2, 1); umma::mbar_init(&sm.bar_m3, 1);
    umma::mbar_init_fence();
  }
  if (warp == 3) umma::tmem_alloc(&sm.tmem_base, TM_COLS);
  __syncthreads();
  // image rows 4c .. 4c+7 of every sample: 224 contiguous pixels
  constexpr int NU8 = (MS * 14 + NT - 1) / NT, NF4 = (MS * 56 + NT - 1) / NT;
  uint4 pu[NU8]; float4 pf[NF4];
  if (a.x_is_u8) {
#pragma unroll
    for (int i = 0; i < NU8; ++i) {
      const int o = tid + i * NT;
      pu[i] = make_uint4(0, 0, 0, 0);
      if (o < MS * 14) {
        const int s = o / 14, q = o - s * 14;
        if (sm.valid[s] != 0.f)
          pu[i] = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.x) + (size_t)sm.sidx[s] * 784 + 112 * c + 16 * q);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NF4; ++i) {
      const int o = tid + i * NT;
      pf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o < MS * 56) {
        const int s = o / 56, q = o - s * 56;
        if (sm.valid[s] != 0.f)
          pf[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + (size_t)sm.sidx[s] * 784 + 112 * c + 4 * q);
      }
    }
  }
  stamp(prof, 1, tid);

  pdl_wait();                 // the parameters of this step are final
  pdl_launch_dependents();
  stamp(prof, 2, tid);

  // ---- W1 slice by TMA: three boxes [64 rows x 32 cols], one per channel, columns ch*144 + 24c .. +31 ---------------
  if (tid == 0) {
    mbarrier_expect_tx(&sm.bar_w, 3 * SLAB);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) tma_load_3d(sm.w_hi + ch * SLAB, &w1_map, ch * NPOOL + CELLS * c, 0, l, &sm.bar_w);
  }
  // small tensors
  for (int o = tid; o < NCLS * HID; o += NT) sm.w2[o] = __ldcg(th + a.off_w2 + o);
  if (tid < 75) sm.wc[tid] = __ldcg(th + a.off_wc + tid);
  else if (tid < 78) sm.wc[tid] = __ldcg(th + a.off_bc + (tid - 75));
  else if (tid >= 96 && tid < 96 + HID) sm.b1[tid - 96] = __ldcg(th + a.off_b1 + (tid - 96));
  else if (tid >= 160 && tid < 160 + NCLS) sm.b2[tid - 160] = __ldcg(th + a.off_b2 + (tid - 160));
  // zero columns 24..31 of every A atom (K padding) for all 64 tile rows
  {
    const int buf = tid / 384, r = tid % 384, row = r / 6, k = r % 6, ch = k >> 1, chunk = 6 + (k & 1);
    unsigned char* base = (buf ? sm.a_lo : sm.a_hi) + ch * SLAB;
    *reinterpret_cast<float4*>(base + umma::swz_chunk_off(row, chunk)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // pixels -> normalised fp32 in even/odd column planes.  A sample's 8 x 28 slab is contiguous and 28 is even, so pixel p
  // lands at plane index p / 2: a run of 16 pixels is two aligned runs of 8 floats (vector stores)
  if (a.x_is_u8) {
#pragma unroll
    for (int i = 0; i < NU8; ++i) {
      const int o = tid + i * NT;
      if (o < MS * 14) {
        const int s = o / 14, q = o - s * 14;
        const bool ok = sm.valid[s] != 0.f;
        const uint32_t w[4] = {pu[i].x, pu[i].y, pu[i].z, pu[i].w};
        float ev[8], od[8];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = ok ? (((w[j >> 2] >> (8 * (j & 3))) & 0xff) * (1.f / 255.f) - a.mean) * a.inv_std : 0.f;
          if (j & 1) od[j >> 1] = v; else ev[j >> 1] = v;
        }
        float4* de = reinterpret_cast<float4*>(sm.xe + s * XP + 8 * q);
        float4* dd = reinterpret_cast<float4*>(sm.xo + s * XP + 8 * q);
        de[0] = make_float4(ev[0], ev[1], ev[2], ev[3]); de[1] = make_float4(ev[4], ev[5], ev[6], ev[7]);
        dd[0] = make_float4(od[0], od[1], od[2], od[3]); dd[1] = make_float4(od[4], od[5], od[6], od[7]);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < NF4; ++i) {
      const int o = tid + i * NT;
      if (o < MS * 56) {
        const int s = o / 56, q = o - s * 56;
        *reinterpret_cast<float2*>(sm.xe + s * XP + 2 * q) = make_float2(pf[i].x, pf[i].z);
        *reinterpret_cast<float2*>(sm.xo + s * XP + 2 * q) = make_float2(pf[i].y, pf[i].w);
      }
    }
  }
  __syncthreads();
  stamp(prof, 3, tid);

  // ---- conv + ReLU + maxpool: one (sample, pooled cell) per item, all three channels from one 6x6 patch ------------
  for (int it = tid; it < MS * CELLS; it += NT) {
    const int s = it / CELLS, cell = it - s * CELLS;
    const int pr = cell / PHW, px = cell - pr * PHW;
    float patch[6][6];
    const float* e = sm.xe + s * XP + (2 * pr) * 14 + px;
    const float* o = sm.xo + s * XP + (2 * pr) * 14 + px;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { patch[r][2 * q] = e[r * 14 + q]; patch[r][2 * q + 1] = o[r * 14 + q]; }
    }
    const bool ok = sm.valid[s] != 0.f;
#pragma unroll
    for (int ch = 0; ch < F; ++ch) {
      float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float w = sm.wc[ch * 25 + ky * 5 + kx];
          acc[0][0] = fmaf(w, patch[ky][kx], acc[0][0]);
          acc[0][1] = fmaf(w, patch[ky][kx + 1], acc[0][1]);
          acc[1][0] = fmaf(w, patch[ky + 1][kx], acc[1][0]);
          acc[1][1] = fmaf(w, patch[ky + 1][kx + 1], acc[1][1]);
        }
      }
      float m = acc[0][0]; int ai = 0;                 // first maximum wins, like ATen's max_pool2d
      if (acc[0][1] > m) { m = acc[0][1]; ai = 1; }
      if (acc[1][0] > m) { m = acc[1][0]; ai = 2; }
      if (acc[1][1] > m) { m = acc[1][1]; ai = 3; }
      m = ok ? fmaxf(m + sm.wc[75 + ch], 0.f) : 0.f;
      float hi, lo;
      split(m, hi, lo);
      const uint32_t off = swz(s, cell);
      *reinterpret_cast<float*>(sm.a_hi + ch * SLAB + off) = hi;
      *reinterpret_cast<float*>(sm.a_lo + ch * SLAB + off) = lo;
      sm.arg[s * KC + ch * CELLS + cell] = (unsigned char)(ai | (m > 0.f ? 4 : 0));
    }
  }
  stamp(prof, 4, tid);
  // ---- W1 slice: raw fp32 -> hi (in place) + lo -----------------------------------------------------------------------
  mbarrier_wait_parity(&sm.bar_w, 0);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int o = (tid + i * NT) * 16;                   // 1536 float4 in the three slabs
    float4 v = *reinterpret_cast<const float4*>(sm.w_hi + o), h4, l4;
    split(v.x, h4.x, l4.x); split(v.y, h4.y, l4.y); split(v.z, h4.z, l4.z); split(v.w, h4.w, l4.w);
    *reinterpret_cast<float4*>(sm.w_hi + o) = h4;
    *reinterpret_cast<float4*>(sm.w_lo + o) = l4;
  }
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t A_HI = umma::smem_u32(sm.a_hi), A_LO = umma::smem_u32(sm.a_lo), W_HI = umma::smem_u32(sm.w_hi),
                 W_LO = umma::smem_u32(sm.w_lo), DH_HI = umma::smem_u32(sm.dh_hi), DH_LO = umma::smem_u32(sm.dh_lo);

  // ---- MMA 1: H_c[64 s x 64 j] = A_c . W1_c^T (both K-major, K = 96), 3xTF32 ------------------------------------------
  if (tid == 0) {
    constexpr uint32_t id = idesc_tf32(64, 64, false, false);
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t A = pass == 0 ? A_LO : A_HI, W = pass == 1 ? W_LO : W_HI;
#pragma unroll
      for (int kk = 0; kk < 12; ++kk) { mma_tf32(tmem + TM_D1, kdesc(A, kk), kdesc(W, kk), id, acc); acc = true; }
    }
    umma::commit(&sm.bar_m1);
  }
  umma::mbar_wait(&sm.bar_m1, 0);
  umma::fence_after_sync();
  stamp(prof, 5, tid);
  if (warp < 8) {                                        // TMEM lane of row s: 32 (s / 16) + s % 16
    const int q = warp & 3, half = warp >> 2;
    float v[32];
    umma::tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + TM_D1 + 32 * half, v);
    if (lane < 16 && 16 * q + lane < MS) {
      float* dst = sm.hpart + (16 * q + lane) * HP_STRIDE + 32 * half;
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[i] = v[i];
    }
  }
  umma::fence_before_sync();
  cluster_sync();                                        // #1: all six partial H are in shared memory
  umma::fence_after_sync();
  stamp(prof, 6, tid);
  if (c == 0 && tid == 0 && a.calls != nullptr) {
    // every CTA of this cluster has read the draw counter; the last cluster of the node to get here advances it
    if (a.arrive == nullptr || nsplit == 1) a.calls[l] = call + 1;
    else if (atomicAdd(a.arrive + l, 1u) == (unsigned)nsplit - 1) { a.arrive[l] = 0; a.calls[l] = call + 1; }
  }

  // ---- reduce-scatter of H over the cluster + fc2 / loss / their backward for this CTA's samples -----------------------
  const int s0 = own_lo(c), ns = own_lo(c + 1) - s0;
  if (tid < ns * HID) {
    const int sl = tid >> 6, j = tid & 63;
    const float* src = sm.hpart + (s0 + sl) * HP_STRIDE + j;
    float v = sm.b1[j];
#pragma unroll
    for (int r = 0; r < CL; ++r) v += ld_dsmem(map_to(src, (uint32_t)r));
    sm.h_loc[tid] = fmaxf(v, 0.f);
  }
  __syncthreads();
  {
    const int o = tid >> 2, part = tid & 3;              // 4 lanes per logit
    const bool live = o < ns * NCLS;
    const int sl = live ? o / NCLS : 0, cc = live ? o - sl * NCLS : 0;
    float v = 0.f;
    if (live) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) { const int j = part * 16 + jj; v = fmaf(sm.h_loc[sl * HID + j], sm.w2[cc * HID + j], v); }
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (live && part == 0) sm.z[sl * 16 + cc] = v + sm.b2[cc];
  }
  __syncthreads();
  if (tid < ns) {
    const int sl = tid, s = s0 + sl;
    float mx = sm.z[sl * 16];
#pragma unroll
    for (int cc = 1; cc < NCLS; ++cc) mx = fmaxf(mx, sm.z[sl * 16 + cc]);
    float se = 0.f;
#pragma unroll
    for (int cc = 0; cc < NCLS; ++cc) se += __expf(sm.z[sl * 16 + cc] - mx);
    const float lse = mx + __logf(se);
    const int y = sm.label[s];
    const float ok = sm.valid[s];
#pragma unroll
    for (int cc = 0; cc < NCLS; ++cc)
      sm.dz[sl * 16 + cc] = ok * bg.inv_bs * (__expf(sm.z[sl * 16 + cc] - lse) - (cc == y ? 1.f : 0.f));
    sm.red[sl] = ok * (lse - sm.z[sl * 16 + y]);
  }
  __syncthreads();
  if (tid < ns * HID) {
    const int sl = tid >> 6, j = tid & 63;
    float v = 0.f;
#pragma unroll
    for (int cc = 0; cc < NCLS; ++cc) v = fmaf(sm.dz[sl * 16 + cc], sm.w2[cc * HID + j], v);
    sm.dh_loc[tid] = sm.h_loc[tid] > 0.f ? v : 0.f;
  }
  __syncthreads();
  // this CTA's share of the fc2 / b1 gradients and of the loss
  for (int o = tid; o < NCLS * HID; o += NT) {
    const int cc = o >> 6, j = o & 63;
    float v = 0.f;
    for (int sl = 0; sl < ns; ++sl) v = fmaf(sm.dz[sl * 16 + cc], sm.h_loc[sl * HID + j], v);
    sm.part[PART_W2 + o] = v;
  }
  if (tid < HID) {
    float v = 0.f;
    for (int sl = 0; sl < ns; ++sl) v += sm.dh_loc[sl * HID + tid];
    sm.part[PART_B1 + tid] = v;
  } else if (tid >= 64 && tid < 64 + NCLS) {
    float v = 0.f;
    for (int sl = 0; sl < ns; ++sl) v += sm.dz[sl * 16 + (tid - 64)];
    sm.part[PART_B2 + (tid - 64)] = v;
  } else if (tid == 96) {
    float v = 0.f;
    for (int sl = 0; sl < ns; ++sl) v += sm.red[sl];
    sm.part[PART_LOSS] = v * bg.inv_bs;
  }
  stamp(prof, 7, tid);
  cluster_sync();                                        // #2: every owner's dH rows are final; hpart is free again
  stamp(prof, 8, tid);
  float* gp = a.grad_part + ((size_t)l * nsplit + bsplit) * a.n_pad;
  // the fc2 / b1 / loss shares of all six CTAs are final: CTA c reduces its sixth of them (DSMEM reads) and writes the
  // result; the peers stay resident until the last cluster barrier, so nothing else has to wait for this
  {
    constexpr int NE = PART_N - PART_B1, PER = (NE + CL - 1) / CL;          // 715 values
    const int o = PART_B1 + c * PER + tid;
    if (tid < PER && o < PART_N) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < CL; ++r) v += ld_dsmem(map_to(sm.part + o, (uint32_t)r));
      if (o < PART_W2) gp[a.off_b1 + (o - PART_B1)] = v;
      else if (o < PART_B2) gp[a.off_w2 + (o - PART_W2)] = v;
      else if (o < PART_LOSS) gp[a.off_b2 + (o - PART_B2)] = v;
      else {
        a.loss_part[l * nsplit + bsplit] = v;
        if (a.loss_mirror != nullptr) a.loss_mirror[l * nsplit + bsplit] = v;     // zero-copy store to pinned host memory
      }
    }
  }

  // ---- gather all MS dH rows from their owners (DSMEM reads) into the local K-major operand tile (hi / lo), and
  //      re-swizzle the W1 slice for its MN-major use (W is dead as a K-major operand: MMA 1 has completed) -------------
  for (int o = tid; o < MS * 16; o += NT) {
    const int s = o >> 4, ch4 = o & 15;
    int r = 0;
#pragma unroll
    for (int q = 1; q < CL; ++q) r += (s >= own_lo(q)) ? 1 : 0;
    const uint32_t src = map_to(sm.dh_loc + (s - own_lo(r)) * HID + 4 * ch4, (uint32_t)r);
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src) : "memory");
    float4 h4, l4;
    split(v.x, h4.x, l4.x); split(v.y, h4.y, l4.y); split(v.z, h4.z, l4.z); split(v.w, h4.w, l4.w);
    const uint32_t off = (uint32_t)(ch4 >> 3) * SLAB + umma::swz_chunk_off(s, ch4 & 7);
    *reinterpret_cast<float4*>(sm.dh_hi + off) = h4;
    *reinterpret_cast<float4*>(sm.dh_lo + off) = l4;
  }
  reswizzle_k_to_mn<3>(sm.w_hi, (uint32_t)(sm.w_lo - sm.w_hi), tid);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  stamp(prof, 9, tid);

  // ---- MMA 2: da1_c[64 s x 96] = dH . W1_c (A K-major, B MN-major) ------------------------------------------------------
  if (tid == 0) {
    constexpr uint32_t id2 = idesc_tf32(64, 96, false, true);
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t D = pass == 0 ? DH_LO : DH_HI, W = pass == 1 ? W_LO : W_HI;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) { mma_tf32(tmem + TM_D2, kdesc(D, kk), mn32desc(W, kk), id2, acc); acc = true; }
    }
    umma::commit(&sm.bar_m2);
  }
  // under MMA 2: the A tile (dead as a K-major operand) moves to the MN-major chunk order for MMA 3
  reswizzle_k_to_mn<3>(sm.a_hi, (uint32_t)(sm.a_lo - sm.a_hi), tid);
  umma::mbar_wait(&sm.bar_m2, 0);
  umma::fence_after_sync();
  // MMA 2 is done with the K-major dH: same tile, MN-major order, for dW1 = dH^T . A
  reswizzle_k_to_mn<2>(sm.dh_hi, (uint32_t)(sm.dh_lo - sm.dh_hi), tid);
  umma::fence_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  stamp(prof, 10, tid);

  // ---- MMA 3: dW1_c[64 j x 96] = dH^T . A_c (both MN-major; K = the MS sample rows) ---------------------------------------
  if (tid == 0) {
    constexpr uint32_t id3 = idesc_tf32(64, 96, true, true);
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t D = pass == 0 ? DH_LO : DH_HI, A = pass == 1 ? A_LO : A_HI;
#pragma unroll
      for (int kk = 0; kk < MS / 8; ++kk) { mma_tf32(tmem + TM_D3, mn32desc(D, kk), mn32desc(A, kk), id3, acc); acc = true; }
    }
    umma::commit(&sm.bar_m3);
  }
  // 24 warps: TMEM quarter q, 16 accumulator columns each: channel atom ch, half of its 32 columns
  const int eq = warp & 3, esub = warp >> 2, ech = esub >> 1, ehalf = esub & 1;
  const int erow = 16 * eq + lane;                       // valid for lane < 16
  const int nvalid = ehalf ? 8 : 16;                     // cells 16..23 of the second half, 24..31 are padding
  {
    // under MMA 3: da1 (D2) -> shared memory, masked by ReLU'(a1)
    float v2[16];
    umma::tmem_ld16(tmem + ((uint32_t)(32 * eq) << 16) + TM_D2 + 32 * ech + 16 * ehalf, v2);
    if (lane < 16 && erow < MS) {
      float* d1 = sm.hpart + erow * KC + ech * CELLS + 16 * ehalf;
      const unsigned char* ag = sm.arg + erow * KC + ech * CELLS + 16 * ehalf;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        if (i < nvalid) {
          float4 d;
          d.x = (ag[i] & 4) ? v2[i] : 0.f; d.y = (ag[i + 1] & 4) ? v2[i + 1] : 0.f;
          d.z = (ag[i + 2] & 4) ? v2[i + 2] : 0.f; d.w = (ag[i + 3] & 4) ? v2[i + 3] : 0.f;
          *reinterpret_cast<float4*>(d1 + i) = d;
        }
      }
    }
  }
  __syncthreads();        // da1 complete
  stamp(prof, 11, tid);
  // ---- conv grads (under MMA 3): each pooled cell routes da1 to its argmax conv position; 3 groups of 256 threads, one
  //      per channel; partial sums are folded over 4 neighbouring lanes and transposed through the dead W slabs -----------
  float cacc[26];
#pragma unroll
  for (int i = 0; i < 26; ++i) cacc[i] = 0.f;
  const int cg = tid / CGROUP, ct = tid - cg * CGROUP;
  if (cg < F) {
    for (int it = ct; it < MS * CELLS; it += CGROUP) {
      const int s = it / CELLS, cell = it - s * CELLS;
      const float g = sm.hpart[s * KC + cg * CELLS + cell];
      if (g != 0.f) {
        const int ai = sm.arg[s * KC + cg * CELLS + cell] & 3;
        const int pr = cell / PHW, px = cell - pr * PHW;
        const int par = ai & 1, base = s * XP + (2 * pr + (ai >> 1)) * 14 + px;
        const float* pA = (par ? sm.xo : sm.xe) + base;        // taps kx = 0, 2, 4
        const float* pB = (par ? sm.xe + 1 : sm.xo) + base;    // taps kx = 1, 3
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
          for (int kx = 0; kx < KS; ++kx)
            cacc[ky * 5 + kx] = fmaf(g, ((kx & 1) ? pB : pA)[ky * 14 + (kx >> 1)], cacc[ky * 5 + kx]);
        cacc[25] += g;
      }
    }
  }
  float* scratch = reinterpret_cast<float*>(sm.w_hi);   // [78][64]: 20 KB of the W slabs (dead since MMA 2 completed)
#pragma unroll
  for (int i = 0; i < 26; ++i) {
    float v = cacc[i];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (cg < F && (lane & 3) == 0) scratch[(cg * 26 + i) * 64 + (ct >> 2)] = v;
  }
  __syncthreads();
  for (int o = warp; o < 78; o += NT / 32) {
    float v = scratch[o * 64 + lane] + scratch[o * 64 + lane + 32];
    v = warp_sum(v);
    if (lane == 0) {
      const int ch = o / 26, i = o - ch * 26;
      // this CTA's share of the conv gradients goes straight into rank 0's collection buffer
      const int dst = (i < 25) ? ch * 25 + i : 75 + ch;
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(map_to(sm.cpart + c * 80 + dst, 0u)), "f"(v) : "memory");
    }
  }
  stamp(prof, 12, tid);
  umma::mbar_wait(&sm.bar_m3, 0);
  umma::fence_after_sync();
  {
    float v3[16];
    umma::tmem_ld16(tmem + ((uint32_t)(32 * eq) << 16) + TM_D3 + 32 * ech + 16 * ehalf, v3);
    if (lane < 16) {
      float* gw = gp + a.off_w1 + erow * FC1_IN + ech * NPOOL + CELLS * c + 16 * ehalf;
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        if (i < nvalid) *reinterpret_cast<float4*>(gw + i) = make_float4(v3[i], v3[i + 1], v3[i + 2], v3[i + 3]);
    }
  }
  umma::fence_before_sync();
  stamp(prof, 13, tid);
  cluster_sync();                                        // #3: all six conv-gradient shares are in rank 0's buffer
  if (c == 0 && tid < 78) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) v += sm.cpart[r * 80 + tid];
    gp[tid < 75 ? a.off_wc + tid : a.off_bc + (tid - 75)] = v;
  }
  if (warp == 3) umma::tmem_dealloc(tmem, TM_COLS);
  stamp(prof, 14, tid);
}

}  // namespace tc

// ---- host side ------------------------------------------------------------------------------------------------------
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t make_w1_tensor_map(const float* theta, int n_pad, int L, int off_w1, void* out_map128) {
  static EncodeTiled encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  // W1 of node l: [64 rows x 432 cols] fp32 at theta + l * n_pad + off_w1  ->  3-D tensor (col, row, node)
  const cuuint64_t dims[3] = {(cuuint64_t)FC1_IN, (cuuint64_t)HID, (cuuint64_t)L};
  const cuuint64_t strides[2] = {(cuuint64_t)FC1_IN * 4, (cuuint64_t)n_pad * 4};
  const cuuint32_t box[3] = {32, 64, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(reinterpret_cast<CUtensorMap*>(out_map128), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                      const_cast<float*>(theta + off_w1), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int MS>
static cudaError_t launch_tc_ms(const Args& a, const void* w1_map128, cudaStream_t st) {
  static cudaError_t prep = cudaFuncSetAttribute(tc::mnist_tc_train_kernel<MS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)sizeof(tc::Smem) + 1024);
  if (prep != cudaSuccess) return prep;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tc::CL, 64 / MS, a.L); cfg.blockDim = dim3(tc::NT);
  cfg.dynamicSmemBytes = sizeof(tc::Smem) + 1024; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("NNDT_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  CUtensorMap map;
  memcpy(&map, w1_map128, sizeof(map));
  return cudaLaunchKernelEx(&cfg, tc::mnist_tc_train_kernel<MS>, a, map);
}

// `nsplit` batch splits per node (1, 2 or 4): 6 * nsplit CTAs per node, `nsplit` gradient partial rows
cudaError_t launch_train_tc(const Args& a, const void* w1_map128, int nsplit, cudaStream_t st) {
  switch (nsplit) {
    case 1: return launch_tc_ms<64>(a, w1_map128, st);
    case 2: return launch_tc_ms<32>(a, w1_map128, st);
    case 4: return launch_tc_ms<16>(a, w1_map128, st);
  }
  return cudaErrorInvalidValue;
}

int tc_max_active_clusters() {
  if (cudaFuncSetAttribute(tc::mnist_tc_train_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(tc::Smem) + 1024) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tc::CL, 1, 1); cfg.blockDim = dim3(tc::NT); cfg.dynamicSmemBytes = sizeof(tc::Smem) + 1024;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, tc::mnist_tc_train_kernel<64>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

}  // namespace mnist
}  // namespace nndt
