// Minimal hand-written tcgen05 / TMEM / mbarrier layer for sm_100a (inline PTX only).
//
// Conventions used by every kernel in this repo:
//  * operands are bf16 tiles of [rows][64 elements] = rows x 128 B "slabs", 1024 B aligned,
//    stored with the 128-byte swizzle (16 B chunk index XOR (row & 7)).  The same slab is a
//    valid K-major operand (K = the 64 contiguous elements, rows = M/N) and a valid MN-major
//    operand (MN = the 64 contiguous elements, rows = K) — only the descriptor differs, which
//    is what lets the backward GEMMs (dW = dZ^T . H, dH = dZ . W) reuse the forward tiles with
//    no transposed copies.
//  * accumulators are fp32 in TMEM, M = 128 (TMEM lane == tile row), cta_group::1.
//  * one elected thread issues tcgen05.mma and commits to an mbarrier; everybody waits on it.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace nndt {
namespace umma {

constexpr int kSlabRowBytes = 128;   // 64 bf16
constexpr int kSlabAtomBytes = 1024; // 8 rows

NNDT_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a swizzled slab
NNDT_DEVINL uint32_t swz_chunk_off(int row, int chunk) {
  return (uint32_t)row * kSlabRowBytes + (uint32_t)((chunk ^ (row & 7)) << 4);
}

// ---- shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout) ----------------
//  [0,14)  start address >> 4      [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//  [46,48) version = 1 (sm_100)    [61,64) layout type (2 = SWIZZLE_128B)
NNDT_DEVINL uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major operand: rows = M (or N), K along the 128-byte row.  One MMA consumes K = 16 bf16 = 32 B;
// step k16 inside a slab advances the start address by 32 B, the next slab holds the next 64 K.
NNDT_DEVINL uint64_t desc_kmajor(uint32_t slab_addr, int k16_in_slab) {
  return make_desc(slab_addr + 32u * (uint32_t)k16_in_slab, 16, kSlabAtomBytes);
}
// MN-major operand: MN along the 128-byte row (64 per slab, further slabs `lbo_bytes` apart),
// K = rows: one MMA consumes 16 rows = 2 swizzle atoms = 2048 B.
NNDT_DEVINL uint64_t desc_mnmajor(uint32_t slab_addr, int k16, uint32_t lbo_bytes) {
  return make_desc(slab_addr + 2048u * (uint32_t)k16, lbo_bytes, kSlabAtomBytes);
}

// ---- instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16, bf16 x bf16 -> fp32 ------
NNDT_DEVINL constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                       // D format  = F32
         | (1u << 7) | (1u << 10)        // A, B format = BF16
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

NNDT_DEVINL void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete
NNDT_DEVINL void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
NNDT_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
NNDT_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t cols) {    // same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
NNDT_DEVINL void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
NNDT_DEVINL void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand fetch)
NNDT_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane_base + t)
NNDT_DEVINL void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive fp32 columns
NNDT_DEVINL void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier --------------------------------------------------------------------------------------
NNDT_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
NNDT_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
NNDT_DEVINL void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// pack 8 fp32 -> 8 bf16 (16 bytes)
NNDT_DEVINL uint4 pack_bf16x8(const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]);
  __nv_bfloat162 d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c);
  o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}

}  // namespace umma
}  // namespace nndt
