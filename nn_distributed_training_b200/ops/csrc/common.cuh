// Shared device helpers for the sm_100a kernels of nn_distributed_training_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include <utility>

#define NNDT_DEVINL __device__ __forceinline__

namespace nndt {

constexpr int kWarp = 32;

NNDT_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
NNDT_DEVINL double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- system-scope flags for cross-GPU producer/consumer sync -----------------
// A rank publishes round k by (1) writing its rows, (2) __threadfence_system(),
// (3) st.release.sys of k into the *reader's* flag slot (remote store over NVLink),
// so readers spin on local memory only.
NNDT_DEVINL void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
NNDT_DEVINL int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
NNDT_DEVINL int ld_relaxed_sys(const int* p) {
  int v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// streaming 128-bit load that bypasses L1 allocation (peer / read-once data)
NNDT_DEVINL float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// Kernels of a round are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may
// start while its predecessor drains.  Convention in this repo: every kernel executes pdl_wait() before it
// touches anything the *immediately preceding* kernel wrote, and only then pdl_launch_dependents() — so when
// a kernel's pre-wait prologue runs, everything two or more kernels back is complete and visible.
NNDT_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
NNDT_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  static const bool no_pdl = getenv("NNDT_NO_PDL") != nullptr;   // debugging / A-B switch
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- TMA bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier transaction tracking ---------------------
NNDT_DEVINL uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
NNDT_DEVINL void mbarrier_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
NNDT_DEVINL void mbarrier_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
// one thread: copy `bytes` (multiple of 16, 16 B aligned both sides) global -> shared, completing on `bar`
NNDT_DEVINL void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_addr_u32(bar))
               : "memory");
}
NNDT_DEVINL void mbarrier_wait_parity(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr_u32(bar);
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

// cp.async helpers (LDGSTS)
NNDT_DEVINL void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
NNDT_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
NNDT_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace nndt
