// tile_common.cuh -- pieces shared by the per-nonzero kernels of strategy RECOMPUTE (recompute_tiles.cu: general tiles,
// recompute_uniform.cu: class-uniform tiles): form codes, per-form sizes, mbarrier / bulk-copy (TMA) wrappers.
#pragma once
#include "common.cuh"

namespace gf {

enum { TF_LAPLACE = 0, TF_ELAST = 1, TF_MASS = 2 };

static inline int tf_of(int family) {
  return family == GFGPU_LAPLACE ? TF_LAPLACE : family == GFGPU_ELASTICITY ? TF_ELAST : family == GFGPU_MASS ? TF_MASS : -1;
}

__host__ __device__ constexpr int tl_clog2(int v) {
  int b = 0;
  while ((1 << b) < v) ++b;
  return b;
}

template <int N, int RF>
struct TlCfg {
  static constexpr int GSZ = RF == TF_ELAST ? N * N : RF == TF_LAPLACE ? N * (N + 1) / 2 : 1;  // doubles per element
  static constexpr int MT = GSZ;                                                               // table entries per (j,i)
  static constexpr int ACC = RF == TF_ELAST ? N * N : 1;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


}  // namespace gf
