// mbarrier / TMA helpers shared by the tcgen05 GEMM and the skinny streaming GEMM.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ftcf {

// 2-D row-major [rows, cols] tensor of `elem`-byte elements; box = [box_rows, 128 bytes]; SWIZZLE_128B; out-of-bounds
// elements read as zero.  Host side, defined in gemm_tcgen05.cu (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint).
int make_tensor_map_2d(CUtensorMap* map, const void* base, int rows, int cols, int elem, int box_rows);

#ifdef __CUDACC__
namespace tma {
constexpr long long kSpinLimit = 1ll << 22;   // bounded waits: a pipeline bug traps instead of hanging the GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// One lane of a fully converged warp (elect.sync).  Single-thread instructions that take UNIFORM operands (tcgen05.mma,
// tcgen05.commit, cp.async.bulk.tensor) must be issued under this predicate from warp-uniform control flow: behind a plain
// `if (lane == 0)` the compiler cannot prove the operands uniform and wraps every such instruction in a per-lane "waterfall" loop
// (R2UR + ELECT + BRA.U.ANY) -- measured 122 cycles per tcgen05.mma issue instead of ~10 (tools/umma_probe.cu, profiles/).
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// 128-bit shared-memory load through the shared window (a generic `*ptr` compiles to LD.E.128 when the compiler cannot prove the
// address space -- longer latency than LDS.128)
__device__ __forceinline__ uint4 lds_128(uint32_t smem_addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
    return v;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t addr, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    if (mbar_try_wait(addr, parity)) return;          // the common case in a running pipeline: no loop bookkeeping at all
    for (int spin = 0; !mbar_try_wait(addr, parity); ++spin)
        if (spin > (1 << 22)) __trap();               // bounded: a pipeline bug traps instead of hanging the GPU
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }
// one box of the tensor map -> shared memory; c0 = element column, c1 = row
__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// same, with an L2 cache policy (createpolicy): weights are read once per token -- evict_first keeps them from flushing the
// activations and the KV rows that were prefetched for the attention kernel
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void load_2d_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
        : "memory");
}
// L2 prefetch of one box of the tensor map (no shared-memory destination, no barrier)
__device__ __forceinline__ void prefetch_2d(const CUtensorMap* map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
}  // namespace tma
#endif

}  // namespace ftcf
