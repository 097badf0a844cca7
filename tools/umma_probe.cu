// Micro-probe (one CTA): cycles per tcgen05.mma for TS / SS form, N = 16 / 32 / 128, same accumulator vs alternating accumulators,
// and the latency of tcgen05.st + wait::st.  Operand contents are irrelevant (zeros).  Build: nvcc -arch=sm_100a -O3 -I../fastertransformer4codefuse_b200/csrc
#include <cstdio>
#include "umma.cuh"
using namespace ftcf;
using namespace ftcf::umma;

__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma, int N, int ts, int alt, int uniform)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    // zero the A region of TMEM (cols 256..511) so that values stay finite
    {
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = 0;
        const int q = threadIdx.x >> 5;
        for (int c = 256; c < 512; c += 32) tmem_st_x32(tmem + ((uint32_t)(q * 32) << 16) + c, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) {
        // warp-uniform control flow, single-thread instructions under elect.sync
        const uint32_t idesc = umma_idesc_f16(N);
        const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 16384);
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            if (uniform) {
                if (tma::elect_one_sync()) {
#pragma unroll 8
                    for (int i = 0; i < n_mma; ++i) {
                        const uint32_t d = tmem + (alt ? (i & 1) * 128 : 0);
                        if (ts) mma_ts(d, tmem + 256 + (i & 7) * 8, umma_desc_k128(b_s + (i & 3) * 32), idesc, 1);
                        else mma_ss(d, umma_desc_k128(a_s + (i & 3) * 32), umma_desc_k128(b_s + (i & 3) * 32), idesc, 1);
                    }
                }
                __syncwarp();
            } else if (threadIdx.x == 0) {
                for (int i = 0; i < n_mma; ++i) {
                    const uint32_t d = tmem + (alt ? (i & 1) * 128 : 0);
                    if (ts) mma_ts(d, tmem + 256 + (i & 7) * 8, umma_desc_k128(b_s + (i & 3) * 32), idesc, 1);
                    else mma_ss(d, umma_desc_k128(a_s + (i & 3) * 32), umma_desc_k128(b_s + (i & 3) * 32), idesc, 1);
                }
            }
            __syncwarp();
            const long long t1 = clock64();
            if (tma::elect_one_sync()) tc_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, rep & 1);
            const long long t2 = clock64();
            if (threadIdx.x == 0) {
                out[rep * 2] = t1 - t0;
                out[rep * 2 + 1] = t2 - t0;
            }
        }
    }
    __syncthreads();
    // tcgen05.st latency: x32 store + wait, per warp
    if (threadIdx.x % 32 == 0 || true) {
        uint32_t r[32];
        for (int i = 0; i < 32; ++i) r[i] = threadIdx.x;
        const int q = threadIdx.x >> 5;
        __syncthreads();
        const long long t0 = clock64();
        for (int it = 0; it < 8; ++it) {
            tmem_st_x32(tmem + ((uint32_t)(q * 32) << 16) + 256 + (it & 3) * 32, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[6] = (t1 - t0) / 8;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main()
{
    long long* d;
    cudaMalloc(&d, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int n_mma = 64;
    for (int uniform = 0; uniform <= 1; ++uniform)
        for (int ts = 0; ts <= 1; ++ts)
            for (int N : {16, 128})
                for (int alt = 0; alt <= 1; ++alt) {
                    probe<<<1, 128, 64 * 1024>>>(d, n_mma, N, ts, alt, uniform);
                    long long h[8];
                    cudaError_t e = cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("%s %s N=%3d %s: issue %.1f cyc/mma, issue+complete %.1f cyc/mma (%d mma) | st.x32+wait %lld cyc\n",
                           uniform ? "elect.sync  " : "if(lane==0) ", ts ? "TS" : "SS", N, alt ? "alternating D" : "same D       ", h[4] / (double)n_mma,
                           h[5] / (double)n_mma, n_mma, h[6]);
                }
    return 0;
}
