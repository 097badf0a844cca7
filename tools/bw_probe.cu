// Read-bandwidth probe for the B200 box: what a READ-ONLY stream can reach (the decode step is one), with plain vector loads
// and with cp.async.bulk rings of different depths.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bw_probe.cu -o tools/bw_probe.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int UNROLL>
__global__ void __launch_bounds__(512) ldg_read(const uint4* __restrict__ in, size_t n16, unsigned* out)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(in + i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// each CTA streams its contiguous share of the buffer through an NS-deep ring of SB-byte stages; COPIES bulk copies per stage
__global__ void __launch_bounds__(64) bulk_read(const uint8_t* __restrict__ in, size_t bytes_per_cta, int ns, int sb, int copies, unsigned* out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[32], empty[32];
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* src = in + (size_t)blockIdx.x * bytes_per_cta;
    const int nst = (int)(bytes_per_cta / sb);
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) {
            const int s = i % ns;
            mbar_wait(&empty[s], ((i / ns) & 1) ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(sb) : "memory");
            const int cb = sb / copies;
            for (int c = 0; c < copies; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)s * sb + c * cb)),
                             "l"(src + (size_t)i * sb + c * cb), "r"(cb), "r"(smem_u32(&full[s]))
                             : "memory");
        }
    } else if (threadIdx.x == 32) {
        unsigned acc = 0;
        for (int i = 0; i < nst; ++i) {
            const int s = i % ns;
            mbar_wait(&full[s], (i / ns) & 1);
            acc ^= *reinterpret_cast<const unsigned*>(smem + (size_t)s * sb);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

// The streaming GEMM's producer pattern: per stage 4 tensor boxes of 32 rows x 128 bytes (SWIZZLE_128B) from a K-major
// matrix [n][k] -- rows k bytes apart -- or from a TILED copy where each (32-row, 512-byte) block is 16 KB contiguous.
// mode 0: 2-D map, coordinates (k byte, row).  mode 1: 4-D map (byte in chunk, row in tile, k chunk, row tile).
__global__ void __launch_bounds__(64) tma_box_read(const __grid_constant__ CUtensorMap map, int mode, int tiles_per_cta, int chunks, int ns, unsigned* out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[32], empty[32];
    if (threadIdx.x == 0) {
        for (int s = 0; s < ns; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nst = tiles_per_cta * chunks;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) {
            const int s = i % ns;
            const int tile = blockIdx.x * tiles_per_cta + i / chunks, kc = i % chunks;
            mbar_wait(&empty[s], ((i / ns) & 1) ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(16384) : "memory");
            for (int j = 0; j < 4; ++j) {
                const uint32_t dst = smem_u32(smem + (size_t)s * 16384 + j * 4096);
                if (mode == 0)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                                 "l"(&map), "r"(smem_u32(&full[s])), "r"(kc * 512 + j * 128), "r"(tile * 32) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                                 "l"(&map), "r"(smem_u32(&full[s])), "r"(j * 128), "r"(0), "r"(kc), "r"(tile) : "memory");
            }
        }
    } else if (threadIdx.x == 32) {
        unsigned acc = 0;
        for (int i = 0; i < nst; ++i) {
            const int s = i % ns;
            mbar_wait(&full[s], (i / ns) & 1);
            acc ^= *reinterpret_cast<const unsigned*>(smem + (size_t)s * 16384);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const size_t bytes = (size_t)12 << 30;
    uint8_t* buf;
    unsigned* out;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time_it = [&](auto&& launch, const char* name) {
        launch();
        CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("%-60s %8.3f ms  %7.1f GB/s\n", name, best, bytes / (best * 1e-3) / 1e9);
    };
    char name[128];
    for (int ctas : {296})
        for (int thr : {512}) {
            snprintf(name, sizeof name, "ldg.128 x4 in flight, %d CTAs x %d thr", ctas, thr);
            time_it([&] { ldg_read<4><<<ctas, thr>>>((const uint4*)buf, bytes / 16, out); }, name);
            snprintf(name, sizeof name, "ldg.128 x8 in flight, %d CTAs x %d thr", ctas, thr);
            time_it([&] { ldg_read<8><<<ctas, thr>>>((const uint4*)buf, bytes / 16, out); }, name);
        }
    CK(cudaFuncSetAttribute(bulk_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int ctas : {148, 296})
        for (int sb : {16384})
            for (int ns : {4, 8})
                for (int copies : {1, 4}) {
                    const size_t smem = (size_t)ns * sb;
                    if (smem > (size_t)(ctas == 148 ? 196 : 98) * 1024 || ns > 32) continue;
                    const size_t per = bytes / ctas / sb * sb;
                    snprintf(name, sizeof name, "bulk ring: %d CTAs, %d x %d B stages, %d copies/stage", ctas, ns, sb, copies);
                    time_it([&] { bulk_read<<<ctas, 64, smem>>>(buf, per, ns, sb, copies, out); }, name);
                }
    // ---- tensor-box patterns
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
    CK(cudaFuncSetAttribute(tma_box_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t use = (size_t)8 << 30;
    for (int k : {5120, 20480}) {
        const size_t n = use / k / 32 * 32;
        const int chunks = k / 512;
        for (int mode : {0, 1}) {
            CUtensorMap map;
            CUresult r;
            if (mode == 0) {
                const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)n};
                const cuuint64_t strides[1] = {(cuuint64_t)k};
                const cuuint32_t box[2] = {128, 32}, es[2] = {1, 1};
                r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            } else {
                const cuuint64_t dims[4] = {512, 32, (cuuint64_t)chunks, (cuuint64_t)(n / 32)};
                const cuuint64_t strides[3] = {512, 16384, (cuuint64_t)k * 32};
                const cuuint32_t box[4] = {128, 32, 1, 1}, es[4] = {1, 1, 1, 1};
                r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            for (int ctas : {148, 296})
                for (int ns : {4, 6}) {
                    if ((size_t)ns * 16384 + 2048 > (size_t)(ctas == 148 ? 196 : 98) * 1024) continue;
                    const int tiles_per_cta = (int)(n / 32 / ctas);
                    const double moved = (double)tiles_per_cta * ctas * 32 * k;
                    snprintf(name, sizeof name, "tensor boxes 4 x (32 x 128 B)/stage, %s, k=%d: %d CTAs, %d stages", mode ? "TILED 4-D" : "K-major 2-D", k, ctas, ns);
                    launch_and_time:
                    {
                        auto launch = [&] { tma_box_read<<<ctas, 64, (size_t)ns * 16384 + 1024>>>(map, mode, tiles_per_cta, chunks, ns, out); };
                        launch();
                        CK(cudaDeviceSynchronize());
                        float best = 1e9f;
                        for (int rr = 0; rr < 3; ++rr) {
                            cudaEventRecord(e0);
                            launch();
                            cudaEventRecord(e1);
                            CK(cudaDeviceSynchronize());
                            float ms;
                            cudaEventElapsedTime(&ms, e0, e1);
                            best = ms < best ? ms : best;
                        }
                        printf("%-78s %8.3f ms  %7.1f GB/s\n", name, best, moved / (best * 1e-3) / 1e9);
                    }
                }
        }
    }
    return 0;
}
