// Standalone microbenchmark (no Python): streams the INT8 weights of a 13B decode token through ftcf_gemm_w8a16 in different
// launch arrangements to separate kernel efficiency from launch-boundary cost.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/gemm_chain_bench.cu -Iinclude -Lfastertransformer4codefuse_b200/lib -lftcf -o /tmp/gcb
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ftcf.h"

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } \
    } while (0)
#define FK(x)                                                    \
    do {                                                         \
        int s = (x);                                             \
        if (s) { printf("%s: %s\n", #x, ftcf_last_error()); exit(1); } \
    } while (0)

int main(int argc, char** argv)
{
    const int m = argc > 1 ? atoi(argv[1]) : 1;
    const int h = 5120, inter = 20480, L = 40;
    const int ks[4] = {h, h, h, inter}, ns[4] = {3 * h, h, inter, h};
    size_t total = 0;
    for (int i = 0; i < 4; ++i) total += (size_t)ks[i] * ns[i];
    uint8_t* w;
    CK(cudaMalloc(&w, total * L));
    CK(cudaMemset(w, 129, total * L));
    __half *x, *y, *scale;
    CK(cudaMalloc(&x, (size_t)64 * inter * 2));
    CK(cudaMalloc(&y, (size_t)64 * inter * 2 * 4));
    CK(cudaMalloc(&scale, (size_t)L * 4 * inter * 2));
    CK(cudaMemset(x, 0, (size_t)64 * inter * 2));
    CK(cudaMemset(scale, 0, (size_t)L * 4 * inter * 2));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    auto chain = [&]() {
        size_t off = 0;
        for (int l = 0; l < L; ++l)
            for (int i = 0; i < 4; ++i) {
                const size_t bytes = (size_t)ks[i] * ns[i];
                const int ni = (i + 1) % 4;
                ftcf_prefetch_hint hint{w + (off + bytes) % (total * L), ns[ni], ks[ni]};
                FK(ftcf_gemm_w8a16_ex(x, w + off, scale, nullptr, y, m, ns[i], ks[i], 0, 1, &hint, st));
                off += bytes;
            }
    };
    auto timed = [&](const char* name, auto&& fn, double bytes, int reps) {
        fn();
        CK(cudaStreamSynchronize(st));
        CK(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; ++r) fn();
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        printf("%-58s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / (ms * 1e-3) / 1e9);
    };

    // per-shape: 40 launches of ONE shape (different weights each), graph + PDL -> GB/s of that shape in a dependent chain.
    // The extra shapes test the per-SM balance theory: n = 148*32 and 296*32 rows give every SM the same number of CTAs.
    const int tn[] = {15360, 5120, 20480, 5120, 4736, 9472, 4736, 18944, 14208};
    const int tk[] = {5120, 5120, 5120, 20480, 20480, 20480, 5120, 5120, 5120};
    for (int ctas : {296})
        for (int i = 0; i < 9; ++i) {
            ftcf_set_tunable("skinny_prefetch_rows", 0);
            ftcf_set_tunable("skinny_pf_ahead", 0);
            ftcf_set_tunable("pdl", 1);
            ftcf_set_tunable("skinny_target_ctas", ctas);
            const size_t bytes = (size_t)tk[i] * tn[i];
            cudaGraph_t g;
            cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int l = 0; l < L; ++l) FK(ftcf_gemm_w8a16(x, w + (size_t)l * total, scale, nullptr, y, m, tn[i], tk[i], 0, 1, st));
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            char name[128];
            snprintf(name, sizeof(name), "m=%d ctas=%d: 40 x (n=%d, k=%d) chain, graph+pdl", m, ctas, tn[i], tk[i]);
            timed(name, [&]() { CK(cudaGraphLaunch(ge, st)); }, (double)bytes * L, 10);
            cudaGraphExecDestroy(ge);
            cudaGraphDestroy(g);
        }
    return 0;
}
