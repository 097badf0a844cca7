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

    for (int pf : {0, 16, 32, 64, 128}) {
        for (int ctas : {296, 444}) {
            const int pdl = 0;
            ftcf_set_tunable("skinny_prefetch_rows", pf);
            ftcf_set_tunable("pdl", pdl);
            ftcf_set_tunable("skinny_target_ctas", ctas);
            char name[128];
            // (a) one launch per matrix shape but covering ALL layers' rows: n = L * n_i (same k)  -> 4 big launches
            snprintf(name, sizeof(name), "m=%d pf=%d ctas=%d: 3 big launches (k=5120, all layers)", m, pf, ctas);
            timed(name, [&]() {
                // treat the whole buffer as [rows, 5120]: total bytes / 5120 rows (valid because every matrix is K-major bytes)
                const size_t rows = (size_t)(3 * h + h + inter) * L;   // k = 5120 part: 40960 rows per layer
                const size_t per = rows / 3;
                for (int j = 0; j < 3; ++j) FK(ftcf_gemm_w8a16(x, w + j * per * 5120, scale, nullptr, y, m, (int)per > 20480 * 16 ? 20480 * 16 : (int)per, 5120, 0, 1, st));
            }, 3.0 * 20480 * 16 * 5120, 3);
            // (b) the real chain, plain stream launches
            snprintf(name, sizeof(name), "m=%d pf=%d ctas=%d: 160-launch chain, stream", m, pf, ctas);
            timed(name, chain, (double)total * L, 5);
            // (c) the real chain, captured into a graph
            cudaGraph_t g;
            cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            chain();
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            snprintf(name, sizeof(name), "m=%d pf=%d ctas=%d: 160-launch chain, graph", m, pf, ctas);
            timed(name, [&]() { CK(cudaGraphLaunch(ge, st)); }, (double)total * L, 10);
            cudaGraphExecDestroy(ge);
            cudaGraphDestroy(g);
        }
    }
    return 0;
}
