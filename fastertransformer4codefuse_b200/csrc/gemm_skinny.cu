// Skinny weight-streaming GEMM for the decode step (m <= 32 tokens per pass): y[m,n] = x[m,k] . W[k,n].
//
// Stands in for the reference's fpA_intB CUTLASS GEMM at decode shapes
// (kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:461-570; tile heuristic cutlass_heuristic.cc:128-212)
// and for cuBLAS on the fp16 / LM-head GEMMs (utils/cublasMMWrapper.cc:154-328, models/gptneox/GptNeoX.cc:869-912).
//
// Roofline: HBM.  At m <= 32 every weight byte is used for <= 64 flop, so the only thing that matters is streaming
// W once at full bandwidth.  Design:
//   * W is stored K-major (W^T, [n][k]); a CTA owns 16*RT output features and ALL of k, its 8 warps stride over k
//     in 128-byte-per-row steps (so every DRAM sector is consumed whole), 4*RT independent 16-byte loads in
//     flight per lane per step, no shared-memory staging of W (it is touched exactly once);
//   * u8 -> fp16 in registers: PRMT builds 0x64xx (1024 + b), one HSUB2 subtracts 1152 -> exact (b - 128);
//   * the 16 x (8*MT) x k product runs on mma.sync.m16n8k16 with the weights as the 16-row operand and the tokens
//     as the 8-column operand; fp32 accumulators; the contraction index is relabelled so that each lane's operand
//     fragments come from ONE contiguous 16-byte piece of a weight row (no shuffles, no transposes);
//   * cross-warp (split-k inside the CTA) reduction through shared memory in a fixed order -> deterministic;
//   * epilogue: per-column scale (fp32), bias, tanh-GELU, fp16 (or fp32 logits) store.
#include "common.cuh"

namespace ftcf {

enum { EPI_W8 = 0, EPI_F16 = 1, EPI_F32 = 2 };

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// four biased bytes -> (b0-128, b1-128) and (b2-128, b3-128) as half2 bit patterns
__device__ __forceinline__ void u8x4_to_h2x2(uint32_t w, uint32_t& lo, uint32_t& hi)
{
    lo = __byte_perm(w, 0x64646464u, 0x4140);
    hi = __byte_perm(w, 0x64646464u, 0x4342);
    const uint32_t magic = 0x64806480u;   // 1152.0 = 1024 + 128, twice
    asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(lo), "r"(magic));
    asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(hi), "r"(magic));
}

__device__ __forceinline__ uint32_t u4_get(const uint4& v, int i)
{
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

template <typename WT, int RT, int MT, int EPI>
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const __half* __restrict__ x, const WT* __restrict__ w, const __half* __restrict__ scale,
                   const __half* __restrict__ bias, void* __restrict__ y, int m, int n, int k, int ldy, int act)
{
    constexpr int EPC = 16 / sizeof(WT);   // k-elements per 16-byte chunk
    constexpr int KSTEP = 8 * EPC;         // k-elements a warp consumes per step (4 lanes x 2 chunks)
    constexpr int NMMA = EPC / 4;          // mma per chunk
    constexpr int XV = EPC / 4;            // uint4 per lane per token per step (2*EPC halves)
    constexpr int NF = 16 * RT, NT = 8 * MT, PITCH = NF + 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.x * NF, m0 = blockIdx.y * NT;

    float acc[RT][MT][4];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][mt][i] = 0.f;

    const WT* wrow[RT][2];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = min(n0 + r * 16 + g + 8 * hh, n - 1);
            wrow[r][hh] = w + (size_t)row * k + t * 2 * EPC;
        }
    const __half* xrow[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int tok = min(m0 + mt * 8 + g, m - 1);
        xrow[mt] = x + (size_t)tok * k + t * 2 * EPC;
    }

    const int iters = k / KSTEP;
#pragma unroll 2
    for (int it = warp; it < iters; it += 8) {
        const int kb = it * KSTEP;
        uint4 wv[RT][2][2];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int c = 0; c < 2; ++c) wv[r][hh][c] = ld_stream_16(wrow[r][hh] + kb + c * EPC);
        uint4 xv[MT][XV];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int i = 0; i < XV; ++i) xv[mt][i] = ld_ro_16(xrow[mt] + kb + i * 8);

#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int j = 0; j < NMMA; ++j) {
                const int pi = c * (EPC / 2) + 2 * j;   // 32-bit word index into the lane's x block
                uint32_t a[RT][4];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    if constexpr (sizeof(WT) == 1) {
                        u8x4_to_h2x2(u4_get(wv[r][0][c], j), a[r][0], a[r][2]);
                        u8x4_to_h2x2(u4_get(wv[r][1][c], j), a[r][1], a[r][3]);
                    } else {
                        a[r][0] = u4_get(wv[r][0][c], 2 * j);
                        a[r][2] = u4_get(wv[r][0][c], 2 * j + 1);
                        a[r][1] = u4_get(wv[r][1][c], 2 * j);
                        a[r][3] = u4_get(wv[r][1][c], 2 * j + 1);
                    }
                }
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t b0 = u4_get(xv[mt][pi / 4], pi % 4);
                    const uint32_t b1 = u4_get(xv[mt][(pi + 1) / 4], (pi + 1) % 4);
#pragma unroll
                    for (int r = 0; r < RT; ++r) mma_16816(acc[r][mt], a[r][0], a[r][1], a[r][2], a[r][3], b0, b1);
                }
            }
    }

    // ---- reduce the 8 warps' partial sums (fixed order), epilogue, store
    __shared__ float red[8][NT][PITCH];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int f = r * 16 + g, tok = mt * 8 + 2 * t;
            red[warp][tok][f] = acc[r][mt][0];
            red[warp][tok + 1][f] = acc[r][mt][1];
            red[warp][tok][f + 8] = acc[r][mt][2];
            red[warp][tok + 1][f + 8] = acc[r][mt][3];
        }
    __syncthreads();
    for (int o = threadIdx.x; o < NF * NT; o += 256) {
        const int f = o % NF, tok = o / NF;
        const int col = n0 + f, row = m0 + tok;
        if (col >= n || row >= m) continue;
        float v = 0.f;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) v += red[wi][tok][f];
        if constexpr (EPI == EPI_W8) {
            v *= __half2float(scale[col]);
            if (bias != nullptr) v += __half2float(bias[col]);
            if (act == 1) v = gelu_tanh_f32(v);
            reinterpret_cast<__half*>(y)[(size_t)row * ldy + col] = __float2half_rn(v);
        } else if constexpr (EPI == EPI_F16) {
            __half hv = __float2half_rn(v);
            if (bias != nullptr) hv = __hadd(hv, bias[col]);
            if (act == 1) hv = gelu_tanh_half_ref(hv);
            reinterpret_cast<__half*>(y)[(size_t)row * ldy + col] = hv;
        } else {
            reinterpret_cast<float*>(y)[(size_t)row * ldy + col] = v;
        }
    }
}

template <typename WT, int EPI>
static int launch_skinny(const void* x, const void* w, const void* scale, const void* bias, void* y, int m, int n, int k,
                         int ldy, int act, cudaStream_t st)
{
    constexpr int EPC = 16 / sizeof(WT);
    FTCF_REQUIRE(k % (8 * EPC) == 0, FTCF_ERR_UNSUPPORTED, "skinny gemm: k=%d must be a multiple of %d", k, 8 * EPC);
    FTCF_REQUIRE(m > 0 && n > 0, FTCF_ERR_INVALID, "skinny gemm: empty problem m=%d n=%d", m, n);
    const int mt = m >= 25 ? 4 : ceil_div(m, 8);
    const bool wide = n >= 16 * 2 * 148 * 2;
    const dim3 block(256);
    const __half* xs = static_cast<const __half*>(x);
    const WT* ws = static_cast<const WT*>(w);
    const __half* sc = static_cast<const __half*>(scale);
    const __half* bs = static_cast<const __half*>(bias);
#define FTCF_SK(RT_, MT_)                                                                                      \
    gemm_skinny_kernel<WT, RT_, MT_, EPI><<<dim3(ceil_div(n, 16 * RT_), ceil_div(m, 8 * MT_)), block, 0, st>>>( \
        xs, ws, sc, bs, y, m, n, k, ldy, act)
    if (wide) {
        switch (mt) {
            case 1: FTCF_SK(2, 1); break;
            case 2: FTCF_SK(2, 2); break;
            case 3: FTCF_SK(2, 3); break;
            default: FTCF_SK(2, 4); break;
        }
    } else {
        switch (mt) {
            case 1: FTCF_SK(1, 1); break;
            case 2: FTCF_SK(1, 2); break;
            case 3: FTCF_SK(1, 3); break;
            default: FTCF_SK(1, 4); break;
        }
    }
#undef FTCF_SK
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

int gemm_w8a16_skinny(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k,
                      int act, cudaStream_t st)
{
    return launch_skinny<uint8_t, EPI_W8>(x, w_nk, scale, bias, y, m, n, k, n, act, st);
}

int gemm_f16_skinny(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                    int out_f32, cudaStream_t st)
{
    if (out_f32) return launch_skinny<__half, EPI_F32>(x, w_nk, nullptr, bias, y, m, n, k, ldy, act, st);
    return launch_skinny<__half, EPI_F16>(x, w_nk, nullptr, bias, y, m, n, k, ldy, act, st);
}

// ---------------------------------------------------------------- fp16 [k,n] -> [n,k]
__global__ void transpose_f16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int k, int n)
{
    __shared__ __half tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx over n, by over k
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = by + i, c = bx + threadIdx.x;
        if (r < k && c < n) tile[i][threadIdx.x] = in[(size_t)r * n + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = bx + i, c = by + threadIdx.x;          // out row = n index, col = k index
        if (r < n && c < k) out[(size_t)r * k + c] = tile[threadIdx.x][i];
    }
}

}  // namespace ftcf

extern "C" int ftcf_transpose_f16(const void* in_kn, void* out_nk, int k, int n, void* stream)
{
    using namespace ftcf;
    FTCF_REQUIRE(k > 0 && n > 0, FTCF_ERR_INVALID, "transpose: empty matrix");
    transpose_f16_kernel<<<dim3(ceil_div(n, 32), ceil_div(k, 32)), dim3(32, 8), 0, as_stream(stream)>>>(
        static_cast<const __half*>(in_kn), static_cast<__half*>(out_nk), k, n);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}
