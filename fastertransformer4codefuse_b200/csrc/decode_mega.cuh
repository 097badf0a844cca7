// Shared declarations of the persistent decode-step kernel (decode_mega.cu) and its caller (engine.cu).
#pragma once
#include "common.cuh"

namespace ftcf {
namespace mg {

constexpr int CW = 8;                       // consumer warps (threads 0..255 = warpgroups 0 and 1)
constexpr int CT = CW * 32;
constexpr int THREADS = CT + 128;           // + the producer warpgroup (warp 8 works; it exists as a warpgroup so that
                                            // setmaxnreg can hand its registers to the consumers)
constexpr int ROWS = 16;                    // output features per weight unit (one MMA M tile)
constexpr int STAGE_K = 1024;               // bytes of each weight row per ring stage
constexpr int STAGE_BYTES = ROWS * STAGE_K;     // 16384 == 64 keys x 256 bytes
constexpr int ATT_TILE = 64;                // keys per K (or V) stage
constexpr int ATT_UNIT = 160;               // keys per attention work unit (~80 KB of K+V at dh = 128)
constexpr int MAX_B = 8;
constexpr int RED_PITCH = 20;
struct LayerDev {
    const void* w[4];          // qkv, o, ffn1, ffn2 in the TILED layout (mega_retile): u8 (value q + 128) or fp16
    const __half* scale[4];    // per-output-feature dequant scales (int8 only)
    const __half *ln1_g, *ln1_b, *ln2_g, *ln2_b, *qkv_b, *ffn1_b, *res_b;
};

struct Params {
    const LayerDev* layers;
    int l0, l1;                // layers [l0, l1) run in this launch
    int embed;                 // 1: the input of layer 0 is wte[out_ids[step - 1]]
    int lm_rows;               // rows of the LM head computed after the last layer (0: none)
    int B, h, Hl, hl, inter, dh, rot, max_len, max_in, tp, vocab;
    int ks;                    // FFN2 split-k factor (k chunks of h elements)
    int att_max_units;
    int ns;                    // ring stages
    int inflight;              // bulk copies per SM allowed in flight at once (<= ns)
    int opnd_pitch;            // bytes per token row of the operand buffer
    float eps, inv_sqrt_dh;
    const __half *wte, *lnf_g, *lnf_b;
    const void* lm_head;       // fp16 [lm_rows][h] in the tiled layout
    float* logits;
    int ld_logits;
    __half *x, *qkv, *inter_buf, *ctx;
    float *ffn_part, *att_part;
    int32_t* att_cnt;
    __half* kv;
    size_t kv_layer_elems;
    const int32_t *out_ids, *step, *seq_len, *input_len, *pad_count;
    const uint8_t* finished;
    unsigned* gbar;
    int dbg;                   // timing experiments only (results are wrong when non-zero): see decode_mega.cu
};

}  // namespace mg

// host side (decode_mega.cu)
// Tiled weight layout read by the kernel.  A K-major matrix [n][kbytes] is cut into blocks of 16 rows x 1024 bytes of k (the
// last block of a row may be narrower); block (rt, kc) is stored contiguously at rt * 16 * kbytes + kc * 16384 as
// [16 rows][kw bytes] with the 16-byte chunk c of row r at chunk c ^ (r & 7): exactly the shared-memory image one ring stage
// wants (bank-conflict-free fragment reads), so a stage is ONE contiguous cp.async.bulk.  Rows are zero-padded to 16.
size_t mega_tiled_bytes(int n, int kbytes);
int mega_retile(const void* w_nk, void* out, int n, int kbytes, cudaStream_t st);
bool mega_supported(int B, int h, int hl, int inter, int dh, int rot, bool w8, int tp, bool parallel_residual);
int mega_plan(mg::Params& p, bool w8);          // fills ks, opnd_pitch, att_max_units, ns from the dimensions in p
size_t mega_smem_bytes(const mg::Params& p);
int mega_launch(const mg::Params& p, bool w8, cudaStream_t st);
extern std::atomic<int> g_mega_dbg, g_mega_ns, g_mega_inflight;   // tunables "mega_dbg" (timing experiments), "mega_ns" (ring depth cap)   // zeroes p.gbar, then one launch on all SMs

}  // namespace ftcf
