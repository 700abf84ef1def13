// Persistent decode step: ONE cooperative kernel per generated token (batch 1..2 per replica).
//
// Why (VERDICT r01 / profiles/r01_fused_gemv_ncu_full_summary.txt): the 5-launch-per-layer chain spends two thirds of
// every stage in a glue prologue that ~130 CTAs recompute behind block barriers, and each dependent launch costs
// >= 1.1 us, while the weights of a stage need only 0.3-1.7 us of HBM time. Here a decode step is a single grid of
// one CTA per SM that lives for the whole step:
//   * weight stream decoupled from the dependency chain: a dedicated TMA warp per CTA walks the static schedule of
//     the CTA's 16-row weight tiles (all 4 x L BitLinear stages, then its lm_head rows) and keeps a ~150 KB shared-
//     memory ring full (cp.async.bulk + mbarrier), up to a whole layer ahead of the compute warps;
//   * no kernel boundaries and no grid barriers between stages: producers publish into L2-resident exchange
//     buffers whose words carry their own validity (Lamport style: a reserved sentinel pattern means "not written
//     yet"), consumers poll the data itself, so a hop is one store -> L2 -> load trip;
//   * every element of glue (LayerNorm of bitnet.py:118, residual, RMSNorm, SiLU*up, *input_factor, quantisation
//     to MMA digits) is computed ONCE, by the CTA that owns the row, and published directly in MMA B-fragment
//     order; consumers only copy 4 bytes per column into shared memory;
//   * the bit-plane IMMA scheme of imma_gemv.cuh is kept (w & (0x01010101 << j) is an int8x4 A fragment), with
//     balanced base-255 digits (bytes in [-127,127]) so that 0x80 never occurs in a digit word and 0x80808080 can
//     serve as the "not written" sentinel at byte granularity.
// Reference being mirrored: one q_len = 1 forward of BitLlamaForCausalLMInf
// (transformers/src/transformers/models/bitllama/modeling_bitllama.py:1512-1611; layer :856-930; attention
// :431-583; MLP :223-257; RMSNorm :67-81; BitLinearInf bitnet.py:112-122) + greedy argmax generation/utils.py:2540.
#pragma once
#include "common.cuh"

namespace onebit {
namespace persist {

constexpr int kCW = 16;              // compute warps
constexpr int kCT = kCW * 32;        // compute threads
constexpr int kThreads = kCT + 32;   // + one TMA producer warp
constexpr int kMaxTok = 2;           // sequences per replica served by this kernel
constexpr int kNB = 16;              // ring chunks in flight (mbarrier pairs)
constexpr int kMaxTiles = 12;        // 16-row weight tiles of one stage per CTA
constexpr int kStatW = 12;           // floats per (CTA, token) statistics record
constexpr int kHeadDim = 128;
constexpr int kRedBytes = 4 * kMaxTiles * 16 * 8 * 4;  // K-split partial sums [KG<=4 | 8][tiles][16][8] int32
constexpr uint32_t kSentD = 0x80808080u;  // digit words: a byte 0x80 (= -128) is never a valid base-255 digit
constexpr uint32_t kSentF = 0xFFFFFFFFu;  // float words: this NaN pattern is never produced (canonicalised away)
constexpr int kTracePoints = 192;  // per trace row (one row per layer, + step start, + lm_head)
constexpr int kTracers = 2;       // CTA 0 and the last CTA record

struct BLDev {
    const uint8_t* w;   // [N][K/8] packed signs
    const void* g;      // [N] weight_scale (param dtype)
    const void* h;      // [K] input_factor (param dtype)
};

struct LayerDev {
    BLDev q, k, v, o, gate, up, down;
    const void* ln_in;    // input_layernorm.weight
    const void* ln_post;  // post_attention_layernorm.weight
    // static power-of-two bounds of the BitLinear inputs: |h * x| < 2^e (RMSNorm / LayerNorm outputs are bounded by
    // sqrt(width)), so the quantiser needs no amax reduction; down_proj's input gets a dynamic bound (see D1).
    int e_qkv[3];
    int e_o;
    int e_gu[2];
    float hmax_down;
};

struct Params {
    int H, I, L, heads, V, max_seq, max_batch, pdt, ncta, trace_on;
    float rms_eps, ln_eps;
    double inv_H, inv_I;
    const LayerDev* layers;
    const int* lscal;  // [L + 1][8] per-layer scalars for the parameter blocks: e_q, e_k, e_v, e_o, e_gate, e_up, max|down.h| bits, 0
    const __half* embed;
    const void* final_norm;
    const __half* lm_head;
    const float* rope_cos;
    const float* rope_sin;
    __half* kcache;
    __half* vcache;
    uint32_t* xch[2];  // exchange arenas (two parity sets: step s uses s & 1 and re-arms the other)
    // word offsets inside one layer's slice of an arena
    size_t o_xA, o_qkv, o_qst, o_xC, o_cst, o_xD1, o_dst, o_xD2, o_d2st, per_layer;
    size_t o_xfin, o_amax, total_words;
    unsigned long long* step_counter;
    int* abort_flag;
    unsigned long long* trace;  // [kTracers][L + 2][kTracePoints] globaltimer stamps (always on: a few stores per stage)
    long long* ids;
    int* pos;
};

struct Geometry {  // host-side plan shared by the launcher
    int ring_bytes, dbuf_bytes, smem_bytes;
};

}  // namespace persist

// host API (persist_step.cu)
struct PersistState;
int persist_create(PersistState** out, const onebit_decoder_config& cfg, const onebit_layer_params* layers,
                   const void* embed, const void* final_norm, const void* lm_head, const float* rope_cos,
                   const float* rope_sin, __half* kcache, __half* vcache, long long* ids, int* pos);
bool persist_supported(const onebit_decoder_config& cfg);
int persist_step(PersistState* S, int batch, const long long* ids_in, float* logits, cudaStream_t s);
int persist_read_trace(PersistState* S, unsigned long long* out, int n);
int persist_abort_flag(PersistState* S, int* out);
void persist_destroy(PersistState* S);

}  // namespace onebit
