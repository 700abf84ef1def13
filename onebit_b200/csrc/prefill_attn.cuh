// Prompt attention (prefill_attn.cu).
#pragma once
#include "common.cuh"

namespace onebit {

struct PrefillAttnArgs {
    const float* t_q; const float* t_k; const float* t_v;  // [M = B*T][ld] fp32 BitLinear outputs (already * weight_scale)
    int M, ld, n_ln;      // tokens, row stride of t_q/t_k/t_v, LayerNorm denominator (rows of the full q/k/v layers)
    int B, T, pos0, n_heads, max_seq;
    const float* rope_cos; const float* rope_sin;  // [max_seq][64]
    __half* kcache; __half* vcache;                // [B_cache][n_heads][max_seq][128] of this layer
    __half* q16;                                   // scratch [B][n_heads][T][128]
    __half* out16; int out_ld;                     // [M][out_ld] fp16 attention output
    float ln_eps;
};
int launch_prefill_attention(const PrefillAttnArgs& A, cudaStream_t s);

}  // namespace onebit
