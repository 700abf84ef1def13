// Fused decode step around the bit-plane IMMA GEMV — the callers either side of the hot path
// (SURVEY.md §8f-1). Mirrors one q_len = 1 forward of the reference's BitLlamaForCausalLMInf
// (transformers/src/transformers/models/bitllama/modeling_bitllama.py):
//   embed_tokens :1202 | LlamaRMSNorm :67-81 | q/k/v :522-524 | RoPE :176-181 | cache + attention :536-563 |
//   o_proj :580 | residual :912 | MLP down(silu(gate) * up) :257 | final norm :1315 | lm_head :1610 |
//   greedy argmax generation/utils.py:2540
// Every BitLinearInf (bitnet.py:112-122) = IMMA GEMV (imma_gemv.cuh) + the LayerNorm of bitnet.py:118, whose
// statistics come from per-CTA partial sums the GEMV emits and are applied in the consumer's prologue.
//
// Launches per layer, by path:
//   batch 1..2, one GPU   5: fused stage q,k,v | attention | fused o | fused gate,up | fused down   (fused_gemv2.cuh)
//   batch 1..2, tp > 1    5 + 2 statistics reductions + 4 small all-reduces                          (fused_gemv.cuh)
//   batch 3..4            9: glue | GEMV qkv | attention | glue | GEMV o | glue | GEMV gate,up | glue | GEMV down
//   batch 5..64           9 (one GPU) / 13 (tp > 1) on the tcgen05 decode tile, see run_tc5_layers
//   ONEBIT_PERSIST=1      the whole step is one cooperative kernel (persist_step.cu)
// All state that changes between steps (token ids, positions) lives in device memory, so a step is one
// CUDA graph replay. The residual stream is kept in fp32.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "fused_gemv.cuh"
#include "fused_gemv2.cuh"
#include "imma_gemv.cuh"
#include "p2p_allreduce.cuh"
#include "prefill_attn.cuh"
#include "persist_step.cuh"

namespace onebit {

int launch_imma_gemv(const imma::Args& a, int param_dtype, cudaStream_t s);  // matvec_mma.cu

namespace {

using imma::QMeta;
constexpr int kGlueThreads = 512;
constexpr int kHeadDim = 128;
constexpr int kMaxBatch = 64;     // sequences per replica (tcgen05 decode tile: 64 tokens)
constexpr int kKSplitMax = 8;    // split-K of the batched-decode projections (t_* buffers hold that many partials)

enum GlueMode {
    GLUE_EMBED_NORM = 0,   // resid_out = embed[ids];                x = RMSNorm(resid_out) * ln_w
    GLUE_RESID_NORM = 1,   // resid_out = resid_in + LN_a(t_a);      x = RMSNorm(resid_out) * ln_w
    GLUE_SILU_MUL = 2,     // x = silu(LN_a(t_a)) * LN_b(t_b)
    GLUE_PLAIN = 3,        // x = x_plain
};

struct GlueArgs {
    int mode, M, K, nprob, write_x_f16;
    int stats_from_data;  // RESID_NORM / SILU_MUL: (sum, sumsq) of t_a (t_b) from the data itself (all-reduced inputs; prompt pass)
    int n_ln;             // LayerNorm denominator of the producer(s) when it is not K (tensor-parallel shards); 0 = K
    // batched decode, single GPU: t_a (t_b) still are `ksplit` split-K partial slabs `slab` floats apart — summed while
    // loading (fixed order), statistics from the data: no separate reduce + statistics launch between GEMM and glue
    int ksplit; long long slab;
    const float* t_a; const float* stats_a; int ncta_a;   // previous GEMV output (already * g), [ncta][M][2] partials
    const float* t_b; const float* stats_b; int ncta_b;
    const float* resid_in; float* resid_out;              // [M][K] fp32
    const __half* embed; const long long* ids;            // GLUE_EMBED_NORM
    const void* ln_w;                                      // RMSNorm weight (TP)
    const float* x_plain;
    float ln_eps, rms_eps;
    const void* h[3]; uint8_t* digits[3]; QMeta* qmeta[3];  // per problem: input_factor, outputs
    __half* x_f16;                                          // optional [M][K] (lm_head input)
};

// ---- block-wide helpers for the latency-critical single-CTA kernels --------------------------------------
// All global loads a thread needs are issued up front (one exposed L2 round trip per kernel instead of one per
// loop iteration); reductions carry up to four values through one shared-memory exchange.
template <int NVAL>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NVAL], double* sh /*[32][NVAL] + NVAL*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NVAL; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    __syncthreads();  // protect sh from the previous use
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NVAL; ++i) sh[warp * NVAL + i] = v[i];
    __syncthreads();
    if (warp == 0) {  // second level in one warp (fixed order: deterministic), result broadcast through sh
#pragma unroll
        for (int i = 0; i < NVAL; ++i) {
            double r = lane < nw ? sh[lane * NVAL + i] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (lane == 0) sh[32 * NVAL + i] = r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NVAL; ++i) v[i] = sh[32 * NVAL + i];
}

template <typename TP>
__device__ __forceinline__ float4 load_param4(const TP* p, int i4);
template <>
__device__ __forceinline__ float4 load_param4<float>(const float* p, int i4) {
    return reinterpret_cast<const float4*>(p)[i4];
}
template <>
__device__ __forceinline__ float4 load_param4<__half>(const __half* p, int i4) {
    const uint2 r = reinterpret_cast<const uint2*>(p)[i4];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 load_param4<__nv_bfloat16>(const __nv_bfloat16* p, int i4) {
    const uint2 r = reinterpret_cast<const uint2*>(p)[i4];
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// (sum, sum of squares) partials of one BitLinear output: this thread's share of the [ncta][M][2] array
__device__ __forceinline__ void load_stat_partials(const float* stats, int ncta, int M, int m, double& s, double& q) {
    for (int c = threadIdx.x; c < ncta; c += blockDim.x) {
        const float2 p = *reinterpret_cast<const float2*>(stats + ((size_t)c * M + m) * 2);
        s += (double)p.x;
        q += (double)p.y;
    }
}
__device__ __forceinline__ void finish_ln(double s, double q, int n_rows, float eps, float& mean, float& rstd) {
    const double mu = s / (double)n_rows;
    const double var = fmax(q / (double)n_rows - mu * mu, 0.0);  // biased variance (nn.LayerNorm)
    mean = (float)mu;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// One CTA per (token, problem): builds x (the BitLinear input) in shared memory, then quantises h_p * x.
// NV4 = ceil(K / 4 / kGlueThreads): float4 values each thread keeps in registers.
template <typename TP, int NV4>
__global__ void __launch_bounds__(kGlueThreads, 1) glue_kernel(const __grid_constant__ GlueArgs A) {
    extern __shared__ __align__(16) float xs[];  // [K]
    __shared__ double shd[33 * 4];
    __shared__ long long scratch[kGlueThreads / 32];
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    const int m = blockIdx.x, p = blockIdx.y, K = A.K, K4 = K >> 2, tid = threadIdx.x;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 va[NV4], vb[NV4];
    fused::Raw4<TP> vw[NV4], vh[NV4];
    fused::Raw4<__half> ve[NV4];
    const bool need_h = !A.write_x_f16;
    const bool norm_mode = A.mode == GLUE_EMBED_NORM || A.mode == GLUE_RESID_NORM;
    // ---- every global load, up front
    double st[4] = {0.0, 0.0, 0.0, 0.0};
    const float4* a4 = nullptr;
    const float4* b4 = nullptr;
    const __half* erow = nullptr;
    if (A.mode == GLUE_EMBED_NORM) erow = A.embed + (size_t)A.ids[m] * K;
    else if (A.mode == GLUE_RESID_NORM) { a4 = reinterpret_cast<const float4*>(A.t_a + (size_t)m * K); b4 = reinterpret_cast<const float4*>(A.resid_in + (size_t)m * K); }
    else if (A.mode == GLUE_SILU_MUL) { a4 = reinterpret_cast<const float4*>(A.t_a + (size_t)m * K); b4 = reinterpret_cast<const float4*>(A.t_b + (size_t)m * K); }
    else a4 = reinterpret_cast<const float4*>(A.x_plain + (size_t)m * K);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int i4 = i * kGlueThreads + tid;
        const bool ok = i4 < K4;
        va[i] = zero4; vb[i] = zero4;
        if (ok) {  // parameter vectors stay raw until used (a conversion right here would serialise the round trips)
            if (erow) ve[i] = fused::ldraw4<__half>(erow, i4);
            if (a4) va[i] = a4[i4];
            if (b4) vb[i] = b4[i4];
            if (norm_mode) vw[i] = fused::ldraw4<TP>(static_cast<const TP*>(A.ln_w), i4);
            if (need_h) vh[i] = fused::ldraw4<TP>(static_cast<const TP*>(A.h[p]), i4);
        }
    }
    // ---- split-K partial slabs of the producer GEMM (batched decode, single GPU): summed in slab order, ZG slabs of loads
    //      in flight at a time
    if (A.ksplit > 1 && (A.mode == GLUE_RESID_NORM || A.mode == GLUE_SILU_MUL)) {
        constexpr int ZG = NV4 <= 2 ? 4 : (NV4 <= 3 ? 2 : 1);
        const long long slab4 = A.slab >> 2;
        for (int z0 = 1; z0 < A.ksplit; z0 += ZG) {
            float4 wa[ZG][NV4], wb[ZG][NV4];
#pragma unroll
            for (int zz = 0; zz < ZG; ++zz)
#pragma unroll
                for (int i = 0; i < NV4; ++i) {
                    const int i4 = i * kGlueThreads + tid;
                    wa[zz][i] = zero4; wb[zz][i] = zero4;
                    if (i4 < K4 && z0 + zz < A.ksplit) {
                        wa[zz][i] = a4[(z0 + zz) * slab4 + i4];
                        if (A.mode == GLUE_SILU_MUL) wb[zz][i] = b4[(z0 + zz) * slab4 + i4];
                    }
                }
#pragma unroll
            for (int zz = 0; zz < ZG; ++zz)
#pragma unroll
                for (int i = 0; i < NV4; ++i) {
                    va[i].x += wa[zz][i].x; va[i].y += wa[zz][i].y; va[i].z += wa[zz][i].z; va[i].w += wa[zz][i].w;
                    vb[i].x += wb[zz][i].x; vb[i].y += wb[zz][i].y; vb[i].z += wb[zz][i].z; vb[i].w += wb[zz][i].w;
                }
        }
    }
    // ---- LayerNorm statistics of the producer(s) (bitnet.py:118), from the GEMV's per-CTA partials
    // (loaded after the big loads in program order: their fp64 conversion stalls the warp until they return)
    if ((A.mode == GLUE_RESID_NORM || A.mode == GLUE_SILU_MUL) && !A.stats_from_data)
        load_stat_partials(A.stats_a, A.ncta_a, A.M, m, st[0], st[1]);
    if (A.mode == GLUE_SILU_MUL && !A.stats_from_data) load_stat_partials(A.stats_b, A.ncta_b, A.M, m, st[2], st[3]);
    if (A.mode == GLUE_SILU_MUL && A.stats_from_data) {  // the rows are in registers anyway: no separate statistics pass
#pragma unroll
        for (int i = 0; i < NV4; ++i)
            if (i * kGlueThreads + tid < K4) {
                st[0] += (double)((va[i].x + va[i].y) + (va[i].z + va[i].w));
                st[1] += (double)((va[i].x * va[i].x + va[i].y * va[i].y) + (va[i].z * va[i].z + va[i].w * va[i].w));
                st[2] += (double)((vb[i].x + vb[i].y) + (vb[i].z + vb[i].w));
                st[3] += (double)((vb[i].x * vb[i].x + vb[i].y * vb[i].y) + (vb[i].z * vb[i].z + vb[i].w * vb[i].w));
            }
    }
    if (A.mode == GLUE_RESID_NORM && A.stats_from_data) {
#pragma unroll
        for (int i = 0; i < NV4; ++i)
            if (i * kGlueThreads + tid < K4) {
                st[0] += (double)((va[i].x + va[i].y) + (va[i].z + va[i].w));
                st[1] += (double)((va[i].x * va[i].x + va[i].y * va[i].y) + (va[i].z * va[i].z + va[i].w * va[i].w));
            }
    }
    float mean_a = 0.f, rstd_a = 1.f, mean_b = 0.f, rstd_b = 1.f;
    if (A.mode == GLUE_RESID_NORM || A.mode == GLUE_SILU_MUL) {
        block_reduce_sum<4>(st, shd);
        const int n_ln = A.n_ln > 0 ? A.n_ln : K;
        finish_ln(st[0], st[1], n_ln, A.ln_eps, mean_a, rstd_a);
        if (A.mode == GLUE_SILU_MUL) finish_ln(st[2], st[3], n_ln, A.ln_eps, mean_b, rstd_b);
    }
    // ---- x
    if (norm_mode) {
        double ss[1] = {0.0};
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
            const int i4 = i * kGlueThreads + tid;
            if (i4 < K4) {
                float4 r = A.mode == GLUE_EMBED_NORM ? fused::cvt4(ve[i]) : va[i];
                if (A.mode == GLUE_RESID_NORM) {  // residual + LayerNorm(o / down output), :912 / :918
                    r.x = vb[i].x + (va[i].x - mean_a) * rstd_a; r.y = vb[i].y + (va[i].y - mean_a) * rstd_a;
                    r.z = vb[i].z + (va[i].z - mean_a) * rstd_a; r.w = vb[i].w + (va[i].w - mean_a) * rstd_a;
                }
                if (p == 0) reinterpret_cast<float4*>(A.resid_out + (size_t)m * K)[i4] = r;
                va[i] = r;
                part += r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
            }
        }
        ss[0] = (double)part;
        block_reduce_sum<1>(ss, shd);
        const float rr = rsqrtf((float)(ss[0] / (double)K) + A.rms_eps);  // LlamaRMSNorm :77-78
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
            if (i * kGlueThreads + tid < K4) {
                const float4 w4 = fused::cvt4(vw[i]);
                va[i].x *= rr * w4.x; va[i].y *= rr * w4.y; va[i].z *= rr * w4.z; va[i].w *= rr * w4.w;
            }
        }
    } else if (A.mode == GLUE_SILU_MUL) {
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
            const float a[4] = {va[i].x, va[i].y, va[i].z, va[i].w}, b[4] = {vb[i].x, vb[i].y, vb[i].z, vb[i].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float g = (a[e] - mean_a) * rstd_a, u = (b[e] - mean_b) * rstd_b;
                o[e] = __fdividef(g, 1.f + __expf(-g)) * u;  // act_fn(gate) * up, modeling_bitllama.py:257
            }
            va[i] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    if (A.write_x_f16) {
        if (p == 0)
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                const int i4 = i * kGlueThreads + tid;
                if (i4 < K4) {
                    const __half2 lo = __floats2half2_rn(va[i].x, va[i].y), hi = __floats2half2_rn(va[i].z, va[i].w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                    reinterpret_cast<uint2*>(A.x_f16 + (size_t)m * K)[i4] = pk;
                }
            }
        return;
    }
    // ---- x' = h_p * x (bitnet.py:113), amax, digits
    float am = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int i4 = i * kGlueThreads + tid;
        if (i4 < K4) {
            const float4 h4 = fused::cvt4(vh[i]);
            const float4 v = make_float4(va[i].x * h4.x, va[i].y * h4.y, va[i].z * h4.z, va[i].w * h4.w);
            reinterpret_cast<float4*>(xs)[i4] = v;
            am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
    __syncthreads();
    imma::quantize_from_smem(xs, K, am, A.digits[p] + (size_t)m * (K / imma::kUnitCols) * imma::kUnitBytes,
                             A.qmeta[p] + m, scratch);
}

// ---- attention for one new token per sequence: LN(q,k,v) -> RoPE -> cache append -> softmax(QK^T/sqrt(d)) V ----
struct AttnArgs {
    const float* t_q; const float* t_k; const float* t_v;  // [M][H] fp32 (already * g)
    const float* stats_q; const float* stats_k; const float* stats_v; int ncta;  // per-CTA partials of each projection
    int M, H, n_heads, max_seq;   // H = row stride of t_q/t_k/t_v (local width under tensor parallelism)
    int n_ln, out_ld;            // rows of the FULL q/k/v layers (LayerNorm denominator); row stride of `out`
    double inv_nln;              // 1 / n_ln (filled by the launcher)
    const int* pos;                 // [M] device
    const float* rope_cos; const float* rope_sin;  // [max_seq][kHeadDim/2]
    __half* kcache; __half* vcache; // [M_max][n_heads][max_seq][kHeadDim] for this layer
    float* out;                     // [M][H] fp32
    float ln_eps;
    int t_cap;                      // cache rows that fit in shared memory (bulk-copied up front); longer contexts stream
    // split-KV (flash decoding): grid.z CTAs share one (sequence, head), each a contiguous slice of the context; the last
    // one to finish (ticket) merges the partial (max, sum, numerator) records in slice order
    int nsplit;
    float* part;                    // [M][n_heads][nsplit][kHeadDim + 2]
    int* tickets;                   // [M][n_heads], zero between launches
    float* amax;                    // [M][n_heads] max |out| of every (sequence, head): quantiser bound of the o_proj stage, or nullptr
    __half* out16;                  // optional fp16 copy of `out` with row stride out_ld (the batched path's o_proj reads it through TMA)
};

__global__ void __launch_bounds__(kHeadDim) attn_kernel(const __grid_constant__ AttnArgs A) {
    extern __shared__ __align__(16) float sc[];  // [max(max_seq, 512)] scores, then K rows, then V rows (fp16)
    __shared__ __align__(8) uint64_t kv_bar;
    __shared__ __align__(16) float qs[kHeadDim];
    __shared__ double shd[33 * 6];
    __shared__ float red[kHeadDim / 32];
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    const int m = blockIdx.x, hd = blockIdx.y, d = threadIdx.x, lane = d & 31, warp = d >> 5;
    const int pos = min(max(A.pos[m], 0), A.max_seq - 1), T = pos + 1;  // host refuses earlier; never index past the cache
    // this CTA's slice of the context (the whole context when nsplit == 1)
    const int nsplit = A.nsplit, zi = blockIdx.z, chunk = (T + nsplit - 1) / nsplit;
    const int j_lo = min(zi * chunk, T), j_hi = min(T, j_lo + chunk);
    const bool owns_new = pos >= j_lo && pos < j_hi;  // the slice that holds the new token appends it to the cache
    __half* kc = A.kcache + (((size_t)m * A.n_heads + hd) * A.max_seq) * kHeadDim;
    __half* vc = A.vcache + (((size_t)m * A.n_heads + hd) * A.max_seq) * kHeadDim;
    // The cached rows 0..pos-1 of this (sequence, head) are contiguous: pull them into shared memory with two bulk
    // (TMA) copies issued before anything else, so the score and P.V loops never wait on a global load.
    const bool in_smem = T <= A.t_cap;
    __half* Ks = reinterpret_cast<__half*>(sc + max(A.max_seq, 4 * kHeadDim));
    __half* Vs = Ks + (size_t)A.t_cap * kHeadDim;
    if (d == 0) {
        imma::mbar_init(&kv_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (in_smem && pos > 0) {
            imma::mbar_expect_tx(&kv_bar, (uint32_t)(2 * pos * kHeadDim * 2));
            imma::bulk_g2s(Ks, kc, (uint32_t)(pos * kHeadDim * 2), &kv_bar);
            imma::bulk_g2s(Vs, vc, (uint32_t)(pos * kHeadDim * 2), &kv_bar);
        }
    }
    // raw q / k / v of this thread's dimension and its rotate_half partner, RoPE factors: in flight while the LayerNorm sums
    // are reduced (one exposed L2 round trip instead of two)
    const size_t col = (size_t)m * A.H + hd * kHeadDim;
    const int half = kHeadDim / 2, dp = d < half ? d + half : d - half, fi = d < half ? d : d - half;
    const float c = A.rope_cos[(size_t)pos * half + fi], s = A.rope_sin[(size_t)pos * half + fi];
    const float tq0 = A.t_q[col + d], tq1 = A.t_q[col + dp], tk0 = A.t_k[col + d], tk1 = A.t_k[col + dp], tv0 = A.t_v[col + d];
    float mq, rq, mk, rk, mv, rv;
    {
        // every warp sums the per-CTA partials on its own (lane c takes CTAs c, c + 32, ...; all loads of a round of 160
        // CTAs are issued before the first use; butterfly in fp64): no block barrier, no fp64 division on the chain
        double st[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const float* sp[3] = {A.stats_q, A.stats_k, A.stats_v};
        for (int c0 = 0; c0 < A.ncta; c0 += 160) {
            float2 p[3][5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int c = c0 + lane + 32 * i;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    p[j][i] = make_float2(0.f, 0.f);
                    if (c < A.ncta) p[j][i] = *reinterpret_cast<const float2*>(sp[j] + ((size_t)c * A.M + m) * 2);
                }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) { st[2 * j] += (double)p[j][i].x; st[2 * j + 1] += (double)p[j][i].y; }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) st[j] = fused2::wsum(st[j]);
        const double inv_n = A.inv_nln;
        float* mo[3] = {&mq, &mk, &mv};
        float* ro[3] = {&rq, &rk, &rv};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double mu = st[2 * j] * inv_n;
            const double var = fma(-mu, mu, st[2 * j + 1] * inv_n);  // biased variance (nn.LayerNorm)
            *mo[j] = (float)mu;
            *ro[j] = rsqrtf(fmaxf((float)var, 0.f) + A.ln_eps);
        }
    }
    const float q0 = (tq0 - mq) * rq, q1 = (tq1 - mq) * rq;
    const float k0 = (tk0 - mk) * rk, k1 = (tk1 - mk) * rk;
    // rotate_half: (-x2, x1)  (modeling_bitllama.py:168-181)
    const float qr = d < half ? q0 * c - q1 * s : q0 * c + q1 * s;
    const float kr = d < half ? k0 * c - k1 * s : k0 * c + k1 * s;
    const float vv = (tv0 - mv) * rv;
    if (owns_new) {
        kc[(size_t)pos * kHeadDim + d] = __float2half_rn(kr);
        vc[(size_t)pos * kHeadDim + d] = __float2half_rn(vv);
    }
    if (in_smem) {
        Ks[(size_t)pos * kHeadDim + d] = __float2half_rn(kr);
        Vs[(size_t)pos * kHeadDim + d] = __float2half_rn(vv);
    }
    qs[d] = qr * 0.08838834764831845f;  // 1/sqrt(128), :546
    __syncthreads();
    if (in_smem && pos > 0) imma::mbar_wait(&kv_bar, 0);
    const __half* kr_base = in_smem ? Ks : kc;
    const __half* vr_base = in_smem ? Vs : vc;
    // scores: one warp per cached position, lanes over d (4 each); 4 positions in flight per warp
    const float4 qv = *reinterpret_cast<const float4*>(&qs[4 * lane]);
    constexpr int NW = kHeadDim / 32;
    for (int j0 = j_lo + warp; j0 < j_hi; j0 += 4 * NW) {
        uint2 raw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * NW;
            raw[u] = make_uint2(0u, 0u);
            if (j < j_hi) raw[u] = *reinterpret_cast<const uint2*>(kr_base + (size_t)j * kHeadDim + 4 * lane);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * NW;
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].y));
            float dot = qv.x * a.x + qv.y * a.y + qv.z * b.x + qv.w * b.y;
            dot = warp_sum(dot);
            if (lane == 0 && j < j_hi) sc[j - j_lo] = dot;
        }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = j_lo + d; j < j_hi; j += kHeadDim) mx = fmaxf(mx, sc[j - j_lo]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int j = j_lo + d; j < j_hi; j += kHeadDim) {
        const float e = expf(sc[j - j_lo] - mx);
        sc[j - j_lo] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float total = red[0] + red[1] + red[2] + red[3];
    const float inv = 1.f / total;
    // P.V: warp w takes positions w, w+4, ...; lane holds dims 4*lane .. 4*lane+3; partials combined through smem
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = j_lo + warp; j0 < j_hi; j0 += 4 * NW) {
        uint2 raw[4];
        float pj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u * NW;
            raw[u] = make_uint2(0u, 0u);
            pj[u] = 0.f;
            if (j < j_hi) {
                raw[u] = *reinterpret_cast<const uint2*>(vr_base + (size_t)j * kHeadDim + 4 * lane);
                pj[u] = sc[j - j_lo];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].y));
            pv.x += pj[u] * a.x; pv.y += pj[u] * a.y; pv.z += pj[u] * b.x; pv.w += pj[u] * b.y;
        }
    }
    __syncthreads();  // sc is dead from here on: reuse it for the cross-warp partials [NW][kHeadDim]
    *reinterpret_cast<float4*>(&sc[warp * kHeadDim + 4 * lane]) = pv;
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) acc += sc[w * kHeadDim + d];
    if (nsplit == 1) {
        const float o = acc * inv;
        A.out[(size_t)m * A.out_ld + hd * kHeadDim + d] = o;
        if (A.out16 != nullptr) A.out16[(size_t)m * A.out_ld + hd * kHeadDim + d] = __float2half_rn(o);
        if (A.amax != nullptr) {
            const float wm = fused2::wmax(fabsf(o));
            if (lane == 0) red[warp] = wm;  // (last read of red was before two barriers)
            __syncthreads();
            if (d == 0) A.amax[m * A.n_heads + hd] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
        }
        return;
    }
    // ---- split-KV: publish this slice's (max, sum, numerator); the last slice to arrive merges all of them in slice order
    float* rec = A.part + (((size_t)m * A.n_heads + hd) * nsplit + zi) * (kHeadDim + 2);
    rec[2 + d] = acc;
    if (d == 0) { rec[0] = j_hi > j_lo ? mx : -INFINITY; rec[1] = j_hi > j_lo ? total : 0.f; }
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (d == 0) s_last = atomicAdd(&A.tickets[m * A.n_heads + hd], 1) == nsplit - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = A.part + ((size_t)m * A.n_heads + hd) * nsplit * (kHeadDim + 2);
    float gm = -INFINITY;
    for (int z = 0; z < nsplit; ++z) gm = fmaxf(gm, __ldcg(base + z * (kHeadDim + 2)));
    float num = 0.f, den = 0.f;
    for (int z = 0; z < nsplit; ++z) {
        const float mz = __ldcg(base + z * (kHeadDim + 2));
        const float f = mz == -INFINITY ? 0.f : expf(mz - gm);
        den += __ldcg(base + z * (kHeadDim + 2) + 1) * f;
        num += __ldcg(base + z * (kHeadDim + 2) + 2 + d) * f;
    }
    const float o = num / den;
    A.out[(size_t)m * A.out_ld + hd * kHeadDim + d] = o;
    if (A.out16 != nullptr) A.out16[(size_t)m * A.out_ld + hd * kHeadDim + d] = __float2half_rn(o);
    if (A.amax != nullptr) {
        const float wm = fused2::wmax(fabsf(o));
        if (lane == 0) red[warp] = wm;
        __syncthreads();
        if (d == 0) A.amax[m * A.n_heads + hd] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    }
    if (d == 0) A.tickets[m * A.n_heads + hd] = 0;
}

// ---- lm_head: logits[m][v] = sum_k W[v][k] * x[m][k], fp16 weights, fp32 accumulate (:1610-1611) ----
template <int MT>
__global__ void __launch_bounds__(256) lm_head_kernel(const __half* __restrict__ W, const __half* __restrict__ x,
                                                      float* __restrict__ logits, int V, int H, int M) {
    extern __shared__ __align__(16) __half xsh[];  // [M][H]
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    for (int i = threadIdx.x; i < M * H / 8; i += 256)
        reinterpret_cast<uint4*>(xsh)[i] = reinterpret_cast<const uint4*>(x)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = blockIdx.x * 8 + warp;
    if (v >= V) return;
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
    const uint4* wr = reinterpret_cast<const uint4*>(W + (size_t)v * H);
    for (int i = lane; i < H / 8; i += 32) {
        uint4 wv;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(wv.x), "=r"(wv.y), "=r"(wv.z), "=r"(wv.w) : "l"(wr + i));
        const __half2* w2 = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m < M) {
                const uint4 xv = reinterpret_cast<const uint4*>(xsh + (size_t)m * H)[i];
                const __half2* x2 = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 a = __half22float2(w2[q]), b = __half22float2(x2[q]);
                    acc[m] += a.x * b.x + a.y * b.y;
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const float r = warp_sum(acc[m]);
        if (lane == 0 && m < M) logits[(size_t)m * V + v] = r;
    }
}

// greedy argmax (first maximum wins, like torch.argmax on ties in practice) + advance ids / positions
__global__ void __launch_bounds__(1024) argmax_advance_kernel(const float* __restrict__ logits, int V,
                                                             long long* __restrict__ ids, int* __restrict__ pos) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    const int m = blockIdx.x;
    const float* row = logits + (size_t)m * V;
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int v = threadIdx.x; v < V; v += 1024) {
        const float x = row[v];
        if (x > best) { best = x; idx = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
        ids[m] = idx;
        pos[m] += 1;
    }
}

__global__ void copy_ids_kernel(const long long* src, long long* dst, int n) {
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
}

// Sum the per-CTA (sum, sumsq) partials of `nproj` projections into [nproj][M][2] floats (fixed order) so that a
// tensor-parallel run can all-reduce a handful of floats; consumers then read them as a single "CTA".
__global__ void __launch_bounds__(256) reduce_stats_kernel(const float* __restrict__ stats, int proj_stride, int ncta,
                                                          int M, float* __restrict__ out) {
    __shared__ double shd[33 * 2];
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    const int p = blockIdx.x / M, m = blockIdx.x % M;
    double st[2] = {0.0, 0.0};
    for (int c = threadIdx.x; c < ncta; c += blockDim.x) {
        const float2 v = *reinterpret_cast<const float2*>(stats + (size_t)p * proj_stride + ((size_t)c * M + m) * 2);
        st[0] += (double)v.x;
        st[1] += (double)v.y;
    }
    block_reduce_sum<2>(st, shd);
    if (threadIdx.x == 0) *reinterpret_cast<float2*>(out + ((size_t)p * M + m) * 2) = make_float2((float)st[0], (float)st[1]);
}

// Batched decode on the tcgen05 path (M > 8): sum the split-K partial outputs of up to 3 projections in place
// (t[0] += t[1..S-1]) and emit each token's LayerNorm (sum, sum of squares) in the "single CTA of partials" format
// the glue / attention kernels read ([projection][M][2] floats). One CTA per (token, projection), fixed order.
constexpr int kReduceSlices = 4;  // column slices per (token, projection): the statistics come out as 4 partial records
struct ReduceArgs {
    float* t[3];
    int N[3];
    float* stats;  // [nprob][kReduceSlices][M][2]
    int M, S;
};
__global__ void __launch_bounds__(256) tc5_reduce_stats_kernel(const __grid_constant__ ReduceArgs A) {
    __shared__ double shd[33 * 2];
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    const int m = blockIdx.x, p = blockIdx.y, z = blockIdx.z, N = A.N[p];
    const int n4 = N >> 2, lo = (int)(((long long)n4 * z) / kReduceSlices), hi = (int)(((long long)n4 * (z + 1)) / kReduceSlices);
    float* base = A.t[p] + (size_t)m * N;
    const size_t split_stride = (size_t)A.M * N;
    double st[2] = {0.0, 0.0};
    for (int i4 = lo + threadIdx.x; i4 < hi; i4 += 256) {
        float4 v = reinterpret_cast<const float4*>(base)[i4];
        for (int zz = 1; zz < A.S; ++zz) {
            const float4 w = reinterpret_cast<const float4*>(base + zz * split_stride)[i4];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        if (A.S > 1) reinterpret_cast<float4*>(base)[i4] = v;
        st[0] += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
        st[1] += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    block_reduce_sum<2>(st, shd);
    if (threadIdx.x == 0)
        *reinterpret_cast<float2*>(A.stats + (((size_t)p * kReduceSlices + z) * A.M + m) * 2) = make_float2((float)st[0], (float)st[1]);
}

// prompt pass helpers: last-token rows of the normalised stream -> [B][H] for lm_head; positions for the decode steps after
__global__ void gather_last_rows_kernel(const __half* __restrict__ x, __half* __restrict__ out, int T, int H) {
    const int b = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)b * T + T - 1) * H);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)b * H);
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
}
__global__ void set_positions_kernel(int* pos, int value, int n) {
    if ((int)threadIdx.x < n) pos[threadIdx.x] = value;
}

template <typename... KArgs, typename... Args>
int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    ONEBIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args...));
    return ONEBIT_OK;
}

}  // namespace
}  // namespace onebit

using namespace onebit;

struct onebit_decoder {
    onebit_decoder_config cfg;
    std::vector<onebit_layer_params> layers;
    const __half* embed = nullptr;
    const void* final_norm = nullptr;
    const __half* lm_head = nullptr;
    const float* rope_cos = nullptr;
    const float* rope_sin = nullptr;
    onebit_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    // device state / scratch (one cudaMalloc arena)
    char* arena = nullptr;
    float* resid[2] = {nullptr, nullptr};
    float *t_qkv = nullptr, *t_o = nullptr, *t_gu = nullptr, *t_d = nullptr, *attn_out = nullptr, *logits = nullptr;
    float *st_qkv = nullptr, *st_o = nullptr, *st_gu = nullptr, *st_d = nullptr;
    uint8_t *dg_qkv = nullptr, *dg_o = nullptr, *dg_gu = nullptr, *dg_d = nullptr;
    imma::QMeta *qm_qkv = nullptr, *qm_o = nullptr, *qm_gu = nullptr, *qm_d = nullptr;
    __half *kcache = nullptr, *vcache = nullptr, *x_f16 = nullptr;
    long long* ids = nullptr;
    long long* ids_stage = nullptr;
    int* pos = nullptr;
    int launches = 0;
    // tensor-parallel geometry (tp = 1: Hl = Hk = H, Il = Ik = I)
    int tp = 1, Hl = 0, Hk = 0, Il = 0, Ik = 0, heads_l = 0;
    float *red_qkv = nullptr, *red_gu = nullptr;  // [nproj][max_batch][2] all-reduced (sum, sumsq)
    // batched decode on the tcgen05 path (max_batch > 8): split-K factor the t_* buffers are sized for, fp16 copies of the
    // input_factor vectors (the MMA A operand is fp16), fp16 activations of the widest layer, per-token statistics
    int ksplit_max = 1;
    std::vector<const __half*> h16;  // [L][7]: q k v o gate up down
    __half* h16_store = nullptr;
    __half* xI_f16 = nullptr;        // [B][I]
    float *red_o = nullptr, *red_d = nullptr;  // [max_batch][2]
    int tc5_ks_d = 1;  // split-K factor of the last down_proj launch (single GPU: its partials are summed by the consuming glue)
    // split-KV decode attention: slices per (sequence, head) (fixed by max_seq_len so that a captured graph stays valid),
    // partial records, merge tickets
    int attn_nsplit = 1;
    float* attn_part = nullptr;
    int* attn_tickets = nullptr;
    // persistent single-kernel step (persist_step.cu): batch <= 2, no tensor parallelism
    PersistState* persist = nullptr;
    // prompt pass (onebit_decoder_prefill): workspace for `pf_cap` tokens, allocated on first use
    char* pf_ws = nullptr;
    size_t pf_cap = 0;
    std::vector<const __half*> pf_h16;  // fp16 input_factor copies (shared with the batched decode path when it has them)
    __half* pf_h16_store = nullptr;
    // one-shot all-reduce over NVLink peer memory (p2p_allreduce.cu); enabled by onebit_decoder_enable_p2p_allreduce
    bool p2p_on = false;
    P2PComm p2p = {};
    unsigned* p2p_state = nullptr;  // call counter, CTA ticket, error flag, last counts
    // host-side upper bound of every sequence's position (reset value + steps enqueued): a step that could write past
    // max_seq_len is refused instead of overrunning the KV cache
    int pos_hi = 0;
    // second-generation fused stages (fused_gemv2.cuh; tp == 1): fp32 side tables in quantiser-item order
    // [L][7] (q k v o gate up down): static input factors (RMSNorm weight folded in for q/k/v/gate/up), weight scales
    bool v2 = false;
    float* tab_store = nullptr;
    std::vector<const float*> tab_fp, tab_g32;
    std::vector<float> tab_fmax;
    float *ext_qkv = nullptr, *ext_o = nullptr, *ext_gu = nullptr, *ext_d = nullptr;  // extension records of the producers
    float* attn_amax = nullptr;  // [max_batch][heads]
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename TP, int NV4>
int glue_launch_inst(GlueArgs& g, cudaStream_t s) {
    auto kern = glue_kernel<TP, NV4>;
    static bool configured[64] = {false};
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured[dev] = true;
    }
    return launch_pdl(kern, dim3(g.M, g.nprob), dim3(kGlueThreads), (size_t)g.K * sizeof(float), s, g);
}

int glue_launch(const onebit_decoder* D, GlueArgs& g, cudaStream_t s) {
    const int nv4 = (g.K / 4 + kGlueThreads - 1) / kGlueThreads;
    return dispatch_dtype(D->cfg.param_dtype, [&](auto pt) {
        using TP = decltype(pt);
        if (nv4 <= 2) return glue_launch_inst<TP, 2>(g, s);
        if (nv4 <= 3) return glue_launch_inst<TP, 3>(g, s);
        if (nv4 <= 6) return glue_launch_inst<TP, 6>(g, s);
        return glue_launch_inst<TP, 7>(g, s);
    });
}

size_t attn_smem_bytes(int max_seq) {
    static bool configured[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        configured[dev] = true;
    }
    return (size_t)std::max(max_seq, 4 * kHeadDim) * sizeof(float) + (size_t)2 * std::min(max_seq, 384) * kHeadDim * 2;
}

// split-KV set-up shared by every decode path: fills the slice fields of `at` and returns the dynamic shared memory to launch with
size_t attn_finish_args(const onebit_decoder* D, AttnArgs& at, int M, bool many_ctas) {
    at.nsplit = D->attn_nsplit;
    at.part = D->attn_part;
    at.tickets = D->attn_tickets;
    if (at.nsplit > 1 || many_ctas) {  // slices (or many small CTAs) stream the cached rows from L2: no shared-memory staging
        at.t_cap = 0;
        return (size_t)std::max(D->cfg.max_seq_len, 4 * kHeadDim) * sizeof(float);
    }
    at.t_cap = std::min(D->cfg.max_seq_len, 384);
    (void)M;
    return attn_smem_bytes(D->cfg.max_seq_len);
}

// ---- fused glue + GEMV stage (fused_gemv.cuh) -----------------------------------------------------------
bool fused_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ONEBIT_FUSED");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// rows per CTA so that the whole launch is about one wave of fat CTAs; returns 0 if this shape cannot use the stage
int fused_rows_per_cta(int total_rows, int M, int K) {
    const int sms = num_sms();
    int rows = ((total_rows + sms - 1) / sms + 31) / 32 * 32;
    if (rows < 32) rows = 32;
    if (rows > 192) rows = 192;
    if (M > 2) return 0;
    while (rows > 32 && fused::smem_bytes(M, K, rows) > 224 * 1024) rows -= 32;  // large K: fewer rows, a second wave
    if (fused::smem_bytes(M, K, rows) > 224 * 1024) return 0;
    return rows;
}

template <typename TP, int TILES, int NV4>
int fused_launch_inst2(const fused::Args& a, int ctas, cudaStream_t s) {
    auto kern = fused::fused_gemv_kernel<TP, 1, TILES, NV4>;
    static bool configured[64] = {false};
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        configured[dev] = true;
    }
    return launch_pdl(kern, dim3(ctas), dim3(fused::kThreads), fused::smem_bytes(a.M, a.K, a.rows_per_cta), s, a);
}

template <typename TP, int TILES>
int fused_launch_inst(const fused::Args& a, int ctas, cudaStream_t s) {
    const int nv4 = (a.K / 4 + fused::kThreads - 1) / fused::kThreads;
    if (nv4 <= 2) return fused_launch_inst2<TP, TILES, 2>(a, ctas, s);
    if (nv4 <= 3) return fused_launch_inst2<TP, TILES, 3>(a, ctas, s);
    if (nv4 <= 6) return fused_launch_inst2<TP, TILES, 6>(a, ctas, s);
    return fused_launch_inst2<TP, TILES, 7>(a, ctas, s);
}

// fills cta_begin, launches; returns the number of CTAs of problem 0 through *ctas_per_problem
int fused_launch(fused::Args a, int param_dtype, cudaStream_t s, int* ctas_per_problem) {
    int ctas = 0;
    for (int i = 0; i < a.nprob; ++i) {
        a.p[i].cta_begin = ctas;
        ctas += (a.p[i].n_rows + a.rows_per_cta - 1) / a.rows_per_cta;
    }
    *ctas_per_problem = (a.p[0].n_rows + a.rows_per_cta - 1) / a.rows_per_cta;
    return dispatch_dtype(param_dtype, [&](auto pt) {
        using TP = decltype(pt);
        switch (a.rows_per_cta / 16) {
            case 2: return fused_launch_inst<TP, 2>(a, ctas, s);
            case 4: return fused_launch_inst<TP, 4>(a, ctas, s);
            case 6: return fused_launch_inst<TP, 6>(a, ctas, s);
            case 8: return fused_launch_inst<TP, 8>(a, ctas, s);
            case 10: return fused_launch_inst<TP, 10>(a, ctas, s);
            case 12: return fused_launch_inst<TP, 12>(a, ctas, s);
            default: return fail(ONEBIT_ERR_INVALID_ARGUMENT, "fused stage: unsupported rows per CTA");
        }
    });
}

int launch_attn(dim3 grid, dim3 block, size_t smem, cudaStream_t s, AttnArgs at) {
    at.inv_nln = 1.0 / (double)at.n_ln;
    return launch_pdl(attn_kernel, grid, block, smem, s, at);
}

int allreduce(const onebit_decoder* D, float* data, int64_t count, cudaStream_t s) {
    if (D->tp <= 1) return ONEBIT_OK;
    if (D->p2p_on) return p2p_allreduce(D->p2p, data, count, s);
    if (!D->allreduce) return fail(ONEBIT_ERR_INVALID_ARGUMENT, "tensor-parallel decoder without an all-reduce callback");
    const int rc = D->allreduce(D->allreduce_user, data, count, s);
    return rc == 0 ? ONEBIT_OK : fail(ONEBIT_ERR_CUDA, "all-reduce callback failed with code " + std::to_string(rc));
}

// The fused-stage layer loop (5 launches per layer; + 4 small collectives per layer under tensor parallelism:
// (sum,sumsq) of q/k/v, partial sums of o_proj, (sum,sumsq) of gate/up, partial sums of down_proj — SURVEY.md §8e).
// Returns false through *ok if this shape cannot use the fused stages.
int run_fused_layers(onebit_decoder* D, int M, cudaStream_t s, bool with_attention, bool first_is_embed, int* launches,
                     int* cur_io, int* nc_d_io, bool* ok) {
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, pd = C.param_dtype, B = C.max_batch;
    const int Hl = D->Hl, Hk = D->Hk, Il = D->Il, Ik = D->Ik, tp = D->tp;
    const int uH = H / imma::kUnitCols, uHk = Hk / imma::kUnitCols, uIk = Ik / imma::kUnitCols;
    const int cHl = (Hl + imma::kRows - 1) / imma::kRows, cIl = (Il + imma::kRows - 1) / imma::kRows;
    const int rq = fused_rows_per_cta(3 * Hl, M, H), ro = fused_rows_per_cta(H, M, Hk);
    const int rg = fused_rows_per_cta(2 * Il, M, H), rd = fused_rows_per_cta(H, M, Ik);
    *ok = fused_enabled() && rq && ro && rg && rd;
    if (!*ok) return ONEBIT_OK;
    int cur = *cur_io, nc_d = *nc_d_io, rc;
    for (int l = 0; l < C.num_layers; ++l) {
        const onebit_layer_params& P = D->layers[l];
        int nc_q = 0, nc_o = 0, nc_g = 0;
        // ---- stage 1: (embed | resid + LN(down)) -> RMSNorm -> q,k,v (column-parallel: local rows)
        fused::Args f = {};
        f.nprob = 3; f.M = M; f.K = H; f.units = uH; f.rows_per_cta = rq;
        f.mode = (l == 0 && first_is_embed) ? fused::EMBED_NORM : fused::RESID_NORM;
        f.t_a = D->t_d; f.stats_a = D->st_d; f.ncta_a = nc_d; f.stats_from_data = tp > 1;
        f.resid_in = D->resid[cur]; f.resid_out = D->resid[cur ^ 1];
        f.embed = D->embed; f.ids = D->ids; f.ln_w = P.input_layernorm; f.ln_eps = C.ln_eps; f.rms_eps = C.rms_eps;
        const onebit_bitlinear_params* qkv[3] = {&P.q, &P.k, &P.v};
        for (int i = 0; i < 3; ++i) {
            f.p[i].w = reinterpret_cast<const uint8_t*>(qkv[i]->weight); f.p[i].g = qkv[i]->weight_scale;
            f.p[i].h = qkv[i]->input_factor; f.p[i].t = D->t_qkv + (size_t)i * B * Hl;
            f.p[i].stats = D->st_qkv + (size_t)i * cHl * B * 2; f.p[i].n_rows = Hl; f.p[i].ld_t = Hl;
        }
        rc = fused_launch(f, pd, s, &nc_q); if (rc) return rc; ++*launches;
        cur ^= 1;
        const float* sq = f.p[0].stats; const float* sk = f.p[1].stats; const float* sv = f.p[2].stats;
        int nc_attn = nc_q;
        if (tp > 1) {  // LayerNorm spans the full N: all-reduce 3 x (sum, sumsq) per token
            rc = launch_pdl(reduce_stats_kernel, dim3(3 * M), dim3(256), 0, s, (const float*)D->st_qkv, cHl * B * 2, nc_q, M,
                            D->red_qkv);
            if (rc) return rc; ++*launches;
            rc = allreduce(D, D->red_qkv, (int64_t)3 * M * 2, s); if (rc) return rc;
            sq = D->red_qkv; sk = D->red_qkv + (size_t)M * 2; sv = D->red_qkv + (size_t)2 * M * 2; nc_attn = 1;
        }
        // ---- stage 2: attention over the local heads
        if (with_attention) {
            AttnArgs at = {};
            at.t_q = f.p[0].t; at.t_k = f.p[1].t; at.t_v = f.p[2].t;
            at.stats_q = sq; at.stats_k = sk; at.stats_v = sv; at.ncta = nc_attn;
            at.M = M; at.H = Hl; at.n_ln = H; at.out_ld = Hk; at.n_heads = D->heads_l; at.max_seq = C.max_seq_len;
            at.pos = D->pos; at.rope_cos = D->rope_cos; at.rope_sin = D->rope_sin;
            const size_t layer_cache = (size_t)B * D->heads_l * C.max_seq_len * kHeadDim;
            at.kcache = D->kcache + l * layer_cache; at.vcache = D->vcache + l * layer_cache;
            at.out = D->attn_out; at.ln_eps = C.ln_eps;
            const size_t asmem = attn_finish_args(D, at, M, false);
            rc = launch_attn( dim3(M, D->heads_l, at.nsplit), dim3(kHeadDim), asmem, s, at);
            if (rc) return rc; ++*launches;
        }
        // ---- stage 3: attention output -> o_proj (row-parallel: local K slice, zero-padded to a multiple of 256)
        f = {};
        f.nprob = 1; f.M = M; f.K = Hk; f.units = uHk; f.rows_per_cta = ro; f.mode = fused::PLAIN; f.x_plain = D->attn_out;
        f.p[0].w = reinterpret_cast<const uint8_t*>(P.o.weight); f.p[0].g = P.o.weight_scale; f.p[0].h = P.o.input_factor;
        f.p[0].t = D->t_o; f.p[0].stats = D->st_o; f.p[0].n_rows = H; f.p[0].ld_t = H;
        rc = fused_launch(f, pd, s, &nc_o); if (rc) return rc; ++*launches;
        rc = allreduce(D, D->t_o, (int64_t)M * H, s); if (rc) return rc;  // partial sums of the K shards
        // ---- stage 4: resid + LN(o) -> RMSNorm -> gate, up (column-parallel)
        f = {};
        f.nprob = 2; f.M = M; f.K = H; f.units = uH; f.rows_per_cta = rg; f.mode = fused::RESID_NORM;
        f.t_a = D->t_o; f.stats_a = D->st_o; f.ncta_a = nc_o; f.stats_from_data = tp > 1;
        f.resid_in = D->resid[cur]; f.resid_out = D->resid[cur ^ 1]; f.ln_w = P.post_attention_layernorm;
        f.ln_eps = C.ln_eps; f.rms_eps = C.rms_eps;
        const onebit_bitlinear_params* gu[2] = {&P.gate, &P.up};
        for (int i = 0; i < 2; ++i) {
            f.p[i].w = reinterpret_cast<const uint8_t*>(gu[i]->weight); f.p[i].g = gu[i]->weight_scale;
            f.p[i].h = gu[i]->input_factor; f.p[i].t = D->t_gu + (size_t)i * B * Ik;
            f.p[i].stats = D->st_gu + (size_t)i * cIl * B * 2; f.p[i].n_rows = Il; f.p[i].ld_t = Ik;
        }
        rc = fused_launch(f, pd, s, &nc_g); if (rc) return rc; ++*launches;
        cur ^= 1;
        const float* sg = f.p[0].stats; const float* su = f.p[1].stats;
        int nc_gu = nc_g;
        if (tp > 1) {
            rc = launch_pdl(reduce_stats_kernel, dim3(2 * M), dim3(256), 0, s, (const float*)D->st_gu, cIl * B * 2, nc_g, M,
                            D->red_gu);
            if (rc) return rc; ++*launches;
            rc = allreduce(D, D->red_gu, (int64_t)2 * M * 2, s); if (rc) return rc;
            sg = D->red_gu; su = D->red_gu + (size_t)M * 2; nc_gu = 1;
        }
        // ---- stage 5: silu(LN(gate)) * LN(up) -> down_proj (row-parallel)
        fused::Args f5 = {};
        f5.nprob = 1; f5.M = M; f5.K = Ik; f5.units = uIk; f5.rows_per_cta = rd; f5.mode = fused::SILU_MUL;
        f5.t_a = f.p[0].t; f5.stats_a = sg; f5.ncta_a = nc_gu; f5.t_b = f.p[1].t; f5.stats_b = su; f5.ncta_b = nc_gu;
        f5.ln_eps = C.ln_eps; f5.n_ln = I;
        f5.p[0].w = reinterpret_cast<const uint8_t*>(P.down.weight); f5.p[0].g = P.down.weight_scale;
        f5.p[0].h = P.down.input_factor; f5.p[0].t = D->t_d; f5.p[0].stats = D->st_d; f5.p[0].n_rows = H; f5.p[0].ld_t = H;
        rc = fused_launch(f5, pd, s, &nc_d); if (rc) return rc; ++*launches;
        rc = allreduce(D, D->t_d, (int64_t)M * H, s); if (rc) return rc;
    }
    *cur_io = cur;
    *nc_d_io = nc_d;
    return ONEBIT_OK;
}

// ---- second-generation fused stage (fused_gemv2.cuh) -------------------------------------------------------
bool fused2_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ONEBIT_FUSED_V2");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

int fused2_rows_per_cta(int total_rows, int M, int K) {
    const int sms = num_sms();
    int rows = ((total_rows + sms - 1) / sms + 31) / 32 * 32;
    if (rows < 32) rows = 32;
    if (rows > 192) rows = 192;
    if (M > 2) return 0;
    while (rows > 32 && fused2::smem_bytes(M, K, rows) > 224 * 1024) rows -= 32;
    if (fused2::smem_bytes(M, K, rows) > 224 * 1024) return 0;
    return rows;
}

template <int TILES>
int fused2_launch_inst(const fused2::Args& a, int ctas, cudaStream_t s) {
    auto kern = fused2::fused_gemv2_kernel<TILES>;
    static bool configured[64] = {false};
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        configured[dev] = true;
    }
    return launch_pdl(kern, dim3(ctas), dim3(fused2::kThreads), fused2::smem_bytes(a.M, a.K, a.rows_per_cta), s, a);
}

int fused2_launch(fused2::Args a, int ext_rep_stride, cudaStream_t s, int* ctas_per_problem) {
    a.ext_rep_stride = ext_rep_stride;
    a.inv_k = 1.0 / (double)a.K;
    a.inv_nln = 1.0 / (double)(a.n_ln > 0 ? a.n_ln : a.K);
    int ctas = 0;
    for (int i = 0; i < a.nprob; ++i) {
        a.p[i].cta_begin = ctas;
        ctas += (a.p[i].n_rows + a.rows_per_cta - 1) / a.rows_per_cta;
    }
    *ctas_per_problem = (a.p[0].n_rows + a.rows_per_cta - 1) / a.rows_per_cta;
    switch (a.rows_per_cta / 16) {
        case 2: return fused2_launch_inst<2>(a, ctas, s);
        case 4: return fused2_launch_inst<4>(a, ctas, s);
        case 6: return fused2_launch_inst<6>(a, ctas, s);
        case 8: return fused2_launch_inst<8>(a, ctas, s);
        case 10: return fused2_launch_inst<10>(a, ctas, s);
        case 12: return fused2_launch_inst<12>(a, ctas, s);
        default: return fail(ONEBIT_ERR_INVALID_ARGUMENT, "fused stage: unsupported rows per CTA");
    }
}

int build_v2_tables(onebit_decoder* D) {
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, L = C.num_layers;
    const size_t per_layer = (size_t)(3 * H + H + 2 * H + I) + (size_t)(3 * H + H + 2 * I + H);
    ONEBIT_CUDA_TRY(cudaMalloc(&D->tab_store, (per_layer * L + (size_t)7 * L) * sizeof(float)));
    float* fm_dev = D->tab_store + per_layer * L;
    D->tab_fp.resize((size_t)L * 7);
    D->tab_g32.resize((size_t)L * 7);
    D->tab_fmax.resize((size_t)L * 7);
    float* p = D->tab_store;
    const int rc = dispatch_dtype(C.param_dtype, [&](auto pt) {
        using TP = decltype(pt);
        for (int l = 0; l < L; ++l) {
            const onebit_layer_params& P = D->layers[l];
            const onebit_bitlinear_params* bl[7] = {&P.q, &P.k, &P.v, &P.o, &P.gate, &P.up, &P.down};
            fused2::TableJobs J = {};
            for (int i = 0; i < 7; ++i) {
                fused2::TableJob& T = J.j[i];
                T.k = i == 6 ? I : H;
                T.n = (i == 4 || i == 5) ? I : H;
                T.h = bl[i]->input_factor;
                T.lnw = i < 3 ? P.input_layernorm : ((i == 4 || i == 5) ? P.post_attention_layernorm : nullptr);
                T.g = bl[i]->weight_scale;
                T.fp = p; p += T.k;
                T.g32 = p; p += T.n;
                T.fmax = fm_dev + (size_t)l * 7 + i;
                D->tab_fp[(size_t)l * 7 + i] = T.fp;
                D->tab_g32[(size_t)l * 7 + i] = T.g32;
            }
            fused2::side_tables_kernel<TP><<<7, 1024>>>(J);
        }
        return (int)ONEBIT_OK;
    });
    if (rc) return rc;
    ONEBIT_CUDA_TRY(cudaGetLastError());
    ONEBIT_CUDA_TRY(cudaMemcpy(D->tab_fmax.data(), fm_dev, (size_t)7 * L * sizeof(float), cudaMemcpyDeviceToHost));
    return ONEBIT_OK;
}

// The layer loop on the second-generation stages (tp == 1): same 5 launches per layer as run_fused_layers, but every
// stage's prologue takes its scalars from the producer's records (fused_gemv2.cuh).
int run_fused2_layers(onebit_decoder* D, int M, cudaStream_t s, bool with_attention, bool first_is_embed, int* launches,
                      int* cur_io, int* nc_d_io, bool* ok) {
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, B = C.max_batch;
    const int uH = H / imma::kUnitCols, uI = I / imma::kUnitCols;
    const int cH = (H + imma::kRows - 1) / imma::kRows, cI = (I + imma::kRows - 1) / imma::kRows;
    const int rq = fused2_rows_per_cta(3 * H, M, H), ro = fused2_rows_per_cta(H, M, H);
    const int rg = fused2_rows_per_cta(2 * I, M, H), rd = fused2_rows_per_cta(H, M, I);
    *ok = D->v2 && fused_enabled() && fused2_enabled() && rq && ro && rg && rd;
    if (!*ok) return ONEBIT_OK;
    int cur = *cur_io, nc_d = *nc_d_io, rc;
    const char* tre = getenv("ONEBIT_TRACE_STAGE");  // (trace builds) which stage of the last layer records its clocks: 1, 3, 4, 5
    const int trace_stage = tre ? atoi(tre) : 5;
    const int ext_rep = (5 * cH + 2 * cI) * B * fused2::kExt;  // floats of one replica of the record block
    auto set_problem = [&](fused2::Problem& p, int l, int i, const onebit_bitlinear_params& bp, float* t, float* stats, float* ext,
                           int n_rows) {
        p.w = reinterpret_cast<const uint8_t*>(bp.weight);
        p.g32 = D->tab_g32[(size_t)l * 7 + i];
        p.fp = D->tab_fp[(size_t)l * 7 + i];
        p.fmax = D->tab_fmax[(size_t)l * 7 + i];
        p.t = t; p.stats = stats; p.ext = ext; p.n_rows = n_rows; p.ld_t = n_rows;
    };
    for (int l = 0; l < C.num_layers; ++l) {
        const onebit_layer_params& P = D->layers[l];
        int nc_q = 0, nc_o = 0, nc_g = 0;
        // ---- stage 1: (embed | resid + LN(down)) -> RMSNorm -> q,k,v
        fused2::Args f = {};
        f.nprob = 3; f.M = M; f.K = H; f.units = uH; f.rows_per_cta = rq;
        f.mode = (l == 0 && first_is_embed) ? fused::EMBED_NORM : fused::RESID_NORM;
        f.t_a = D->t_d; f.stats_a = D->st_d; f.ext_a = D->ext_d; f.ncta_a = nc_d;
        f.resid_in = D->resid[cur]; f.resid_out = D->resid[cur ^ 1];
        f.embed = D->embed; f.ids = D->ids; f.ln_eps = C.ln_eps; f.rms_eps = C.rms_eps;
        const onebit_bitlinear_params* qkv[3] = {&P.q, &P.k, &P.v};
        for (int i = 0; i < 3; ++i)
            set_problem(f.p[i], l, i, *qkv[i], D->t_qkv + (size_t)i * B * H, D->st_qkv + (size_t)i * cH * B * 2,
                        D->ext_qkv + (size_t)i * cH * B * fused2::kExt, H);
        f.ext_stride_in = f.ext_stride_out = cH * B * 4;
        f.a_perm = 1; f.rin_perm = 1; f.rout_perm = 1; f.t_perm = 0;  // q/k/v feed the attention kernel: natural order
        f.trace = l == C.num_layers - 1 && trace_stage == 1;
        rc = fused2_launch(f, ext_rep, s, &nc_q); if (rc) return rc; ++*launches;
        cur ^= 1;
        // ---- stage 2: attention (also records max |out| per head for the o_proj quantiser)
        if (with_attention) {
            AttnArgs at = {};
            at.t_q = f.p[0].t; at.t_k = f.p[1].t; at.t_v = f.p[2].t;
            at.stats_q = f.p[0].stats; at.stats_k = f.p[1].stats; at.stats_v = f.p[2].stats; at.ncta = nc_q;
            at.M = M; at.H = H; at.n_ln = H; at.out_ld = H; at.n_heads = C.num_heads; at.max_seq = C.max_seq_len;
            at.pos = D->pos; at.rope_cos = D->rope_cos; at.rope_sin = D->rope_sin;
            const size_t layer_cache = (size_t)B * C.num_heads * C.max_seq_len * kHeadDim;
            at.kcache = D->kcache + l * layer_cache; at.vcache = D->vcache + l * layer_cache;
            at.out = D->attn_out; at.ln_eps = C.ln_eps; at.amax = D->attn_amax;
            const size_t asmem = attn_finish_args(D, at, M, false);
            rc = launch_attn( dim3(M, C.num_heads, at.nsplit), dim3(kHeadDim), asmem, s, at);
            if (rc) return rc; ++*launches;
        }
        // ---- stage 3: attention output -> o_proj; its records carry the residual terms for stage 4
        f = {};
        f.nprob = 1; f.M = M; f.K = H; f.units = uH; f.rows_per_cta = ro; f.mode = fused::PLAIN; f.x_plain = D->attn_out;
        // (without attention — the projection-only timing chain — the records are those of the last real step: stale but valid)
        f.x_amax = D->attn_amax; f.n_amax = C.num_heads;
        f.resid_next = D->resid[cur]; f.resid_ld = H; f.ext_stride_out = cH * B * 4;
        set_problem(f.p[0], l, 3, P.o, D->t_o, D->st_o, D->ext_o, H);
        f.rnext_perm = 1; f.t_perm = 1;
        f.trace = l == C.num_layers - 1 && trace_stage == 3;
        rc = fused2_launch(f, ext_rep, s, &nc_o); if (rc) return rc; ++*launches;
        // ---- stage 4: resid + LN(o) -> RMSNorm -> gate, up
        f = {};
        f.nprob = 2; f.M = M; f.K = H; f.units = uH; f.rows_per_cta = rg; f.mode = fused::RESID_NORM;
        f.t_a = D->t_o; f.stats_a = D->st_o; f.ext_a = D->ext_o; f.ncta_a = nc_o;
        f.resid_in = D->resid[cur]; f.resid_out = D->resid[cur ^ 1]; f.ln_eps = C.ln_eps; f.rms_eps = C.rms_eps;
        const onebit_bitlinear_params* gu[2] = {&P.gate, &P.up};
        for (int i = 0; i < 2; ++i)
            set_problem(f.p[i], l, 4 + i, *gu[i], D->t_gu + (size_t)i * B * I, D->st_gu + (size_t)i * cI * B * 2,
                        D->ext_gu + (size_t)i * cI * B * fused2::kExt, I);
        f.ext_stride_in = cH * B * 4; f.ext_stride_out = cI * B * 4;
        const int last = l == C.num_layers - 1;  // the final glue kernel reads the residual stream and t_d in natural order
        f.a_perm = 1; f.rin_perm = 1; f.rout_perm = !last; f.t_perm = 1;
        f.trace = l == C.num_layers - 1 && trace_stage == 4;
        rc = fused2_launch(f, ext_rep, s, &nc_g); if (rc) return rc; ++*launches;
        cur ^= 1;
        // ---- stage 5: silu(LN(gate)) * LN(up) -> down_proj; records carry the residual terms for the next layer
        fused2::Args f5 = {};
        f5.nprob = 1; f5.M = M; f5.K = I; f5.units = uI; f5.rows_per_cta = rd; f5.mode = fused::SILU_MUL;
        f5.t_a = f.p[0].t; f5.stats_a = f.p[0].stats; f5.ext_a = f.p[0].ext; f5.ncta_a = nc_g;
        f5.t_b = f.p[1].t; f5.stats_b = f.p[1].stats; f5.ext_b = f.p[1].ext; f5.ncta_b = nc_g;
        f5.ln_eps = C.ln_eps; f5.n_ln = I;
        f5.resid_next = D->resid[cur]; f5.resid_ld = H; f5.ext_stride_out = cH * B * 4;
        set_problem(f5.p[0], l, 6, P.down, D->t_d, D->st_d, D->ext_d, H);
        f5.a_perm = 1; f5.rnext_perm = !last; f5.t_perm = !last;
        f5.trace = l == C.num_layers - 1 && trace_stage == 5;
        rc = fused2_launch(f5, ext_rep, s, &nc_d); if (rc) return rc; ++*launches;
    }
    *cur_io = cur;
    *nc_d_io = nc_d;
    return ONEBIT_OK;
}

// split-K factor of a batched-decode projection launch: the decode tile runs 3 CTAs per SM (registers, shared memory);
// fill one wave of those slots, never overshoot it (a second, mostly empty wave costs a whole CTA lifetime), at least 4 K
// chunks per CTA
int tc5_ksplit(int row_tiles, int K, int ksplit_max) {
    const int want = std::max(1, 3 * num_sms() / row_tiles);
    return std::max(1, std::min(std::min(want, ksplit_max), K / 64 / 4));
}

// Batched decode (5..64 sequences per replica): every BitLinear runs on the tcgen05 path (prefill_tc5.cu, decode tile
// configuration: weights as the 128-row UMMA M operand with input_factor folded in, the M tokens as UMMA N, split-K over
// grid.z), the glue kernels hand it fp16 activations, a reduce pass sums the K splits and emits the LayerNorm statistics.
// Per layer: glue | q,k,v (one launch) | reduce | attention | o | glue | gate,up (one launch) | glue | down = 9 launches on a
// single GPU (the glue after o / gate,up / down sums the split-K partials itself and takes the LayerNorm statistics from
// the data; the attention kernel writes o_proj's fp16 activations); tensor-parallel shards keep a reduce pass before each
// all-reduce and a glue that zero-pads the attention output: 13 launches.
int run_tc5_layers(onebit_decoder* D, int M, cudaStream_t s, int* launches, int* cur_io, bool only_proj = false) {
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, pd = C.param_dtype, B = C.max_batch;
    // tensor parallelism (Megatron layout, tp.py): q/k/v and gate/up hold Hl / Il of the rows, o / down a K slice of Hk / Ik
    // columns (zero padded); 4 small all-reduces per layer as on the fused path
    const int Hl = D->Hl, Hk = D->Hk, Il = D->Il, Ik = D->Ik, tp = D->tp;
    int cur = *cur_io, rc;
    auto reduce = [&](float* t0, float* t1, float* t2, int n, int nprob, int S, float* stats) -> int {
        if (only_proj) return ONEBIT_OK;
        ReduceArgs r = {};
        r.t[0] = t0; r.t[1] = t1; r.t[2] = t2; r.N[0] = r.N[1] = r.N[2] = n; r.stats = stats; r.M = M; r.S = S;
        ++*launches;
        return launch_pdl(tc5_reduce_stats_kernel, dim3(M, nprob, kReduceSlices), dim3(256), 0, s, r);
    };
    for (int l = 0; l < C.num_layers; ++l) {
        const onebit_layer_params& P = D->layers[l];
        const __half* const* h16 = &D->h16[(size_t)l * 7];
        // ---- glue 1: (embed | resid + LN(down of the previous layer)) -> RMSNorm -> fp16 x
        GlueArgs g = {};
        g.mode = l == 0 ? GLUE_EMBED_NORM : GLUE_RESID_NORM;
        g.M = M; g.K = H; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = D->x_f16;
        g.t_a = D->t_d; g.stats_a = D->red_d; g.ncta_a = kReduceSlices; g.stats_from_data = 1;
        if (tp == 1 && l > 0) { g.ksplit = D->tc5_ks_d; g.slab = (long long)M * H; }
        g.resid_in = D->resid[cur]; g.resid_out = D->resid[cur ^ 1];
        g.embed = D->embed; g.ids = D->ids; g.ln_w = P.input_layernorm; g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
        if (!only_proj) { rc = glue_launch(D, g, s); if (rc) return rc; ++*launches; }
        cur ^= 1;
        // ---- q, k, v (column-parallel: local rows)
        Tc5Launch t = {};
        t.x16 = D->x_f16; t.M = M; t.K = H; t.nprob = 3; t.param_dtype = pd;
        t.ksplit = tc5_ksplit(3 * ((Hl + 127) / 128), H, D->ksplit_max);
        const onebit_bitlinear_params* qkv[3] = {&P.q, &P.k, &P.v};
        float* tq[3];
        for (int i = 0; i < 3; ++i) {
            tq[i] = D->t_qkv + (size_t)i * B * H * D->ksplit_max;
            t.p[i].w = static_cast<const int8_t*>(qkv[i]->weight); t.p[i].h16 = h16[i]; t.p[i].g = qkv[i]->weight_scale;
            t.p[i].t = tq[i]; t.p[i].N = Hl;
        }
        rc = launch_tc5(t, s); if (rc) return rc; ++*launches;
        rc = reduce(tq[0], tq[1], tq[2], Hl, 3, t.ksplit, D->red_qkv); if (rc) return rc;
        if (!only_proj) { rc = allreduce(D, D->red_qkv, (int64_t)3 * kReduceSlices * M * 2, s); if (rc) return rc; }
        // ---- attention over the local heads
        AttnArgs at = {};
        at.t_q = tq[0]; at.t_k = tq[1]; at.t_v = tq[2];
        const size_t pstride = (size_t)kReduceSlices * M * 2;  // statistics of one projection: kReduceSlices partial records
        at.stats_q = D->red_qkv; at.stats_k = D->red_qkv + pstride; at.stats_v = D->red_qkv + 2 * pstride; at.ncta = kReduceSlices;
        at.M = M; at.H = Hl; at.n_ln = H; at.out_ld = Hk; at.n_heads = D->heads_l; at.max_seq = C.max_seq_len; at.pos = D->pos;
        at.rope_cos = D->rope_cos; at.rope_sin = D->rope_sin;
        const size_t layer_cache = (size_t)B * D->heads_l * C.max_seq_len * kHeadDim;
        at.kcache = D->kcache + l * layer_cache; at.vcache = D->vcache + l * layer_cache;
        // many (sequence, head) CTAs: stream the cached rows from L2 instead of staging them (several CTAs per SM)
        at.out = D->attn_out; at.ln_eps = C.ln_eps;
        // no padding columns (single GPU): the attention kernel writes the fp16 activations of o_proj itself
        const bool attn_writes_x16 = Hk == Hl;
        if (attn_writes_x16) at.out16 = D->x_f16;
        const size_t asmem = attn_finish_args(D, at, M, true);
        if (!only_proj) {
            rc = launch_attn( dim3(M, D->heads_l, at.nsplit), dim3(kHeadDim), asmem, s, at);
            if (rc) return rc; ++*launches;
        }
        // ---- glue 2: attention output [M][Hk] (pad columns stay zero) -> fp16
        g = {};
        g.mode = GLUE_PLAIN; g.M = M; g.K = Hk; g.nprob = 1; g.x_plain = D->attn_out; g.write_x_f16 = 1; g.x_f16 = D->x_f16;
        if (!only_proj && !attn_writes_x16) { rc = glue_launch(D, g, s); if (rc) return rc; ++*launches; }
        // ---- o_proj (row-parallel: K slice) + all-reduce of the partial sums
        t = {};
        t.x16 = D->x_f16; t.M = M; t.K = Hk; t.nprob = 1; t.param_dtype = pd; t.ksplit = tc5_ksplit((H + 127) / 128, Hk, D->ksplit_max);
        t.p[0].w = static_cast<const int8_t*>(P.o.weight); t.p[0].h16 = h16[3]; t.p[0].g = P.o.weight_scale; t.p[0].t = D->t_o; t.p[0].N = H;
        rc = launch_tc5(t, s); if (rc) return rc; ++*launches;
        const int ks_o = t.ksplit;
        if (tp > 1) {  // (single GPU: the consuming glue sums the split-K partials and takes the statistics from the data)
            rc = reduce(D->t_o, nullptr, nullptr, H, 1, t.ksplit, D->red_o); if (rc) return rc;
            if (!only_proj) { rc = allreduce(D, D->t_o, (int64_t)M * H, s); if (rc) return rc; }
        }
        // ---- glue 3: resid + LN(o) -> RMSNorm -> fp16 x (statistics from the all-reduced data under tensor parallelism)
        g = {};
        g.mode = GLUE_RESID_NORM; g.M = M; g.K = H; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = D->x_f16;
        g.t_a = D->t_o; g.stats_a = D->red_o; g.ncta_a = kReduceSlices; g.stats_from_data = 1;
        if (tp == 1) { g.ksplit = ks_o; g.slab = (long long)M * H; }
        g.resid_in = D->resid[cur]; g.resid_out = D->resid[cur ^ 1]; g.ln_w = P.post_attention_layernorm;
        g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
        if (!only_proj) { rc = glue_launch(D, g, s); if (rc) return rc; ++*launches; }
        cur ^= 1;
        // ---- gate, up (column-parallel); outputs laid out with the padded width the down_proj shard consumes
        t = {};
        t.x16 = D->x_f16; t.M = M; t.K = H; t.nprob = 2; t.param_dtype = pd;
        t.ksplit = tc5_ksplit(2 * ((Il + 127) / 128), H, D->ksplit_max);
        const onebit_bitlinear_params* gu[2] = {&P.gate, &P.up};
        float* tg[2];
        for (int i = 0; i < 2; ++i) {
            tg[i] = D->t_gu + (size_t)i * B * I * D->ksplit_max;
            t.p[i].w = static_cast<const int8_t*>(gu[i]->weight); t.p[i].h16 = h16[4 + i]; t.p[i].g = gu[i]->weight_scale;
            t.p[i].t = tg[i]; t.p[i].N = Il; t.p[i].ldt = Ik;
        }
        rc = launch_tc5(t, s); if (rc) return rc; ++*launches;
        const int ks_gu = t.ksplit;
        if (tp > 1) {
            rc = reduce(tg[0], tg[1], nullptr, Ik, 2, t.ksplit, D->red_gu); if (rc) return rc;
            if (!only_proj) { rc = allreduce(D, D->red_gu, (int64_t)2 * kReduceSlices * M * 2, s); if (rc) return rc; }
        }
        // ---- glue 4: silu(LN(gate)) * LN(up) -> fp16 [M][Ik]
        g = {};
        g.mode = GLUE_SILU_MUL; g.M = M; g.K = Ik; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = D->xI_f16; g.n_ln = I;
        g.t_a = tg[0]; g.stats_a = D->red_gu; g.ncta_a = kReduceSlices;
        g.t_b = tg[1]; g.stats_b = D->red_gu + (size_t)kReduceSlices * M * 2; g.ncta_b = kReduceSlices;
        if (tp == 1) { g.ksplit = ks_gu; g.slab = (long long)M * Ik; g.stats_from_data = 1; }
        g.ln_eps = C.ln_eps;
        if (!only_proj) { rc = glue_launch(D, g, s); if (rc) return rc; ++*launches; }
        // ---- down_proj (row-parallel) + all-reduce of the partial sums
        t = {};
        t.x16 = D->xI_f16; t.M = M; t.K = Ik; t.nprob = 1; t.param_dtype = pd; t.ksplit = tc5_ksplit((H + 127) / 128, Ik, D->ksplit_max);
        t.p[0].w = static_cast<const int8_t*>(P.down.weight); t.p[0].h16 = h16[6]; t.p[0].g = P.down.weight_scale; t.p[0].t = D->t_d; t.p[0].N = H;
        rc = launch_tc5(t, s); if (rc) return rc; ++*launches;
        D->tc5_ks_d = t.ksplit;
        if (tp > 1) {
            rc = reduce(D->t_d, nullptr, nullptr, H, 1, t.ksplit, D->red_d); if (rc) return rc;
            if (!only_proj) { rc = allreduce(D, D->t_d, (int64_t)M * H, s); if (rc) return rc; }
        }
    }
    *cur_io = cur;
    return ONEBIT_OK;
}

}  // namespace

extern "C" {

void onebit_decoder_destroy(onebit_decoder* D) {
    if (!D) return;
    persist_destroy(D->persist);
    cudaFree(D->p2p_state);
    cudaFree(D->pf_ws);
    cudaFree(D->pf_h16_store);
    cudaFree(D->tab_store);
    cudaFree(D->arena);
    delete D;
}

int onebit_decoder_create(onebit_decoder** out, const onebit_decoder_config* cfg, const onebit_layer_params* layers,
                          const void* embed_tokens_f16, const void* final_norm, const void* lm_head_f16,
                          const float* rope_cos, const float* rope_sin, onebit_allreduce_fn allreduce,
                          void* allreduce_user) {
    ONEBIT_REQUIRE(out && cfg && layers && embed_tokens_f16 && final_norm && lm_head_f16 && rope_cos && rope_sin,
                   "decoder_create: NULL argument");
    *out = nullptr;
    const int H = cfg->hidden_size, I = cfg->intermediate_size, L = cfg->num_layers, B = cfg->max_batch;
    ONEBIT_REQUIRE(dtype_ok(cfg->param_dtype), "decoder_create: bad param_dtype");
    ONEBIT_REQUIRE(H > 0 && I > 0 && L > 0 && cfg->num_heads > 0 && cfg->vocab_size > 0 && cfg->max_seq_len > 0,
                   "decoder_create: bad sizes");
    ONEBIT_REQUIRE(H == cfg->num_heads * kHeadDim, "decoder_create: only head_dim 128 (LLaMA-7B/13B) is built");
    ONEBIT_REQUIRE(H % imma::kUnitCols == 0 && I % imma::kUnitCols == 0 && I <= 14336 && H <= 14336,
                   "decoder_create: hidden/intermediate size must be multiples of 256 and <= 14336");
    ONEBIT_REQUIRE(B >= 1 && B <= kMaxBatch, "decoder_create: a replica serves batch 1..64 (1..4 on the bit-plane GEMV, 5..64 on the tcgen05 path)");
    const int tp = cfg->tp_size > 1 ? cfg->tp_size : 1;
    ONEBIT_REQUIRE(cfg->num_heads % tp == 0 && I % tp == 0 && (I / tp) % 8 == 0,
                   "decoder_create: heads and intermediate size must divide by tp_size");
    ONEBIT_REQUIRE(tp == 1 || allreduce != nullptr, "decoder_create: tp_size > 1 needs an all-reduce callback");
    ONEBIT_REQUIRE(tp == 1 || (cfg->tp_rank >= 0 && cfg->tp_rank < tp), "decoder_create: bad tp_rank");
    ONEBIT_REQUIRE(H % 8 == 0, "decoder_create: hidden must be a multiple of 8");
    ONEBIT_REQUIRE(cfg->max_seq_len <= 6144, "decoder_create: max_seq_len up to 6144 (attention keeps the score row + 384 cached K/V rows in shared memory)");
    onebit_decoder* D = new (std::nothrow) onebit_decoder();
    ONEBIT_REQUIRE(D, "decoder_create: out of host memory");
    D->cfg = *cfg;
    D->layers.assign(layers, layers + L);
    D->embed = static_cast<const __half*>(embed_tokens_f16);
    D->final_norm = final_norm;
    D->lm_head = static_cast<const __half*>(lm_head_f16);
    D->rope_cos = rope_cos;
    D->rope_sin = rope_sin;
    D->allreduce = allreduce;
    D->allreduce_user = allreduce_user;
    D->tp = tp;
    D->heads_l = cfg->num_heads / tp;
    D->Hl = H / tp;
    D->Il = I / tp;
    D->Hk = (D->Hl + imma::kUnitCols - 1) / imma::kUnitCols * imma::kUnitCols;  // K of the o_proj shard (zero padded)
    D->Ik = (D->Il + imma::kUnitCols - 1) / imma::kUnitCols * imma::kUnitCols;  // K of the down_proj shard

    const int uH = H / imma::kUnitCols, uI = I / imma::kUnitCols;
    const int cH = (H + imma::kRows - 1) / imma::kRows, cI = (I + imma::kRows - 1) / imma::kRows;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const bool big = B > 4;  // batches the bit-plane GEMV cannot hold: tcgen05 path with split-K partials
    const size_t ks = big ? kKSplitMax : 1;
    D->ksplit_max = (int)ks;
    const size_t o_res0 = take((size_t)B * H * 4), o_res1 = take((size_t)B * H * 4);
    const size_t o_tqkv = take((size_t)3 * B * H * 4 * ks), o_to = take((size_t)B * H * 4 * ks);
    const size_t o_tgu = take((size_t)2 * B * I * 4 * ks), o_td = take((size_t)B * H * 4 * ks);
    const size_t o_att = take((size_t)B * H * 4), o_log = take((size_t)B * cfg->vocab_size * 4);
    const size_t o_sqkv = take((size_t)3 * cH * B * 8), o_so = take((size_t)cH * B * 8);
    const size_t o_sgu = take((size_t)2 * cI * B * 8), o_sd = take((size_t)cH * B * 8);
    const size_t o_dqkv = take((size_t)3 * B * uH * imma::kUnitBytes), o_do = take((size_t)B * uH * imma::kUnitBytes);
    const size_t o_dgu = take((size_t)2 * B * uH * imma::kUnitBytes), o_dd = take((size_t)B * uI * imma::kUnitBytes);
    const size_t o_qqkv = take(3 * B * sizeof(imma::QMeta)), o_qo = take(B * sizeof(imma::QMeta));
    const size_t o_qgu = take(2 * B * sizeof(imma::QMeta)), o_qd = take(B * sizeof(imma::QMeta));
    const size_t cache_elems = (size_t)L * B * D->heads_l * cfg->max_seq_len * kHeadDim;
    const size_t o_kc = take(cache_elems * 2), o_vc = take(cache_elems * 2);
    const size_t o_x16 = take((size_t)B * H * 2), o_ids = take(B * 8), o_ids2 = take(B * 8), o_pos = take(B * 4);
    const size_t o_rq = take((size_t)3 * kReduceSlices * B * 2 * 4), o_rg = take((size_t)2 * kReduceSlices * B * 2 * 4);
    const size_t o_ro = take((size_t)kReduceSlices * B * 2 * 4), o_rd = take((size_t)kReduceSlices * B * 2 * 4);
    D->attn_nsplit = std::max(1, std::min(16, cfg->max_seq_len / 512));
    const size_t o_ap = take((size_t)B * D->heads_l * D->attn_nsplit * (kHeadDim + 2) * 4), o_at = take((size_t)B * D->heads_l * 4);
    const size_t o_xI = take(big ? (size_t)B * std::max(I, D->Ik) * 2 : 16);
    // record arrays of the second-generation fused stages: one contiguous block, fused2::kReplicas copies of it
    const size_t ext_one = ((size_t)5 * cH + (size_t)2 * cI) * B * fused2::kExt * 4;
    const size_t o_eq = take(ext_one * fused2::kReplicas);
    const size_t o_eo = o_eq + (size_t)3 * cH * B * fused2::kExt * 4, o_eg = o_eo + (size_t)cH * B * fused2::kExt * 4;
    const size_t o_ed = o_eg + (size_t)2 * cI * B * fused2::kExt * 4;
    const size_t o_am = take((size_t)B * cfg->num_heads * 4);
    const bool need_h16 = big && cfg->param_dtype != ONEBIT_F16;
    const size_t o_h16 = take(need_h16 ? (size_t)L * (5 * (size_t)H + D->Hk + D->Ik) * 2 : 16);
    cudaError_t e = cudaMalloc(&D->arena, off);
    if (e != cudaSuccess) {
        delete D;
        return fail(ONEBIT_ERR_CUDA, std::string("decoder_create: cudaMalloc: ") + cudaGetErrorString(e));
    }
    cudaMemset(D->arena, 0, off);
    char* a = D->arena;
    D->resid[0] = (float*)(a + o_res0); D->resid[1] = (float*)(a + o_res1);
    D->t_qkv = (float*)(a + o_tqkv); D->t_o = (float*)(a + o_to); D->t_gu = (float*)(a + o_tgu); D->t_d = (float*)(a + o_td);
    D->attn_out = (float*)(a + o_att); D->logits = (float*)(a + o_log);
    D->st_qkv = (float*)(a + o_sqkv); D->st_o = (float*)(a + o_so); D->st_gu = (float*)(a + o_sgu); D->st_d = (float*)(a + o_sd);
    D->dg_qkv = (uint8_t*)(a + o_dqkv); D->dg_o = (uint8_t*)(a + o_do); D->dg_gu = (uint8_t*)(a + o_dgu); D->dg_d = (uint8_t*)(a + o_dd);
    D->qm_qkv = (imma::QMeta*)(a + o_qqkv); D->qm_o = (imma::QMeta*)(a + o_qo);
    D->qm_gu = (imma::QMeta*)(a + o_qgu); D->qm_d = (imma::QMeta*)(a + o_qd);
    D->kcache = (__half*)(a + o_kc); D->vcache = (__half*)(a + o_vc); D->x_f16 = (__half*)(a + o_x16);
    D->ids = (long long*)(a + o_ids); D->ids_stage = (long long*)(a + o_ids2); D->pos = (int*)(a + o_pos);
    D->red_qkv = (float*)(a + o_rq); D->red_gu = (float*)(a + o_rg);
    D->red_o = (float*)(a + o_ro); D->red_d = (float*)(a + o_rd);
    D->attn_part = (float*)(a + o_ap); D->attn_tickets = (int*)(a + o_at);
    D->xI_f16 = (__half*)(a + o_xI);
    D->ext_qkv = (float*)(a + o_eq); D->ext_o = (float*)(a + o_eo); D->ext_gu = (float*)(a + o_eg); D->ext_d = (float*)(a + o_ed);
    D->attn_amax = (float*)(a + o_am);
    if (tp == 1) {
        const int rc = build_v2_tables(D);
        if (rc != ONEBIT_OK) { onebit_decoder_destroy(D); return rc; }
        D->v2 = true;
    }
    D->h16_store = (__half*)(a + o_h16);
    if (big) {  // the tcgen05 A operand folds input_factor in as fp16: one-time copies when the parameters are bf16 / fp32
        D->h16.resize((size_t)L * 7);
        __half* hp = D->h16_store;
        for (int l = 0; l < L; ++l) {
            const onebit_bitlinear_params* bl[7] = {&layers[l].q, &layers[l].k, &layers[l].v, &layers[l].o, &layers[l].gate, &layers[l].up, &layers[l].down};
            for (int i = 0; i < 7; ++i) {
                const int k = i == 6 ? D->Ik : (i == 3 ? D->Hk : H);  // local K of the shard (padded for o / down)
                if (need_h16) {
                    const int rc = launch_to_half(bl[i]->input_factor, hp, k, cfg->param_dtype, nullptr);
                    if (rc != ONEBIT_OK) { onebit_decoder_destroy(D); return rc; }
                    D->h16[(size_t)l * 7 + i] = hp;
                    hp += k;
                } else {
                    D->h16[(size_t)l * 7 + i] = static_cast<const __half*>(bl[i]->input_factor);
                }
            }
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { onebit_decoder_destroy(D); return fail(ONEBIT_ERR_CUDA, "decoder_create: fp16 input_factor copies failed"); }
    }
    if (persist_supported(*cfg)) {
        const int rc = persist_create(&D->persist, *cfg, layers, embed_tokens_f16, final_norm, lm_head_f16, rope_cos, rope_sin,
                                      D->kcache, D->vcache, D->ids, D->pos);
        if (rc != ONEBIT_OK) {
            onebit_decoder_destroy(D);
            return rc;
        }
    }
    *out = D;
    return ONEBIT_OK;
}

int onebit_decoder_reset(onebit_decoder* D, const int64_t* ids_host, const int32_t* pos_host, int batch, void* stream) {
    ONEBIT_REQUIRE(D && ids_host && pos_host && batch >= 1 && batch <= D->cfg.max_batch, "decoder_reset: bad arguments");
    int hi = 0;
    for (int b = 0; b < batch; ++b) {
        ONEBIT_REQUIRE(ids_host[b] >= 0 && ids_host[b] < D->cfg.vocab_size, "decoder_reset: token id outside [0, vocab_size)");
        ONEBIT_REQUIRE(pos_host[b] >= 0 && pos_host[b] < D->cfg.max_seq_len, "decoder_reset: position outside [0, max_seq_len)");
        hi = std::max(hi, (int)pos_host[b]);
    }
    D->pos_hi = hi;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ONEBIT_CUDA_TRY(cudaMemcpyAsync(D->ids, ids_host, (size_t)batch * 8, cudaMemcpyHostToDevice, s));
    ONEBIT_CUDA_TRY(cudaMemcpyAsync(D->pos, pos_host, (size_t)batch * 4, cudaMemcpyHostToDevice, s));
    ONEBIT_CUDA_TRY(cudaStreamSynchronize(s));
    return ONEBIT_OK;
}

const int64_t* onebit_decoder_next_ids(onebit_decoder* D) { return D ? reinterpret_cast<const int64_t*>(D->ids) : nullptr; }
const int32_t* onebit_decoder_positions(onebit_decoder* D) { return D ? D->pos : nullptr; }
int onebit_decoder_kernel_launches_per_step(onebit_decoder* D) { return D ? D->launches : 0; }

static int decoder_step_impl(onebit_decoder* D, int batch, const int64_t* forced_ids_dev, float* logits_dev, void* stream);

int onebit_decoder_step(onebit_decoder* D, int batch, const int64_t* forced_ids_dev, float* logits_dev, void* stream) {
    ONEBIT_REQUIRE(D && batch >= 1 && batch <= D->cfg.max_batch, "decoder_step: bad arguments");
    // the step writes cache row `pos`: refuse once any sequence may have reached the end of the static KV cache
    // (callers that replay a captured graph must keep the same count themselves: BitLlamaDecoderB200 does)
    if (D->pos_hi >= D->cfg.max_seq_len)
        return fail(ONEBIT_ERR_INVALID_ARGUMENT, "decoder_step: a sequence has reached max_seq_len (" + std::to_string(D->cfg.max_seq_len) +
                                                     "); create the decoder with a longer cache");
    D->pos_hi += 1;
    if (D->persist && batch <= persist::kMaxTok) {
        D->launches = 1;
        return persist_step(D->persist, batch, forced_ids_dev ? reinterpret_cast<const long long*>(forced_ids_dev) : D->ids,
                            logits_dev ? logits_dev : D->logits, static_cast<cudaStream_t>(stream));
    }
    // Tensor-parallel steps interleave NCCL kernels (another stream, event-ordered): keep plain stream order there, so
    // that no early-launched CTA of ours can sit on an SM waiting for a collective that needs that SM.
    // (with the one-shot peer-memory all-reduce every kernel of the step is ours: programmatic dependent launch stays on)
    const bool suspend = D->tp > 1 && !D->p2p_on;
    if (suspend) pdl_suspend(true);
    const int rc = decoder_step_impl(D, batch, forced_ids_dev, logits_dev, stream);
    if (suspend) pdl_suspend(false);
    return rc;
}

static int decoder_step_impl(onebit_decoder* D, int batch, const int64_t* forced_ids_dev, float* logits_dev, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, M = batch, pd = C.param_dtype;
    const int uH = H / imma::kUnitCols, uI = I / imma::kUnitCols;
    const int cH = (H + imma::kRows - 1) / imma::kRows, cI = (I + imma::kRows - 1) / imma::kRows;
    const size_t dgH = (size_t)C.max_batch * uH * imma::kUnitBytes;
    int rc, launches = 0;
    if (forced_ids_dev) {
        rc = launch_pdl(copy_ids_kernel, dim3(1), dim3(kMaxBatch), 0, s, reinterpret_cast<const long long*>(forced_ids_dev),
                        D->ids, M);
        if (rc) return rc;
        ++launches;
    }
    // batches the bit-plane GEMV cannot hold (it keeps every token's digits of the widest layer in shared memory) run
    // on the tcgen05 path when the decoder was created for them (max_batch > 4)
    const int nt_gemv = M <= 2 ? 1 : (M <= 4 ? 2 : 4);
    const bool gemv_fits = M <= imma::kMaxTokens && imma::gemv_smem_bytes(M, std::max(uH, uI), nt_gemv) <= 224 * 1024;
    const bool use_tc5 = D->ksplit_max > 1 && (M > 4 || !gemv_fits);
    if (!use_tc5 && !gemv_fits)
        return fail(ONEBIT_ERR_INVALID_ARGUMENT,
                    "decoder_step: batch " + std::to_string(M) + " does not fit the decode GEMV for this model width "
                    "(LLaMA-7B: up to 4, LLaMA2-13B: up to 3 sequences); create the decoder with max_batch > 4 for the "
                    "batched tcgen05 path");
    int cur = 0;  // residual ping-pong index holding the current stream
    int nc_d = cH;  // CTAs that wrote the (sum, sumsq) partials of the last down_proj
    bool use_fused = false;
    if (use_tc5) {
        rc = run_tc5_layers(D, M, s, &launches, &cur);
        if (rc) return rc;
        use_fused = true;  // (skips the split-chain loop below)
    } else {
        rc = run_fused2_layers(D, M, s, /*with_attention=*/true, /*first_is_embed=*/true, &launches, &cur, &nc_d, &use_fused);
        if (rc) return rc;
        if (!use_fused) {
            rc = run_fused_layers(D, M, s, /*with_attention=*/true, /*first_is_embed=*/true, &launches, &cur, &nc_d, &use_fused);
            if (rc) return rc;
        }
    }
    if (!use_fused && D->tp > 1)
        return fail(ONEBIT_ERR_INVALID_ARGUMENT, "tensor-parallel decode needs the fused stages (batch <= 2, ONEBIT_FUSED != 0) or a decoder created with max_batch > 4 (batched tcgen05 path)");
    for (int l = 0; !use_fused && l < C.num_layers; ++l) {
        const onebit_layer_params& P = D->layers[l];
        // ---- glue 1: (embed | resid + LN(down of previous layer)) -> RMSNorm -> q/k/v digits
        GlueArgs g = {};
        g.mode = l == 0 ? GLUE_EMBED_NORM : GLUE_RESID_NORM;
        g.M = M; g.K = H; g.nprob = 3;
        g.t_a = D->t_d; g.stats_a = D->st_d; g.ncta_a = cH;
        g.resid_in = D->resid[cur]; g.resid_out = D->resid[cur ^ 1];
        g.embed = D->embed; g.ids = D->ids; g.ln_w = P.input_layernorm;
        g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
        g.h[0] = P.q.input_factor; g.h[1] = P.k.input_factor; g.h[2] = P.v.input_factor;
        for (int i = 0; i < 3; ++i) { g.digits[i] = D->dg_qkv + i * dgH; g.qmeta[i] = D->qm_qkv + i * C.max_batch; }
        rc = glue_launch(D, g, s); if (rc) return rc; ++launches;
        cur ^= 1;
        // ---- GEMV q,k,v
        imma::Args a = {};
        a.nprob = 3; a.M = M; a.K = H; a.units = uH;
        const onebit_bitlinear_params* qkv[3] = {&P.q, &P.k, &P.v};
        for (int i = 0; i < 3; ++i) {
            a.p[i].w = reinterpret_cast<const uint8_t*>(qkv[i]->weight); a.p[i].g = qkv[i]->weight_scale;
            a.p[i].digits = D->dg_qkv + i * dgH; a.p[i].qmeta = D->qm_qkv + i * C.max_batch;
            a.p[i].t = D->t_qkv + (size_t)i * C.max_batch * H; a.p[i].stats = D->st_qkv + (size_t)i * cH * C.max_batch * 2;
            a.p[i].n_rows = H; a.p[i].ld_t = H;
        }
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc; ++launches;
        // ---- attention
        AttnArgs at = {};
        at.t_q = a.p[0].t; at.t_k = a.p[1].t; at.t_v = a.p[2].t;
        at.stats_q = a.p[0].stats; at.stats_k = a.p[1].stats; at.stats_v = a.p[2].stats; at.ncta = cH;
        at.M = M; at.H = H; at.n_ln = H; at.out_ld = H; at.n_heads = C.num_heads; at.max_seq = C.max_seq_len; at.pos = D->pos;
        at.rope_cos = D->rope_cos; at.rope_sin = D->rope_sin;
        const size_t layer_cache = (size_t)C.max_batch * D->heads_l * C.max_seq_len * kHeadDim;
        at.kcache = D->kcache + l * layer_cache; at.vcache = D->vcache + l * layer_cache;
        at.out = D->attn_out; at.ln_eps = C.ln_eps;
        const size_t asmem2 = attn_finish_args(D, at, M, false);
        rc = launch_attn( dim3(M, C.num_heads, at.nsplit), dim3(kHeadDim), asmem2, s, at);
        if (rc) return rc; ++launches;
        // ---- glue 2: attention output -> o digits
        g = {};
        g.mode = GLUE_PLAIN; g.M = M; g.K = H; g.nprob = 1; g.x_plain = D->attn_out;
        g.h[0] = P.o.input_factor; g.digits[0] = D->dg_o; g.qmeta[0] = D->qm_o;
        rc = glue_launch(D, g, s); if (rc) return rc; ++launches;
        // ---- GEMV o
        a = {};
        a.nprob = 1; a.M = M; a.K = H; a.units = uH;
        a.p[0].w = reinterpret_cast<const uint8_t*>(P.o.weight); a.p[0].g = P.o.weight_scale;
        a.p[0].digits = D->dg_o; a.p[0].qmeta = D->qm_o; a.p[0].t = D->t_o; a.p[0].stats = D->st_o;
        a.p[0].n_rows = H; a.p[0].ld_t = H;
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc; ++launches;
        // ---- glue 3: resid + LN(o) -> RMSNorm -> gate/up digits
        g = {};
        g.mode = GLUE_RESID_NORM; g.M = M; g.K = H; g.nprob = 2;
        g.t_a = D->t_o; g.stats_a = D->st_o; g.ncta_a = cH;
        g.resid_in = D->resid[cur]; g.resid_out = D->resid[cur ^ 1]; g.ln_w = P.post_attention_layernorm;
        g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
        g.h[0] = P.gate.input_factor; g.h[1] = P.up.input_factor;
        for (int i = 0; i < 2; ++i) { g.digits[i] = D->dg_gu + i * dgH; g.qmeta[i] = D->qm_gu + i * C.max_batch; }
        rc = glue_launch(D, g, s); if (rc) return rc; ++launches;
        cur ^= 1;
        // ---- GEMV gate, up
        a = {};
        a.nprob = 2; a.M = M; a.K = H; a.units = uH;
        const onebit_bitlinear_params* gu[2] = {&P.gate, &P.up};
        for (int i = 0; i < 2; ++i) {
            a.p[i].w = reinterpret_cast<const uint8_t*>(gu[i]->weight); a.p[i].g = gu[i]->weight_scale;
            a.p[i].digits = D->dg_gu + i * dgH; a.p[i].qmeta = D->qm_gu + i * C.max_batch;
            a.p[i].t = D->t_gu + (size_t)i * C.max_batch * I; a.p[i].stats = D->st_gu + (size_t)i * cI * C.max_batch * 2;
            a.p[i].n_rows = I; a.p[i].ld_t = I;
        }
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc; ++launches;
        // ---- glue 4: silu(LN(gate)) * LN(up) -> down digits
        g = {};
        g.mode = GLUE_SILU_MUL; g.M = M; g.K = I; g.nprob = 1;
        g.t_a = a.p[0].t; g.stats_a = a.p[0].stats; g.ncta_a = cI;
        g.t_b = a.p[1].t; g.stats_b = a.p[1].stats; g.ncta_b = cI;
        g.ln_eps = C.ln_eps;
        g.h[0] = P.down.input_factor; g.digits[0] = D->dg_d; g.qmeta[0] = D->qm_d;
        rc = glue_launch(D, g, s); if (rc) return rc; ++launches;
        // ---- GEMV down
        a = {};
        a.nprob = 1; a.M = M; a.K = I; a.units = uI;
        a.p[0].w = reinterpret_cast<const uint8_t*>(P.down.weight); a.p[0].g = P.down.weight_scale;
        a.p[0].digits = D->dg_d; a.p[0].qmeta = D->qm_d; a.p[0].t = D->t_d; a.p[0].stats = D->st_d;
        a.p[0].n_rows = H; a.p[0].ld_t = H;
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc; ++launches;
    }
    // ---- final: resid + LN(down) -> RMSNorm(final) -> fp16 x -> lm_head -> argmax
    GlueArgs g = {};
    g.mode = GLUE_RESID_NORM; g.M = M; g.K = H; g.nprob = 1; g.write_x_f16 = 1;
    g.t_a = D->t_d; g.stats_a = use_tc5 ? D->red_d : D->st_d; g.ncta_a = use_tc5 ? kReduceSlices : (use_fused ? nc_d : cH);
    g.stats_from_data = D->tp > 1;
    if (use_tc5 && D->tp == 1) { g.stats_from_data = 1; g.ksplit = D->tc5_ks_d; g.slab = (long long)M * H; }
    g.resid_in = D->resid[cur]; g.resid_out = D->resid[cur ^ 1]; g.ln_w = D->final_norm;
    g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps; g.x_f16 = D->x_f16;
    rc = glue_launch(D, g, s); if (rc) return rc; ++launches;
    float* logits = logits_dev ? logits_dev : D->logits;
    if (M > 8) {  // dense fp16 GEMM on the tcgen05 pipeline (the GEMV form would re-read x from shared memory per row)
        rc = launch_dense_tc5(D->x_f16, D->lm_head, logits, M, H, C.vocab_size, s);
    } else if (M <= 2) {
        rc = launch_pdl(lm_head_kernel<2>, dim3((C.vocab_size + 7) / 8), dim3(256), (size_t)M * H * 2, s, D->lm_head,
                        (const __half*)D->x_f16, logits, C.vocab_size, H, M);
    } else {
        static bool configured[64] = {false};
        int dev = 0;
        ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !configured[dev]) {
            ONEBIT_CUDA_TRY(cudaFuncSetAttribute(lm_head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
            configured[dev] = true;
        }
        rc = launch_pdl(lm_head_kernel<8>, dim3((C.vocab_size + 7) / 8), dim3(256), (size_t)M * H * 2, s, D->lm_head,
                        (const __half*)D->x_f16, logits, C.vocab_size, H, M);
    }
    if (rc) return rc; ++launches;
    rc = launch_pdl(argmax_advance_kernel, dim3(M), dim3(1024), 0, s, (const float*)logits, C.vocab_size, D->ids, D->pos);
    if (rc) return rc; ++launches;
    D->launches = launches;
    return ONEBIT_OK;
}

// Enqueue only the BitLinear GEMV launches of one step (4 per layer), reusing whatever activation digits are
// resident: the weight-streaming kernel chain on its own, for the roofline measurement in bench.py.
// Prompt pass: all T tokens of `batch` sequences at once (what the reference's forward does for q_len > 1,
// modeling_bitllama.py:1217-1315,1546-1611; lm_eval.py:99-124 runs 2048-token windows this way). Every BitLinear is one
// tcgen05 GEMM over M = batch * T tokens (prefill tile for M > 64), attention is the causal flash kernel of prefill_attn.cu,
// K / V land in the static cache at positions pos0 .. pos0 + T - 1, so decode steps continue from there.
//   ids_dev            [batch][T] int64 token ids (device)
//   logits_last_dev    [batch][V] fp32 logits of every sequence's last token, or NULL
//   logits_all_dev     [batch * T][V] fp32 logits of every token (perplexity / parity), or NULL
// On return the decoder's next ids are the greedy continuation of each prompt and its positions are pos0 + T.
int onebit_decoder_prefill(onebit_decoder* D, int batch, int T, int pos0, const int64_t* ids_dev, float* logits_last_dev,
                           float* logits_all_dev, void* stream) {
    ONEBIT_REQUIRE(D && ids_dev && batch >= 1 && batch <= D->cfg.max_batch && T >= 1 && pos0 >= 0, "decoder_prefill: bad arguments");
    ONEBIT_REQUIRE(D->tp == 1, "decoder_prefill: the prompt pass is single-GPU (tensor-parallel decoders feed the prompt step by step)");
    ONEBIT_REQUIRE(pos0 + T <= D->cfg.max_seq_len, "decoder_prefill: the prompt does not fit max_seq_len");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, L = C.num_layers, pd = C.param_dtype, V = C.vocab_size;
    const size_t M = (size_t)batch * T;
    ONEBIT_REQUIRE(M < (1u << 30), "decoder_prefill: too many tokens");
    // ---- workspace
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b_res = up(M * H * 4), b_tq = up(3 * M * H * 4), b_th = up(M * H * 4), b_tg = up(2 * M * I * 4);
    const size_t b_xa = up(M * H * 2), b_xi = up(M * I * 2), b_q = up(M * H * 2), b_st = up((size_t)3 * kReduceSlices * M * 2 * 4);
    const size_t need = 2 * b_res + b_tq + 2 * b_th + b_tg + b_xa + b_xi + b_q + 4 * b_st;
    if (D->pf_cap < need) {
        cudaFree(D->pf_ws);
        D->pf_ws = nullptr;
        D->pf_cap = 0;
        cudaError_t e = cudaMalloc(&D->pf_ws, need);
        if (e != cudaSuccess) return fail(ONEBIT_ERR_CUDA, std::string("decoder_prefill: cudaMalloc of the prompt workspace: ") + cudaGetErrorString(e));
        D->pf_cap = need;
    }
    char* w = D->pf_ws;
    float* resid[2] = {(float*)w, (float*)(w + b_res)}; w += 2 * b_res;
    float* t_qkv = (float*)w; w += b_tq;
    float* t_o = (float*)w; w += b_th;
    float* t_d = (float*)w; w += b_th;
    float* t_gu = (float*)w; w += b_tg;
    __half* xa = (__half*)w; w += b_xa;
    __half* xi = (__half*)w; w += b_xi;
    __half* q16 = (__half*)w; w += b_q;
    float* st_qkv = (float*)w; w += b_st;
    float* st_o = (float*)w; w += b_st;
    float* st_gu = (float*)w; w += b_st;
    float* st_d = (float*)w; w += b_st;
    // ---- fp16 input_factor vectors (the tcgen05 A operand folds them in)
    if (D->pf_h16.empty()) {
        if (!D->h16.empty()) {
            D->pf_h16 = D->h16;
        } else {
            D->pf_h16.resize((size_t)L * 7);
            __half* hp = nullptr;
            if (pd != ONEBIT_F16) {
                ONEBIT_CUDA_TRY(cudaMalloc(&D->pf_h16_store, (size_t)L * (6 * (size_t)H + I) * 2));
                hp = D->pf_h16_store;
            }
            for (int l = 0; l < L; ++l) {
                const onebit_layer_params& P = D->layers[l];
                const onebit_bitlinear_params* bl[7] = {&P.q, &P.k, &P.v, &P.o, &P.gate, &P.up, &P.down};
                for (int i = 0; i < 7; ++i) {
                    const int k = i == 6 ? I : H;
                    if (pd != ONEBIT_F16) {
                        const int rc = launch_to_half(bl[i]->input_factor, hp, k, pd, s);
                        if (rc) return rc;
                        D->pf_h16[(size_t)l * 7 + i] = hp;
                        hp += k;
                    } else {
                        D->pf_h16[(size_t)l * 7 + i] = static_cast<const __half*>(bl[i]->input_factor);
                    }
                }
            }
        }
    }
    int rc, cur = 0;
    const int Mi = (int)M;
    // (no statistics pass: every consumer — glue kernels, q/k/v preparation — holds its token's whole rows and takes the
    // LayerNorm sums from them)
    auto proj = [&](const __half* x, int K, int nprob, const onebit_bitlinear_params* const* bl, const __half* const* h16, float* const* t, int N) -> int {
        Tc5Launch tl = {};
        tl.x16 = x; tl.M = Mi; tl.K = K; tl.nprob = nprob; tl.ksplit = 1; tl.param_dtype = pd;
        for (int i = 0; i < nprob; ++i) {
            tl.p[i].w = static_cast<const int8_t*>(bl[i]->weight); tl.p[i].h16 = h16[i]; tl.p[i].g = bl[i]->weight_scale;
            tl.p[i].t = t[i]; tl.p[i].N = N;
        }
        return launch_tc5(tl, s);
    };
    for (int l = 0; l < L; ++l) {
        const onebit_layer_params& P = D->layers[l];
        const __half* const* h16 = &D->pf_h16[(size_t)l * 7];
        GlueArgs g = {};
        g.mode = l == 0 ? GLUE_EMBED_NORM : GLUE_RESID_NORM;
        g.M = Mi; g.K = H; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = xa;
        g.t_a = t_d; g.stats_a = st_d; g.ncta_a = kReduceSlices; g.stats_from_data = 1;
        g.resid_in = resid[cur]; g.resid_out = resid[cur ^ 1];
        g.embed = D->embed; g.ids = reinterpret_cast<const long long*>(ids_dev); g.ln_w = P.input_layernorm;
        g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
        rc = glue_launch(D, g, s); if (rc) return rc;
        cur ^= 1;
        const onebit_bitlinear_params* qkv[3] = {&P.q, &P.k, &P.v};
        float* tq[3] = {t_qkv, t_qkv + M * H, t_qkv + 2 * M * H};
        rc = proj(xa, H, 3, qkv, h16, tq, H); if (rc) return rc;
        PrefillAttnArgs at = {};
        at.t_q = tq[0]; at.t_k = tq[1]; at.t_v = tq[2]; at.M = Mi; at.ld = H; at.n_ln = H;
        at.B = batch; at.T = T; at.pos0 = pos0; at.n_heads = C.num_heads; at.max_seq = C.max_seq_len;
        at.rope_cos = D->rope_cos; at.rope_sin = D->rope_sin;
        const size_t layer_cache = (size_t)C.max_batch * D->heads_l * C.max_seq_len * kHeadDim;
        at.kcache = D->kcache + l * layer_cache; at.vcache = D->vcache + l * layer_cache;
        at.q16 = q16; at.out16 = xa; at.out_ld = H; at.ln_eps = C.ln_eps;
        rc = launch_prefill_attention(at, s); if (rc) return rc;
        const onebit_bitlinear_params* po[1] = {&P.o};
        float* to[1] = {t_o};
        rc = proj(xa, H, 1, po, h16 + 3, to, H); if (rc) return rc;
        g = {};
        g.mode = GLUE_RESID_NORM; g.M = Mi; g.K = H; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = xa;
        g.t_a = t_o; g.stats_a = st_o; g.ncta_a = kReduceSlices; g.stats_from_data = 1;
        g.resid_in = resid[cur]; g.resid_out = resid[cur ^ 1]; g.ln_w = P.post_attention_layernorm;
        g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
        rc = glue_launch(D, g, s); if (rc) return rc;
        cur ^= 1;
        const onebit_bitlinear_params* gu[2] = {&P.gate, &P.up};
        float* tg[2] = {t_gu, t_gu + M * I};
        rc = proj(xa, H, 2, gu, h16 + 4, tg, I); if (rc) return rc;
        g = {};
        g.mode = GLUE_SILU_MUL; g.M = Mi; g.K = I; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = xi;
        g.t_a = tg[0]; g.stats_a = st_gu; g.ncta_a = kReduceSlices; g.stats_from_data = 1;
        g.t_b = tg[1]; g.stats_b = st_gu + (size_t)kReduceSlices * M * 2; g.ncta_b = kReduceSlices;
        g.ln_eps = C.ln_eps;
        rc = glue_launch(D, g, s); if (rc) return rc;
        const onebit_bitlinear_params* pdn[1] = {&P.down};
        float* td[1] = {t_d};
        rc = proj(xi, I, 1, pdn, h16 + 6, td, H); if (rc) return rc;
    }
    // ---- final norm -> fp16 x of every token
    GlueArgs g = {};
    g.mode = GLUE_RESID_NORM; g.M = Mi; g.K = H; g.nprob = 1; g.write_x_f16 = 1; g.x_f16 = xa;
    g.t_a = t_d; g.stats_a = st_d; g.ncta_a = kReduceSlices; g.stats_from_data = 1;
    g.resid_in = resid[cur]; g.resid_out = resid[cur ^ 1]; g.ln_w = D->final_norm; g.ln_eps = C.ln_eps; g.rms_eps = C.rms_eps;
    rc = glue_launch(D, g, s); if (rc) return rc;
    if (logits_all_dev) {  // every token's logits, 64 tokens per dense tcgen05 launch
        for (size_t m0 = 0; m0 < M; m0 += 64) {
            const int64_t mm = (int64_t)std::min<size_t>(64, M - m0);
            rc = launch_dense_tc5(xa + m0 * H, D->lm_head, logits_all_dev + m0 * V, mm, H, V, s);
            if (rc) return rc;
        }
    }
    // ---- last token of every sequence: logits, greedy next id, positions for the decode steps that follow
    gather_last_rows_kernel<<<batch, 128, 0, s>>>(xa, D->x_f16, T, H);
    ONEBIT_CUDA_TRY(cudaGetLastError());
    float* logits = logits_last_dev ? logits_last_dev : D->logits;
    if (batch > 8) {
        rc = launch_dense_tc5(D->x_f16, D->lm_head, logits, batch, H, V, s);
    } else {
        static bool configured[64] = {false};
        int dev = 0;
        ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !configured[dev]) {
            ONEBIT_CUDA_TRY(cudaFuncSetAttribute(lm_head_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
            configured[dev] = true;
        }
        rc = launch_pdl(lm_head_kernel<8>, dim3((V + 7) / 8), dim3(256), (size_t)batch * H * 2, s, D->lm_head, (const __half*)D->x_f16,
                        logits, V, H, batch);
    }
    if (rc) return rc;
    set_positions_kernel<<<1, kMaxBatch, 0, s>>>(D->pos, pos0 + T - 1, batch);
    ONEBIT_CUDA_TRY(cudaGetLastError());
    rc = launch_pdl(argmax_advance_kernel, dim3(batch), dim3(1024), 0, s, (const float*)logits, V, D->ids, D->pos);
    if (rc) return rc;
    D->pos_hi = pos0 + T;
    return ONEBIT_OK;
}

int onebit_decoder_gemv_only(onebit_decoder* D, int batch, void* stream) {
    ONEBIT_REQUIRE(D && batch >= 1 && batch <= D->cfg.max_batch, "decoder_gemv_only: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const onebit_decoder_config& C = D->cfg;
    const int H = C.hidden_size, I = C.intermediate_size, M = batch, pd = C.param_dtype;
    const int uH = H / imma::kUnitCols, uI = I / imma::kUnitCols;
    const int cH = (H + imma::kRows - 1) / imma::kRows, cI = (I + imma::kRows - 1) / imma::kRows;
    const size_t dgH = (size_t)C.max_batch * uH * imma::kUnitBytes;
    if (D->ksplit_max > 1 && M > 4) {  // batched decode: the projection launches of the tcgen05 path alone
        int launches = 0, cur = 0;
        return run_tc5_layers(D, M, s, &launches, &cur, /*only_proj=*/true);
    }
    {
        int launches = 0, cur = 0, nc_d = cH;
        bool used = false;
        int rc = run_fused2_layers(D, M, s, /*with_attention=*/false, /*first_is_embed=*/false, &launches, &cur, &nc_d, &used);
        if (rc) return rc;
        if (used) return ONEBIT_OK;
        rc = run_fused_layers(D, M, s, /*with_attention=*/false, /*first_is_embed=*/false, &launches, &cur, &nc_d, &used);
        if (rc) return rc;
        if (used) return ONEBIT_OK;
    }
    for (int l = 0; l < C.num_layers; ++l) {
        const onebit_layer_params& P = D->layers[l];
        imma::Args a = {};
        a.nprob = 3; a.M = M; a.K = H; a.units = uH;
        const onebit_bitlinear_params* qkv[3] = {&P.q, &P.k, &P.v};
        for (int i = 0; i < 3; ++i) {
            a.p[i].w = reinterpret_cast<const uint8_t*>(qkv[i]->weight); a.p[i].g = qkv[i]->weight_scale;
            a.p[i].digits = D->dg_qkv + i * dgH; a.p[i].qmeta = D->qm_qkv + i * C.max_batch;
            a.p[i].t = D->t_qkv + (size_t)i * C.max_batch * H; a.p[i].stats = D->st_qkv + (size_t)i * cH * C.max_batch * 2;
            a.p[i].n_rows = H; a.p[i].ld_t = H;
        }
        int rc = launch_imma_gemv(a, pd, s); if (rc) return rc;
        a = {};
        a.nprob = 1; a.M = M; a.K = H; a.units = uH;
        a.p[0].w = reinterpret_cast<const uint8_t*>(P.o.weight); a.p[0].g = P.o.weight_scale;
        a.p[0].digits = D->dg_o; a.p[0].qmeta = D->qm_o; a.p[0].t = D->t_o; a.p[0].stats = D->st_o;
        a.p[0].n_rows = H; a.p[0].ld_t = H;
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc;
        a = {};
        a.nprob = 2; a.M = M; a.K = H; a.units = uH;
        const onebit_bitlinear_params* gu[2] = {&P.gate, &P.up};
        for (int i = 0; i < 2; ++i) {
            a.p[i].w = reinterpret_cast<const uint8_t*>(gu[i]->weight); a.p[i].g = gu[i]->weight_scale;
            a.p[i].digits = D->dg_gu + i * dgH; a.p[i].qmeta = D->qm_gu + i * C.max_batch;
            a.p[i].t = D->t_gu + (size_t)i * C.max_batch * I; a.p[i].stats = D->st_gu + (size_t)i * cI * C.max_batch * 2;
            a.p[i].n_rows = I; a.p[i].ld_t = I;
        }
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc;
        a = {};
        a.nprob = 1; a.M = M; a.K = I; a.units = uI;
        a.p[0].w = reinterpret_cast<const uint8_t*>(P.down.weight); a.p[0].g = P.down.weight_scale;
        a.p[0].digits = D->dg_d; a.p[0].qmeta = D->qm_d; a.p[0].t = D->t_d; a.p[0].stats = D->st_d;
        a.p[0].n_rows = H; a.p[0].ld_t = H;
        rc = launch_imma_gemv(a, pd, s); if (rc) return rc;
    }
    return ONEBIT_OK;
}

int onebit_decoder_step_host(onebit_decoder* D, int batch, const int64_t* ids_host, int64_t* next_ids_host, void* stream) {
    ONEBIT_REQUIRE(D && ids_host && next_ids_host && batch >= 1 && batch <= D->cfg.max_batch, "decoder_step_host: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ONEBIT_CUDA_TRY(cudaMemcpyAsync(D->ids_stage, ids_host, (size_t)batch * 8, cudaMemcpyHostToDevice, s));
    int rc = onebit_decoder_step(D, batch, reinterpret_cast<const int64_t*>(D->ids_stage), nullptr, stream);
    if (rc != ONEBIT_OK) return rc;
    ONEBIT_CUDA_TRY(cudaMemcpyAsync(next_ids_host, D->ids, (size_t)batch * 8, cudaMemcpyDeviceToHost, s));
    ONEBIT_CUDA_TRY(cudaStreamSynchronize(s));
    return ONEBIT_OK;
}

int onebit_decoder_enable_p2p_allreduce(onebit_decoder* D, int rank, int nranks, void* const* peer_buffers, size_t buffer_bytes) {
    ONEBIT_REQUIRE(D && peer_buffers && D->tp > 1 && nranks == D->tp && rank >= 0 && rank < nranks && nranks <= kP2PMaxRanks,
                   "decoder_enable_p2p_allreduce: bad arguments (needs a tensor-parallel decoder, nranks == tp_size <= 8)");
    const size_t cap = buffer_bytes / sizeof(float) / (3 * (size_t)nranks);
    const size_t need = (size_t)D->cfg.max_batch * D->cfg.hidden_size;
    ONEBIT_REQUIRE(cap >= need, "decoder_enable_p2p_allreduce: symmetric buffer too small (needs 3 * nranks * max_batch * hidden floats)");
    if (!D->p2p_state) {
        ONEBIT_CUDA_TRY(cudaMalloc(&D->p2p_state, 8 * sizeof(unsigned)));
        ONEBIT_CUDA_TRY(cudaMemset(D->p2p_state, 0, 8 * sizeof(unsigned)));
    }
    P2PComm c = {};
    c.rank = rank; c.n = nranks; c.cap = cap;
    for (int r = 0; r < nranks; ++r) {
        ONEBIT_REQUIRE(peer_buffers[r] != nullptr, "decoder_enable_p2p_allreduce: NULL peer buffer");
        c.peer[r] = static_cast<float*>(peer_buffers[r]);
    }
    c.call_counter = D->p2p_state; c.cta_ticket = D->p2p_state + 1; c.error_flag = reinterpret_cast<int*>(D->p2p_state + 2);
    c.last_count = D->p2p_state + 4;
    D->p2p = c;
    D->p2p_on = true;
    return ONEBIT_OK;
}

int onebit_decoder_status(onebit_decoder* D, int* code) {
    ONEBIT_REQUIRE(D && code, "decoder_status: bad arguments");
    *code = 0;
    if (D->p2p_on) {  // 3 = a peer's data did not arrive in the one-shot all-reduce
        int e = 0;
        ONEBIT_CUDA_TRY(cudaMemcpy(&e, D->p2p.error_flag, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) { *code = 3; return ONEBIT_OK; }
    }
    if (D->persist) return persist_abort_flag(D->persist, code);
    return ONEBIT_OK;
}

int onebit_decoder_read_trace(onebit_decoder* D, uint64_t* out, int n) {
    ONEBIT_REQUIRE(D && out && n > 0, "decoder_read_trace: bad arguments");
    if (!D->persist) return 0;
    return persist_read_trace(D->persist, reinterpret_cast<unsigned long long*>(out), n);
}

int onebit_decoder_is_persistent(onebit_decoder* D) { return D && D->persist ? 1 : 0; }

}  // extern "C"

// Debug only (side builds with -DONEBIT_TRACE): stamps written by the kernels instantiated in THIS translation unit
// (fused stages); matvec_mma.cu has its own copy for the stand-alone GEMV.
extern "C" __attribute__((visibility("default"))) int onebit_debug_read_trace_decoder(long long* out8) {
#ifdef ONEBIT_TRACE
    return cudaMemcpyFromSymbol(out8, onebit::imma::g_trace, sizeof(long long) * 8) == cudaSuccess ? 0 : -2;
#else
    (void)out8;
    return -1;
#endif
}
extern "C" __attribute__((visibility("default"))) int onebit_debug_read_trace16(long long* out16) {
#ifdef ONEBIT_TRACE
    return cudaMemcpyFromSymbol(out16, onebit::imma::g_trace, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
#else
    (void)out16;
    return -1;
#endif
}
