// Fused "glue + bit-plane IMMA GEMV" stage of the decode step: ONE kernel per BitLinear group.
//
// The stand-alone GEMV (imma_gemv.cuh) gets its activation digits from a separate single-CTA glue kernel, so a
// decoder layer is 9 dependent launches of ~6 us each (profiles/r01_launches_multikernel_v1.csv). Measured on
// B200, a dependent kernel boundary costs ~1.1 us and every extra stage adds its own L2 round trips and
// reductions, so this kernel folds the glue INTO the GEMV:
//   * <= ~148 fat CTAs (one wave): each owns TILES x 16 output rows of one problem over the whole of K, with its
//     slice of the sign matrix bulk-copied (TMA) into shared memory BEFORE the programmatic-dependency wait;
//   * every CTA rebuilds the BitLinear input redundantly from the producer's fp32 outputs (LayerNorm-apply from
//     the per-CTA partial sums, residual add, RMSNorm, SiLU*up, * input_factor), quantises it to the 23-bit
//     integer digits in shared memory (~14 instructions per column per CTA: cheap now that there are <= 148
//     CTAs instead of 128-688), then runs the IMMA loop with the B fragments held in registers across all of
//     its row tiles (B traffic / TILES), and emits g*t plus LayerNorm partials for the next stage.
// A decoder layer becomes 5 launches: qkv | attention | o | gate,up | down.
#pragma once
#include "imma_gemv.cuh"

namespace onebit {
namespace fused {

using imma::QMeta;
constexpr int kThreads = 512;
constexpr int kWarps = 16;
// bytes of padding per weight row in shared memory: pitch = 32 bytes past a multiple of 128, so that the LDS.64 of 8 rows x
// 4 lanes covers all 32 banks (2 wavefronts) at every K (a fixed 32-byte pad does that for K = 4096 only)
__host__ __device__ inline int row_pad(int kb) { return (32 - (kb & 127) + 128) & 127; }

// shared-memory staging of x' is padded by 8 floats per 32 so that the quantiser's stride-8 gathers are conflict-free
__host__ __device__ inline int xs_pad(int k) { return k + ((k >> 5) << 3); }
constexpr int kDigBlk = 288;  // bytes per (unit, plane-pair) block of digits in shared memory (256 + pad: conflict-free stores)

enum Mode { EMBED_NORM = 0, RESID_NORM = 1, SILU_MUL = 2, PLAIN = 3 };

struct Problem {
    const uint8_t* w;  // [n_rows][K/8]
    const void* g;     // [n_rows] TP
    const void* h;     // [K] TP  (input_factor of this projection)
    float* t;          // [M][n_rows] fp32 out (= g * S @ (h*x))
    float* stats;      // [ctas of this problem][M][2]
    int n_rows;
    int ld_t;          // row stride of t (>= n_rows; tensor-parallel shards pad it to the consumer's K)
    int cta_begin;
};

struct Args {
    Problem p[3];
    int nprob, M, K, units, rows_per_cta;
    int mode;
    const float* t_a; const float* stats_a; int ncta_a;  // producer A: g*t and its per-CTA (sum, sumsq) partials
    const float* t_b; const float* stats_b; int ncta_b;  // producer B (up_proj) for SILU_MUL
    const float* resid_in; float* resid_out;             // fp32 residual stream (ping-pong)
    const __half* embed; const long long* ids;           // EMBED_NORM
    const void* ln_w;                                     // RMSNorm weight (TP)
    const float* x_plain;                                 // PLAIN
    float ln_eps, rms_eps;
    int n_ln;             // rows of the producer's FULL layer (LayerNorm denominator); 0 = K
    int stats_from_data;  // RESID_NORM: t_a was all-reduced across ranks -> take (sum, sumsq) from the data itself
};

inline size_t smem_bytes(int M, int K, int rows_per_cta) {
    const size_t wbytes = (size_t)rows_per_cta * (K / 8 + row_pad(K / 8));
    const size_t dig = (size_t)M * (K / 256) * 4 * kDigBlk;
    const size_t xs = (size_t)xs_pad(K) * 4 + 64;
    const size_t red = (size_t)kWarps * rows_per_cta * 8 * 4;  // aliases the weight region after the main loop
    return (wbytes > red ? wbytes : red) + dig + xs + 64;
}

// block-wide sum of NVAL doubles (fixed order: deterministic); sh needs 33 * NVAL doubles
template <int NVAL>
__device__ __forceinline__ void block_sum(double (&v)[NVAL], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NVAL; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NVAL; ++i) sh[warp * NVAL + i] = v[i];
    __syncthreads();
    if (warp == 0) {
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
__device__ __forceinline__ float block_max(float v, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < nw; ++w) r = fmaxf(r, sh[w]);
    return r;
}

// Parameter vectors are loaded RAW (4 elements) and converted at the point of use: converting right after the load
// makes every load wait for its own round trip (measured: 13 k cycles for 11 x 3 loads instead of ~2 k).
template <typename TP> struct Raw4;
template <> struct Raw4<float> { float4 v; };
template <> struct Raw4<__half> { uint2 v; };
template <> struct Raw4<__nv_bfloat16> { uint2 v; };
template <typename TP>
__device__ __forceinline__ Raw4<TP> ldraw4(const TP* p, int i4) {
    Raw4<TP> r;
    r.v = reinterpret_cast<const decltype(r.v)*>(p)[i4];
    return r;
}
__device__ __forceinline__ float4 cvt4(const Raw4<float>& r) { return r.v; }
__device__ __forceinline__ float4 cvt4(const Raw4<__half>& r) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 cvt4(const Raw4<__nv_bfloat16>& r) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.v.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

template <typename TP>
__device__ __forceinline__ float4 ldp4(const TP* p, int i4);
template <>
__device__ __forceinline__ float4 ldp4<float>(const float* p, int i4) { return reinterpret_cast<const float4*>(p)[i4]; }
template <>
__device__ __forceinline__ float4 ldp4<__half>(const __half* p, int i4) {
    const uint2 r = reinterpret_cast<const uint2*>(p)[i4];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 ldp4<__nv_bfloat16>(const __nv_bfloat16* p, int i4) {
    const uint2 r = reinterpret_cast<const uint2*>(p)[i4];
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ void stat_partials(const float* stats, int ncta, int M, int m, double& s, double& q) {
    for (int c = threadIdx.x; c < ncta; c += blockDim.x) {
        const float2 p = *reinterpret_cast<const float2*>(stats + ((size_t)c * M + m) * 2);
        s += (double)p.x;
        q += (double)p.y;
    }
}
__device__ __forceinline__ void finish_ln(double s, double q, int n, float eps, float& mean, float& rstd) {
    const double mu = s / (double)n;
    const double var = fmax(q / (double)n - mu * mu, 0.0);
    mean = (float)mu;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// TILES = rows_per_cta / 16 (compile-time: the accumulators of every row tile live in registers across K);
// NV4 = ceil(K / 4 / 256): float4 values per thread of each prologue array.
template <typename TP, int NT, int TILES, int NV4>
__global__ void __launch_bounds__(kThreads, 1) fused_gemv_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ double shd[33 * 4];
    __shared__ float shf[kWarps];
    __shared__ long long shl[kWarps];
    __shared__ float s_inv[2 * NT];
    __shared__ long long s_qtot[2 * NT];
    __shared__ double s_invd[2 * NT];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int M = A.M, K = A.K, K4 = K >> 2, Kb = K >> 3, pitch = Kb + row_pad(Kb);
    constexpr int kRowsCta = TILES * 16;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < 3; ++i)
        if (i < A.nprob && (int)blockIdx.x >= A.p[i].cta_begin) pi = i;
    const Problem& P = A.p[pi];
    const int cta = (int)blockIdx.x - P.cta_begin;
    const int row0 = cta * kRowsCta;
    const int rows_here = min(kRowsCta, P.n_rows - row0);

    const size_t wregion = max((size_t)kRowsCta * pitch, (size_t)kWarps * kRowsCta * 8 * NT * 4);
    unsigned char* Ws = smem;                                        // [rows][pitch] packed signs
    unsigned char* Bs = smem + ((wregion + 15) & ~(size_t)15);       // [M][units][1024] digits
    float* xs = reinterpret_cast<float*>(Bs + (size_t)M * A.units * 4 * kDigBlk);  // [xs_pad(K)] staging of x'
    TR(0);
    // ---- 1. this CTA's slice of the sign matrix -> shared memory (static data: before the dependency wait) ----
    if (tid == 0) {
        imma::mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) imma::mbar_expect_tx(&s_bar, (uint32_t)(rows_here * Kb));
        __syncwarp();
        for (int r = lane; r < rows_here; r += 32)
            imma::bulk_g2s(Ws + (size_t)r * pitch, P.w + (size_t)(row0 + r) * Kb, (uint32_t)Kb, &s_bar);
    }
    // static parameter vectors (input_factor, RMSNorm weight) are loaded raw BEFORE the dependency wait as well
    const TP* hptr = static_cast<const TP*>(P.h);
    const TP* lnw = static_cast<const TP*>(A.ln_w);
    const bool norm_mode = A.mode == EMBED_NORM || A.mode == RESID_NORM;
    Raw4<TP> vh[NV4], vw[NV4];
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int i4 = i * kThreads + tid;
        if (i4 < K4) {
            vh[i] = ldraw4<TP>(hptr, i4);
            if (norm_mode) vw[i] = ldraw4<TP>(lnw, i4);
        }
    }
    TP graw = from_f32<TP>(1.f);  // weight_scale of the row this thread finalises (static: load now, convert at use)
    if (tid < rows_here) graw = static_cast<const TP*>(P.g)[row0 + tid];
    imma::pdl_launch_dependents();
    imma::pdl_wait();
    TR(1);

    // ---- 2. rebuild the BitLinear input (glue), per token: x' -> digits in shared memory ----
    // Every global load of a token is issued before the first use (one exposed L2 round trip), NV4 float4 per array.
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < M; ++m) {
        float4 va[NV4], vb[NV4];
        Raw4<__half> ve[NV4];
        double st[4] = {0.0, 0.0, 0.0, 0.0};
        {
            const __half* erow = A.mode == EMBED_NORM ? A.embed + (size_t)A.ids[m] * K : nullptr;
            const float4* a4 = reinterpret_cast<const float4*>(
                (A.mode == PLAIN ? A.x_plain : A.t_a) + (size_t)m * K);
            const float4* b4 = A.mode == RESID_NORM ? reinterpret_cast<const float4*>(A.resid_in + (size_t)m * K)
                               : (A.mode == SILU_MUL ? reinterpret_cast<const float4*>(A.t_b + (size_t)m * K) : nullptr);
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                const int i4 = i * kThreads + tid;
                va[i] = z4; vb[i] = z4;
                if (i4 < K4) {
                    if (erow) ve[i] = ldraw4<__half>(erow, i4);
                    else va[i] = a4[i4];
                    if (b4) vb[i] = b4[i4];
                }
            }
        }
        // (the partial-sum loads come after the big loads in program order: their fp64 conversion stalls the warp)
        if ((A.mode == RESID_NORM && !A.stats_from_data) || A.mode == SILU_MUL) stat_partials(A.stats_a, A.ncta_a, M, m, st[0], st[1]);
        if (A.mode == SILU_MUL) stat_partials(A.stats_b, A.ncta_b, M, m, st[2], st[3]);
        float mean_a = 0.f, rstd_a = 1.f, mean_b = 0.f, rstd_b = 1.f;
        if (A.mode == RESID_NORM || A.mode == SILU_MUL) {
            if (A.mode == RESID_NORM && A.stats_from_data) {  // tensor parallel: statistics of the reduced vector
                st[0] = st[1] = 0.0;
#pragma unroll
                for (int i = 0; i < NV4; ++i)
                    if (i * kThreads + tid < K4) {
                        st[0] += (double)((va[i].x + va[i].y) + (va[i].z + va[i].w));
                        st[1] += (double)((va[i].x * va[i].x + va[i].y * va[i].y) + (va[i].z * va[i].z + va[i].w * va[i].w));
                    }
            }
            block_sum<4>(st, shd);
            const int nln = A.n_ln > 0 ? A.n_ln : K;
            finish_ln(st[0], st[1], nln, A.ln_eps, mean_a, rstd_a);
            if (A.mode == SILU_MUL) finish_ln(st[2], st[3], nln, A.ln_eps, mean_b, rstd_b);
        }
        TR(2);
        float am = 0.f;
        if (norm_mode) {
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                const int i4 = i * kThreads + tid;
                if (i4 < K4) {
                    float4 r = A.mode == EMBED_NORM ? cvt4(ve[i]) : va[i];
                    if (A.mode == RESID_NORM)  // residual + LayerNorm(o / down output)
                        r = make_float4(vb[i].x + (r.x - mean_a) * rstd_a, vb[i].y + (r.y - mean_a) * rstd_a,
                                        vb[i].z + (r.z - mean_a) * rstd_a, vb[i].w + (r.w - mean_a) * rstd_a);
                    if (blockIdx.x == 0) reinterpret_cast<float4*>(A.resid_out + (size_t)m * K)[i4] = r;
                    va[i] = r;
                    part += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
                }
            }
            double ss[1] = {(double)part};
            block_sum<1>(ss, shd);
            const float rr = rsqrtf((float)(ss[0] / (double)K) + A.rms_eps);  // LlamaRMSNorm
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                if (i * kThreads + tid < K4) {
                    const float4 w4 = cvt4(vw[i]);
                    va[i].x *= rr * w4.x; va[i].y *= rr * w4.y; va[i].z *= rr * w4.z; va[i].w *= rr * w4.w;
                }
            }
        } else if (A.mode == SILU_MUL) {
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                const float ga[4] = {(va[i].x - mean_a) * rstd_a, (va[i].y - mean_a) * rstd_a, (va[i].z - mean_a) * rstd_a,
                                     (va[i].w - mean_a) * rstd_a};
                const float ub[4] = {(vb[i].x - mean_b) * rstd_b, (vb[i].y - mean_b) * rstd_b, (vb[i].z - mean_b) * rstd_b,
                                     (vb[i].w - mean_b) * rstd_b};
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = __fdividef(ga[e], 1.f + __expf(-ga[e])) * ub[e];  // silu(gate) * up
                va[i] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
            const int i4 = i * kThreads + tid;
            if (i4 < K4) {
                const float4 h4 = cvt4(vh[i]);
                const float4 v = make_float4(va[i].x * h4.x, va[i].y * h4.y, va[i].z * h4.z, va[i].w * h4.w);
                *reinterpret_cast<float4*>(xs + xs_pad(4 * i4)) = v;
                am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        }
        TR(3);
        am = block_max(am, shf);  // includes the barrier that publishes xs
        int e = 0;
        if (am > 0.f && am < 3.0e38f) frexpf(am, &e);
        const float S = ldexpf(1.0f, 22 - e);
        // quantise: item = (unit, t, word, plane j) -> 4 columns -> 4 digit registers (see imma_gemv.cuh)
        int qs = 0;
        unsigned char* dg = Bs + (size_t)m * A.units * 4 * kDigBlk;
        const int items = A.units * 64;
        for (int it = tid; it < items; it += kThreads) {
            const int j = it & 7, ws = (it >> 3) & 1, tt = (it >> 4) & 3, u = it >> 6;
            const float* xr = xs + xs_pad(u * 256 + 64 * tt + 32 * ws) + j;  // the 32 columns of one word: one pad group
            uint32_t dw[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int q = __float2int_rn(xr[8 * b] * S);
                qs += q;
                const int v = (j == 7) ? -q : (q << (7 - j));
                dw[b] = ((uint32_t)v + 0x00808080u) ^ 0x00808080u;
            }
            const uint32_t t0 = __byte_perm(dw[0], dw[1], 0x5140), t1 = __byte_perm(dw[2], dw[3], 0x5140);
            const uint32_t t2 = __byte_perm(dw[0], dw[1], 0x7362), t3 = __byte_perm(dw[2], dw[3], 0x7362);
            uint32_t* dst = reinterpret_cast<uint32_t*>(dg + ((size_t)u * 4 + (j >> 1)) * kDigBlk) + tt * 4 + (j & 1) * 2 + ws;
            dst[0] = __byte_perm(t0, t1, 0x5410);
            dst[16] = __byte_perm(t0, t1, 0x7632);
            dst[32] = __byte_perm(t2, t3, 0x5410);
            dst[48] = __byte_perm(t2, t3, 0x7632);
        }
        TR(4);
        long long q64 = qs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q64 += __shfl_xor_sync(0xffffffffu, q64, o);
        __syncthreads();  // xs is rewritten by the next token; shl reuse
        if (lane == 0) shl[warp] = q64;
        __syncthreads();
        if (tid == 0) {
            long long tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += shl[w];
            s_qtot[m] = tot;
            s_invd[m] = ldexp(1.0, e - 22);
        }
    }
    __syncthreads();                 // digits + per-token meta visible
    imma::mbar_wait(&s_bar, 0);      // weights landed (issued long ago)
    TR(5);

    // ---- 3. IMMA loop: this warp's K units; B fragments in registers across all row tiles ----
    int acc[TILES][NT][4];
#pragma unroll
    for (int r = 0; r < TILES; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][nt][i] = 0;
    for (int u = warp; u < A.units; u += kWarps) {
        uint4 bv[4][NT];
#pragma unroll
        for (int jp = 0; jp < 4; ++jp)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int m = 2 * nt + (g >> 2);
                bv[jp][nt] = make_uint4(0u, 0u, 0u, 0u);
                if (m < M)
                    bv[jp][nt] = *reinterpret_cast<const uint4*>(Bs + ((size_t)(m * A.units + u) * 4 + jp) * kDigBlk +
                                                                 ((g & 3) * 4 + t4) * 16);
            }
        const unsigned char* wu = Ws + (size_t)u * 32 + 8 * t4;
#pragma unroll
        for (int r = 0; r < TILES; ++r) {
            const uint2 w0 = *reinterpret_cast<const uint2*>(wu + (size_t)(16 * r + g) * pitch);
            const uint2 w1 = *reinterpret_cast<const uint2*>(wu + (size_t)(16 * r + g + 8) * pitch);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const uint32_t mask = 0x01010101u << (2 * jp + jj);
                    const uint32_t a0 = imma::plane(w0.x, mask), a1 = imma::plane(w1.x, mask);
                    const uint32_t a2 = imma::plane(w0.y, mask), a3 = imma::plane(w1.y, mask);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        imma::imma16832(acc[r][nt], a0, a1, a2, a3, jj ? bv[jp][nt].z : bv[jp][nt].x,
                                        jj ? bv[jp][nt].w : bv[jp][nt].y);
                }
        }
    }

    // ---- 4. combine the K split across warps (red aliases the weight region), finalise, store, partial stats ----
    TR(6);
    __syncthreads();
    int* red = reinterpret_cast<int*>(smem);  // [kWarps][kRowsCta][8 * NT]
    constexpr int kCols = 8 * NT;
#pragma unroll
    for (int r = 0; r < TILES; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            int* base = red + ((size_t)warp * kRowsCta + 16 * r) * kCols + 8 * nt + 2 * t4;
            *reinterpret_cast<int2*>(base + (size_t)g * kCols) = make_int2(acc[r][nt][0], acc[r][nt][1]);
            *reinterpret_cast<int2*>(base + (size_t)(g + 8) * kCols) = make_int2(acc[r][nt][2], acc[r][nt][3]);
        }
    __syncthreads();
    for (int m = 0; m < M; ++m) {
        double st[2] = {0.0, 0.0};
        if (tid < rows_here) {
            const int r = tid;
            int4 a = make_int4(0, 0, 0, 0);
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int4 v = *reinterpret_cast<const int4*>(red + ((size_t)w * kRowsCta + r) * kCols + 4 * m);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            const long long V = (((long long)a.w * 256 + a.z) * 256 + a.y) * 256 + a.x;  // 128 * sum_{bit=1} q
            const float val = (float)((double)(s_qtot[m] - 2 * (V >> 7)) * s_invd[m]) * to_f32(graw);
            P.t[(size_t)m * P.ld_t + row0 + r] = val;
            st[0] = (double)val;
            st[1] = (double)val * (double)val;
        }
        block_sum<2>(st, shd);
        if (tid == 0) *reinterpret_cast<float2*>(P.stats + ((size_t)cta * M + m) * 2) = make_float2((float)st[0], (float)st[1]);
    }
    (void)s_inv;
    TR(7);
}

}  // namespace fused
}  // namespace onebit
