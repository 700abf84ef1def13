// Fused "glue + bit-plane IMMA GEMV" stage, second generation: nothing block-wide between the dependency wait and the
// IMMA loop except ONE barrier, and nothing static left behind the wait.
//
// What the in-kernel clocks of the first generation (fused_gemv.cuh) showed on B200 (tools/trace_gemv2.py,
// profiles/r02_fused2_stage_trace.txt): of the ~16 k cycles of a stage, the IMMA loop is 1-4 k; the rest is a serial
// program on the dependency chain — three block reductions per token (LayerNorm sums, RMSNorm sum of squares, max |x'|),
// fp64 divisions and square roots in the statistics, frexp / ldexp calls, a shared-memory staging pass of x', launch
// arguments fetched cold from the constant bank, static vectors fetched from DRAM after the wait, and a warp's bulk
// copies serialised lane by lane by the compiler. This kernel removes them one by one:
//   * the PRODUCER's epilogue emits, per CTA and token, everything its consumer needs as partial results over the rows
//     the CTA owns anyway: base record (sum t, sum t^2, max t, min t) and resid record (sum r*t, sum r, sum r^2, max |r|),
//     r = the residual row the consumer will add. With mu / rstd from the first two,
//         sum (r + (t - mu) rstd)^2 = sum r^2 + 2 rstd (sum r t - mu sum r) + rstd^2 (sum t^2 - 2 mu sum t + N mu^2)
//     is the RMSNorm denominator of modeling_bitllama.py:67-81 without touching the vector, and
//         max |x'| <= (max |r| + max(|max t - mu|, |min t - mu|) rstd) * rms * max |ln_w * h|
//     bounds the quantiser range: the scale stays a power of two >= the exact one, so the 23-bit integers lose a few
//     low bits at most (tests: LLaMA-7B/13B widths, outlier channels) and can never saturate. The attention kernel
//     emits max |out| per head for the o_proj stage. Every warp reduces the <= ~150 records with shuffles on its own
//     (measured faster than one warp + broadcast), fp32 sums, fp64 only for the multiply-adds of E[t^2] - mu^2;
//     every record array exists kReplicas times so that 128+ CTAs do not poll the same few L2 lines;
//   * vectors between fused stages (t of o / gate / up / down, the residual stream) live in ITEM ORDER: the four
//     columns 8b + j of one 32-bit weight word a quantiser item needs are adjacent, one 16-byte load per item, x' goes
//     from registers to digits without a shared-memory pass (producers permute their 4-byte stores for free);
//   * static factors (input_factor, RMSNorm weight * input_factor, weight_scale) come from fp32 side tables built at
//     decoder creation in item order; they are loaded (round 0) or prefetched into L2 (later rounds) BEFORE the wait;
//   * the launch arguments are copied to shared memory once; all address arithmetic sits above the wait;
//   * every warp issues its own share of the sign-slice bulk copies; sum_k q goes through one shared-memory integer
//     atomic per warp; rounding uses the 1.5 * 2^23 constant instead of F2I; the epilogue needs one barrier for its
//     records; the shared-memory pitch of the sign rows is conflict-free at every K (fused::row_pad).
// Measured (profiles/r02_bench_default_line.json): LLaMA-7B batch 1 1.455 -> 1.207 ms/step, 9.33 -> 7.59 us per stage.
//
// Only the first stage of a step (token embedding, no producer) keeps one block reduction. Reference semantics are those
// of fused_gemv.cuh (bitnet.py:112-122 around modeling_bitllama.py:229-231,451-454,522-524,580,257).
#pragma once
#include "fused_gemv.cuh"

namespace onebit {
namespace fused2 {

using fused::kDigBlk;
using fused::EMBED_NORM;
using fused::PLAIN;
using fused::RESID_NORM;
using fused::SILU_MUL;

constexpr int kExt = 8;       // floats per (CTA, token) record pair: base (sum t, sum t^2, max t, min t) | resid (sum r*t, sum r, sum r^2, max |r|)
constexpr int kReplicas = 4;  // copies of every record array (readers pick one by CTA index)
constexpr int kRecLanes = 5;  // records per lane and round: 160 producer CTAs in one round trip
constexpr int kMaxRowWarps = 8;  // rows per CTA <= 256

struct Problem {
    const uint8_t* w;   // [n_rows][K/8]
    const float* g32;   // [n_rows] weight_scale as fp32
    const float* fp;    // [K] static input factor in item order (ln_w * h for the norm modes, h otherwise)
    float fmax;         // max |fp|
    float* t;           // [M][ld_t] fp32 out (= g * S @ x')
    float* stats;       // [ctas of this problem][M][2]  (sum t, sum t^2)
    float* ext;         // [2][max ctas][M][4]: base records, then resid records (ext_stride floats apart)
    int n_rows, ld_t, cta_begin;
};

struct Args {
    Problem p[3];
    int nprob, M, K, units, rows_per_cta, mode;
    const float* t_a; const float* stats_a; const float* ext_a; int ncta_a;
    const float* t_b; const float* stats_b; const float* ext_b; int ncta_b;
    const float* resid_in; float* resid_out;
    const __half* embed; const long long* ids;
    const float* x_plain; const float* x_amax; int n_amax;  // PLAIN: per-(token, head) max |x| records, or nullptr
    int ext_rep_stride;  // floats between two replicas of the record arrays
    int ext_stride_in, ext_stride_out;  // floats between the base and the resid record arrays of ext_a / of every p[].ext
    const float* resid_next; int resid_ld;  // residual rows the consumer of this stage will add to LN(t), or nullptr
    float ln_eps, rms_eps;
    int n_ln;  // rows of the producer's full layer (LayerNorm denominator); 0 = K
    int trace;  // ONEBIT_TRACE builds: this launch records CTA 0's stage clocks
    // vectors in ITEM ORDER (element 4 * it + b <-> column item_col(it) + 8 b): one 16-byte load per quantiser item
    int a_perm;      // t_a / t_b
    int rin_perm;    // resid_in
    int rout_perm;   // resid_out
    int rnext_perm;  // resid_next
    int t_perm;      // p[].t as written by this launch
    double inv_nln, inv_k;  // 1 / LayerNorm rows of the producer, 1 / K (no fp64 division on the device)
};

// position of column c in item order
__host__ __device__ __forceinline__ int perm_index(int c) { return ((((c >> 5) << 3) + (c & 7)) << 2) + ((c >> 3) & 3); }

// LayerNorm statistics from (sum, sum of squares): fp64 multiply-adds only (the cancellation in E[t^2] - mu^2 is the one
// place that needs them), reciprocal square root in fp32
__device__ __forceinline__ void finish_ln_fast(float s, float q, double inv_n, float eps, float& mean, float& rstd) {
    const double mu = (double)s * inv_n;
    const double var = fma(-mu, mu, (double)q * inv_n);
    mean = (float)mu;
    rstd = rsqrtf(fmaxf((float)var, 0.f) + eps);
}

#ifdef ONEBIT_TRACE
#define TR2(i) do { if (A.trace && blockIdx.x == 0 && threadIdx.x == 0) ::onebit::imma::g_trace[i] = clock64(); } while (0)
#else
#define TR2(i)
#endif

using fused::row_pad;  // K = 11008 (1376 bytes per row) with a fixed 32-byte pad landed every row on the same banks

constexpr int kThreads = 512;  // one CTA per SM (a 256-thread, two-CTAs-per-SM variant was measured: 1.69 vs 1.39 ms/step)
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 1024 / kThreads;  // quantiser items per thread and glue round (a round = 1024 items per CTA)

inline size_t smem_bytes(int M, int K, int rows_per_cta) {
    const size_t wbytes = (size_t)rows_per_cta * (K / 8 + row_pad(K / 8));
    const size_t red = (size_t)kWarps * rows_per_cta * 8 * 4;  // aliases the weight region after the main loop
    const size_t dig = (size_t)M * (K / 256) * 4 * kDigBlk;
    return (((wbytes > red ? wbytes : red) + 15) & ~(size_t)15) + dig + 64;
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float wmin(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// first column of item `it` (its four columns are c, c + 8, c + 16, c + 24)
__host__ __device__ __forceinline__ int item_col(int it) { return ((it >> 3) << 5) + (it & 7); }

// TILES = rows_per_cta / 16 (compile-time: the accumulators of every row tile live in registers across K).
template <int TILES>
__global__ void __launch_bounds__(kThreads, 1) fused_gemv2_kernel(const __grid_constant__ Args Ap) {
    extern __shared__ __align__(16) unsigned char smem[];
    // The launch parameters live in the constant bank, cold at every launch: scattered first touches cost an L2 round trip
    // each, in program order, on the dependency chain. One cooperative copy into shared memory up front instead.
    __shared__ __align__(16) Args sA;
#ifdef ONEBIT_TRACE
    const long long t_start = clock64();
#endif
    {
        static_assert(sizeof(Args) % 4 == 0, "Args is copied word by word");
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&Ap);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sA);
        for (int i = threadIdx.x; i < (int)(sizeof(Args) / 4); i += kThreads) dst[i] = src[i];
    }
    const Args& A = sA;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned long long s_qtot[2];
    __shared__ double s_invd[2];
    __shared__ double s_redd[16];
    __shared__ float s_redf[16];
    __shared__ float s_wrec[2][kMaxRowWarps][8];

    if (threadIdx.x == 0) {
        imma::mbar_init(&s_bar, kWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_qtot[0] = 0ull;
        s_qtot[1] = 0ull;
    }
    __syncthreads();  // sA complete (everything below reads the shared copy)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int M = A.M, K = A.K, Kb = K >> 3, pitch = Kb + row_pad(Kb);
    constexpr int kRowsCta = TILES * 16;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < 3; ++i)
        if (i < A.nprob && (int)blockIdx.x >= A.p[i].cta_begin) pi = i;
    const Problem& P = A.p[pi];
    const int cta = (int)blockIdx.x - P.cta_begin;
    const int row0 = cta * kRowsCta;
    const int rows_here = min(kRowsCta, P.n_rows - row0);
    const int items = A.units * 64;
    const int n_rounds = (items + kItems * kThreads - 1) / (kItems * kThreads);

    const size_t wregion = max((size_t)kRowsCta * pitch, (size_t)kWarps * kRowsCta * 8 * 4);
    unsigned char* Ws = smem;                                   // [rows][pitch] packed signs
    unsigned char* Bs = smem + ((wregion + 15) & ~(size_t)15);  // [M][units][4][kDigBlk] digits
    TR2(14);
    // ---- 1. static data before the dependency wait: this CTA's slice of the sign matrix, one bulk (TMA) copy per padded
    //         row. The compiler serialises a warp's bulk copies lane by lane (~35 cycles each), so every warp issues its own
    //         share (rows warp, warp + kWarps, ...): <= 12 copies per warp, asynchronous from then on ----
    {
        const int my_rows = rows_here > warp ? (rows_here - warp + kWarps - 1) / kWarps : 0;
        if (lane == 0) imma::mbar_expect_tx(&s_bar, (uint32_t)(my_rows * Kb));  // (also this warp's arrival)
        __syncwarp();
        if (lane < my_rows) {
            const int r = warp + lane * kWarps;
            imma::bulk_g2s(Ws + (size_t)r * pitch, P.w + (size_t)(row0 + r) * Kb, (uint32_t)Kb, &s_bar);
        }
    }
    TR2(15);
#ifdef ONEBIT_TRACE
    if (A.trace && blockIdx.x == 0 && threadIdx.x == 0) ::onebit::imma::g_trace[0] = t_start;
#endif
    float graw = 1.f;
    if (tid < rows_here) graw = P.g32[row0 + tid];
    // static factors of this thread's items: round 0 into registers, later rounds prefetched into L2 (the side tables are
    // touched once per step: without this their DRAM latency sits on the chain after the wait)
    const float4* fp4 = reinterpret_cast<const float4*>(P.fp);
    float4 fv0[kItems];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int it = i * kThreads + tid;
        fv0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (it < items) fv0[i] = fp4[it];
    }
    for (int it = kItems * kThreads + tid; it < items; it += 8 * kThreads)  // one 128-byte line per 8 items
        asm volatile("prefetch.global.L2 [%0];" ::"l"(fp4 + (it & ~7)));
    // everything the glue needs that does not depend on the producer: resolved before the wait
    const int mode = A.mode;
    const bool norm_mode = mode == EMBED_NORM || mode == RESID_NORM;
    const bool has_rec = mode == RESID_NORM || mode == SILU_MUL;
    const bool silu = mode == SILU_MUL;
    const float* a_base = mode == PLAIN ? A.x_plain : A.t_a;
    const float* b_base = mode == RESID_NORM ? A.resid_in : (silu ? A.t_b : nullptr);
    const bool a_perm = A.a_perm != 0, b_perm = (mode == RESID_NORM ? A.rin_perm : A.a_perm) != 0;
    // every record array exists kReplicas times: 128+ CTAs polling the same few L2 lines serialise on their slices
    const size_t rep_off = (size_t)(blockIdx.x & (kReplicas - 1)) * A.ext_rep_stride;
    const float4* rec_a = reinterpret_cast<const float4*>(A.ext_a + rep_off) + (size_t)lane * M;
    const float4* rec_b = reinterpret_cast<const float4*>((silu ? A.ext_b : A.ext_a + A.ext_stride_in) + rep_off) + (size_t)lane * M;
    const int n_a = A.ncta_a, n_b = silu ? A.ncta_b : A.ncta_a;
    const int qmul = (tid & 7) == 7 ? -1 : (1 << (7 - (tid & 7)));  // plane scale of this thread's items (j = tid % 8)
    imma::pdl_launch_dependents();
    TR2(12);
    imma::pdl_wait();
    TR2(1);
#ifdef ONEBIT_TRACE
    if (A.trace && blockIdx.x == 0 && threadIdx.x == 0) {  // when does the first global load after the wait come back?
        const float probe = *reinterpret_cast<const volatile float*>(P.fp);
        if (probe == 123.456f) ::onebit::imma::g_trace[15] = 0;
        ::onebit::imma::g_trace[13] = clock64();
    }
#endif

    // ---- 2. glue, per token: round 0 of the vector loads goes in flight, warp 0 turns the producer's records into the
    //         stage scalars (one barrier), then rounds of (x', quantise) with the next round's loads in flight ----
    struct Rnd { float va[kItems][4], vb[kItems][4]; float4 fv[kItems]; };
    for (int m = 0; m < M; ++m) {
        const float* a32 = a_base + (size_t)m * K;
        const unsigned short* erow =
            mode == EMBED_NORM ? reinterpret_cast<const unsigned short*>(A.embed + (size_t)A.ids[m] * K) : nullptr;
        const float* b32 = b_base ? b_base + (size_t)m * K : nullptr;
        auto load_round = [&](int rd, Rnd& R) {
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const int it = (rd * kItems + i) * kThreads + tid, c0 = item_col(it);
                R.fv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int b = 0; b < 4; ++b) { R.va[i][b] = 0.f; R.vb[i][b] = 0.f; }
                if (it < items) {
                    if (rd > 0) R.fv[i] = fp4[it];
                    if (erow) {
#pragma unroll
                        for (int b = 0; b < 4; ++b) R.va[i][b] = __uint_as_float((uint32_t)erow[c0 + 8 * b]);  // raw fp16 bits
                    } else if (a_perm) {
                        const float4 v = reinterpret_cast<const float4*>(a32)[it];
                        R.va[i][0] = v.x; R.va[i][1] = v.y; R.va[i][2] = v.z; R.va[i][3] = v.w;
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; ++b) R.va[i][b] = a32[c0 + 8 * b];
                    }
                    if (b32) {
                        if (b_perm) {
                            const float4 v = reinterpret_cast<const float4*>(b32)[it];
                            R.vb[i][0] = v.x; R.vb[i][1] = v.y; R.vb[i][2] = v.z; R.vb[i][3] = v.w;
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; ++b) R.vb[i][b] = b32[c0 + 8 * b];
                        }
                    }
                }
            }
        };
        Rnd cur;
        load_round(0, cur);
#pragma unroll
        for (int i = 0; i < kItems; ++i) cur.fv[i] = fv0[i];
        TR2(8);
        float mean_a = 0.f, rstd_a = 1.f, mean_b = 0.f, rstd_b = 1.f, bound = 0.f, rr = 1.f;
        if (has_rec || (mode == PLAIN && A.x_amax != nullptr)) {
            // every warp reduces the records on its own (no barrier: measured faster than one warp + broadcast)
            {
                if (has_rec) {
                    // acc_a = (sum t, sum t^2, max t, min t) of producer A; acc_b = RESID: (sum r t, sum r, sum r^2, max |r|),
                    // SILU: the base record of producer B. All loads of a round are issued before the first use.
                    float aa[4] = {0.f, 0.f, -INFINITY, INFINITY};
                    float ab[4] = {0.f, 0.f, silu ? -INFINITY : 0.f, silu ? INFINITY : 0.f};
                    for (int c0 = 0; c0 < max(n_a, n_b); c0 += 32 * kRecLanes) {
                        float4 ra[kRecLanes], rb[kRecLanes];
#pragma unroll
                        for (int i = 0; i < kRecLanes; ++i) {
                            const int c = c0 + lane + 32 * i;
                            ra[i] = make_float4(0.f, 0.f, -INFINITY, INFINITY);
                            rb[i] = make_float4(0.f, 0.f, silu ? -INFINITY : 0.f, silu ? INFINITY : 0.f);
                            if (c < n_a) ra[i] = rec_a[(size_t)(c0 + 32 * i) * M + m];
                            if (c < n_b) rb[i] = rec_b[(size_t)(c0 + 32 * i) * M + m];
                        }
#pragma unroll
                        for (int i = 0; i < kRecLanes; ++i) {
                            aa[0] += ra[i].x; aa[1] += ra[i].y; aa[2] = fmaxf(aa[2], ra[i].z); aa[3] = fminf(aa[3], ra[i].w);
                            ab[0] += rb[i].x; ab[1] += rb[i].y;
                            if (silu) { ab[2] = fmaxf(ab[2], rb[i].z); ab[3] = fminf(ab[3], rb[i].w); }
                            else { ab[2] += rb[i].z; ab[3] = fmaxf(ab[3], rb[i].w); }
                        }
                    }
                    TR2(9);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        aa[0] += __shfl_xor_sync(0xffffffffu, aa[0], o);
                        aa[1] += __shfl_xor_sync(0xffffffffu, aa[1], o);
                        aa[2] = fmaxf(aa[2], __shfl_xor_sync(0xffffffffu, aa[2], o));
                        aa[3] = fminf(aa[3], __shfl_xor_sync(0xffffffffu, aa[3], o));
                        ab[0] += __shfl_xor_sync(0xffffffffu, ab[0], o);
                        ab[1] += __shfl_xor_sync(0xffffffffu, ab[1], o);
                        const float o2 = __shfl_xor_sync(0xffffffffu, ab[2], o), o3 = __shfl_xor_sync(0xffffffffu, ab[3], o);
                        if (silu) { ab[2] = fmaxf(ab[2], o2); ab[3] = fminf(ab[3], o3); }
                        else { ab[2] += o2; ab[3] = fmaxf(ab[3], o3); }
                    }
                    TR2(10);
                    const int nln = A.n_ln > 0 ? A.n_ln : K;
                    finish_ln_fast(aa[0], aa[1], A.inv_nln, A.ln_eps, mean_a, rstd_a);
                    const float dev_a = fmaxf(aa[2] - mean_a, mean_a - aa[3]) * rstd_a;  // max |LN(t_a)|
                    if (!silu) {
                        const double mu = (double)mean_a, rho = (double)rstd_a;
                        const double c1 = fma(-mu, (double)ab[1], (double)ab[0]);                                      // sum r t - mu sum r
                        const double c2 = fma(mu * mu, (double)nln, fma(-2.0 * mu, (double)aa[0], (double)aa[1]));  // sum (t - mu)^2
                        const double ss = fma(rho * rho, c2, fma(2.0 * rho, c1, (double)ab[2]));
                        rr = rsqrtf(fmaxf((float)(ss * A.inv_k), 0.f) + A.rms_eps);  // LlamaRMSNorm
                        bound = (ab[3] + dev_a) * rr * P.fmax;
                    } else {
                        finish_ln_fast(ab[0], ab[1], A.inv_nln, A.ln_eps, mean_b, rstd_b);
                        const float dev_b = fmaxf(ab[2] - mean_b, mean_b - ab[3]) * rstd_b;
                        // |silu(x)| <= max(x_max, 0.2785) for x <= x_max
                        bound = fmaxf((aa[2] - mean_a) * rstd_a, 0.2785f) * dev_b * P.fmax;
                    }
                } else {
                    float xm = 0.f;
                    for (int c = lane; c < A.n_amax; c += 32) xm = fmaxf(xm, A.x_amax[(size_t)m * A.n_amax + c]);
                    bound = wmax(xm) * P.fmax;
                }
                TR2(11);
            }
        } else {
            // no producer records (token embedding; o_proj input without attention records): one pass over the vector and
            // the one block reduction of the step
            float part = 0.f, am = 0.f;
            for (int it = tid; it < items; it += kThreads) {
                const int c0 = item_col(it);
                const float4 f = fp4[it];
                const float f4[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float r = erow ? __half2float(__ushort_as_half(erow[c0 + 8 * b])) : a32[A.a_perm ? 4 * it + b : c0 + 8 * b];
                    part += r * r;
                    am = fmaxf(am, fabsf(r * f4[b]));
                }
            }
            double pd = wsum((double)part);
            am = wmax(am);
            if (lane == 0) { s_redd[warp] = pd; s_redf[warp] = am; }
            __syncthreads();
            pd = 0.0; am = 0.f;
            for (int w = 0; w < kWarps; ++w) { pd += s_redd[w]; am = fmaxf(am, s_redf[w]); }
            __syncthreads();
            if (mode == EMBED_NORM) rr = rsqrtf((float)(pd * A.inv_k) + A.rms_eps);
            bound = am * rr;
        }
        TR2(2);
        bound *= 1.0001f;  // rounding slack of the bound arithmetic (a violation below 2x is harmless: see imma_gemv.cuh)
        // power-of-two scale: bound = f * 2^e with f in [0.5, 1) (exponent field; no frexp / ldexp calls on the chain)
        int e = 0;
        if (bound > 0.f && bound < 3.0e38f) e = min(max((int)((__float_as_uint(bound) >> 23) & 0xffu) - 126, -100), 100);
        const float S = __uint_as_float((uint32_t)(127 + 22 - e) << 23);
        const float cS = rr * S;
        int qs = 0;
        unsigned char* dg = Bs + (size_t)m * A.units * 4 * kDigBlk;
#pragma unroll 1
        for (int rd = 0; rd < n_rounds; ++rd) {
            Rnd nxt;
            if (rd + 1 < n_rounds) load_round(rd + 1, nxt);
            auto& va = cur.va;
            auto& vb = cur.vb;
            auto& fv = cur.fv;
            TR2(3);
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const int it = (rd * kItems + i) * kThreads + tid, c0 = item_col(it);
                const float f4[4] = {fv[i].x, fv[i].y, fv[i].z, fv[i].w};
                float y[4];  // x' * S
                if (norm_mode) {
                    float r[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (mode == EMBED_NORM) r[b] = __half2float(__ushort_as_half((unsigned short)__float_as_uint(va[i][b])));
                        else r[b] = vb[i][b] + (va[i][b] - mean_a) * rstd_a;  // residual + LayerNorm(o / down output)
                        y[b] = r[b] * cS * f4[b];
                    }
                    if (blockIdx.x == 0 && it < items) {  // the new residual stream
                        float* ro = A.resid_out + (size_t)m * K;
                        if (A.rout_perm) reinterpret_cast<float4*>(ro)[it] = make_float4(r[0], r[1], r[2], r[3]);
                        else {
#pragma unroll
                            for (int b = 0; b < 4; ++b) ro[c0 + 8 * b] = r[b];
                        }
                    }
                } else if (mode == SILU_MUL) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float ga = (va[i][b] - mean_a) * rstd_a, ub = (vb[i][b] - mean_b) * rstd_b;
                        y[b] = __fdividef(ga, 1.f + __expf(-ga)) * ub * f4[b] * S;  // silu(gate) * up * input_factor
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b) y[b] = va[i][b] * f4[b] * S;
                }
                // quantise straight from registers: item = (unit, t, word, plane j) -> 4 columns -> 4 digit registers
                if (it < items) {
                    const int j = it & 7, ws = (it >> 3) & 1, tt = (it >> 4) & 3, u = it >> 6;
                    uint32_t dw[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        // round to nearest even through the 1.5 * 2^23 magic constant (|y| <= 2^22): no F2I
                        const int q = __float_as_int(y[b] + 12582912.f) - 0x4B400000;
                        qs += q;
                        dw[b] = ((uint32_t)(q * qmul) + 0x00808080u) ^ 0x00808080u;  // q << (7 - j), or -q for plane 7
                    }
                    const uint32_t t0 = __byte_perm(dw[0], dw[1], 0x5140), t1 = __byte_perm(dw[2], dw[3], 0x5140);
                    const uint32_t t2 = __byte_perm(dw[0], dw[1], 0x7362), t3 = __byte_perm(dw[2], dw[3], 0x7362);
                    uint32_t* dst = reinterpret_cast<uint32_t*>(dg + ((size_t)u * 4 + (j >> 1)) * kDigBlk) + tt * 4 + (j & 1) * 2 + ws;
                    dst[0] = __byte_perm(t0, t1, 0x5410);
                    dst[16] = __byte_perm(t0, t1, 0x7632);
                    dst[32] = __byte_perm(t2, t3, 0x5410);
                    dst[48] = __byte_perm(t2, t3, 0x7632);
                }
            }
            if (rd + 1 < n_rounds) cur = nxt;
        }
        long long q64 = qs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q64 += __shfl_xor_sync(0xffffffffu, q64, o);
        if (lane == 0) atomicAdd(&s_qtot[m], (unsigned long long)q64);  // integer: order-independent
        // a non-finite bound (NaN / Inf anywhere upstream reaches it through the records) poisons the outputs of this token,
        // as the reference's fp arithmetic would, instead of being quantised into finite garbage
        if (tid == 0) s_invd[m] = bound == bound && bound < 3.0e38f ? (double)__uint_as_float((uint32_t)(127 + e - 22) << 23) : (double)NAN;
        TR2(4);
    }
    // residual rows of the consumer (resid records): in flight across the IMMA loop
    float rnext[2] = {0.f, 0.f};
    if (A.resid_next != nullptr && tid < rows_here) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
            if (m < M) rnext[m] = A.resid_next[(size_t)m * A.resid_ld + (A.rnext_perm ? perm_index(row0 + tid) : row0 + tid)];
    }
    __syncthreads();                 // digits + per-token meta visible
    imma::mbar_wait(&s_bar, 0);      // signs landed (issued long ago)
    TR2(5);

    // ---- 3. IMMA loop: this warp's K units; B fragments in registers across all row tiles. Few row tiles per CTA (o_proj,
    //         down_proj) mean few independent accumulator chains per warp: even / odd planes get their own there ----
    constexpr int NACC = TILES <= 2 ? 2 : 1;
    int accs[NACC][TILES][4];
#pragma unroll
    for (int z = 0; z < NACC; ++z)
#pragma unroll
        for (int r = 0; r < TILES; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) accs[z][r][i] = 0;
    for (int u = warp; u < A.units; u += kWarps) {
        uint4 bv[4];
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            const int m = g >> 2;
            bv[jp] = make_uint4(0u, 0u, 0u, 0u);
            if (m < M)
                bv[jp] = *reinterpret_cast<const uint4*>(Bs + ((size_t)(m * A.units + u) * 4 + jp) * kDigBlk + ((g & 3) * 4 + t4) * 16);
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
                    imma::imma16832(accs[NACC == 2 ? jj : 0][r], a0, a1, a2, a3, jj ? bv[jp].z : bv[jp].x, jj ? bv[jp].w : bv[jp].y);
                }
        }
    }
    int acc[TILES][4];
#pragma unroll
    for (int r = 0; r < TILES; ++r)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[r][i] = NACC == 2 ? accs[0][r][i] + accs[NACC - 1][r][i] : accs[0][r][i];

    // ---- 4. combine the K split across warps (red aliases the weight region), finalise, store, records ----
    TR2(6);
    __syncthreads();
    int* red = reinterpret_cast<int*>(smem);  // [kWarps][kRowsCta][8]
#pragma unroll
    for (int r = 0; r < TILES; ++r) {
        int* base = red + ((size_t)warp * kRowsCta + 16 * r) * 8 + 2 * t4;
        *reinterpret_cast<int2*>(base + (size_t)g * 8) = make_int2(acc[r][0], acc[r][1]);
        *reinterpret_cast<int2*>(base + (size_t)(g + 8) * 8) = make_int2(acc[r][2], acc[r][3]);
    }
    __syncthreads();
    const int row_warps = (rows_here + 31) >> 5;
    if (warp < row_warps) {
        for (int m = 0; m < M; ++m) {
            const bool act = tid < rows_here;
            float val = 0.f;
            if (act) {
                int4 a = make_int4(0, 0, 0, 0);
#pragma unroll
                for (int w = 0; w < kWarps; ++w) {
                    const int4 v = *reinterpret_cast<const int4*>(red + ((size_t)w * kRowsCta + tid) * 8 + 4 * m);
                    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                }
                const long long V = (((long long)a.w * 256 + a.z) * 256 + a.y) * 256 + a.x;  // 128 * sum_{bit=1} q
                val = (float)((double)((long long)s_qtot[m] - 2 * (V >> 7)) * s_invd[m]) * graw;
                P.t[(size_t)m * P.ld_t + (A.t_perm ? perm_index(row0 + tid) : row0 + tid)] = val;
            }
            const float r = rnext[m];
            float v0 = val, v1 = val * val, v2 = r * val, v3 = r, v4 = r * r;
            float vmx = act ? val : -INFINITY, vmn = act ? val : INFINITY, vrm = fabsf(r);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 += __shfl_xor_sync(0xffffffffu, v0, o);
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                v4 += __shfl_xor_sync(0xffffffffu, v4, o);
                vmx = fmaxf(vmx, __shfl_xor_sync(0xffffffffu, vmx, o));
                vmn = fminf(vmn, __shfl_xor_sync(0xffffffffu, vmn, o));
                vrm = fmaxf(vrm, __shfl_xor_sync(0xffffffffu, vrm, o));
            }
            if (lane == 0) {
                float* w8 = s_wrec[m][warp];
                w8[0] = v0; w8[1] = v1; w8[2] = v2; w8[3] = v3; w8[4] = v4; w8[5] = vmx; w8[6] = vmn; w8[7] = vrm;
            }
        }
    }
    __syncthreads();
    if (tid < M * kReplicas) {
        const int m = tid % M, rep = tid / M;
        double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        float mx = -INFINITY, mn = INFINITY, rm = 0.f;
        for (int w = 0; w < row_warps; ++w) {  // fixed order: deterministic
            const float* w8 = s_wrec[m][w];
#pragma unroll
            for (int i = 0; i < 5; ++i) s[i] += (double)w8[i];
            mx = fmaxf(mx, w8[5]); mn = fminf(mn, w8[6]); rm = fmaxf(rm, w8[7]);
        }
        if (rep == 0) *reinterpret_cast<float2*>(P.stats + ((size_t)cta * M + m) * 2) = make_float2((float)s[0], (float)s[1]);
        float* eb = P.ext + (size_t)rep * A.ext_rep_stride;
        reinterpret_cast<float4*>(eb)[(size_t)cta * M + m] = make_float4((float)s[0], (float)s[1], mx, mn);
        *reinterpret_cast<float4*>(eb + A.ext_stride_out + ((size_t)cta * M + m) * 4) = make_float4((float)s[2], (float)s[3], (float)s[4], rm);
    }
    TR2(7);
}

// side tables of one decoder layer (one block per BitLinear): fp[4 * it + b] = h[c] (* ln_w[c]) for column
// c = item_col(it) + 8 b, fmax = max |fp|, g32 = weight_scale as fp32
struct TableJob {
    const void* h; const void* lnw; const void* g;
    float* fp; float* g32; float* fmax;
    int k, n;
};
struct TableJobs { TableJob j[7]; };
template <typename TP>
__global__ void __launch_bounds__(1024) side_tables_kernel(const __grid_constant__ TableJobs J) {
    __shared__ float sh[32];
    const TableJob& T = J.j[blockIdx.x];
    const TP* h = static_cast<const TP*>(T.h);
    const TP* lnw = static_cast<const TP*>(T.lnw);
    const TP* g = static_cast<const TP*>(T.g);
    float am = 0.f;
    for (int idx = threadIdx.x; idx < T.k; idx += blockDim.x) {
        const int c = item_col(idx >> 2) + 8 * (idx & 3);
        float v = to_f32(h[c]);
        if (lnw != nullptr) v *= to_f32(lnw[c]);
        T.fp[idx] = v;
        am = fmaxf(am, fabsf(v));
    }
    for (int i = threadIdx.x; i < T.n; i += blockDim.x) T.g32[i] = to_f32(g[i]);
    am = wmax(am);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = am;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
        *T.fmax = r;
    }
}

}  // namespace fused2
}  // namespace onebit
