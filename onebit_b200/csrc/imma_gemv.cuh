// Bit-plane IMMA packed-sign GEMV (decode path, M <= 8 tokens) — shared between the stand-alone C-ABI
// entry points (matvec_mma.cu) and the fused decode step (decoder.cu).
//
//     t[m][n] = sum_k s(n,k) * x'[m][k],   x' = h * x  (bitnet.py:113-115), optionally * g[n] (:116)
//
// The sign matrix is read from HBM once at 1 bit/element and is never expanded in memory.
//
// Why integer tensor cores (profiles/r01_ubench_mma_sync.txt, measured on B200): the warp-level mma.sync
// path sustains 0.47 HMMA.16816 or 0.48 IMMA.16832 per clock per SM. With 16 weight rows as the MMA's M
// side that is 120 weight-bits/clk/SM in fp16 (67 % of what HBM can deliver) but 246 in int8 (137 %);
// CUDA cores need >= 1 lane-op per bit (~35 %). Only the int8 shape leaves the kernel HBM-bound.
//
// Bit-plane trick (no shifts: 0.25 ALU op per weight). For bit j of every byte of a 32-bit weight word,
// `w & (0x01010101 << j)` already is an int8x4 A-fragment register: byte b = bit(8b+j) * 2^j (-128*bit for
// j = 7). The activation side absorbs the plane scale: x' is quantised per token to a 23-bit integer q with
// a power-of-two scale (error 2^-23 of the token's max), column k (plane j = k % 8) carries v = q << (7-j)
// (v = -q for j = 7), and v is split into four balanced base-256 digits that occupy four B columns of the
// MMA. Every plane then accumulates into ONE int32 accumulator per digit and
//     sum_d 256^d * acc_d = 128 * sum_{bit=1} q   exactly,   sum_k s*q = sum_k q - 2 * sum_{bit=1} q.
// Integer arithmetic end to end: bit-reproducible and independent of how K is split across warps.
//
// Activation digits are produced ONCE per token by the producer of x (quantize_tokens_kernel or a fused
// glue kernel of the decoder) directly in MMA B-fragment order, 1 KB per (token, 256-column unit); the GEMV
// CTAs pull them into shared memory with one bulk (TMA) copy. Re-quantising in every CTA would cost
// ~14 instructions per column per CTA, more than the 0.25 op/weight main loop.
//
// Work split: a CTA owns 32 consecutive output rows (two 16-row MMA tiles sharing each B fragment) over the
// whole of K, so row sums are final inside the CTA (no atomics; LayerNorm partial sums are emitted). Its 8
// warps interleave over 256-column units; each warp first puts ALL its weight words in flight (<= 7 units
// x 8 registers) and only then waits for the producer (programmatic dependent launch), so the weight stream
// overlaps the tail of the previous kernel.
#pragma once
#include "common.cuh"

namespace onebit {
namespace imma {

#ifdef ONEBIT_TRACE
__device__ long long g_trace[16];
#define TR(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) ::onebit::imma::g_trace[i] = clock64(); } while (0)
#else
#define TR(i)
#endif

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kRows = 32;           // rows per CTA
constexpr int kUnitCols = 256;      // columns per warp work unit (one uint2 per row per thread)
constexpr int kUnitBytes = 1024;    // B-fragment digits of one (token, unit)
constexpr int kMaxProblems = 4;     // fused launches: q/k/v (3), gate/up (2)
constexpr int kMaxTokens = 8;

struct __align__(16) QMeta {  // per token, written by the quantiser
    double inv_scale;         // x' ~= q * inv_scale
    long long qtot;           // sum_k q
};

struct Problem {
    const uint8_t* w;       // [n_rows][K/8]
    const void* g;          // [n_rows] (TP) or nullptr
    const uint8_t* digits;  // [M][units][1024]
    const QMeta* qmeta;     // [M]
    float* t;               // [M][ld_t] fp32 out
    float* stats;           // [ctas of this problem][M][2] per-CTA (sum, sum sq) of the stored values, or nullptr
    int n_rows;
    int ld_t;               // row stride of t (elements)
    int cta_begin;          // first CTA of this problem
};

struct Args {
    Problem p[kMaxProblems];
    int nprob;
    int M;
    int K;
    int units;  // K / 256
};

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// `w & mask` as an opaque instruction: keeps the compiler from hoisting / CSE-ing hundreds of masked values.
__device__ __forceinline__ uint32_t plane(uint32_t w, uint32_t mask) {
    uint32_t r;
    asm volatile("and.b32 %0, %1, %2;" : "=r"(r) : "r"(w), "r"(mask));
    return r;
}

__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ---- mbarrier + 1-D bulk (TMA) copy global -> shared ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Dynamic shared memory the GEMV kernel needs.
inline size_t gemv_smem_bytes(int M, int units, int NT) {
    return (size_t)M * units * kUnitBytes + (size_t)kWarps * kRows * 8 * NT * sizeof(int);
}

// UPW = ceil(units / kWarps) compile-time bound of the weight-word register file (2, 3, 6 or 7).
template <typename TP, int NT, int UPW>
__global__ void __launch_bounds__(kThreads, (NT == 1 && UPW <= 3) ? 5 : 2) gemv_kernel(const __grid_constant__ Args A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_bar[UPW];  // one per round of 8 units: the IMMA loop starts on the first 8 KB
    const int M = A.M;
    unsigned char* Bs = smem;                                                          // [M][units][1024]
    int* red = reinterpret_cast<int*>(smem + (size_t)M * A.units * kUnitBytes);       // [kWarps][32][8*NT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kMaxProblems; ++i)
        if (i < A.nprob && (int)blockIdx.x >= A.p[i].cta_begin) pi = i;
    const Problem& P = A.p[pi];
    const int cta = (int)blockIdx.x - P.cta_begin;
    const int row0 = cta * kRows;
    const int Kb = A.K >> 3;

    TR(0);
    // ---- 1. every weight word of this warp goes in flight (touches only static data) ----
    uint2 wreg[UPW][4];
    {
        const uint8_t* rp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)  // tail rows re-read the last row (results discarded)
            rp[i] = P.w + (size_t)min(row0 + g + 8 * i, P.n_rows - 1) * Kb + 8 * t4 + 32 * warp;
#pragma unroll
        for (int s = 0; s < UPW; ++s) {
            const bool ok = warp + s * kWarps < A.units;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                wreg[s][i] = ok ? ldg_stream_u2(rp[i] + s * (32 * kWarps)) : make_uint2(0u, 0u);
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int r8 = 0; r8 < UPW; ++r8) mbar_init(&s_bar[r8], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // epilogue operands: this thread finalises (row = tid & 31, token = tid >> 5); g is static -> load it now
    const int er = tid & 31, em = tid >> 5;
    TP graw = from_f32<TP>(1.f);  // converted at use: a conversion here would wait for the load
    if (P.g != nullptr && row0 + er < P.n_rows) graw = static_cast<const TP*>(P.g)[row0 + er];
    TR(1);
    pdl_launch_dependents();
    pdl_wait();  // producer's digits / qmeta are now visible
    TR(2);
    QMeta qm;
    qm.inv_scale = 0.0;
    qm.qtot = 0;
    if (em < M) qm = P.qmeta[em];

    // ---- 2. activation digits: bulk (TMA) copies into shared memory, one mbarrier per round of 8 units ----
    if (tid == 0) {
#pragma unroll
        for (int r8 = 0; r8 < UPW; ++r8) {
            const int u0 = r8 * kWarps, nu = min(kWarps, A.units - u0);
            if (nu > 0) {
                mbar_expect_tx(&s_bar[r8], (uint32_t)(M * nu * kUnitBytes));
                for (int m = 0; m < M; ++m)
                    bulk_g2s(Bs + (size_t)(m * A.units + u0) * kUnitBytes, P.digits + (size_t)(m * A.units + u0) * kUnitBytes,
                             (uint32_t)(nu * kUnitBytes), &s_bar[r8]);
            }
        }
    }
    __syncthreads();  // barrier inits visible to all waiters
    TR(3);

    // two independent accumulator sets per row tile (even / odd planes): 4 interleaved IMMA dependency chains
    int acc2[2][2][NT][4];
#pragma unroll
    for (int z = 0; z < 2; ++z)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc2[z][r][nt][i] = 0;

    // ---- 3. main loop: 4 LOP3 + 1 IMMA per (16 rows x 32 columns x plane) ----
#pragma unroll
    for (int s = 0; s < UPW; ++s) {
        const int u = warp + s * kWarps;
        if (u < A.units) {
            mbar_wait(&s_bar[s], 0);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint4 bv[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int m = 2 * nt + (g >> 2);
                    bv[nt] = make_uint4(0u, 0u, 0u, 0u);
                    if (m < M)
                        bv[nt] = *reinterpret_cast<const uint4*>(Bs + ((size_t)(m * A.units + u) * 4 + jp) * 256 +
                                                                 ((g & 3) * 4 + t4) * 16);
                }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const uint32_t mask = 0x01010101u << (2 * jp + jj);
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const uint32_t a0 = plane(wreg[s][2 * r].x, mask), a1 = plane(wreg[s][2 * r + 1].x, mask);
                        const uint32_t a2 = plane(wreg[s][2 * r].y, mask), a3 = plane(wreg[s][2 * r + 1].y, mask);
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt)
                            imma16832(acc2[jj][r][nt], a0, a1, a2, a3, jj ? bv[nt].z : bv[nt].x,
                                      jj ? bv[nt].w : bv[nt].y);
                    }
                }
            }
        }
    }

    TR(4);
    int acc[2][NT][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][nt][i] = acc2[0][r][nt][i] + acc2[1][r][nt][i];

    // ---- 4. combine the K split across warps, undo the quantisation, scale, store ----
    constexpr int kCols = 8 * NT;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            int* base = red + ((size_t)warp * kRows + 16 * r) * kCols + 8 * nt + 2 * t4;
            *reinterpret_cast<int2*>(base + (size_t)g * kCols) = make_int2(acc[r][nt][0], acc[r][nt][1]);
            *reinterpret_cast<int2*>(base + (size_t)(g + 8) * kCols) = make_int2(acc[r][nt][2], acc[r][nt][3]);
        }
    __syncthreads();
    const int r = tid & 31, m = tid >> 5;  // warp <-> token: the stats reduction is a plain warp_sum
    const int n = row0 + r;
    float su = 0.f, sq = 0.f;
    if (m < M && m < 2 * NT) {
        int4 a = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const int4 v = *reinterpret_cast<const int4*>(red + ((size_t)w * kRows + r) * kCols + 4 * m);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        const long long V = (((long long)a.w * 256 + a.z) * 256 + a.y) * 256 + a.x;  // = 128 * sum_{bit=1} q
        float val = (float)((double)(qm.qtot - 2 * (V >> 7)) * qm.inv_scale);
        if (n < P.n_rows) {
            val *= to_f32(graw);
            P.t[(size_t)m * P.ld_t + n] = val;
            su = val;
            sq = val * val;
        }
    }
    if (P.stats != nullptr && m < 2 * NT) {
        su = warp_sum(su);
        sq = warp_sum(sq);
        if (lane == 0 && m < M) *reinterpret_cast<float2*>(P.stats + ((size_t)cta * M + m) * 2) = make_float2(su, sq);
    }
    TR(5);
}

// ------------------------------------------------------------------------------------------------------
// Quantiser building block (device): one CTA turns one token's x' (fp32, in shared memory) into digits.
// xs: x'[K] in shared memory; digits: [units][1024] for this token; returns through qmeta.
// `scratch` needs 2 * (blockDim/32) floats worth of shared memory (reused as long long).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void quantize_from_smem(const float* xs, int K, float amax_thread, uint8_t* digits,
                                                   QMeta* qmeta, void* scratch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float* fs = reinterpret_cast<float*>(scratch);
    float am = amax_thread;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
    if (lane == 0) fs[warp] = am;
    __syncthreads();
    am = 0.f;
    for (int w = 0; w < nw; ++w) am = fmaxf(am, fs[w]);
    int e = 0;
    if (am > 0.f && am < 3.0e38f) frexpf(am, &e);  // am = f * 2^e, f in [0.5, 1)  =>  |x'| < 2^e
    const float S = ldexpf(1.0f, 22 - e);           // |q| <= 2^22
    __syncthreads();                                // fs is reused below

    // item = (unit, t, word, plane j) -> the four columns 8b + j (b = 0..3) of one 32-bit weight word.
    // Balanced digits come as bytes: with u = v + 0x00808080, byte_i(u) ^ 0x80 (i < 3) and byte_3(u) are the s8
    // digits of v, because sum_i (byte_i(u) - 128 [i<3]) * 256^i = u - 0x808080 = v.
    int qs = 0;
    const int items = (K / kUnitCols) * 64;
    for (int it = tid; it < items; it += blockDim.x) {
        const int j = it & 7, ws = (it >> 3) & 1, tt = (it >> 4) & 3, u = it >> 6;
        const float* xr = xs + u * kUnitCols + 64 * tt + 32 * ws + j;
        uint32_t dw[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int q = __float2int_rn(xr[8 * b] * S);
            qs += q;
            const int v = (j == 7) ? -q : (q << (7 - j));
            dw[b] = ((uint32_t)v + 0x00808080u) ^ 0x00808080u;  // byte d = digit d of column b
        }
        // 4x4 byte transpose: register d <- digit d of the four columns
        const uint32_t t0 = __byte_perm(dw[0], dw[1], 0x5140), t1 = __byte_perm(dw[2], dw[3], 0x5140);
        const uint32_t t2 = __byte_perm(dw[0], dw[1], 0x7362), t3 = __byte_perm(dw[2], dw[3], 0x7362);
        uint32_t* dst = reinterpret_cast<uint32_t*>(digits + ((size_t)u * 4 + (j >> 1)) * 256) + tt * 4 + (j & 1) * 2 + ws;
        dst[0] = __byte_perm(t0, t1, 0x5410);   // digit 0 -> lanes g&3 = 0
        dst[16] = __byte_perm(t0, t1, 0x7632);  // digit 1
        dst[32] = __byte_perm(t2, t3, 0x5410);  // digit 2
        dst[48] = __byte_perm(t2, t3, 0x7632);  // digit 3
    }
    long long q64 = qs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q64 += __shfl_xor_sync(0xffffffffu, q64, o);
    long long* ls = reinterpret_cast<long long*>(scratch);
    if (lane == 0) ls[warp] = q64;
    __syncthreads();
    if (tid == 0) {
        long long tot = 0;
        for (int w = 0; w < nw; ++w) tot += ls[w];
        QMeta qm;
        qm.inv_scale = am < 3.0e38f ? ldexp(1.0, e - 22) : (double)NAN;  // Inf input: the token's outputs are NaN, not finite garbage
        qm.qtot = tot;
        *qmeta = qm;
    }
}

}  // namespace imma
}  // namespace onebit
