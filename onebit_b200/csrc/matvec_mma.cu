// Tensor-core packed-sign mat-vec for decode-size batches (M <= 8):
//     t[m][n] = sum_k s(n,k) * h[k] * x[m][k]                                  (bitnet.py:113-116)
// with the sign matrix read from HBM exactly once at 1 bit/element and never expanded in memory.
//
// Why integer tensor cores. Measured on B200 (profiles/r01_ubench_mma_sync.txt): the warp-level
// mma.sync path issues 0.47 HMMA.16816 or 0.48 IMMA.16832 per clock per SM. With the 16 weight rows
// as the M side of the MMA that is 120 weight-bits/clk/SM in fp16 (67 % of what HBM delivers) but
// 246 in int8 (137 %), so only the int8 shape keeps the kernel HBM-bound. CUDA cores alone need
// >= 1 lane-op per bit and top out near 35 %.
//
// Bit-plane trick (no shifts, 0.25 ALU op per weight). For bit position j of every byte of a 32-bit
// weight word, `w & (0x01010101 << j)` is already a valid int8x4 A-fragment register: byte b holds
// bit(8b+j) * 2^j (and -128 * bit for j = 7). The activation side absorbs the plane scale: h*x is
// quantised per token to a 23-bit integer q (power-of-two scale, so the only error is the 2^-22
// rounding), column k with plane j = k % 8 carries v = q << (7-j) (v = -q for j = 7), and v is split
// into four balanced base-256 digits that sit in four B columns of the MMA. All planes then
// accumulate into ONE int32 accumulator per digit: sum_d 256^d * acc_d = 128 * sum_{bit=1} q exactly,
// and  sum_k s*q = sum_k q - 2 * sum_{bit=1} q.  Integer arithmetic end to end => bit-reproducible
// and independent of the split of K across warps.
//
// Work split. A CTA owns 32 consecutive output rows (two 16-row MMA tiles sharing every B fragment)
// over the whole of K, so row sums are final inside the CTA (no atomics, LayerNorm partials can be
// emitted). Its 8 warps interleave over 256-column units; each warp first issues the global loads
// of ALL its weight words (<= 7 units x 8 registers) and only then quantises x, so the weight
// stream is in flight during the prologue — and, under programmatic dependent launch, during the
// tail of the previous kernel.
#include "common.cuh"

namespace onebit {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kRows = 32;             // rows per CTA (2 MMA row tiles)
constexpr int kUnitCols = 256;        // columns per warp work unit (uint2 per row per thread)
constexpr int kMaxUnitsPerWarp = 7;   // weight words kept in registers: 7 * 8 = 56 registers
constexpr int kBsTokUnitBytes = 4 * 288;  // B fragments of one (token, unit): 4 plane pairs, padded to 288 B
constexpr int kBsBudget = 160 * 1024;
constexpr int kSmemLimit = 224 * 1024;  // dynamic shared memory we allow ourselves (227 KB per CTA on sm_100)

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// `w & mask` as an opaque instruction: the weight words are loop-invariant across K chunks, and without this the
// compiler hoists all 448 masked values out of the chunk loop and spills them.
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

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __half2* h2 = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h2[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
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
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Balanced base-256 digits of v (|v| < 2^30): v = d0 + 256 d1 + 65536 d2 + 2^24 d3, every digit in [-128,127].
__device__ __forceinline__ void digits4(int v, int (&d)[4]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        d[i] = (int)(signed char)(v & 0xFF);
        v = (v - d[i]) >> 8;
    }
    d[3] = v;
}

struct MatvecArgs {
    const void* x;
    const uint8_t* w;
    const void* g;
    const void* h;
    float* t;        // [M][N] fp32
    float* stats;    // optional [gridDim.x][M][2]: per-CTA (sum u, sum u^2) of the rows it owns, or nullptr
    int64_t M, K, N;
    int scale_by_g;
    int units;            // K / 256
    int units_per_chunk;  // units whose B fragments fit in shared memory at once
    int staged;           // 1: x and h are bulk-copied into shared memory first (single chunk only)
};

template <typename TX, typename TP, int NT>
__global__ void __launch_bounds__(kThreads, 1) matvec_imma_kernel(const __grid_constant__ MatvecArgs A) {
    constexpr int kTok = 2 * NT;  // token slots (2 per n-tile: 4 digit columns each)
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* Bs = smem;                                             // [M][units_per_chunk][4][288]
    int* red = reinterpret_cast<int*>(smem + (size_t)A.M * A.units_per_chunk * kBsTokUnitBytes);       // [kWarps][32][8*NT]
    __shared__ float s_amax[kWarps][kTok];
    __shared__ long long s_qsum[kWarps][kTok];
    __shared__ float s_scale[kTok];      // S  (q = rint(x' * S))
    __shared__ double s_inv[kTok];       // 1 / S
    __shared__ long long s_qtot[kTok];   // sum_k q
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int64_t Kb = A.K >> 3;
    const int64_t row0 = (int64_t)blockIdx.x * kRows;
    const int M = (int)A.M;

    // Staged mode: h (static, like the weights) is bulk-copied to shared memory right away, x after the dependency
    // wait. One elected thread issues the copies; everybody waits on the mbarriers. This turns the 8+ dependent
    // L2 round trips of a load-use loop into one.
    const TX* x = static_cast<const TX*>(A.x);
    const TP* h = static_cast<const TP*>(A.h);
    unsigned char* stage = smem + (size_t)M * A.units_per_chunk * kBsTokUnitBytes + (size_t)kWarps * kRows * 8 * NT * sizeof(int);
    if (A.staged) {
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&s_bar[0], (uint32_t)(A.K * sizeof(TP)));
            bulk_g2s(stage + (size_t)M * A.K * sizeof(TX), A.h, (uint32_t)(A.K * sizeof(TP)), &s_bar[0]);
        }
    }

    // ---- 1. put every weight word of this warp in flight (independent of x: may overlap the producer) ----
    uint2 wreg[kMaxUnitsPerWarp][4];
    {
        const uint8_t* wbase = A.w + 8 * t4;
        int64_t rr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rr[i] = min(row0 + g + 8 * i, A.N - 1) * Kb;  // clamp: tail rows re-read the last row
#pragma unroll
        for (int s = 0; s < kMaxUnitsPerWarp; ++s) {
            const int u = warp + s * kWarps;
            if (u < A.units) {
#pragma unroll
                for (int i = 0; i < 4; ++i) wreg[s][i] = ldg_stream_u2(wbase + rr[i] + (int64_t)u * 32);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) wreg[s][i] = make_uint2(0u, 0u);
            }
        }
    }

    // Programmatic dependent launch: everything above touched only the (static) weights. Let the next kernel in
    // the stream start its own weight prefetch now, then wait until our producer's writes to x are visible.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (A.staged) {
        if (tid == 0) {
            mbar_expect_tx(&s_bar[1], (uint32_t)(M * A.K * sizeof(TX)));
            bulk_g2s(stage, A.x, (uint32_t)(M * A.K * sizeof(TX)), &s_bar[1]);
        }
        __syncthreads();  // the barrier inits by thread 0 are visible to every waiter
        mbar_wait(&s_bar[0], 0);
        mbar_wait(&s_bar[1], 0);
        x = reinterpret_cast<const TX*>(stage);
        h = reinterpret_cast<const TP*>(stage + (size_t)M * A.K * sizeof(TX));
    }

    // ---- 2. per-token scale: amax over K of |h*x| ----
    {
        float am[kTok];
#pragma unroll
        for (int m = 0; m < kTok; ++m) am[m] = 0.f;
        for (int64_t v8 = tid; v8 < (A.K >> 3); v8 += kThreads) {
            float hv[8];
            load8(h + 8 * v8, hv);
#pragma unroll
            for (int m = 0; m < kTok; ++m)
                if (m < M) {
                    float xv[8];
                    load8(x + (int64_t)m * A.K + 8 * v8, xv);
#pragma unroll
                    for (int i = 0; i < 8; ++i) am[m] = fmaxf(am[m], fabsf(xv[i] * hv[i]));
                }
        }
#pragma unroll
        for (int m = 0; m < kTok; ++m) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) am[m] = fmaxf(am[m], __shfl_xor_sync(0xffffffffu, am[m], o));
            if (lane == 0) s_amax[warp][m] = am[m];
        }
        __syncthreads();
        if (tid < kTok) {
            float a = 0.f;
            for (int w = 0; w < kWarps; ++w) a = fmaxf(a, s_amax[w][tid]);
            int e = 0;
            if (a > 0.f && a < 3.0e38f) frexpf(a, &e);  // a = f * 2^e, f in [0.5, 1)  =>  a <= 2^e
            s_scale[tid] = ldexpf(1.0f, 22 - e);         // |q| <= 2^22
            s_inv[tid] = ldexp(1.0, e - 22);
        }
        __syncthreads();
    }

    int acc[2][NT][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][nt][i] = 0;
    int qsum[kTok];
#pragma unroll
    for (int m = 0; m < kTok; ++m) qsum[m] = 0;

    // ---- 3. chunks of K: quantise h*x into MMA B-fragment order, then consume with the resident weight words ----
    const int UC = A.units_per_chunk;
    for (int c0 = 0; c0 < A.units; c0 += UC) {
        const int cu = min(UC, A.units - c0);
        if (c0 > 0) __syncthreads();
        // quantisation: one item = (token, unit, t, word, plane j) -> 4 columns (byte b = 0..3) -> 4 digit registers.
        // Balanced digits come for free as bytes: with u = v + 0x00808080, byte_i(u) ^ 0x80 (i < 3) and byte_3(u)
        // are the s8 digits of v, because sum_i (byte_i(u) - 128 [i<3]) * 256^i = u - 0x808080 = v.
        const int items = M * cu * 64;
        for (int it = tid; it < items; it += kThreads) {
            const int j = it & 7, ws = (it >> 3) & 1, tt = (it >> 4) & 3;
            const int u = (it >> 6) % cu, m = (it >> 6) / cu;
            const int64_t kbase = (int64_t)(c0 + u) * kUnitCols + 64 * tt + 32 * ws + j;
            const float S = s_scale[m];
            const TX* xr = x + (int64_t)m * A.K + kbase;
            const TP* hr = h + kbase;
            uint32_t dw[4];
            int qs = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int q = __float2int_rn(to_f32(xr[8 * b]) * to_f32(hr[8 * b]) * S);
                qs += q;
                const int v = (j == 7) ? -q : (q << (7 - j));
                dw[b] = ((uint32_t)v + 0x00808080u) ^ 0x00808080u;  // byte d = digit d of column b
            }
#pragma unroll
            for (int mm = 0; mm < kTok; ++mm)
                if (mm == m) qsum[mm] += qs;
            // 4x4 byte transpose: register d <- digit d of the four columns
            const uint32_t t0 = __byte_perm(dw[0], dw[1], 0x5140), t1 = __byte_perm(dw[2], dw[3], 0x5140);
            const uint32_t t2 = __byte_perm(dw[0], dw[1], 0x7362), t3 = __byte_perm(dw[2], dw[3], 0x7362);
            uint32_t* dst = reinterpret_cast<uint32_t*>(Bs + ((size_t)(m * UC + u) * 4 + (j >> 1)) * 288) +
                            tt * 4 + (j & 1) * 2 + ws;
            dst[0] = __byte_perm(t0, t1, 0x5410);
            dst[16] = __byte_perm(t0, t1, 0x7632);
            dst[32] = __byte_perm(t2, t3, 0x5410);
            dst[48] = __byte_perm(t2, t3, 0x7632);
        }
        __syncthreads();

        // consume: this warp's units inside the chunk
#pragma unroll
        for (int s = 0; s < kMaxUnitsPerWarp; ++s) {
            const int u = warp + s * kWarps;  // global unit index
            if (u >= c0 && u < c0 + cu) {
                const int ul = u - c0;
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    uint4 bv[NT];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int m = 2 * nt + (g >> 2);
                        bv[nt] = make_uint4(0u, 0u, 0u, 0u);
                        if (m < M)
                            bv[nt] = *reinterpret_cast<const uint4*>(Bs + ((size_t)(m * UC + ul) * 4 + jp) * 288 +
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
                                imma16832(acc[r][nt], a0, a1, a2, a3, jj ? bv[nt].z : bv[nt].x, jj ? bv[nt].w : bv[nt].y);
                        }
                    }
                }
            }
        }
    }

    // ---- 4. combine the K-split across warps, undo the quantisation, scale, store ----
#pragma unroll
    for (int m = 0; m < kTok; ++m) {
        long long q = qsum[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) s_qsum[warp][m] = q;
    }
    constexpr int kCols = 8 * NT;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            int* base = red + ((size_t)warp * kRows + 16 * r) * kCols + 8 * nt + 2 * t4;
            base[(size_t)g * kCols] = acc[r][nt][0];
            base[(size_t)g * kCols + 1] = acc[r][nt][1];
            base[(size_t)(g + 8) * kCols] = acc[r][nt][2];
            base[(size_t)(g + 8) * kCols + 1] = acc[r][nt][3];
        }
    __syncthreads();
    if (tid < kTok) {
        long long q = 0;
        for (int w = 0; w < kWarps; ++w) q += s_qsum[w][tid];
        s_qtot[tid] = q;
    }
    __syncthreads();
    float su = 0.f, sq = 0.f;  // LayerNorm partials of this thread's output (one row, one token)
    const int r = tid & 31, m = tid >> 5;  // warp <-> token keeps the partial reduction a plain warp_sum
    const int64_t n = row0 + r;
    if (m < M && m < kTok) {
        long long V = 0;
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            long long a = 0;
            for (int w = 0; w < kWarps; ++w) a += red[((size_t)w * kRows + r) * kCols + 4 * m + d];
            V = V * 256 + a;
        }
        const long long sel = V >> 7;  // = sum_{bit=1} q  (V is an exact multiple of 128)
        float val = (float)((double)(s_qtot[m] - 2 * sel) * s_inv[m]);
        if (A.scale_by_g && n < A.N) val *= to_f32(static_cast<const TP*>(A.g)[n]);
        if (n < A.N) {
            A.t[(int64_t)m * A.N + n] = val;
            su = val;
            sq = val * val;
        }
    }
    if (A.stats != nullptr && m < kTok) {
        su = warp_sum(su);
        sq = warp_sum(sq);
        if (lane == 0 && m < M) {
            float* st = A.stats + ((size_t)blockIdx.x * M + m) * 2;
            st[0] = su;
            st[1] = sq;
        }
    }
}

template <typename TX, typename TP, int NT>
int launch_nt(const MatvecArgs& a, size_t smem, cudaStream_t s) {
    auto kern = matvec_imma_kernel<TX, TP, NT>;
    static bool configured[64] = {false};  // per template instance and device; the attribute is sticky
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((a.N + kRows - 1) / kRows));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    ONEBIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    return ONEBIT_OK;
}

}  // namespace

bool matvec_mma_supported(int64_t m, int64_t k, int64_t n, int act_dtype) {
    (void)act_dtype;
    (void)n;
    return m >= 1 && m <= 8 && k % kUnitCols == 0 && k / kUnitCols <= kMaxUnitsPerWarp * kWarps;
}

int launch_matvec_mma(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m, int64_t k,
                      int64_t n, int act_dtype, int param_dtype, bool scale_by_g, cudaStream_t s) {
    if (m == 0 || n == 0) return ONEBIT_OK;
    ONEBIT_REQUIRE(matvec_mma_supported(m, k, n, act_dtype), "matvec_mma: unsupported shape");
    MatvecArgs a;
    a.x = x;
    a.w = reinterpret_cast<const uint8_t*>(w);
    a.g = g;
    a.h = h;
    a.t = t;
    a.stats = nullptr;
    a.M = m;
    a.K = k;
    a.N = n;
    a.scale_by_g = scale_by_g ? 1 : 0;
    a.units = (int)(k / kUnitCols);
    const int nt = m <= 2 ? 1 : (m <= 4 ? 2 : 4);
    const int ktok = 2 * nt;
    (void)ktok;
    const size_t red_bytes = (size_t)kWarps * kRows * 8 * nt * sizeof(int);
    const size_t stage_bytes = (size_t)m * k * dtype_size(act_dtype) + (size_t)k * dtype_size(param_dtype);
    const size_t bs_all = (size_t)m * a.units * kBsTokUnitBytes;
    size_t smem;
    if (bs_all + red_bytes + stage_bytes <= (size_t)kSmemLimit && (size_t)m * k * dtype_size(act_dtype) < (1u << 20)) {
        a.staged = 1;
        a.units_per_chunk = a.units;
        smem = bs_all + red_bytes + stage_bytes;
    } else {
        a.staged = 0;
        a.units_per_chunk = (int)std::min<int64_t>(a.units, kBsBudget / ((int64_t)m * kBsTokUnitBytes));
        smem = (size_t)m * a.units_per_chunk * kBsTokUnitBytes + red_bytes;
    }
    return dispatch_dtype(act_dtype, [&](auto xt) {
        using TX = decltype(xt);
        return dispatch_dtype(param_dtype, [&](auto pt) {
            using TP = decltype(pt);
            switch (nt) {
                case 1: return launch_nt<TX, TP, 1>(a, smem, s);
                case 2: return launch_nt<TX, TP, 2>(a, smem, s);
                default: return launch_nt<TX, TP, 4>(a, smem, s);
            }
        });
    });
}

}  // namespace onebit
