// CUDA-core packed-sign mat-vec: t[m][n] = sum_k s(n,k) * h[k] * x[m][k]   (bitnet.py:113-116).
//
// Generic variant: any K % 8 == 0, any N, any M, any of f16/bf16/f32 activations. It is the
// correctness anchor and the fallback for shapes the tensor-core variants do not take (odd K,
// tensor-parallel shards whose rows are not 16-byte multiples). Signs are consumed straight from the
// 1-bit layout: a weight bit is shifted into the IEEE sign position and XOR-ed onto h*x, so no +-1
// value is ever materialised outside a register.
//
// Mapping: a CTA owns 32 output rows (lane <-> row, so there is no cross-lane reduction) and splits K
// eight ways across its warps; the per-warp partials are combined in a fixed order (deterministic).
// h*x is staged in shared memory as fp32 in K-chunks and read as warp-wide broadcasts.
#include "common.cuh"

namespace onebit {
namespace {

constexpr int kRowsPerCta = 32;
constexpr int kSlices = 8;
constexpr int kThreads = kRowsPerCta * kSlices;
constexpr int kChunkCols = 2048;  // K columns staged per pass
constexpr int kTok = 4;           // tokens per CTA pass

__device__ __forceinline__ float flip(float v, uint32_t signbit) {
    return __uint_as_float(__float_as_uint(v) ^ signbit);
}

template <typename TX, typename TP>
__global__ void __launch_bounds__(kThreads)
matvec_simt_kernel(const TX* __restrict__ x, const uint8_t* __restrict__ w, const TP* __restrict__ g,
                   const TP* __restrict__ h, float* __restrict__ t, int64_t M, int64_t K, int64_t N,
                   int scale_by_g) {
    __shared__ __align__(16) float xs[kTok][kChunkCols];
    __shared__ float red[kSlices][kTok][kRowsPerCta];

    const int lane = threadIdx.x & 31;
    const int ks = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * kRowsPerCta + lane;
    const int64_t m0 = (int64_t)blockIdx.y * kTok;
    const bool row_ok = n < N;
    const int64_t Kb = K >> 3;
    const uint8_t* wrow = w + (row_ok ? n : 0) * Kb;
    const bool word_path = (Kb & 3) == 0;

    float acc[kTok];
#pragma unroll
    for (int m = 0; m < kTok; ++m) acc[m] = 0.f;

    for (int64_t kc0 = 0; kc0 < K; kc0 += kChunkCols) {
        const int cols = (int)min((int64_t)kChunkCols, K - kc0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < kTok * kChunkCols; idx += kThreads) {
            const int m = idx / kChunkCols, kk = idx % kChunkCols;
            float v = 0.f;
            if (kk < cols && m0 + m < M) v = to_f32(x[(m0 + m) * K + kc0 + kk]) * to_f32(h[kc0 + kk]);
            xs[m][kk] = v;
        }
        __syncthreads();
        if (!row_ok) continue;
        const int cbytes = cols >> 3;
        const uint8_t* wc = wrow + (kc0 >> 3);
        int done_bytes = 0;
        if (word_path) {
            const int nwords = cbytes >> 2;
            const int per = (nwords + kSlices - 1) / kSlices;  // contiguous words per slice
            const int j0 = ks * per, j1 = min(nwords, j0 + per);
            for (int j = j0; j < j1; ++j) {
                const uint32_t wv = __ldg(reinterpret_cast<const uint32_t*>(wc) + j);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t s0 = (wv << (31 - 4 * q)) & 0x80000000u;
                    const uint32_t s1 = (wv << (30 - 4 * q)) & 0x80000000u;
                    const uint32_t s2 = (wv << (29 - 4 * q)) & 0x80000000u;
                    const uint32_t s3 = (wv << (28 - 4 * q)) & 0x80000000u;
#pragma unroll
                    for (int m = 0; m < kTok; ++m) {
                        const float4 xv = *reinterpret_cast<const float4*>(&xs[m][32 * j + 4 * q]);
                        acc[m] += flip(xv.x, s0);
                        acc[m] += flip(xv.y, s1);
                        acc[m] += flip(xv.z, s2);
                        acc[m] += flip(xv.w, s3);
                    }
                }
            }
            done_bytes = nwords << 2;
        }
        // byte tail (and the whole row when rows are not 4-byte multiples)
        for (int b = done_bytes + ks; b < cbytes; b += kSlices) {
            const uint32_t wv = wc[b];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t s = (wv << (31 - i)) & 0x80000000u;
#pragma unroll
                for (int m = 0; m < kTok; ++m) acc[m] += flip(xs[m][8 * b + i], s);
            }
        }
    }

#pragma unroll
    for (int m = 0; m < kTok; ++m) red[ks][m][lane] = acc[m];
    __syncthreads();
    if (ks == 0 && row_ok) {
        const float gs = scale_by_g ? to_f32(g[n]) : 1.f;
#pragma unroll
        for (int m = 0; m < kTok; ++m) {
            if (m0 + m >= M) break;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < kSlices; ++q) s += red[q][m][lane];
            t[(m0 + m) * N + n] = s * gs;
        }
    }
}

}  // namespace

int launch_matvec_simt(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m,
                       int64_t k, int64_t n, int act_dtype, int param_dtype, bool scale_by_g, cudaStream_t s) {
    if (m == 0 || n == 0) return ONEBIT_OK;
    dim3 grid((unsigned)((n + kRowsPerCta - 1) / kRowsPerCta), (unsigned)((m + kTok - 1) / kTok));
    ONEBIT_REQUIRE(grid.y <= 65535, "matvec_simt: M too large for this variant (max 262140 tokens)");
    return dispatch_dtype(act_dtype, [&](auto xt) {
        using TX = decltype(xt);
        return dispatch_dtype(param_dtype, [&](auto pt) {
            using TP = decltype(pt);
            matvec_simt_kernel<TX, TP><<<grid, kThreads, 0, s>>>(
                static_cast<const TX*>(x), reinterpret_cast<const uint8_t*>(w), static_cast<const TP*>(g),
                static_cast<const TP*>(h), t, m, k, n, scale_by_g ? 1 : 0);
            ONEBIT_CUDA_TRY(cudaGetLastError());
            return ONEBIT_OK;
        });
    });
}

}  // namespace onebit
