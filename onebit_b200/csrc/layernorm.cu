// Scale-by-g + LayerNorm over the N outputs of each token (bitnet.py:116-120).
//
// nn.LayerNorm(N, elementwise_affine=False): mean and BIASED variance over N, eps inside the sqrt.
// One CTA per token; the row (<= 16K fp32) is held in registers between the passes so HBM/L2 sees one
// read of t and one write of y. Statistics are two-pass (mean, then centred squares) in fp32 with a
// fixed reduction order, so results are deterministic run to run.
#include "common.cuh"

namespace onebit {
namespace {

constexpr int kLnThreads = 512;
constexpr int kLnRegs = 32;  // row elements cached per thread -> rows up to 16384 stay in registers

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();  // protect sh from the previous use
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    float r = 0.f;
    const int nw = blockDim.x >> 5;
#pragma unroll 1
    for (int i = 0; i < nw; ++i) r += sh[i];  // every thread sums in the same order
    return r;
}

template <typename T>
__device__ __forceinline__ float4 load4_as_f32(const T* p, int64_t i4);
template <>
__device__ __forceinline__ float4 load4_as_f32<float>(const float* p, int64_t i4) {
    return reinterpret_cast<const float4*>(p)[i4];
}
template <>
__device__ __forceinline__ float4 load4_as_f32<__half>(const __half* p, int64_t i4) {
    const uint2 r = reinterpret_cast<const uint2*>(p)[i4];
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 load4_as_f32<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i4) {
    const uint2 r = reinterpret_cast<const uint2*>(p)[i4];
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void store4_from_f32(T* p, int64_t i4, float4 v);
template <>
__device__ __forceinline__ void store4_from_f32<float>(float* p, int64_t i4, float4 v) {
    reinterpret_cast<float4*>(p)[i4] = v;
}
template <>
__device__ __forceinline__ void store4_from_f32<__half>(__half* p, int64_t i4, float4 v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(p)[i4] = pk;
}
template <>
__device__ __forceinline__ void store4_from_f32<__nv_bfloat16>(__nv_bfloat16* p, int64_t i4, float4 v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(p)[i4] = pk;
}

// Vectorised variant for N % 4 == 0 and N <= 4 * NV4 * blockDim: the row lives in NV4 float4 registers per
// thread, all loads are issued before the first reduction (this is the variant every LLaMA shape takes).
template <typename TY, typename TP, int NV4, int THREADS>
__global__ void __launch_bounds__(THREADS)
scale_layernorm_vec_kernel(const float* __restrict__ t, const TP* __restrict__ g, const TP* __restrict__ bias,
                           TY* __restrict__ y, int64_t N, float eps) {
    __shared__ float sh[THREADS / 32];
    const int64_t m = blockIdx.x, N4 = N >> 2;
    const float* row = t + m * N;
    float4 u[NV4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int64_t i4 = (int64_t)i * THREADS + threadIdx.x;
        u[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i4 < N4) {
            float4 v = reinterpret_cast<const float4*>(row)[i4];
            if (g) {
                const float4 gv = load4_as_f32<TP>(g, i4);
                v.x *= gv.x; v.y *= gv.y; v.z *= gv.z; v.w *= gv.w;
            }
            u[i] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    }
    const float mean = block_sum(s, sh) / (float)N;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int64_t i4 = (int64_t)i * THREADS + threadIdx.x;
        if (i4 < N4) {
            const float a = u[i].x - mean, b = u[i].y - mean, c = u[i].z - mean, d = u[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float var = block_sum(q, sh) / (float)N;
    const float rstd = rsqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int64_t i4 = (int64_t)i * THREADS + threadIdx.x;
        if (i4 < N4) {
            float4 v = make_float4((u[i].x - mean) * rstd, (u[i].y - mean) * rstd, (u[i].z - mean) * rstd,
                                   (u[i].w - mean) * rstd);
            if (bias) {
                const float4 bv = load4_as_f32<TP>(bias, i4);
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            }
            store4_from_f32<TY>(y + m * N, i4, v);
        }
    }
}

template <typename TY, typename TP>
__global__ void __launch_bounds__(kLnThreads)
scale_layernorm_kernel(const float* __restrict__ t, const TP* __restrict__ g, const TP* __restrict__ bias,
                       TY* __restrict__ y, int64_t N, float eps) {
    __shared__ float sh[kLnThreads / 32];
    const int64_t m = blockIdx.x;
    const float* row = t + m * N;
    TY* out = y + m * N;
    const bool cached = N <= (int64_t)kLnThreads * kLnRegs;
    float u[kLnRegs];
    float s = 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < kLnRegs; ++i) {
            const int64_t n = (int64_t)i * kLnThreads + threadIdx.x;
            float v = 0.f;
            if (n < N) v = row[n] * (g ? to_f32(g[n]) : 1.f);
            u[i] = v;
            s += v;
        }
    } else {
        for (int64_t n = threadIdx.x; n < N; n += kLnThreads) s += row[n] * (g ? to_f32(g[n]) : 1.f);
    }
    const float mean = block_sum(s, sh) / (float)N;
    float q = 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < kLnRegs; ++i) {
            const int64_t n = (int64_t)i * kLnThreads + threadIdx.x;
            if (n < N) {
                const float d = u[i] - mean;
                q += d * d;
            }
        }
    } else {
        for (int64_t n = threadIdx.x; n < N; n += kLnThreads) {
            const float d = row[n] * (g ? to_f32(g[n]) : 1.f) - mean;
            q += d * d;
        }
    }
    const float var = block_sum(q, sh) / (float)N;
    const float rstd = rsqrtf(var + eps);
    if (cached) {
#pragma unroll
        for (int i = 0; i < kLnRegs; ++i) {
            const int64_t n = (int64_t)i * kLnThreads + threadIdx.x;
            if (n < N) {
                float v = (u[i] - mean) * rstd;
                if (bias) v += to_f32(bias[n]);
                out[n] = from_f32<TY>(v);
            }
        }
    } else {
        for (int64_t n = threadIdx.x; n < N; n += kLnThreads) {
            float v = (row[n] * (g ? to_f32(g[n]) : 1.f) - mean) * rstd;
            if (bias) v += to_f32(bias[n]);
            out[n] = from_f32<TY>(v);
        }
    }
}

// Column-parallel shards: per-token (sum u, sum u^2) over the local rows, in fp64.
template <typename TP>
__global__ void __launch_bounds__(kLnThreads)
partial_stats_kernel(const float* __restrict__ t, const TP* __restrict__ g, double* __restrict__ stats, int64_t N) {
    __shared__ double sh[2][kLnThreads / 32];
    const int64_t m = blockIdx.x;
    const float* row = t + m * N;
    double s = 0.0, q = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += kLnThreads) {
        const double v = (double)(row[n] * (g ? to_f32(g[n]) : 1.f));
        s += v;
        q += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = s;
        sh[1][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ss = 0.0, qq = 0.0;
        for (int i = 0; i < kLnThreads / 32; ++i) {
            ss += sh[0][i];
            qq += sh[1][i];
        }
        stats[2 * m] = ss;
        stats[2 * m + 1] = qq;
    }
}

template <typename TY, typename TP>
__global__ void __launch_bounds__(256)
apply_stats_kernel(const float* __restrict__ t, const TP* __restrict__ g, const TP* __restrict__ bias,
                   const double* __restrict__ stats, TY* __restrict__ y, int64_t N, int64_t n_global, float eps) {
    const int64_t m = blockIdx.y;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double mean = stats[2 * m] / (double)n_global;
    const double var = fmax(stats[2 * m + 1] / (double)n_global - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float v = (t[m * N + n] * (g ? to_f32(g[n]) : 1.f) - (float)mean) * rstd;
    if (bias) v += to_f32(bias[n]);
    y[m * N + n] = from_f32<TY>(v);
}

}  // namespace

int launch_scale_layernorm(const float* t, const void* g, const void* bias, void* y, int64_t m, int64_t n,
                           int act_dtype, int param_dtype, float eps, cudaStream_t s) {
    if (m == 0 || n == 0) return ONEBIT_OK;
    return dispatch_dtype(act_dtype, [&](auto yt) {
        using TY = decltype(yt);
        return dispatch_dtype(param_dtype, [&](auto pt) {
            using TP = decltype(pt);
            const TP* gp = static_cast<const TP*>(g);
            const TP* bp = static_cast<const TP*>(bias);
            TY* yp = static_cast<TY*>(y);
            const int64_t n4 = n / 4;
            if (n % 4 == 0 && n4 <= 2 * 256)
                scale_layernorm_vec_kernel<TY, TP, 2, 256><<<(unsigned)m, 256, 0, s>>>(t, gp, bp, yp, n, eps);
            else if (n % 4 == 0 && n4 <= 4 * 256)
                scale_layernorm_vec_kernel<TY, TP, 4, 256><<<(unsigned)m, 256, 0, s>>>(t, gp, bp, yp, n, eps);
            else if (n % 4 == 0 && n4 <= 8 * 512)
                scale_layernorm_vec_kernel<TY, TP, 8, 512><<<(unsigned)m, 512, 0, s>>>(t, gp, bp, yp, n, eps);
            else
                scale_layernorm_kernel<TY, TP><<<(unsigned)m, kLnThreads, 0, s>>>(t, gp, bp, yp, n, eps);
            ONEBIT_CUDA_TRY(cudaGetLastError());
            return ONEBIT_OK;
        });
    });
}

int launch_scale_partial_stats(const float* t, const void* g, double* stats, int64_t m, int64_t n, int param_dtype,
                               cudaStream_t s) {
    if (m == 0) return ONEBIT_OK;
    return dispatch_dtype(param_dtype, [&](auto pt) {
        using TP = decltype(pt);
        partial_stats_kernel<TP><<<(unsigned)m, kLnThreads, 0, s>>>(t, static_cast<const TP*>(g), stats, n);
        ONEBIT_CUDA_TRY(cudaGetLastError());
        return ONEBIT_OK;
    });
}

int launch_layernorm_apply_stats(const float* t, const void* g, const void* bias, const double* stats, void* y,
                                 int64_t m, int64_t n_local, int64_t n_global, int act_dtype, int param_dtype,
                                 float eps, cudaStream_t s) {
    if (m == 0 || n_local == 0) return ONEBIT_OK;
    ONEBIT_REQUIRE(m <= 65535, "layernorm_apply_stats: M too large");
    dim3 grid((unsigned)((n_local + 255) / 256), (unsigned)m);
    return dispatch_dtype(act_dtype, [&](auto yt) {
        using TY = decltype(yt);
        return dispatch_dtype(param_dtype, [&](auto pt) {
            using TP = decltype(pt);
            apply_stats_kernel<TY, TP><<<grid, 256, 0, s>>>(t, static_cast<const TP*>(g), static_cast<const TP*>(bias),
                                                             stats, static_cast<TY*>(y), n_local, n_global, eps);
            ONEBIT_CUDA_TRY(cudaGetLastError());
            return ONEBIT_OK;
        });
    });
}

}  // namespace onebit
