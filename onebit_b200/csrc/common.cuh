// Shared device/host helpers for libonebit_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/onebit_b200.h"

namespace onebit {

// ---- error plumbing (thread-local message behind onebit_last_error) -------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define ONEBIT_CUDA_TRY(expr)                                                                      \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::onebit::fail(ONEBIT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define ONEBIT_REQUIRE(cond, msg)                                                 \
    do {                                                                          \
        if (!(cond)) return ::onebit::fail(ONEBIT_ERR_INVALID_ARGUMENT, (msg));   \
    } while (0)

inline size_t dtype_size(int dt) { return dt == ONEBIT_F32 ? 4 : 2; }
inline bool dtype_ok(int dt) { return dt == ONEBIT_F16 || dt == ONEBIT_BF16 || dt == ONEBIT_F32; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

void pdl_suspend(bool on);  // per-thread: tensor-parallel steps interleave NCCL kernels from another stream
bool pdl_enabled();  // programmatic dependent launch on our own kernel chain (ONEBIT_PDL=0 disables)
int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device

// ---- dtype conversion -------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Dispatch a (runtime dtype) -> (template type) call. `F` is a generic lambda taking a value of the type.
template <typename F>
inline int dispatch_dtype(int dt, F&& f) {
    switch (dt) {
        case ONEBIT_F16: return f(__half{});
        case ONEBIT_BF16: return f(__nv_bfloat16{});
        case ONEBIT_F32: return f(float{});
        default: return fail(ONEBIT_ERR_INVALID_ARGUMENT, "unknown dtype code " + std::to_string(dt));
    }
}

// ---- kernel launchers implemented in the .cu files --------------------------------------------------
// t[m][n] = sum_k sign(n,k) * h[k] * x[m][k]   (optionally * g[n]); fp32 output.
int launch_matvec_simt(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m,
                       int64_t k, int64_t n, int act_dtype, int param_dtype, bool scale_by_g, cudaStream_t s);
bool matvec_mma_supported(int64_t m, int64_t k, int64_t n, int act_dtype);
size_t matvec_mma_workspace_bytes(int64_t m, int64_t k);
int launch_matvec_mma(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m,
                      int64_t k, int64_t n, int act_dtype, int param_dtype, bool scale_by_g, void* workspace,
                      cudaStream_t s);
bool prefill_tc5_supported(int64_t m, int64_t k, int64_t n);
size_t prefill_tc5_workspace_bytes(int64_t m, int64_t k, int act_dtype, int param_dtype);
int launch_prefill_tc5(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m, int64_t k,
                       int64_t n, int act_dtype, int param_dtype, bool scale_by_g, void* workspace, cudaStream_t s);
// tcgen05 path shared with the decoder: projections over the same fp16 activations, optional split-K (decode batches)
struct Tc5LaunchProblem {
    const int8_t* w;    // [N][K/8]
    const __half* h16;  // [K] input_factor as fp16
    const void* g;      // [N] weight_scale (param dtype) or nullptr
    float* t;           // [ksplit][M][ldt]
    int N;
    int ldt;            // leading dimension of t; 0 = N
};
struct Tc5Launch {
    const __half* x16;  // [M][K]
    int M, K, nprob, ksplit, param_dtype;
    int bf16;           // x16 and h16 hold bfloat16 (else fp16)
    Tc5LaunchProblem p[3];
};
int launch_tc5(const Tc5Launch& L, cudaStream_t s);
int launch_dense_tc5(const __half* x16, const __half* w16, float* out, int64_t m, int64_t k, int64_t n, cudaStream_t s);
int launch_to_half(const void* src, __half* dst, int64_t n, int dtype, cudaStream_t s);
int launch_quantize_tokens(const void* x, const void* h, uint8_t* digits, void* qmeta, int64_t m, int64_t k,
                           int act_dtype, int param_dtype, cudaStream_t s);

int launch_scale_layernorm(const float* t, const void* g, const void* bias, void* y, int64_t m, int64_t n,
                           int act_dtype, int param_dtype, float eps, cudaStream_t s);
int launch_scale_partial_stats(const float* t, const void* g, double* stats, int64_t m, int64_t n, int param_dtype,
                               cudaStream_t s);
int launch_layernorm_apply_stats(const float* t, const void* g, const void* bias, const double* stats, void* y,
                                 int64_t m, int64_t n_local, int64_t n_global, int act_dtype, int param_dtype,
                                 float eps, cudaStream_t s);
int launch_pack(const void* w, int8_t* packed, int64_t n, int64_t k, int dtype, cudaStream_t s);
int launch_unpack(const int8_t* packed, void* out, int64_t n, int64_t k, int dtype, cudaStream_t s);

}  // namespace onebit
