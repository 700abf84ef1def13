// Bit layout converters on the GPU.
//   pack   : scripts/convert_llama_to_infer_ckpt.py:7-15 (fp16_to_int8) — column 8j+i -> bit i of byte j,
//            bit = 1 <=> sign = -1; sign(0) = 0 packs as +1 (bit 0).
//   unpack : bitnet.py:98-110 (int8_to_fp16) — -2*bit + 1 in the requested dtype.
// One thread per packed byte; the eight source/destination values are contiguous.
#include "common.cuh"

namespace onebit {
namespace {

template <typename T>
__global__ void pack_kernel(const T* __restrict__ w, uint8_t* __restrict__ packed, int64_t total_bytes) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_bytes) return;
    const T* src = w + idx * 8;
    unsigned b = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b |= (to_f32(src[i]) <= -1.0f ? 1u : 0u) << i;
    packed[idx] = (uint8_t)b;
}

template <typename T>
__global__ void unpack_kernel(const uint8_t* __restrict__ packed, T* __restrict__ out, int64_t total_bytes) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total_bytes) return;
    const unsigned b = packed[idx];
    T* dst = out + idx * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = from_f32<T>(((b >> i) & 1u) ? -1.0f : 1.0f);
}

}  // namespace

int launch_pack(const void* w, int8_t* packed, int64_t n, int64_t k, int dtype, cudaStream_t s) {
    const int64_t total = n * (k / 8);
    if (total == 0) return ONEBIT_OK;
    return dispatch_dtype(dtype, [&](auto tt) {
        using T = decltype(tt);
        pack_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(static_cast<const T*>(w),
                                                                          reinterpret_cast<uint8_t*>(packed), total);
        ONEBIT_CUDA_TRY(cudaGetLastError());
        return ONEBIT_OK;
    });
}

int launch_unpack(const int8_t* packed, void* out, int64_t n, int64_t k, int dtype, cudaStream_t s) {
    const int64_t total = n * (k / 8);
    if (total == 0) return ONEBIT_OK;
    return dispatch_dtype(dtype, [&](auto tt) {
        using T = decltype(tt);
        unpack_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(packed),
                                                                            static_cast<T*>(out), total);
        ONEBIT_CUDA_TRY(cudaGetLastError());
        return ONEBIT_OK;
    });
}

}  // namespace onebit
