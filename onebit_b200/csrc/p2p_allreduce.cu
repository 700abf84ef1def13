// One-shot all-reduce over NVLink peer memory for the tensor-parallel decode step ("one-shot all-reduce over
// NVSwitch ... NCCL as the baseline"). The reference has no tensor parallelism for BitLinearInf; this implements the two
// collectives its LayerNorm-over-the-full-N forces on a sharded layer (partial sums of o_proj / down_proj, (sum, sum of
// squares) of q/k/v and gate/up) without a library call:
//   * every rank owns a symmetric buffer [3 rotations][n ranks][cap] of floats that every peer can address (the host side
//     passes the peer pointers; torch's symmetric memory does the mapping);
//   * a collective is ONE kernel: each thread PUSHES its element into slot `rank` of every peer's buffer (plain stores to
//     peer pointers: NVLink writes), then polls its own buffer until the element of every rank has arrived, and sums them in
//     rank order (identical bits on every rank). Words carry their own validity (Lamport): -0.0f means "not written yet",
//     real -0.0 values are sent as +0.0;
//   * three rotating regions: call k uses region k % 3 and re-arms region (k + 2) % 3 (= the one call k - 1 used), which no
//     peer can write again before it has received this rank's data of call k + 1. The call counter lives in device memory and
//     is bumped by the last CTA, so a captured CUDA graph replays correctly.
#include <cstdlib>

#include "common.cuh"
#include "p2p_allreduce.cuh"

namespace onebit {
namespace {

constexpr uint32_t kNotYet = 0x80000000u;  // -0.0f

__device__ __forceinline__ uint32_t ldv(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stv(uint32_t* p, uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(P2PComm c, float* __restrict__ data, int count) {
    __shared__ unsigned s_call;
    if (threadIdx.x == 0) s_call = *reinterpret_cast<volatile unsigned*>(c.call_counter);
    __syncthreads();
    const unsigned call = s_call;
    const int rot = (int)(call % 3u), clr = (int)((call + 2u) % 3u);
    const size_t region = (size_t)c.n * c.cap;
    uint32_t* mine = reinterpret_cast<uint32_t*>(c.peer[c.rank]);
    // re-arm the region the previous call used, over the extent THAT call used (its count may have been larger than ours)
    {
        const unsigned prev = c.last_count[clr];
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < prev; i += gridDim.x * blockDim.x)
            for (int r = 0; r < c.n; ++r) stv(mine + (size_t)clr * region + (size_t)r * c.cap + i, kNotYet);
    }
    const unsigned long long deadline = now_ns() + 2000000000ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        uint32_t v = __float_as_uint(data[i]);
        if (v == kNotYet) v = 0u;
#pragma unroll 1
        for (int p = 0; p < c.n; ++p)
            stv(reinterpret_cast<uint32_t*>(c.peer[p]) + (size_t)rot * region + (size_t)c.rank * c.cap + i, v);
        float acc = 0.f;
#pragma unroll 1
        for (int r = 0; r < c.n; ++r) {
            const uint32_t* src = mine + (size_t)rot * region + (size_t)r * c.cap + i;
            uint32_t w = ldv(src);
            int spins = 0;
            while (w == kNotYet) {
                if ((++spins & 1023) == 0 && now_ns() > deadline) {  // a lost peer must end the kernel, not hang the GPU
                    atomicExch(c.error_flag, 1);
                    break;
                }
                w = ldv(src);
            }
            acc += __uint_as_float(w);
        }
        data[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(c.cta_ticket, 1u) + 1u;
        if (done == gridDim.x) {  // last CTA: every CTA has read the call counter
            *c.cta_ticket = 0u;
            c.last_count[clr] = 0u;
            c.last_count[rot] = (unsigned)count;
            *reinterpret_cast<volatile unsigned*>(c.call_counter) = call + 1u;
        }
    }
}

}  // namespace

int p2p_allreduce(const P2PComm& c, float* data, int64_t count, cudaStream_t s) {
    ONEBIT_REQUIRE(c.n >= 2 && c.n <= kP2PMaxRanks && count >= 1 && (size_t)count <= c.cap, "p2p_allreduce: count exceeds the symmetric buffer");
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((count + threads - 1) / threads, 64);
    p2p_allreduce_kernel<<<blocks, threads, 0, s>>>(c, data, (int)count);
    ONEBIT_CUDA_TRY(cudaGetLastError());
    return ONEBIT_OK;
}

}  // namespace onebit
