// Micro-benchmarks that decide the decode-step architecture on B200:
//  (1) dependent chain of tiny kernels in a CUDA graph, with / without programmatic dependent launch
//  (2) grid-wide barrier cost inside one persistent cooperative kernel (148 CTAs)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_launch tools/ubench_launch.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void tiny(float* p, int pdl) {
    if (pdl) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
    if (threadIdx.x == 0) p[blockIdx.x] += 1.0f;
}

// sense-reversing grid barrier on a global counter (all CTAs co-resident)
__device__ __forceinline__ void grid_bar(unsigned* counter, unsigned nblocks, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        __threadfence();
        unsigned target = epoch * nblocks;
        unsigned v = atomicAdd(counter, 1u) + 1;
        while (v < target) { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter)); }
        __threadfence();
    }
    __syncthreads();
}

__global__ void persistent(unsigned* counter, float* data, int nbar, int mode) {
    unsigned epoch = 0;
    cg::grid_group grid = cg::this_grid();
    for (int i = 0; i < nbar; ++i) {
        if (threadIdx.x < 32) data[blockIdx.x * 32 + threadIdx.x] += 1.0f;  // a little global traffic per phase
        if (mode == 0) grid_bar(counter, gridDim.x, epoch);
        else grid.sync();
    }
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* buf; CK(cudaMalloc(&buf, 1 << 20)); CK(cudaMemset(buf, 0, 1 << 20));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int N = 300;
    for (int grid : {1, 128, 688}) for (int pdl = 0; pdl < 2; ++pdl) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < N; ++i) {
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = pdl;
            CK(cudaLaunchKernelEx(&cfg, tiny, buf, pdl));
        }
        CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
        for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("graph chain of %d tiny kernels, grid=%3d, pdl=%d: %.3f us per kernel\n", N, grid, pdl, ms * 1e3 / (10 * N));
    }
    unsigned* counter; CK(cudaMalloc(&counter, 4));
    for (int mode = 0; mode < 2; ++mode) for (int threads : {256, 512}) {
        int nbar = 2000;
        CK(cudaMemset(counter, 0, 4));
        void* args[] = {&counter, &buf, &nbar, &mode};
        CK(cudaLaunchCooperativeKernel((void*)persistent, dim3(sms), dim3(threads), args, 0, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaMemset(counter, 0, 4));
        CK(cudaEventRecord(e0, s));
        CK(cudaLaunchCooperativeKernel((void*)persistent, dim3(sms), dim3(threads), args, 0, s));
        CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("persistent %d CTAs x %d thr, %s barrier: %.3f us per barrier phase\n", sms, threads,
               mode == 0 ? "atomic-counter" : "cg::grid.sync", ms * 1e3 / nbar);
    }
    return 0;
}
