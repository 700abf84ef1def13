// Stand-alone launchers of the bit-plane IMMA GEMV (see imma_gemv.cuh for the algorithm):
//   quantize_tokens_kernel : x, h -> activation digits + per-token meta   (one CTA per token)
//   imma::gemv_kernel      : digits, packed signs -> t = S @ (h*x) in fp32 (+ optional g scale)
#include <algorithm>

#include "imma_gemv.cuh"

namespace onebit {
namespace {

using namespace imma;

constexpr int kQuantThreads = 512;

template <typename TX, typename TP>
__global__ void __launch_bounds__(kQuantThreads)
quantize_tokens_kernel(const TX* __restrict__ x, const TP* __restrict__ h, uint8_t* __restrict__ digits,
                       QMeta* __restrict__ qmeta, int K) {
    extern __shared__ __align__(16) float xs[];  // x'[K]
    __shared__ long long scratch[kQuantThreads / 32];
    pdl_launch_dependents();
    pdl_wait();
    const int m = blockIdx.x;
    const TX* xr = x + (size_t)m * K;
    float am = 0.f;
    for (int k = threadIdx.x; k < K; k += kQuantThreads) {
        const float v = to_f32(xr[k]) * to_f32(h[k]);  // bitnet.py:113
        xs[k] = v;
        am = fmaxf(am, fabsf(v));
    }
    __syncthreads();
    quantize_from_smem(xs, K, am, digits + (size_t)m * (K / kUnitCols) * kUnitBytes, qmeta + m, scratch);
}

template <typename KernT>
int set_smem_limit(KernT kern, size_t bytes) {
    ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return ONEBIT_OK;
}

template <typename TP, int NT, int UPW>
int launch_gemv_inst(const Args& a, int total_ctas, cudaStream_t s) {
    auto kern = gemv_kernel<TP, NT, UPW>;
    static bool configured[64] = {false};  // per instance and device: the attribute is sticky
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        int rc = set_smem_limit(kern, 224 * 1024);
        if (rc != ONEBIT_OK) return rc;
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)total_ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = gemv_smem_bytes(a.M, a.units, NT);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    ONEBIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
    return ONEBIT_OK;
}

template <typename TP, int NT>
int launch_gemv_nt(const Args& a, int total_ctas, cudaStream_t s) {
    const int upw = (a.units + kWarps - 1) / kWarps;
    if (upw <= 2) return launch_gemv_inst<TP, NT, 2>(a, total_ctas, s);
    if (upw <= 3) return launch_gemv_inst<TP, NT, 3>(a, total_ctas, s);
    if (upw <= 6) return launch_gemv_inst<TP, NT, 6>(a, total_ctas, s);
    return launch_gemv_inst<TP, NT, 7>(a, total_ctas, s);
}

}  // namespace

// Launch the GEMV over a prepared problem list (also used by the decoder).
int launch_imma_gemv(const imma::Args& a_in, int param_dtype, cudaStream_t s) {
    imma::Args a = a_in;
    int ctas = 0;
    for (int i = 0; i < a.nprob; ++i) {
        a.p[i].cta_begin = ctas;
        ctas += (a.p[i].n_rows + kRows - 1) / kRows;
    }
    if (ctas == 0 || a.M == 0) return ONEBIT_OK;
    const int nt = a.M <= 2 ? 1 : (a.M <= 4 ? 2 : 4);
    ONEBIT_REQUIRE(gemv_smem_bytes(a.M, a.units, nt) <= 224 * 1024,
                   "imma gemv: activation digits do not fit in shared memory");
    return dispatch_dtype(param_dtype, [&](auto pt) {
        using TP = decltype(pt);
        switch (nt) {
            case 1: return launch_gemv_nt<TP, 1>(a, ctas, s);
            case 2: return launch_gemv_nt<TP, 2>(a, ctas, s);
            default: return launch_gemv_nt<TP, 4>(a, ctas, s);
        }
    });
}

int launch_quantize_tokens(const void* x, const void* h, uint8_t* digits, void* qmeta, int64_t m, int64_t k,
                           int act_dtype, int param_dtype, cudaStream_t s) {
    if (m == 0) return ONEBIT_OK;
    return dispatch_dtype(act_dtype, [&](auto xt) {
        using TX = decltype(xt);
        return dispatch_dtype(param_dtype, [&](auto pt) {
            using TP = decltype(pt);
            auto kern = quantize_tokens_kernel<TX, TP>;
            static bool configured[64] = {false};
            int dev = 0;
            ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 0 && dev < 64 && !configured[dev]) {
                int rc = set_smem_limit(kern, 64 * 1024);
                if (rc != ONEBIT_OK) return rc;
                configured[dev] = true;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)m);
            cfg.blockDim = dim3(kQuantThreads);
            cfg.dynamicSmemBytes = (size_t)k * sizeof(float);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl_enabled() ? 1 : 0;
            ONEBIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, static_cast<const TX*>(x), static_cast<const TP*>(h), digits,
                                               static_cast<QMeta*>(qmeta), (int)k));
            return ONEBIT_OK;
        });
    });
}

bool matvec_mma_supported(int64_t m, int64_t k, int64_t n, int act_dtype) {
    (void)act_dtype;
    if (!(m >= 1 && m <= kMaxTokens && k % kUnitCols == 0 && k <= 14336 && n >= 1 && n < (1 << 30))) return false;
    const int nt = m <= 2 ? 1 : (m <= 4 ? 2 : 4);
    return gemv_smem_bytes((int)m, (int)(k / kUnitCols), nt) <= 224 * 1024;
}

size_t matvec_mma_workspace_bytes(int64_t m, int64_t k) {
    return (size_t)m * (size_t)(k / kUnitCols) * kUnitBytes + (size_t)m * sizeof(QMeta) + 32;
}

int launch_matvec_mma(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m, int64_t k,
                      int64_t n, int act_dtype, int param_dtype, bool scale_by_g, void* workspace, cudaStream_t s) {
    if (m == 0 || n == 0) return ONEBIT_OK;
    ONEBIT_REQUIRE(matvec_mma_supported(m, k, n, act_dtype), "matvec_mma: unsupported shape");
    uint8_t* digits = static_cast<uint8_t*>(workspace);
    QMeta* qmeta = reinterpret_cast<QMeta*>(digits + (size_t)m * (k / kUnitCols) * kUnitBytes);
    int rc = launch_quantize_tokens(x, h, digits, qmeta, m, k, act_dtype, param_dtype, s);
    if (rc != ONEBIT_OK) return rc;
    Args a = {};
    a.nprob = 1;
    a.M = (int)m;
    a.K = (int)k;
    a.units = (int)(k / kUnitCols);
    a.p[0].w = reinterpret_cast<const uint8_t*>(w);
    a.p[0].g = scale_by_g ? g : nullptr;
    a.p[0].digits = digits;
    a.p[0].qmeta = qmeta;
    a.p[0].t = t;
    a.p[0].stats = nullptr;
    a.p[0].n_rows = (int)n;
    a.p[0].ld_t = (int)n;
    return launch_imma_gemv(a, param_dtype, s);
}

}  // namespace onebit

// Debug only (side builds with -DONEBIT_TRACE): clock64() stamps of CTA 0 / thread 0 of the last GEMV launch.
extern "C" __attribute__((visibility("default"))) int onebit_debug_read_trace(long long* out8) {
#ifdef ONEBIT_TRACE
    return cudaMemcpyFromSymbol(out8, onebit::imma::g_trace, sizeof(long long) * 8) == cudaSuccess ? 0 : -2;
#else
    (void)out8;
    return -1;
#endif
}
