// C-ABI entry points of libonebit_b200.so (declared in include/onebit_b200.h): argument checking,
// variant selection and the host-buffer layer handle. No torch types, no hidden global state apart
// from a per-thread error string and a cached SM count.
#include <cstdlib>
#include <new>

#include "common.cuh"

namespace onebit {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

static thread_local int t_pdl_suspended = 0;
void pdl_suspend(bool on) { t_pdl_suspended = on ? 1 : 0; }

bool pdl_enabled() {
    if (t_pdl_suspended) return false;
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ONEBIT_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

static int check_forward_args(const void* x, const int8_t* w, const void* g, const void* h, const void* out,
                              int64_t m, int64_t k, int64_t n, int act_dtype, int param_dtype) {
    ONEBIT_REQUIRE(dtype_ok(act_dtype), "act_dtype must be ONEBIT_F16/BF16/F32");
    ONEBIT_REQUIRE(dtype_ok(param_dtype), "param_dtype must be ONEBIT_F16/BF16/F32");
    ONEBIT_REQUIRE(m >= 0 && k > 0 && n > 0, "m must be >= 0 and k, n > 0");
    ONEBIT_REQUIRE(k % 8 == 0, "in_features (K) must be a multiple of 8: the weight holds 8 columns per int8 byte");
    if (m > 0) ONEBIT_REQUIRE(x && out, "x / output pointer is NULL");
    ONEBIT_REQUIRE(w && h, "weight / input_factor pointer is NULL");
    ONEBIT_REQUIRE(aligned16(x) && aligned16(w) && aligned16(g) && aligned16(h) && aligned16(out),
                   "device pointers must be 16-byte aligned");
    return ONEBIT_OK;
}

static int pick_variant(int variant, int64_t m, int64_t k, int64_t n, int act_dtype, int* chosen) {
    switch (variant) {
        case ONEBIT_VARIANT_AUTO:  // decode-size batches: bit-plane IMMA GEMV; larger: tcgen05; odd shapes: CUDA cores
            // (small batches at widths whose digits do not fit the GEMV's shared memory — e.g. 5..8 tokens at K = 11008 —
            //  go to the tcgen05 tile as well: the CUDA-core kernel is a 2-3 %-of-HBM anchor, not a fallback to land on)
            *chosen = matvec_mma_supported(m, k, n, act_dtype)
                          ? ONEBIT_VARIANT_MMA
                          : (m > 4 && prefill_tc5_supported(m, k, n) ? ONEBIT_VARIANT_TC5 : ONEBIT_VARIANT_SIMT);
            return ONEBIT_OK;
        case ONEBIT_VARIANT_SIMT:
            *chosen = variant;
            return ONEBIT_OK;
        case ONEBIT_VARIANT_MMA:
            ONEBIT_REQUIRE(matvec_mma_supported(m, k, n, act_dtype),
                           "ONEBIT_VARIANT_MMA does not support this shape/dtype (needs K % 256 == 0, K <= 14336, M <= 8)");
            *chosen = variant;
            return ONEBIT_OK;
        case ONEBIT_VARIANT_TC5:
            ONEBIT_REQUIRE(prefill_tc5_supported(m, k, n), "ONEBIT_VARIANT_TC5 does not support this shape (needs K % 64 == 0)");
            *chosen = variant;
            return ONEBIT_OK;
        default:
            return fail(ONEBIT_ERR_INVALID_ARGUMENT, "unknown variant code " + std::to_string(variant));
    }
}

static int matvec_impl(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m, int64_t k,
                       int64_t n, int act_dtype, int param_dtype, bool scale_by_g, void* ws, size_t ws_bytes,
                       int variant, cudaStream_t s) {
    int chosen = 0;
    int rc = pick_variant(variant, m, k, n, act_dtype, &chosen);
    if (rc != ONEBIT_OK) return rc;
    if (chosen == ONEBIT_VARIANT_MMA) {
        if (m > 0 && (!ws || !aligned16(ws) || ws_bytes < matvec_mma_workspace_bytes(m, k)))
            return fail(ONEBIT_ERR_WORKSPACE, "workspace missing, misaligned or smaller than onebit_matvec_workspace_bytes");
        return launch_matvec_mma(x, w, g, h, t, m, k, n, act_dtype, param_dtype, scale_by_g, ws, s);
    }
    if (chosen == ONEBIT_VARIANT_TC5) {
        if (m > 0 && (!ws || !aligned16(ws) || ws_bytes < prefill_tc5_workspace_bytes(m, k, act_dtype, param_dtype)))
            return fail(ONEBIT_ERR_WORKSPACE, "workspace missing, misaligned or smaller than onebit_matvec_workspace_bytes");
        return launch_prefill_tc5(x, w, g, h, t, m, k, n, act_dtype, param_dtype, scale_by_g, ws, s);
    }
    return launch_matvec_simt(x, w, g, h, t, m, k, n, act_dtype, param_dtype, scale_by_g, s);
}

}  // namespace onebit

using namespace onebit;

extern "C" {

const char* onebit_version(void) { return "onebit_b200 0.1.0 (sm_100a)"; }
const char* onebit_last_error(void) { return g_last_error.c_str(); }

int onebit_device_check(int device) {
    int count = 0;
    ONEBIT_CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return fail(ONEBIT_ERR_UNSUPPORTED_DEVICE, "no CUDA device " + std::to_string(device));
    int major = 0;
    ONEBIT_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10)
        return fail(ONEBIT_ERR_UNSUPPORTED_DEVICE,
                    "libonebit_b200 is built for sm_100a only; device has compute capability major " +
                        std::to_string(major));
    return ONEBIT_OK;
}

int onebit_pack_signs(const void* w, int8_t* packed, int64_t n, int64_t k, int dtype, void* stream) {
    ONEBIT_REQUIRE(dtype_ok(dtype), "dtype must be ONEBIT_F16/BF16/F32");
    ONEBIT_REQUIRE(n >= 0 && k >= 0 && k % 8 == 0, "pack_signs: K must be a multiple of 8");
    ONEBIT_REQUIRE((w && packed) || n * k == 0, "pack_signs: NULL pointer");
    return launch_pack(w, packed, n, k, dtype, static_cast<cudaStream_t>(stream));
}

int onebit_unpack_signs(const int8_t* packed, void* out, int64_t n, int64_t k, int dtype, void* stream) {
    ONEBIT_REQUIRE(dtype_ok(dtype), "dtype must be ONEBIT_F16/BF16/F32");
    ONEBIT_REQUIRE(n >= 0 && k >= 0 && k % 8 == 0, "unpack_signs: K must be a multiple of 8");
    ONEBIT_REQUIRE((out && packed) || n * k == 0, "unpack_signs: NULL pointer");
    return launch_unpack(packed, out, n, k, dtype, static_cast<cudaStream_t>(stream));
}

static size_t t_bytes(int64_t m, int64_t n) {  // t = S @ (h*x) in fp32, before scale + LayerNorm; 256-byte multiple
    return (((size_t)m * (size_t)n * sizeof(float)) + 255) & ~(size_t)255;
}

size_t onebit_matvec_workspace_bytes(int64_t m, int64_t k) {
    if (m <= 0 || k <= 0) return 16;
    // digits of the IMMA variant (small m) or the fp16 staging of x / h of the tcgen05 variant (fp32 worst case)
    const size_t a = m <= 8 ? matvec_mma_workspace_bytes(m, k) : 0;
    const size_t b = prefill_tc5_workspace_bytes(m, k, ONEBIT_F32, ONEBIT_F32);
    return a > b ? a : b;
}

size_t onebit_bitlinear_workspace_bytes(int64_t m, int64_t k, int64_t n) {
    if (m <= 0 || n <= 0) return 16;
    return t_bytes(m, n) + onebit_matvec_workspace_bytes(m, k);
}

int onebit_bitlinear_matvec(const void* x, const int8_t* weight, const void* weight_scale, const void* input_factor,
                            float* t, int64_t m, int64_t k, int64_t n, int act_dtype, int param_dtype, int scale_by_g,
                            void* workspace, size_t workspace_bytes, int variant, void* stream) {
    int rc = check_forward_args(x, weight, weight_scale, input_factor, t, m, k, n, act_dtype, param_dtype);
    if (rc != ONEBIT_OK) return rc;
    ONEBIT_REQUIRE(!scale_by_g || weight_scale, "scale_by_g set but weight_scale is NULL");
    return matvec_impl(x, weight, weight_scale, input_factor, t, m, k, n, act_dtype, param_dtype, scale_by_g != 0,
                       workspace, workspace_bytes, variant, static_cast<cudaStream_t>(stream));
}

int onebit_scale_layernorm(const float* t, const void* weight_scale, const void* bias, void* y, int64_t m, int64_t n,
                           int act_dtype, int param_dtype, float eps, void* stream) {
    ONEBIT_REQUIRE(dtype_ok(act_dtype) && dtype_ok(param_dtype), "bad dtype code");
    ONEBIT_REQUIRE(m >= 0 && n > 0, "scale_layernorm: m >= 0, n > 0");
    ONEBIT_REQUIRE((t && y) || m == 0, "scale_layernorm: NULL pointer");
    return launch_scale_layernorm(t, weight_scale, bias, y, m, n, act_dtype, param_dtype, eps,
                                  static_cast<cudaStream_t>(stream));
}

int onebit_scale_partial_stats(const float* t, const void* weight_scale, double* stats, int64_t m, int64_t n,
                               int param_dtype, void* stream) {
    ONEBIT_REQUIRE(dtype_ok(param_dtype), "bad dtype code");
    ONEBIT_REQUIRE(m >= 0 && n > 0 && ((t && stats) || m == 0), "scale_partial_stats: bad arguments");
    return launch_scale_partial_stats(t, weight_scale, stats, m, n, param_dtype, static_cast<cudaStream_t>(stream));
}

int onebit_layernorm_apply_stats(const float* t, const void* weight_scale, const void* bias, const double* stats,
                                 void* y, int64_t m, int64_t n_local, int64_t n_global, int act_dtype, int param_dtype,
                                 float eps, void* stream) {
    ONEBIT_REQUIRE(dtype_ok(act_dtype) && dtype_ok(param_dtype), "bad dtype code");
    ONEBIT_REQUIRE(m >= 0 && n_local > 0 && n_global >= n_local, "layernorm_apply_stats: bad sizes");
    ONEBIT_REQUIRE((t && y && stats) || m == 0, "layernorm_apply_stats: NULL pointer");
    return launch_layernorm_apply_stats(t, weight_scale, bias, stats, y, m, n_local, n_global, act_dtype, param_dtype,
                                        eps, static_cast<cudaStream_t>(stream));
}

int onebit_bitlinear_forward(const void* x, const int8_t* weight, const void* weight_scale, const void* input_factor,
                             const void* bias, void* y, int64_t m, int64_t k, int64_t n, int act_dtype, int param_dtype,
                             float eps, void* workspace, size_t workspace_bytes, int variant, void* stream) {
    int rc = check_forward_args(x, weight, weight_scale, input_factor, y, m, k, n, act_dtype, param_dtype);
    if (rc != ONEBIT_OK) return rc;
    ONEBIT_REQUIRE(weight_scale, "weight_scale pointer is NULL");
    if (m == 0) return ONEBIT_OK;
    if (!workspace || workspace_bytes < onebit_bitlinear_workspace_bytes(m, k, n) || !aligned16(workspace))
        return fail(ONEBIT_ERR_WORKSPACE, "workspace missing, misaligned or smaller than onebit_bitlinear_workspace_bytes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* t = static_cast<float*>(workspace);
    rc = matvec_impl(x, weight, weight_scale, input_factor, t, m, k, n, act_dtype, param_dtype, /*scale_by_g=*/false,
                     static_cast<char*>(workspace) + t_bytes(m, n), workspace_bytes - t_bytes(m, n), variant, s);
    if (rc != ONEBIT_OK) return rc;
    return launch_scale_layernorm(t, weight_scale, bias, y, m, n, act_dtype, param_dtype, eps, s);
}

// ---- host-buffer layer handle ---------------------------------------------------------------------
struct onebit_layer {
    int8_t* w = nullptr;
    void* g = nullptr;
    void* h = nullptr;
    void* bias = nullptr;
    void* x = nullptr;
    void* y = nullptr;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int64_t k = 0, n = 0, max_m = 0;
    int act_dtype = 0, param_dtype = 0;
    float eps = 1e-5f;
};

void onebit_layer_destroy(onebit_layer* L) {
    if (!L) return;
    cudaFree(L->w);
    cudaFree(L->g);
    cudaFree(L->h);
    cudaFree(L->bias);
    cudaFree(L->x);
    cudaFree(L->y);
    cudaFree(L->ws);
    delete L;
}

int onebit_layer_create(onebit_layer** out, const int8_t* weight_host, const void* weight_scale_host,
                        const void* input_factor_host, const void* bias_host, int64_t k, int64_t n, int act_dtype,
                        int param_dtype, float eps, int64_t max_m) {
    ONEBIT_REQUIRE(out, "layer_create: out is NULL");
    *out = nullptr;
    ONEBIT_REQUIRE(dtype_ok(act_dtype) && dtype_ok(param_dtype), "bad dtype code");
    ONEBIT_REQUIRE(k > 0 && n > 0 && k % 8 == 0 && max_m > 0, "layer_create: k % 8 == 0, n > 0, max_m > 0 required");
    ONEBIT_REQUIRE(weight_host && weight_scale_host && input_factor_host, "layer_create: NULL host buffer");
    onebit_layer* L = new (std::nothrow) onebit_layer();
    ONEBIT_REQUIRE(L, "layer_create: out of host memory");
    L->k = k;
    L->n = n;
    L->max_m = max_m;
    L->act_dtype = act_dtype;
    L->param_dtype = param_dtype;
    L->eps = eps;
    const size_t ps = dtype_size(param_dtype), as = dtype_size(act_dtype);
    L->ws_bytes = onebit_bitlinear_workspace_bytes(max_m, k, n);
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dst, const void* src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dst, bytes);
        if (e == cudaSuccess && src) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    up(reinterpret_cast<void**>(&L->w), weight_host, (size_t)n * (size_t)(k / 8));
    up(&L->g, weight_scale_host, (size_t)n * ps);
    up(&L->h, input_factor_host, (size_t)k * ps);
    if (bias_host) up(&L->bias, bias_host, (size_t)n * ps);
    up(&L->x, nullptr, (size_t)max_m * (size_t)k * as);
    up(&L->y, nullptr, (size_t)max_m * (size_t)n * as);
    up(&L->ws, nullptr, L->ws_bytes);
    if (e != cudaSuccess) {
        onebit_layer_destroy(L);
        return fail(ONEBIT_ERR_CUDA, std::string("layer_create: ") + cudaGetErrorString(e));
    }
    *out = L;
    return ONEBIT_OK;
}

int onebit_layer_forward_device(onebit_layer* L, const void* x_dev, void* y_dev, int64_t m, void* stream) {
    ONEBIT_REQUIRE(L, "layer is NULL");
    ONEBIT_REQUIRE(m >= 0 && m <= L->max_m, "layer_forward: m exceeds max_m given at creation");
    return onebit_bitlinear_forward(x_dev, L->w, L->g, L->h, L->bias, y_dev, m, L->k, L->n, L->act_dtype,
                                    L->param_dtype, L->eps, L->ws, L->ws_bytes, ONEBIT_VARIANT_AUTO, stream);
}

int onebit_layer_forward_host(onebit_layer* L, const void* x_host, void* y_host, int64_t m, void* stream) {
    ONEBIT_REQUIRE(L, "layer is NULL");
    ONEBIT_REQUIRE(m >= 0 && m <= L->max_m, "layer_forward: m exceeds max_m given at creation");
    if (m == 0) return ONEBIT_OK;
    ONEBIT_REQUIRE(x_host && y_host, "layer_forward_host: NULL host buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t as = dtype_size(L->act_dtype);
    ONEBIT_CUDA_TRY(cudaMemcpyAsync(L->x, x_host, (size_t)m * (size_t)L->k * as, cudaMemcpyHostToDevice, s));
    int rc = onebit_layer_forward_device(L, L->x, L->y, m, stream);
    if (rc != ONEBIT_OK) return rc;
    ONEBIT_CUDA_TRY(cudaMemcpyAsync(y_host, L->y, (size_t)m * (size_t)L->n * as, cudaMemcpyDeviceToHost, s));
    ONEBIT_CUDA_TRY(cudaStreamSynchronize(s));
    return ONEBIT_OK;
}

}  // extern "C"
