// Torch dispatcher face of the C ABI (SURVEY.md §8b "What a native replacement must export"):
//   onebit_b200::bitlinear(Tensor x, Tensor weight, Tensor weight_scale, Tensor input_factor, Tensor? bias, float eps) -> Tensor
//   onebit_b200::bitlinear_nolayernorm(Tensor x, Tensor weight, Tensor weight_scale, Tensor input_factor, bool scale_by_g) -> Tensor
//   onebit_b200::pack_signs(Tensor w) -> Tensor          onebit_b200::unpack_signs(Tensor packed, ScalarType dtype) -> Tensor
// The CUDA key forwards to libonebit_b200.so (include/onebit_b200.h) on at::cuda::getCurrentCUDAStream(); scratch
// comes from the caching allocator, so calls are CUDA-graph capturable. The CPU key exists so that a CPU tensor gets
// a precise error instead of "no kernel registered": this path has NO CPU implementation by design (the reference's
// own CPU arithmetic lives in oracle/, which is test infrastructure).
// Replaces BitLinearInf.forward / int8_to_fp16 (transformers/src/transformers/models/bitnet.py:98-122) and
// fp16_to_int8 (scripts/convert_llama_to_infer_ckpt.py:7-15).
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "onebit_b200.h"

namespace {

int dtype_code(at::ScalarType t, const char* what) {
    switch (t) {
        case at::kHalf: return ONEBIT_F16;
        case at::kBFloat16: return ONEBIT_BF16;
        case at::kFloat: return ONEBIT_F32;
        default: TORCH_CHECK(false, "onebit_b200: unsupported ", what, " dtype ", t, " (supported: float16, bfloat16, float32)");
    }
    return -1;
}

void check_rc(int rc, const char* what) { TORCH_CHECK(rc == ONEBIT_OK, what, " failed (code ", rc, "): ", onebit_last_error()); }

struct LayerDims { int64_t m, k, n; };

LayerDims check_layer(const at::Tensor& x, const at::Tensor& weight, const at::Tensor& g, const at::Tensor& h,
                      const c10::optional<at::Tensor>& bias) {
    TORCH_CHECK(x.is_cuda(), "onebit_b200: `input` lives on ", x.device(), "; the 1-bit linear path is CUDA (sm_100a) only and has no CPU fallback");
    TORCH_CHECK(weight.device() == x.device() && g.device() == x.device() && h.device() == x.device(),
                "onebit_b200: weight / weight_scale / input_factor must be on the input's device ", x.device());
    TORCH_CHECK(weight.scalar_type() == at::kChar, "onebit_b200: `weight` must be int8 bit-packed signs, got ", weight.scalar_type());
    TORCH_CHECK(weight.dim() == 2 && weight.is_contiguous(), "onebit_b200: `weight` must be a contiguous [out_features, in_features // 8] tensor");
    TORCH_CHECK(x.dim() >= 1, "onebit_b200: input needs at least one dimension");
    const int64_t n = weight.size(0), k = weight.size(1) * 8;
    TORCH_CHECK(x.size(-1) == k, "onebit_b200: input has ", x.size(-1), " features, the packed weight expects ", k);
    TORCH_CHECK(h.numel() == k && g.numel() == n, "onebit_b200: weight_scale / input_factor sizes do not match the packed weight");
    TORCH_CHECK(g.scalar_type() == h.scalar_type(), "onebit_b200: weight_scale and input_factor must share one dtype");
    if (bias.has_value() && bias->defined()) {
        TORCH_CHECK(bias->device() == x.device() && bias->numel() == n && bias->scalar_type() == g.scalar_type(),
                    "onebit_b200: bias must be [out_features] on the input's device with the dtype of weight_scale");
    }
    return {k == 0 ? 0 : x.numel() / k, k, n};
}

std::vector<int64_t> out_sizes(const at::Tensor& x, int64_t n) {
    auto s = x.sizes().vec();
    s.back() = n;
    return s;
}

at::Tensor bitlinear_cuda(const at::Tensor& x, const at::Tensor& weight, const at::Tensor& weight_scale,
                          const at::Tensor& input_factor, const c10::optional<at::Tensor>& bias, double eps) {
    const LayerDims d = check_layer(x, weight, weight_scale, input_factor, bias);
    const c10::cuda::CUDAGuard guard(x.device());
    at::Tensor y = at::empty(out_sizes(x, d.n), x.options());
    if (d.m == 0) return y;
    const at::Tensor xc = x.contiguous(), g = weight_scale.contiguous(), h = input_factor.contiguous();
    at::Tensor b;
    if (bias.has_value() && bias->defined()) b = bias->contiguous();
    const size_t ws_bytes = onebit_bitlinear_workspace_bytes(d.m, d.k, d.n);
    at::Tensor ws = at::empty({(int64_t)ws_bytes}, x.options().dtype(at::kByte));
    check_rc(onebit_bitlinear_forward(xc.data_ptr(), static_cast<const int8_t*>(weight.data_ptr()), g.data_ptr(), h.data_ptr(),
                                      b.defined() ? b.data_ptr() : nullptr, y.data_ptr(), d.m, d.k, d.n,
                                      dtype_code(x.scalar_type(), "activation"), dtype_code(g.scalar_type(), "parameter"), (float)eps,
                                      ws.data_ptr(), ws_bytes, ONEBIT_VARIANT_AUTO, c10::cuda::getCurrentCUDAStream().stream()),
             "onebit_bitlinear_forward");
    return y;
}

at::Tensor bitlinear_nolayernorm_cuda(const at::Tensor& x, const at::Tensor& weight, const at::Tensor& weight_scale,
                                      const at::Tensor& input_factor, bool scale_by_g) {
    const LayerDims d = check_layer(x, weight, weight_scale, input_factor, c10::nullopt);
    const c10::cuda::CUDAGuard guard(x.device());
    at::Tensor t = at::empty(out_sizes(x, d.n), x.options().dtype(at::kFloat));
    if (d.m == 0) return t;
    const at::Tensor xc = x.contiguous(), g = weight_scale.contiguous(), h = input_factor.contiguous();
    const size_t ws_bytes = onebit_matvec_workspace_bytes(d.m, d.k);
    at::Tensor ws = at::empty({(int64_t)ws_bytes}, x.options().dtype(at::kByte));
    check_rc(onebit_bitlinear_matvec(xc.data_ptr(), static_cast<const int8_t*>(weight.data_ptr()), g.data_ptr(), h.data_ptr(),
                                     t.data_ptr<float>(), d.m, d.k, d.n, dtype_code(x.scalar_type(), "activation"),
                                     dtype_code(g.scalar_type(), "parameter"), scale_by_g ? 1 : 0, ws.data_ptr(), ws_bytes,
                                     ONEBIT_VARIANT_AUTO, c10::cuda::getCurrentCUDAStream().stream()),
             "onebit_bitlinear_matvec");
    return t;
}

at::Tensor pack_signs_cuda(const at::Tensor& w) {
    TORCH_CHECK(w.is_cuda(), "onebit_b200: pack_signs is CUDA only (no CPU fallback)");
    TORCH_CHECK(w.dim() == 2 && w.size(1) % 8 == 0, "onebit_b200: pack_signs expects [N, K] with K % 8 == 0");
    const c10::cuda::CUDAGuard guard(w.device());
    const at::Tensor wc = w.contiguous();
    at::Tensor out = at::empty({w.size(0), w.size(1) / 8}, w.options().dtype(at::kChar));
    check_rc(onebit_pack_signs(wc.data_ptr(), static_cast<int8_t*>(out.data_ptr()), w.size(0), w.size(1),
                               dtype_code(w.scalar_type(), "weight"), c10::cuda::getCurrentCUDAStream().stream()),
             "onebit_pack_signs");
    return out;
}

at::Tensor unpack_signs_cuda(const at::Tensor& packed, at::ScalarType dtype) {
    TORCH_CHECK(packed.is_cuda(), "onebit_b200: unpack_signs is CUDA only (no CPU fallback)");
    TORCH_CHECK(packed.scalar_type() == at::kChar && packed.dim() == 2, "onebit_b200: unpack_signs expects an int8 [N, K/8] tensor");
    const c10::cuda::CUDAGuard guard(packed.device());
    const at::Tensor pc = packed.contiguous();
    at::Tensor out = at::empty({packed.size(0), packed.size(1) * 8}, packed.options().dtype(dtype));
    check_rc(onebit_unpack_signs(static_cast<const int8_t*>(pc.data_ptr()), out.data_ptr(), packed.size(0), packed.size(1) * 8,
                                 dtype_code(dtype, "output"), c10::cuda::getCurrentCUDAStream().stream()),
             "onebit_unpack_signs");
    return out;
}

// ---- CPU key: a precise refusal ------------------------------------------------------------------------------------
[[noreturn]] void refuse_cpu(const char* op) {
    TORCH_CHECK(false, "onebit_b200::", op, ": CPU tensors are refused — the 1-bit linear path is CUDA (sm_100a) only and has no CPU fallback. "
                "Move the module and its input to a B200 device.");
    abort();
}
at::Tensor bitlinear_cpu(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, const c10::optional<at::Tensor>&, double) {
    refuse_cpu("bitlinear");
}
at::Tensor bitlinear_nolayernorm_cpu(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, bool) {
    refuse_cpu("bitlinear_nolayernorm");
}
at::Tensor pack_signs_cpu(const at::Tensor&) { refuse_cpu("pack_signs"); }
at::Tensor unpack_signs_cpu(const at::Tensor&, at::ScalarType) { refuse_cpu("unpack_signs"); }

}  // namespace

TORCH_LIBRARY(onebit_b200, m) {
    m.def("bitlinear(Tensor x, Tensor weight, Tensor weight_scale, Tensor input_factor, Tensor? bias=None, float eps=1e-05) -> Tensor");
    m.def("bitlinear_nolayernorm(Tensor x, Tensor weight, Tensor weight_scale, Tensor input_factor, bool scale_by_g=True) -> Tensor");
    m.def("pack_signs(Tensor w) -> Tensor");
    m.def("unpack_signs(Tensor packed, ScalarType dtype) -> Tensor");
}
TORCH_LIBRARY_IMPL(onebit_b200, CUDA, m) {
    m.impl("bitlinear", &bitlinear_cuda);
    m.impl("bitlinear_nolayernorm", &bitlinear_nolayernorm_cuda);
    m.impl("pack_signs", &pack_signs_cuda);
    m.impl("unpack_signs", &unpack_signs_cuda);
}
TORCH_LIBRARY_IMPL(onebit_b200, CPU, m) {
    m.impl("bitlinear", &bitlinear_cpu);
    m.impl("bitlinear_nolayernorm", &bitlinear_nolayernorm_cpu);
    m.impl("pack_signs", &pack_signs_cpu);
    m.impl("unpack_signs", &unpack_signs_cpu);
}
