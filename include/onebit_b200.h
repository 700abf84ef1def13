/*
 * onebit_b200.h — C ABI of the B200-native OneBit 1-bit linear-layer path (libonebit_b200.so).
 *
 * The reference (xuyuzhuang11/OneBit @ 42d6d7b) has no FFI for this path: the hot path is the Python
 * nn.Module `BitLinearInf` (transformers/src/transformers/models/bitnet.py:71-122) built from ATen ops.
 * Each entry point below names the reference lines it replaces. All pointers are plain device (or,
 * where stated, host) pointers; no torch types cross this boundary. `stream` is a cudaStream_t passed
 * as void* (NULL = legacy default stream). Every function returns ONEBIT_OK (0) or a negative error
 * code; onebit_last_error() returns the message of the calling thread's last failure.
 *
 * All device entry points are asynchronous on `stream`, perform no allocation and no host
 * synchronisation, and are CUDA-graph capturable. Weights are read-only.
 *
 * Data layout (bitnet.py:78-80, scripts/convert_llama_to_infer_ckpt.py:7-15):
 *   weight        int8  [N, K/8] row-major; column 8j+i of row n is bit i (LSB first) of byte (n, j);
 *                       bit = 1 <=> sign = -1, bit = 0 <=> sign = +1.
 *   weight_scale  [N]   (g)      input_factor [K] (h)      bias [N] or NULL
 *   x             [M, K] row-major (M = product of the leading dims)   y [M, N]
 */
#ifndef ONEBIT_B200_H
#define ONEBIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ONEBIT_API __attribute__((visibility("default")))
#else
#define ONEBIT_API
#endif

#define ONEBIT_OK 0
#define ONEBIT_ERR_INVALID_ARGUMENT (-1) /* shape / dtype / alignment contract violated            */
#define ONEBIT_ERR_CUDA (-2)             /* a CUDA runtime call or launch failed                    */
#define ONEBIT_ERR_UNSUPPORTED_DEVICE (-3) /* not an sm_100 (B200) device                           */
#define ONEBIT_ERR_WORKSPACE (-4)        /* workspace too small / NULL                              */

typedef enum onebit_dtype {
    ONEBIT_F16 = 0,
    ONEBIT_BF16 = 1,
    ONEBIT_F32 = 2
} onebit_dtype;

/* Kernel selection for onebit_bitlinear_forward (0 = automatic by M/K/N; others force one variant,
 * used by tests and the bench to compare variants). */
typedef enum onebit_variant {
    ONEBIT_VARIANT_AUTO = 0,
    ONEBIT_VARIANT_SIMT = 1, /* CUDA-core packed GEMV, any K % 8 == 0                                */
    ONEBIT_VARIANT_MMA = 2,  /* bit-plane int8 mma.sync (IMMA), K % 256 == 0, M <= 8                  */
    ONEBIT_VARIANT_TC5 = 3   /* tcgen05 + TMEM, large M (prefill)                                    */
} onebit_variant;

ONEBIT_API const char* onebit_version(void);
ONEBIT_API const char* onebit_last_error(void);

/* ONEBIT_OK iff `device` exists and is compute capability 10.x. */
ONEBIT_API int onebit_device_check(int device);

/* ---- bit layout ---------------------------------------------------------------------------------
 * onebit_pack_signs   replaces fp16_to_int8, scripts/convert_llama_to_infer_ckpt.py:7-15:
 *                     bit = (uint8)((1 - w)/2) for w in {+1,-1,0}; in general bit = 1 iff w <= -1.
 * onebit_unpack_signs replaces BitLinearInf.int8_to_fp16, bitnet.py:98-110 (-2*bit + 1 in `dtype`).
 * Requires K % 8 == 0. */
ONEBIT_API int onebit_pack_signs(const void* w, int8_t* packed, int64_t n, int64_t k, int dtype, void* stream);
ONEBIT_API int onebit_unpack_signs(const int8_t* packed, void* out, int64_t n, int64_t k, int dtype, void* stream);

/* ---- forward ------------------------------------------------------------------------------------
 * onebit_bitlinear_forward replaces BitLinearInf.forward, bitnet.py:112-122:
 *     y = LayerNorm_N( g * ( S @ (h * x) ) ) (+ bias),  LayerNorm without affine, biased variance, eps.
 * act_dtype: dtype of x and y. param_dtype: dtype of weight_scale / input_factor / bias.
 * Requires K % 8 == 0, M >= 0, all pointers 16-byte aligned (torch allocations are).
 * workspace: at least onebit_bitlinear_workspace_bytes(m, k, n) bytes of device memory (fp32 pre-LayerNorm
 * tensor + the activation digits of the tensor-core variant), 16-byte aligned. */
ONEBIT_API size_t onebit_bitlinear_workspace_bytes(int64_t m, int64_t k, int64_t n);
ONEBIT_API size_t onebit_matvec_workspace_bytes(int64_t m, int64_t k);
ONEBIT_API int onebit_bitlinear_forward(const void* x, const int8_t* weight, const void* weight_scale,
                             const void* input_factor, const void* bias, void* y, int64_t m, int64_t k,
                             int64_t n, int act_dtype, int param_dtype, float eps, void* workspace,
                             size_t workspace_bytes, int variant, void* stream);

/* The two halves of the forward, exposed for tensor-parallel shards and fused consumers
 * (the LayerNorm of bitnet.py:118 is a reduction over the FULL N, so a row- or column-sharded
 * layer must reduce across ranks between the halves — see DESIGN.md "Multi-GPU").
 *   onebit_bitlinear_matvec : t = S @ (h * x), fp32 [M, N]; `scale_by_g` != 0 multiplies by g
 *                             (bitnet.py:113-116). For a K-shard pass scale_by_g = 0, all-reduce t,
 *                             then finish with onebit_scale_layernorm. workspace: at least
 *                             onebit_matvec_workspace_bytes(m, k) bytes.
 *   onebit_scale_layernorm  : y = LayerNorm_N(g * t) (+ bias); weight_scale == NULL means t is
 *                             already scaled (bitnet.py:116-120). */
ONEBIT_API int onebit_bitlinear_matvec(const void* x, const int8_t* weight, const void* weight_scale,
                            const void* input_factor, float* t, int64_t m, int64_t k, int64_t n,
                            int act_dtype, int param_dtype, int scale_by_g, void* workspace,
                            size_t workspace_bytes, int variant, void* stream);
ONEBIT_API int onebit_scale_layernorm(const float* t, const void* weight_scale, const void* bias, void* y, int64_t m,
                           int64_t n, int act_dtype, int param_dtype, float eps, void* stream);
/* Column-parallel (N-sharded) LayerNorm: per-token partial sums (sum, sum of squares) of g*t over the
 * local N rows -> stats[m][2] (fp64, so that sum-of-squares statistics stay exact enough), to be all-reduced (SUM), then onebit_layernorm_apply_stats with the global
 * N. */
ONEBIT_API int onebit_scale_partial_stats(const float* t, const void* weight_scale, double* stats, int64_t m, int64_t n,
                               int param_dtype, void* stream);
ONEBIT_API int onebit_layernorm_apply_stats(const float* t, const void* weight_scale, const void* bias,
                                 const double* stats, void* y, int64_t m, int64_t n_local, int64_t n_global,
                                 int act_dtype, int param_dtype, float eps, void* stream);

/* ---- host-buffer convenience (what a non-torch caller binds; used for the end-to-end timing) -------
 * A layer handle owns DEVICE copies of weight / weight_scale / input_factor / bias made from HOST
 * buffers once (the reference's `model.to(cuda)`, evaluation/lm_eval.py:68), plus its workspace.
 * onebit_layer_forward_host copies x host->device, runs the forward, copies y device->host and
 * synchronises `stream` before returning. x_host / y_host should be pinned for full speed. */
typedef struct onebit_layer onebit_layer;
ONEBIT_API int onebit_layer_create(onebit_layer** out, const int8_t* weight_host, const void* weight_scale_host,
                        const void* input_factor_host, const void* bias_host, int64_t k, int64_t n,
                        int act_dtype, int param_dtype, float eps, int64_t max_m);
ONEBIT_API int onebit_layer_forward_host(onebit_layer* layer, const void* x_host, void* y_host, int64_t m, void* stream);
ONEBIT_API int onebit_layer_forward_device(onebit_layer* layer, const void* x_dev, void* y_dev, int64_t m, void* stream);
ONEBIT_API void onebit_layer_destroy(onebit_layer* layer);

/* ---- fused decode step (the callers either side of the hot path) -------------------
 * One decode step of the reference's BitLlamaForCausalLMInf (modeling_bitllama.py:1512-1675) for `batch`
 * sequences with a static KV cache: embedding (:1202), per layer RMSNorm (:67-81) -> q/k/v BitLinearInf
 * (:522-524) -> RoPE (:176-181) -> attention over the cache (:543-563) -> o_proj (:580) -> residual ->
 * RMSNorm -> gate/up BitLinearInf -> SiLU*up -> down_proj (:257) -> residual; final RMSNorm (:1315), lm_head
 * (:1610) and greedy argmax (generation/utils.py:2540). Batches of 1..4 sequences run every BitLinear through the
 * bit-plane IMMA GEMV, 5..64 through the tcgen05 decode tile; the glue between them (LayerNorm of bitnet.py:118,
 * residual adds, norms, activation quantisation) is fused into the GEMV stages (5 launches per layer at batch 1..2)
 * or into small kernels (9 per layer otherwise); a step is CUDA-graph capturable (position and token ids live in
 * device memory and are advanced on the device).
 * All pointers in the parameter structs are DEVICE pointers that must outlive the decoder. Floating-point
 * parameters (weight_scale, input_factor, norm weights) are `param_dtype`; embed_tokens / lm_head are fp16. */
typedef struct onebit_bitlinear_params {
    const int8_t* weight;        /* [N, K/8] packed signs */
    const void* weight_scale;    /* [N] */
    const void* input_factor;    /* [K] */
} onebit_bitlinear_params;

typedef struct onebit_layer_params {
    onebit_bitlinear_params q, k, v, o, gate, up, down;
    const void* input_layernorm;           /* [hidden] RMSNorm weight */
    const void* post_attention_layernorm;  /* [hidden] */
} onebit_layer_params;

typedef struct onebit_decoder_config {
    int hidden_size, intermediate_size, num_layers, num_heads, vocab_size, max_seq_len, max_batch;
    int param_dtype;   /* onebit_dtype of weight_scale / input_factor / norm weights */
    float rms_eps;     /* config.rms_norm_eps */
    float ln_eps;      /* 1e-5, nn.LayerNorm default inside BitLinearInf */
    /* tensor parallelism (1 = none). With tp_size > 1 the decoder holds the rank's shard: q/k/v/gate/up are
     * row (N) shards, o/down are byte-column (K) shards; the caller provides an all-reduce callback. */
    int tp_size, tp_rank;
} onebit_decoder_config;

/* All-reduce (SUM, in place) of `count` fp32 values on `stream`, supplied by the host (NCCL in production,
 * nothing when tp_size == 1). Must be stream-ordered and capturable. */
typedef int (*onebit_allreduce_fn)(void* user, float* data, int64_t count, void* stream);

typedef struct onebit_decoder onebit_decoder;
ONEBIT_API int onebit_decoder_create(onebit_decoder** out, const onebit_decoder_config* cfg,
                                     const onebit_layer_params* layers, const void* embed_tokens_f16,
                                     const void* final_norm, const void* lm_head_f16, const float* rope_cos,
                                     const float* rope_sin, onebit_allreduce_fn allreduce, void* allreduce_user);
/* Set the device-resident state: token ids to feed next [batch] and their positions [batch]. Host pointers. */
ONEBIT_API int onebit_decoder_reset(onebit_decoder* dec, const int64_t* ids_host, const int32_t* pos_host, int batch,
                                    void* stream);
/* One step for `batch` sequences: consumes the device-resident ids/positions, writes logits (fp32
 * [batch, vocab], device, optional) and leaves next ids (argmax) + advanced positions on the device.
 * `forced_ids_dev` (optional, device int64 [batch]) overrides the fed ids (teacher forcing / prompt). */
ONEBIT_API int onebit_decoder_step(onebit_decoder* dec, int batch, const int64_t* forced_ids_dev, float* logits_dev,
                                   void* stream);
/* Host-buffer step for the end-to-end timing: ids from host (pinned), next ids back to host, synchronises. */
ONEBIT_API int onebit_decoder_step_host(onebit_decoder* dec, int batch, const int64_t* ids_host,
                                        int64_t* next_ids_host, void* stream);
/* Only the BitLinear GEMV launches of a step (4 per layer) on whatever activation digits are resident — the
 * weight-streaming kernel chain by itself, used by bench.py for the roofline measurement. */
ONEBIT_API int onebit_decoder_gemv_only(onebit_decoder* dec, int batch, void* stream);
ONEBIT_API const int64_t* onebit_decoder_next_ids(onebit_decoder* dec);  /* device int64 [max_batch] */
ONEBIT_API const int32_t* onebit_decoder_positions(onebit_decoder* dec); /* device int32 [max_batch] */
ONEBIT_API int onebit_decoder_kernel_launches_per_step(onebit_decoder* dec);
/* Persistent single-kernel step (experimental, opt-in with ONEBIT_PERSIST=1; batch <= 2, tp_size == 1): 1 if this
 * decoder uses it. */
ONEBIT_API int onebit_decoder_is_persistent(onebit_decoder* dec);
/* Prompt pass (replaces the q_len > 1 forward of BitLlamaForCausalLMInf, modeling_bitllama.py:1217-1315,1546-1611, with the
 * causal mask of :1267-1269): all T tokens of `batch` sequences at once — every BitLinear is one tcgen05 GEMM over
 * batch * T tokens, attention a causal flash kernel; K / V are written to the static cache at positions pos0 .. pos0+T-1.
 * ids_dev: device int64 [batch][T]. logits_last_dev: device fp32 [batch][V] or NULL. logits_all_dev: device fp32
 * [batch*T][V] or NULL. Afterwards the decoder's next ids are each prompt's greedy continuation and its positions
 * pos0 + T, so onebit_decoder_step continues the sequences. Single-GPU decoders only. Not CUDA-graph capturable (it
 * sizes and may allocate its workspace). */
ONEBIT_API int onebit_decoder_prefill(onebit_decoder* dec, int batch, int T, int pos0, const int64_t* ids_dev,
                                       float* logits_last_dev, float* logits_all_dev, void* stream);
/* Tensor parallelism without a library collective: switch the decoder's all-reduces (partial sums of o_proj / down_proj,
 * LayerNorm statistics of q/k/v and gate/up — SURVEY.md 8e; the reference has no tensor parallelism for BitLinearInf) from the
 * callback to a one-shot Lamport all-reduce over NVLink peer memory. peer_buffers[r] = rank r's symmetric buffer of
 * buffer_bytes bytes, addressable from this device (e.g. torch.distributed._symmetric_memory buffer_ptrs), filled with
 * -0.0f on every rank BEFORE any rank steps; buffer_bytes >= 3 * nranks * max_batch * hidden_size * 4. */
ONEBIT_API int onebit_decoder_enable_p2p_allreduce(onebit_decoder* dec, int rank, int nranks, void* const* peer_buffers,
                                                   size_t buffer_bytes);
/* Health of the persistent step (synchronous device read): 0 = fine, 1 = an in-kernel exchange timed out,
 * 2 = a step was asked to decode past max_seq_len (it wrote nothing outside the cache), 3 = a peer's data did not
 * arrive in the one-shot tensor-parallel all-reduce. */
ONEBIT_API int onebit_decoder_status(onebit_decoder* dec, int* code);
/* Device time stamps (ns, %globaltimer) recorded by two CTAs of the persistent step (the first and the last) during
 * the LAST step: layout [2 tracers][L + 2 rows][160 slots]. Row 0: [0] kernel start. Row 1 + l (layer l): [0] layer start,
 * [1] q/k/v done, [2] attention done, [3] o_proj + gate/up inputs done, [4] gate/up done, [5] down_proj done; sub-stage
 * stamps: [6,7] q/k/v inputs copied / MMA done, [8,9,10] o_proj inputs / MMA / statistics exchange, [11,12,13] gate/up,
 * [14,15,16] down_proj, [17,18,19] attention inputs / cache append / softmax, [20..23] weights of the stage in shared memory;
 * [32..95] time of every CTA barrier of the layer in program order, [96..159] the source line of that barrier. Row 1 + L: [0] lm_head start, [1] step end
 * (first tracer only). Returns the number of 64-bit words written (0 when the decoder is not persistent). Synchronous. */
ONEBIT_API int onebit_decoder_read_trace(onebit_decoder* dec, uint64_t* out, int n);
ONEBIT_API void onebit_decoder_destroy(onebit_decoder* dec);

#ifdef __cplusplus
}
#endif
#endif /* ONEBIT_B200_H */
