/*
 * onebit_oracle.c — CPU restatement of the OneBit 1-bit linear layer (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for the CUDA path in onebit_b200/csrc. It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call it. The product path never links or loads this library.
 *
 * Every function restates one piece of the reference (xuyuzhuang11/OneBit @ 42d6d7b):
 *   - bit layout ............ scripts/convert_llama_to_infer_ckpt.py:7-15  (fp16_to_int8)
 *   - unpack ................ transformers/src/transformers/models/bitnet.py:98-110 (int8_to_fp16)
 *   - forward ............... transformers/src/transformers/models/bitnet.py:112-122 (BitLinearInf.forward)
 *   - LayerNorm ............. bitnet.py:86,118  nn.LayerNorm(N, elementwise_affine=False), eps 1e-5, biased var
 *
 * Pinning: oracle/oracle.py checks this restatement against the tests/golden fixtures (.npz), which were produced by
 * executing the reference's own Python code in the build container (tests/golden/gen_golden.py).
 *
 * Arithmetic: fp32 tensors at the points where the reference materialises fp32 tensors, double
 * accumulation inside reductions (the reference's summation order inside MKL/ATen is unspecified,
 * so parity is tolerance-based: rel-L2 <= 1e-3, see tests/).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* scripts/convert_llama_to_infer_ckpt.py:10 — int_tensor = ((0 - t + 1) / 2).to(uint8):
 * +1 -> 0, -1 -> 1, 0 -> 0 (0.5 truncates). :11-13 — 8 consecutive columns per byte,
 * column 8j+i lands in bit i (multiplier [1,2,4,...,128]), stored as int8. */
void onebit_oracle_pack(const float* signs, int8_t* packed, int64_t n_rows, int64_t k_cols) {
    const int64_t kb = k_cols / 8;
    for (int64_t n = 0; n < n_rows; ++n) {
        for (int64_t j = 0; j < kb; ++j) {
            unsigned acc = 0;
            for (int i = 0; i < 8; ++i) {
                float v = (0.0f - signs[n * k_cols + 8 * j + i] + 1.0f) / 2.0f;
                unsigned bit = (unsigned)(uint8_t)v; /* truncation toward zero, as .to(torch.uint8) */
                acc += bit << i;
            }
            packed[n * kb + j] = (int8_t)(uint8_t)acc;
        }
    }
}

/* bitnet.py:98-110 — ((w[...,None] >> arange(8)) & 1) -> view(N, K) -> -2*b + 1.
 * The arithmetic shift of a negative int8 still yields the stored bit after "& 1". */
void onebit_oracle_unpack(const int8_t* packed, float* signs, int64_t n_rows, int64_t k_cols) {
    const int64_t kb = k_cols / 8;
    for (int64_t n = 0; n < n_rows; ++n)
        for (int64_t j = 0; j < kb; ++j) {
            int w = packed[n * kb + j];
            for (int i = 0; i < 8; ++i) signs[n * k_cols + 8 * j + i] = (float)(-2 * ((w >> i) & 1) + 1);
        }
}

/* bitnet.py:112-122 for x [m_tokens, k_cols] (fp32), packed weight [n_rows, k_cols/8]:
 *   :113  x' = x * h                       (fp32 tensor)
 *   :114-115  out = x' @ S^T               (fp32 tensor; here: double accumulation of +-x')
 *   :116  out *= g                         (fp32, in place)
 *   :118  y = LayerNorm_N(out), eps        (mean / biased variance over the N outputs of a token)
 *   :119-120  y += bias (if any)
 * `pre_ln` (optional, may be NULL) receives the :116 tensor, for tests of the un-normalised path. */
void onebit_oracle_forward(const float* x, const int8_t* packed, const float* g, const float* h,
                           const float* bias, float* y, float* pre_ln, int64_t m_tokens, int64_t k_cols,
                           int64_t n_rows, float eps) {
    const int64_t kb = k_cols / 8;
    float* xp = (float*)malloc(sizeof(float) * (size_t)k_cols);
    float* u = (float*)malloc(sizeof(float) * (size_t)n_rows);
    for (int64_t m = 0; m < m_tokens; ++m) {
        double total = 0.0;
        for (int64_t k = 0; k < k_cols; ++k) {
            xp[k] = x[m * k_cols + k] * h[k];
            total += (double)xp[k];
        }
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < n_rows; ++n) {
            /* sum_k s*x' = sum_k x' - 2 * sum_{bit=1} x'  (bit = 1 <=> sign = -1) */
            double neg = 0.0;
            const uint8_t* row = (const uint8_t*)packed + n * kb;
            for (int64_t j = 0; j < kb; ++j) {
                unsigned w = row[j];
                const float* xs = xp + 8 * j;
                double part = 0.0;
                for (int i = 0; i < 8; ++i)
                    if ((w >> i) & 1u) part += (double)xs[i];
                neg += part;
            }
            float out = (float)(total - 2.0 * neg);
            u[n] = out * g[n];
        }
        if (pre_ln) memcpy(pre_ln + m * n_rows, u, sizeof(float) * (size_t)n_rows);
        double mean = 0.0;
        for (int64_t n = 0; n < n_rows; ++n) mean += (double)u[n];
        mean /= (double)n_rows;
        double var = 0.0;
        for (int64_t n = 0; n < n_rows; ++n) {
            double d = (double)u[n] - mean;
            var += d * d;
        }
        var /= (double)n_rows;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        for (int64_t n = 0; n < n_rows; ++n) {
            float v = (float)(((double)u[n] - mean) * rstd);
            if (bias) v += bias[n];
            y[m * n_rows + n] = v;
        }
    }
    free(xp);
    free(u);
}

/* Faithful-cost variant for the CPU baseline: performs the reference's actual sequence of dense
 * steps (bitnet.py:98-118): materialise the full +-1 fp32 matrix on every call, then a dense
 * [m,k]x[k,n] product, scale, LayerNorm. Same results as onebit_oracle_forward up to fp32
 * summation order. `scratch` must hold n_rows*k_cols floats. */
void onebit_oracle_forward_dense(const float* x, const int8_t* packed, const float* g, const float* h,
                                 const float* bias, float* y, float* scratch, int64_t m_tokens,
                                 int64_t k_cols, int64_t n_rows, float eps) {
    const int64_t kb = k_cols / 8;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < n_rows; ++n)
        for (int64_t j = 0; j < kb; ++j) {
            int w = packed[n * kb + j];
            for (int i = 0; i < 8; ++i) scratch[n * k_cols + 8 * j + i] = (float)(-2 * ((w >> i) & 1) + 1);
        }
    float* xp = (float*)malloc(sizeof(float) * (size_t)k_cols);
    float* u = (float*)malloc(sizeof(float) * (size_t)n_rows);
    for (int64_t m = 0; m < m_tokens; ++m) {
        for (int64_t k = 0; k < k_cols; ++k) xp[k] = x[m * k_cols + k] * h[k];
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < n_rows; ++n) {
            const float* srow = scratch + n * k_cols;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int64_t k = 0; k < k_cols; k += 8)
                for (int i = 0; i < 8; ++i) acc[i] += srow[k + i] * xp[k + i];
            float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            u[n] = s * g[n];
        }
        double mean = 0.0, var = 0.0;
        for (int64_t n = 0; n < n_rows; ++n) mean += (double)u[n];
        mean /= (double)n_rows;
        for (int64_t n = 0; n < n_rows; ++n) {
            double d = (double)u[n] - mean;
            var += d * d;
        }
        var /= (double)n_rows;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        for (int64_t n = 0; n < n_rows; ++n) {
            float v = (float)(((double)u[n] - mean) * rstd);
            if (bias) v += bias[n];
            y[m * n_rows + n] = v;
        }
    }
    free(xp);
    free(u);
}

int onebit_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
