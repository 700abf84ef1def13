// Prompt ("prefill") attention for the decoder: causal softmax(Q K^T / sqrt(d)) V over a whole prompt at once
// (modeling_bitllama.py:536-563 with the causal mask of :1267-1269), for LLaMA heads of dimension 128.
//   qkv_prep_kernel   LayerNorm of the q/k/v BitLinear outputs (bitnet.py:118) from per-token statistics, rotate-half RoPE
//                     (:168-181), K and V appended to the fp16 cache, Q staged as fp16 [B][heads][T][128] with
//                     log2(e) / sqrt(128) folded in;
//   prefill_attn_kernel  flash-attention forward on warp-level tensor cores (mma.sync m16n8k16, fp16 in, fp32 accumulate):
//                     one CTA = 64 queries of one (sequence, head), 4 warps x 16 rows, key tiles of 64 through shared memory,
//                     online softmax in registers; writes fp16 [B*T][H] = the activation operand of the o_proj GEMM.
#include "common.cuh"
#include "prefill_attn.cuh"

namespace onebit {
namespace {

constexpr int kD = 128;          // head dimension
constexpr int kBQ = 64;          // queries per CTA
constexpr int kBK = 64;          // keys per tile
constexpr int kPitch = kD + 8;   // shared-memory row pitch in halves (272 B: conflict-free fragment loads, 16 B aligned)

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// grid (T, B), 256 threads: one CTA per token. LayerNorm sums of the token's q / k / v rows straight from the rows
// (coalesced float4 pass), then LayerNorm + RoPE + cache append + Q staging for every head.
__global__ void __launch_bounds__(256) qkv_prep_kernel(const PrefillAttnArgs A) {
    __shared__ double red[8][6];
    __shared__ float s_mean[3], s_rstd[3];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = b * A.T + t, pos = A.pos0 + t, W = A.n_heads * kD;  // W = row width (local heads)
    const float* rows[3] = {A.t_q + (size_t)m * A.ld, A.t_k + (size_t)m * A.ld, A.t_v + (size_t)m * A.ld};
    double st[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int p = 0; p < 3; ++p)
        for (int i = tid; i < W / 4; i += 256) {
            const float4 v = reinterpret_cast<const float4*>(rows[p])[i];
            st[2 * p] += (double)((v.x + v.y) + (v.z + v.w));
            st[2 * p + 1] += (double)((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w));
        }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st[k] += __shfl_xor_sync(0xffffffffu, st[k], o);
        if (lane == 0) red[warp][k] = st[k];
    }
    __syncthreads();
    if (tid < 3) {
        double s = 0.0, q = 0.0;
        for (int w = 0; w < 8; ++w) { s += red[w][2 * tid]; q += red[w][2 * tid + 1]; }
        const double mu = s / (double)A.n_ln, var = fmax(q / (double)A.n_ln - mu * mu, 0.0);
        s_mean[tid] = (float)mu;
        s_rstd[tid] = (float)(1.0 / sqrt(var + (double)A.ln_eps));
    }
    __syncthreads();
    const float mq = s_mean[0], rq = s_rstd[0], mk = s_mean[1], rk = s_rstd[1], mv = s_mean[2], rv = s_rstd[2];
    const int half = kD / 2;
    for (int e = tid; e < W; e += 256) {
        const int hd = e >> 7, d = e & (kD - 1);
        const int dp = d < half ? d + half : d - half, fi = d < half ? d : d - half;
        const float c = A.rope_cos[(size_t)pos * half + fi], s = A.rope_sin[(size_t)pos * half + fi];
        const int col = hd * kD;
        const float q0 = (rows[0][col + d] - mq) * rq, q1 = (rows[0][col + dp] - mq) * rq;
        const float k0 = (rows[1][col + d] - mk) * rk, k1 = (rows[1][col + dp] - mk) * rk;
        const float qr = d < half ? q0 * c - q1 * s : q0 * c + q1 * s;  // rotate_half: (-x2, x1)
        const float kr = d < half ? k0 * c - k1 * s : k0 * c + k1 * s;
        const float vv = (rows[2][col + d] - mv) * rv;
        const size_t crow = (((size_t)b * A.n_heads + hd) * A.max_seq + pos) * kD + d;
        A.kcache[crow] = __float2half_rn(kr);
        A.vcache[crow] = __float2half_rn(vv);
        // 1 / sqrt(128) (:546) and log2(e) (the softmax below uses exp2) folded into Q
        A.q16[(((size_t)b * A.n_heads + hd) * A.T + t) * kD + d] = __float2half_rn(qr * (0.08838834764831845f * 1.4426950408889634f));
    }
}

// grid (ceil(T / 64), heads, B), 128 threads
__global__ void __launch_bounds__(128) prefill_attn_kernel(const PrefillAttnArgs A) {
    __shared__ __align__(16) __half Ks[kBK * kPitch];
    __shared__ __align__(16) __half Vs[kBK * kPitch];
    const int qb = blockIdx.x, hd = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int q0 = qb * kBQ + warp * 16;                 // first query row of this warp (prompt-relative)
    const __half* Q = A.q16 + (((size_t)b * A.n_heads + hd) * A.T) * kD;
    const __half* Kc = A.kcache + (((size_t)b * A.n_heads + hd) * A.max_seq) * kD;
    const __half* Vc = A.vcache + (((size_t)b * A.n_heads + hd) * A.max_seq) * kD;
    // Q fragments of the warp's 16 rows (rows beyond T read row T - 1: their results are never stored)
    uint32_t qf[8][4];
    {
        const int r0 = min(q0 + g, A.T - 1), r1 = min(q0 + g + 8, A.T - 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            qf[ks][0] = *reinterpret_cast<const uint32_t*>(Q + (size_t)r0 * kD + ks * 16 + 2 * t4);
            qf[ks][1] = *reinterpret_cast<const uint32_t*>(Q + (size_t)r1 * kD + ks * 16 + 2 * t4);
            qf[ks][2] = *reinterpret_cast<const uint32_t*>(Q + (size_t)r0 * kD + ks * 16 + 8 + 2 * t4);
            qf[ks][3] = *reinterpret_cast<const uint32_t*>(Q + (size_t)r1 * kD + ks * 16 + 8 + 2 * t4);
        }
    }
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const int qpos0 = A.pos0 + q0 + g, qpos1 = qpos0 + 8;          // absolute positions of this thread's two rows
    const int last_key = min(A.pos0 + qb * kBQ + kBQ - 1, A.pos0 + A.T - 1);  // causal: keys 0 .. position of the last query
    const int ntiles = last_key / kBK + 1;
    for (int kt = 0; kt < ntiles; ++kt) {
        const int key0 = kt * kBK;
        __syncthreads();  // the previous tile has been consumed
        for (int i = tid; i < kBK * (kD / 8); i += 128) {  // 64 rows x 16 chunks of 16 B
            const int r = i >> 4, ch = i & 15;
            const int kp = min(key0 + r, A.max_seq - 1);
            *reinterpret_cast<uint4*>(Ks + r * kPitch + ch * 8) = *reinterpret_cast<const uint4*>(Kc + (size_t)kp * kD + ch * 8);
            *reinterpret_cast<uint4*>(Vs + r * kPitch + ch * 8) = *reinterpret_cast<const uint4*>(Vc + (size_t)kp * kD + ch * 8);
        }
        __syncthreads();
        if (key0 > A.pos0 + q0 + 15) continue;  // (warp-uniform) every key of the tile lies beyond this warp's rows
        // ---- S = Q K^T (16 x 64 per warp)
        float sc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[nt][j] = 0.f;
            const __half* krow = Ks + (nt * 8 + g) * kPitch + 2 * t4;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                mma16816(sc[nt], qf[ks], *reinterpret_cast<const uint32_t*>(krow + ks * 16), *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8));
        }
        // ---- causal mask + online softmax (rows g and g + 8 of the warp's tile; a row is spread over the 4 lanes of a quad)
        float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int kp = key0 + nt * 8 + 2 * t4;
            if (kp > qpos0) sc[nt][0] = -INFINITY;
            if (kp + 1 > qpos0) sc[nt][1] = -INFINITY;
            if (kp > qpos1) sc[nt][2] = -INFINITY;
            if (kp + 1 > qpos1) sc[nt][3] = -INFINITY;
            mx[0] = fmaxf(mx[0], fmaxf(sc[nt][0], sc[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(sc[nt][2], sc[nt][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2], psum[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            corr[r] = mx[r] == -INFINITY ? 1.f : exp2f(mrow[r] - mx[r]);  // (first tile: exp2(-inf) = 0)
            mrow[r] = mx[r];
        }
        uint32_t pf[4][4];  // P as the A operand of the second MMA: k-step j = keys 16j .. 16j + 15 of the tile
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = mx[0] == -INFINITY ? 0.f : exp2f(sc[nt][0] - mx[0]), p1 = mx[0] == -INFINITY ? 0.f : exp2f(sc[nt][1] - mx[0]);
            const float p2 = mx[1] == -INFINITY ? 0.f : exp2f(sc[nt][2] - mx[1]), p3 = mx[1] == -INFINITY ? 0.f : exp2f(sc[nt][3] - mx[1]);
            psum[0] += p0 + p1;
            psum[1] += p2 + p3;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_h2(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * corr[r] + psum[r];  // (quad-partial: reduced at the end)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        // ---- O += P V (16 x 128 per warp): V fragments through ldmatrix.trans from the row-major tile
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half* vrow = Vs + (j * 16 + (lane & 15)) * kPitch;  // lanes 0..15 address the 16 key rows of this k-step
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, vrow + nt * 8);
                mma16816(o[nt], pf[j], b0, b1);
            }
        }
    }
    // ---- normalise and store fp16 [B*T][H]
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
        lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    }
    const float inv0 = 1.f / lrow[0], inv1 = 1.f / lrow[1];
    const int t0 = q0 + g, t1 = q0 + g + 8;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
        const int col = hd * kD + nt * 8 + 2 * t4;
        if (t0 < A.T) *reinterpret_cast<uint32_t*>(A.out16 + ((size_t)b * A.T + t0) * A.out_ld + col) = pack_h2(o[nt][0] * inv0, o[nt][1] * inv0);
        if (t1 < A.T) *reinterpret_cast<uint32_t*>(A.out16 + ((size_t)b * A.T + t1) * A.out_ld + col) = pack_h2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

}  // namespace

int launch_prefill_attention(const PrefillAttnArgs& A, cudaStream_t s) {
    ONEBIT_REQUIRE(A.T >= 1 && A.B >= 1 && A.pos0 >= 0 && A.pos0 + A.T <= A.max_seq, "prefill attention: the prompt does not fit the KV cache");
    qkv_prep_kernel<<<dim3(A.T, A.B), 256, 0, s>>>(A);
    ONEBIT_CUDA_TRY(cudaGetLastError());
    prefill_attn_kernel<<<dim3((A.T + kBQ - 1) / kBQ, A.n_heads, A.B), 128, 0, s>>>(A);
    ONEBIT_CUDA_TRY(cudaGetLastError());
    return ONEBIT_OK;
}

}  // namespace onebit
