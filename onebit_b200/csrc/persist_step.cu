// Persistent decode-step kernel (see persist_step.cuh for the design) + its host side.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "imma_gemv.cuh"
#include "persist_step.cuh"

namespace onebit {
namespace persist {
namespace {

using imma::imma16832;
using imma::mbar_expect_tx;
using imma::mbar_init;
using imma::mbar_wait;
using imma::plane;
using imma::smem_u32;

// ---------------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }

__device__ __forceinline__ uint4 ldv4(const uint32_t* p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldv1(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stv1(uint32_t* p, uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void stv16(void* p, uint32_t v) { asm volatile("st.volatile.global.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

// validity of Lamport words: digit words are invalid while ANY byte still is 0x80; float words while they equal kSentF
__device__ __forceinline__ uint32_t hz(uint32_t x) {
    const uint32_t y = x ^ kSentD;
    return (y - 0x01010101u) & ~y & 0x80808080u;
}
__device__ __forceinline__ bool bad_d(const uint4& v) { return (hz(v.x) | hz(v.y) | hz(v.z) | hz(v.w)) != 0u; }
__device__ __forceinline__ bool bad_f(const uint4& v) { return v.x == kSentF || v.y == kSentF || v.z == kSentF || v.w == kSentF; }
__device__ __forceinline__ uint32_t fbits(float f) {  // publishable bit pattern of a float
    const uint32_t b = __float_as_uint(f);
    return b == kSentF ? 0x7FFFFFFFu : b;
}

// give-up logic of every polling loop: a protocol bug must end the kernel, not hang the GPU. The bound is a
// wall-clock deadline for the whole step (set by every CTA at kernel start), checked every 1024 failed polls.
__shared__ unsigned long long s_deadline;
__shared__ unsigned long long* s_abort_info;  // trace row 0, slots 8..: who gave up first (debug aid)
__shared__ int s_site;                        // layer * 16 + stage of the compute warps (set at stage boundaries)
constexpr int kLmRows = 2;  // lm_head rows per ring chunk
#ifndef ONEBIT_POLL_BACKOFF_NS
#define ONEBIT_POLL_BACKOFF_NS 100
#endif
constexpr unsigned kPollBackoffNs = ONEBIT_POLL_BACKOFF_NS;  // pause between polls of a not-yet-valid exchange word (keeps the hot L2 lines free for the writers)
constexpr unsigned long long kStepBudgetNs = 400ull * 1000ull * 1000ull;
__device__ __noinline__ bool spin_giveup_slow(int spins, int* abort_flag) {
    if (ldv1(reinterpret_cast<const uint32_t*>(abort_flag)) != 0u) return true;
    if (spins >= 1024 && gtime() > s_deadline) {
        if (atomicExch(abort_flag, 1) == 0 && s_abort_info != nullptr) {
            s_abort_info[8] = blockIdx.x;
            s_abort_info[9] = threadIdx.x;
            s_abort_info[10] = (unsigned long long)s_site;
        }
        return true;
    }
    return false;
}
__device__ __forceinline__ bool spin_giveup(int& spins, int* abort_flag) {
    ++spins;
    if (spins != 8 && (spins & 1023) != 0) return false;  // the check at 8 makes an aborted step drain quickly
    return spin_giveup_slow(spins, abort_flag);
}
// bounded mbarrier wait (a lost TMA completion must end the kernel, not hang the GPU); false = gave up
__device__ __noinline__ bool mbar_wait_b(uint64_t* bar, uint32_t parity, int* abort_flag) {
    int spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
        if (spin_giveup(spins, abort_flag)) return false;
    }
}

__device__ __forceinline__ size_t dsize(int pdt) { return pdt == ONEBIT_F32 ? 4 : 2; }
__device__ __noinline__ float ldp(const void* p, int i, int pdt) {  // parameter vector element (runtime dtype)
    if (pdt == ONEBIT_F16) return __half2float(static_cast<const __half*>(p)[i]);
    if (pdt == ONEBIT_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
    return static_cast<const float*>(p)[i];
}

// ---- per-layer parameter block (shared memory, filled by the TMA producer; offsets in ELEMENTS of the param dtype) ----
constexpr int kPBScalarBytes = 64;  // 2 x 8 words: quantiser exponents of layer l and of layer l + 1 (see lscal)
constexpr int PB_gA = 0;      // weight_scale of this CTA's q/k/v tiles, 16 per tile (<= 12 tiles)
constexpr int PB_gO = 192;    // o_proj.weight_scale      [colC0, colC0 + nownC)
constexpr int PB_gDn = 288;   // down_proj.weight_scale   [colC0, ...)
constexpr int PB_lnN = 384;   // next layer's input_layernorm.weight (or the final norm) [colC0, ...)
constexpr int PB_hq = 480, PB_hk = 576, PB_hv = 672;  // next layer's q/k/v input_factor [colC0, ...)
constexpr int PB_lnP = 768;   // post_attention_layernorm.weight [colC0, ...)
constexpr int PB_hg = 864, PB_hu = 960;               // gate/up input_factor [colC0, ...)
constexpr int PB_gG = 1056, PB_gU = 1152;             // gate/up weight_scale [colD0, colD0 + nownD)
constexpr int PB_hD = 1248;   // down_proj.input_factor [colD0, ...)
constexpr int PB_hO = 1344;   // o_proj.input_factor of this CTA's attention head (128)
constexpr int kPBEls = 1472;
constexpr int kPBBytes = kPBScalarBytes + kPBEls * 4;
__device__ __forceinline__ float pbf(const unsigned char* pb, int off_el, int i, int pdt) {
    const unsigned char* p = pb + kPBScalarBytes;
    if (pdt == ONEBIT_F16) return __half2float(reinterpret_cast<const __half*>(p)[off_el + i]);
    if (pdt == ONEBIT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[off_el + i]);
    return reinterpret_cast<const float*>(p)[off_el + i];
}
// scalars of the block: word w of layer l (which = 0) or l + 1 (which = 1): 0..2 e_q/e_k/e_v, 3 e_o, 4..5 e_gate/e_up,
// 6 max|down.input_factor| (float bits)
__device__ __forceinline__ int pbi(const unsigned char* pb, int which, int w) { return reinterpret_cast<const int*>(pb)[which * 8 + w]; }

// balanced base-255 digits of v (|v| <= 2^29): four bytes in [-127, 127], sum_d digit_d * 255^d == v
__device__ __forceinline__ uint32_t digits255(int v) {
    const uint32_t u = (uint32_t)v + 2114125312u;  // + 127 * (1 + 255 + 255^2 + 255^3)
    const uint32_t q1 = __umulhi(u, 0x80808081u) >> 7, e0 = u - q1 * 255u;
    const uint32_t q2 = __umulhi(q1, 0x80808081u) >> 7, e1 = q1 - q2 * 255u;
    const uint32_t q3 = __umulhi(q2, 0x80808081u) >> 7, e2 = q2 - q3 * 255u;
    return ((e0 - 127u) & 0xFFu) | (((e1 - 127u) & 0xFFu) << 8) | (((e2 - 127u) & 0xFFu) << 16) | (((q3 - 127u) & 0xFFu) << 24);
}
// x' (already * input_factor), |x'| < 2^e -> packed digits of the plane-scaled 23-bit integer
// (imma_gemv.cuh: v = q << (7-j), -q for j = 7)
__device__ __forceinline__ uint32_t quant_digits(float xp, int e, int col) {
    int q = __float2int_rn(xp * ldexpf(1.0f, 22 - e));
    q = max(-(1 << 22), min(1 << 22, q));
    const int j = col & 7;
    const int v = (j == 7) ? -q : (q << (7 - j));
    return digits255(v);
}
// mean / 1/sqrt(var + eps) of a LayerNorm (bitnet.py:118) from fp64 (sum, sum of squares) over n = 1/inv_n values
__device__ __forceinline__ void ln_finish(double s, double q, double inv_n, float eps, float* mean, float* rstd) {
    const double mu = s * inv_n;
    const double var = fmax(q * inv_n - mu * mu, 0.0);  // biased variance
    *mean = (float)mu;
    *rstd = __frcp_rn(__fsqrt_rn((float)var + eps));
}
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// word address (inside one [K]-word digit vector) of B-fragment word (plane j, digit d) of the 32-column block at col0
__device__ __forceinline__ int frag_word(int col0, int j, int d) {
    const int u = col0 >> 8, W = (col0 & 255) >> 5;
    return u * 256 + (j >> 1) * 64 + d * 16 + (W >> 1) * 4 + (j & 1) * 2 + (W & 1);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// double as two publishable floats (hi, lo), `stride` words apart. Statistics records are contiguous per CTA
// ([token][cta][word]): one CTA's record is one or two 32-byte sectors, written by one thread (a transposed layout was
// measured: 12 scattered 4-byte stores per CTA into lines shared by 32 writers made every exchange ~4 us slower).
__device__ __forceinline__ void put_hilo(uint32_t* p, uint32_t* pc, int stride, double v) {
    const float hi = (float)v, lo = (float)(v - (double)hi);
    stv1(p, fbits(hi));
    stv1(p + stride, fbits(lo));
    stv1(pc, kSentF);
    stv1(pc + stride, kSentF);
}

// ---- the ring of weight tiles: fixed slots -----------------------------------------------------------------------
// A slot is 2 * H bytes: one re-tiled 16-row tile of a K = H matrix, which is also one fp16 lm_head row. A down_proj
// tile (K = I) takes nsD = ceil(I / H) consecutive slots, an lm_head chunk kLmRows. A chunk never straddles the ring
// end (the slots left over are skipped, by both sides). Producer and consumers walk the same chunk sequence, each
// with a slot counter and one parity bit per slot: the consumers' bits belong to the `full` barriers (flipped when a
// slot is waited on as the first slot of a chunk), the producer's to the `empty` barriers (flipped for every slot it
// takes; the consumers arrive on every slot of a chunk they release). Skipped slots see no barrier traffic at all,
// so every barrier advances exactly one phase per use and neither side can get two phases ahead of the other.
constexpr int kMaxSlots = 32;
struct Ring {
    uint32_t slot, NS, phase;
};
__device__ __forceinline__ uint32_t ring_take(Ring& R, uint32_t n, uint32_t& skip) {
    skip = 0;
    if (R.slot + n > R.NS) {
        skip = R.NS - R.slot;
        R.slot = 0;
    }
    const uint32_t s = R.slot;
    R.slot += n;
    if (R.slot >= R.NS) R.slot = 0;
    return s;
}
__device__ __forceinline__ void block_range(int nblocks, int cta, int ncta, int& b0, int& b1) {
    b0 = (int)(((long long)nblocks * cta) / ncta);
    b1 = (int)(((long long)nblocks * (cta + 1)) / ncta);
}
// Weight tiles live in HBM re-tiled at load time (retile_kernel, 1 bit per element as in the checkpoint): a 16-row tile
// is one contiguous block of 16 * K/8 bytes = [K/256 units][32 lanes][16 B], the 16 B of lane (g, t4) being the
// 8 bytes of row g and the 8 bytes of row g + 8 that form its four A-fragment words of that unit. One bulk copy per tile,
// one conflict-free LDS.128 per lane and unit. `pitch` below is tile_bytes / 16 = K / 8.
__device__ __host__ __forceinline__ int row_pitch(int kb) { return kb; }

// layer-invariant description of one BitLinear stage of this CTA (built once per step, shared memory)
enum { ST_A = 0, ST_C = 1, ST_D1 = 2, ST_D2 = 3 };
struct StageTab {
    int K, T;          // input width, 16-row tiles of this CTA
    int lgKG, tpg;     // 16 warps = TG tile groups x KG = 2^lgKG K groups; tiles per group (<= 3)
    int nsl;           // ring slots per tile
    int p_lo, nsets;   // stage A: first projection (0 = q, 1 = k, 2 = v) this CTA has rows of, and how many it spans
    int aux[kMaxTiles];              // stage A: word offset of the tile's first row inside one token's [3][H] q/k/v record
    unsigned char pslot[kMaxTiles];  // which digit set / scale the tile's rows use
    short goff[kMaxTiles];           // element offset (parameter block) of the weight_scale of the tile's first row
};
// split of the 16 compute warps of a T-tile, `units`-unit stage: fewest (tiles per warp) x (units per warp)
__device__ __forceinline__ void plan_split(int T, int units, int& lgKG, int& tpg) {
    int best = 1 << 30;
    lgKG = 4;
    tpg = max(T, 1);
    for (int lt = 0; lt <= 2; ++lt) {  // TG = 1, 2, 4
        const int TG = 1 << lt, KG = kCW >> lt, tp = (max(T, 1) + TG - 1) / TG;
        if (tp > 3) continue;
        const int cost = tp * ((units + KG - 1) / KG);
        if (cost < best) { best = cost; lgKG = 4 - lt; tpg = tp; }
    }
}

// ---- shared-memory plan ---------------------------------------------------------------------------------------
struct Smem {
    unsigned char* ring;
    uint32_t* dbuf;
    unsigned char* pbuf;          // [2][kPBBytes] parameter blocks (layer parity)
    int* red;
    float* u;                     // [kMaxTok][192]
    uint32_t* stat;               // [M * ncta * kStatW] (>= 3 * ncta * 4)
    uint32_t* stage;              // [kMaxTok][96][3]
    double* redd;                 // [32]
    unsigned long long* q128;     // [2][kMaxTok]
    double* invs;                 // [2][kMaxTok]
    StageTab* tab;                // [4]
    uint32_t* tslot;              // [kMaxTiles] ring slot of tile i of the current stage
    float* fscr;                  // [64]
};

struct Ctx {
    int tid, lane, warp, cta, ncta, M, H, I, pdt, ring_bytes, max_batch;
    int nownC, colC0, om, oc;  // ownership of residual-stream columns: thread (om, oc) owns column colC0 + oc of token om
    bool ownerC;
    Smem S;
    uint64_t* full;
    uint64_t* empty;
    int* abort_flag;
    uint32_t* X;   // this step's exchange arena
    uint32_t* Xc;  // the other parity set: re-armed (sentinels) word for word as we publish
    Ring R;
    unsigned long long* trl;  // trace row of the current layer (tracer CTAs only, else nullptr)
    const unsigned char* pb;  // parameter block of the current layer
    int tseq;                 // next slot of the barrier-level trace (tracer thread only)
};
__device__ __forceinline__ void stamp(const Ctx& c, int slot) {
    if (c.trl != nullptr && c.tid == 0) c.trl[slot] = gtime();
}
// barrier-level trace: slots [32, 96) hold times, [96, 160) the source line of the barrier
__device__ __forceinline__ void csync(Ctx& c, int line) {
    cta_sync();
    if (c.trl != nullptr && c.tid == 0 && c.tseq < 64) {
        c.trl[32 + c.tseq] = gtime();
        c.trl[96 + c.tseq] = (unsigned long long)line;
        ++c.tseq;
    }
}

// ---- Lamport copies global -> shared ------------------------------------------------------------------------
// digit vector of K words: also commits this vector's Q128 = 128 * sum_k q_k to *q128 (integer, order independent)
__device__ __noinline__ void poll_copy_d(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int nchunks, int tid, int lane,
                                         unsigned long long* q128, int* abort_flag) {
    constexpr int U = 6;
    int se = 0, so = 0;
    for (int base = 0; base < nchunks; base += kCT * U) {
        uint4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = base + k * kCT + tid;
            v[k] = make_uint4(0u, 0u, 0u, 0u);
            if (i < nchunks) v[k] = ldv4(src + 4 * (size_t)i);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = base + k * kCT + tid;
            if (i < nchunks) {
                int spins = 0;
                while (bad_d(v[k])) {
                    if (spin_giveup(spins, abort_flag)) break;
                    __nanosleep(kPollBackoffNs);
                    v[k] = ldv4(src + 4 * (size_t)i);
                }
                *reinterpret_cast<uint4*>(dst + 4 * (size_t)i) = v[k];
                se = __dp4a((int)v[k].x, 0x01010101, se);
                se = __dp4a((int)v[k].y, 0x01010101, se);
                so = __dp4a((int)v[k].z, 0x01010101, so);
                so = __dp4a((int)v[k].w, 0x01010101, so);
            }
        }
    }
    // thread tid always sees plane pair (tid>>4)&3 and digit (tid>>2)&3 (chunk index = tid mod 64 pattern)
    const int d = (tid >> 2) & 3, jp = (tid >> 4) & 3;
    const long long p255 = d == 0 ? 1ll : (d == 1 ? 255ll : (d == 2 ? 65025ll : 16581375ll));
    const long long ce = 1ll << (2 * jp), co = (jp == 3) ? -128ll : (1ll << (2 * jp + 1));
    long long c = p255 * ((long long)se * ce + (long long)so * co);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) atomicAdd(q128, (unsigned long long)c);
}
// float records: nchunks uint4 (nchunks may exceed kCT)
__device__ __noinline__ void poll_copy_f(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int nchunks, int tid,
                                         int* abort_flag) {
    for (int i = tid; i < nchunks; i += kCT) {
        uint4 v = ldv4(src + 4 * (size_t)i);
        int spins = 0;
        while (bad_f(v)) {
            if (spin_giveup(spins, abort_flag)) break;
            __nanosleep(kPollBackoffNs);
            v = ldv4(src + 4 * (size_t)i);
        }
        *reinterpret_cast<uint4*>(dst + 4 * (size_t)i) = v;
    }
}

// ---- IMMA phase: this warp's NT tiles x its K group of the stage, partial sums to `red` [tile][kg][16 rows][8] ------
template <int NT>
__device__ __forceinline__ void imma_phase(const unsigned char* ring, uint32_t slot_bytes, const uint32_t* tslot, const StageTab& tb, int t0,
                                           int kg, int KG, const uint32_t* dbuf, int set_words, int M, int* red, int lane) {
    const int K = tb.K, units = K >> 8;
    const int g = lane >> 2, t4 = lane & 3, bm = g >> 2, bd = g & 3;
    int acc[NT][2][4];
    const unsigned char* wp[NT];
    const uint32_t* bp[NT];
#pragma unroll
    for (int ti = 0; ti < NT; ++ti) {
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[ti][b][c] = 0;
        wp[ti] = ring + (size_t)tslot[t0 + ti] * slot_bytes + lane * 16;
        bp[ti] = dbuf + (size_t)tb.pslot[t0 + ti] * set_words + (size_t)bm * K + bd * 16 + t4 * 4;
    }
    const bool have = bm < M;
    for (int u = kg; u < units; u += KG) {
        uint4 w[NT];
#pragma unroll
        for (int ti = 0; ti < NT; ++ti) w[ti] = *reinterpret_cast<const uint4*>(wp[ti] + u * 512);
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            uint4 bv[NT];
            bv[0] = have ? *reinterpret_cast<const uint4*>(bp[0] + u * 256 + jp * 64) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int ti = 1; ti < NT; ++ti) {  // same digit set as tile 0 (the common case): no second load
                bv[ti] = bv[0];
                if (bp[ti] != bp[0] && have) bv[ti] = *reinterpret_cast<const uint4*>(bp[ti] + u * 256 + jp * 64);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const uint32_t mask = 0x01010101u << (2 * jp + jj);
#pragma unroll
                for (int ti = 0; ti < NT; ++ti)
                    imma16832(acc[ti][jj], plane(w[ti].x, mask), plane(w[ti].z, mask), plane(w[ti].y, mask), plane(w[ti].w, mask),
                              jj ? bv[ti].z : bv[ti].x, jj ? bv[ti].w : bv[ti].y);
            }
        }
    }
#pragma unroll
    for (int ti = 0; ti < NT; ++ti) {
        int* base = red + ((t0 + ti) * KG + kg) * 128 + 2 * t4;
        *reinterpret_cast<int2*>(base + g * 8) = make_int2(acc[ti][0][0] + acc[ti][1][0], acc[ti][0][1] + acc[ti][1][1]);
        *reinterpret_cast<int2*>(base + (g + 8) * 8) = make_int2(acc[ti][0][2] + acc[ti][1][2], acc[ti][0][3] + acc[ti][1][3]);
    }
}
// row result of the stage: t = sum_k s(n,k) x'_k for (row r of tile `tile`, token em), from the K-group partial sums
__device__ __forceinline__ float row_value(const int* red, int tile, int KG, int r, int em, long long q128, double invs) {
    int4 a = make_int4(0, 0, 0, 0);
    const int* p = red + tile * KG * 128 + r * 8 + 4 * em;
    for (int kg = 0; kg < KG; ++kg) {
        const int4 v = *reinterpret_cast<const int4*>(p + kg * 128);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    // V = sum_d 255^d a_d = 128 * sum_{bit=1} q, |V| < 2^45: exact in double
    const double V = fma((double)a.w, 16581375.0, fma((double)a.z, 65025.0, fma((double)a.y, 255.0, (double)a.x)));
    return (float)(((double)q128 - 2.0 * V) * invs);  // invs = 2^(e-22) / 128
}

// generic BitLinear stage core: copy the input digit vectors (Lamport poll), wait for the stage's weight tiles (the next
// T chunks of the ring), run the IMMA phase, release the tiles.
__device__ __forceinline__ void stage_core(Ctx& c, int st, int nsets, const uint32_t* x0, const uint32_t* x1, int ts) {
    const StageTab& tb = c.S.tab[st];
    const int K = tb.K, T = tb.T, nsl = tb.nsl;
    const int set_words = c.M * K;
    csync(c, __LINE__);  // S.invs / S.q128 of this stage are in place, the previous stage's readers are done
    if (T > 0) {
        for (int ps = 0; ps < nsets; ++ps)
            for (int m = 0; m < c.M; ++m)
                poll_copy_d((ps ? x1 : x0) + (size_t)m * K, c.S.dbuf + (size_t)ps * set_words + (size_t)m * K, K / 4, c.tid, c.lane,
                            &c.S.q128[ps * kMaxTok + m], c.abort_flag);
    }
    stamp(c, ts);
    uint32_t myslot = 0;
    if (nsl == 1) {  // T consecutive single-slot chunks: closed form
        uint32_t sl = c.R.slot + (uint32_t)c.tid;
        if (sl >= c.R.NS) sl -= c.R.NS;
        if (c.tid < T) {
            myslot = sl;
            c.S.tslot[c.tid] = sl;
            mbar_wait_b(&c.full[sl], (c.R.phase >> sl) & 1u, c.abort_flag);
        }
        const unsigned long long bits = ((1ull << T) - 1ull) << c.R.slot;
        c.R.phase ^= (uint32_t)((bits | (bits >> c.R.NS)) & ((1ull << c.R.NS) - 1ull));
        c.R.slot += (uint32_t)T;
        if (c.R.slot >= c.R.NS) c.R.slot -= c.R.NS;
    } else {
        for (int i = 0; i < T; ++i) {
            uint32_t skip;
            const uint32_t sl = ring_take(c.R, (uint32_t)nsl, skip);  // skipped slots: no barrier traffic on either side
            if (c.tid == i) {
                myslot = sl;
                c.S.tslot[i] = sl;
                mbar_wait_b(&c.full[sl], (c.R.phase >> sl) & 1u, c.abort_flag);
            }
            c.R.phase ^= 1u << sl;
        }
    }
    csync(c, __LINE__);
    stamp(c, ts == 6 ? 20 : (ts == 8 ? 21 : (ts == 11 ? 22 : 23)));  // the stage's weights are in shared memory
    {
        const int KG = 1 << tb.lgKG, tg = c.warp >> tb.lgKG, kg = c.warp & (KG - 1);
        const int t0 = tg * tb.tpg, nt = min(tb.tpg, T - t0);
        const uint32_t slot_bytes = 2u * (uint32_t)c.H;
        if (nt == 1) imma_phase<1>(c.S.ring, slot_bytes, c.S.tslot, tb, t0, kg, KG, c.S.dbuf, set_words, c.M, c.S.red, c.lane);
        else if (nt == 2) imma_phase<2>(c.S.ring, slot_bytes, c.S.tslot, tb, t0, kg, KG, c.S.dbuf, set_words, c.M, c.S.red, c.lane);
        else if (nt == 3) imma_phase<3>(c.S.ring, slot_bytes, c.S.tslot, tb, t0, kg, KG, c.S.dbuf, set_words, c.M, c.S.red, c.lane);
    }
    csync(c, __LINE__);
    if (c.tid < T)
        for (int k = 0; k < nsl; ++k) mbar_arrive(&c.empty[myslot + (uint32_t)k]);
    stamp(c, ts + 1);
}

// publish the digits of `nprob` BitLinear inputs for the owned 32-column blocks: S.stage[(m*nprob+p)*nown + col] holds
// the four packed digits of column col; every (plane, digit) word of the B-fragment layout is assembled from 4 columns
__device__ __forceinline__ void publish32(Ctx& c, size_t xoff, size_t prob_stride, size_t tok_stride, int nprob, int nown, int col0) {
    const int nblk = nown >> 5;
    if (nblk <= 0) return;
    const int per_tok = nprob * nblk, items = kMaxTok * per_tok * 32;
    for (int it = c.tid; it < items; it += kCT) {
        const int jd = it & 31, j = jd >> 2, d = jd & 3;
        const int r = it >> 5, m = r >= per_tok ? 1 : 0, r2 = r - m * per_tok;
        const int p = r2 >= 2 * nblk ? 2 : (r2 >= nblk ? 1 : 0), blk = r2 - p * nblk;
        const size_t a = xoff + (size_t)p * prob_stride + (size_t)m * tok_stride + frag_word(col0 + blk * 32, j, d);
        if (m < c.M) {
            const uint32_t* sp = c.S.stage + (size_t)(m * nprob + p) * nown + blk * 32 + j;
            const uint32_t sel = (uint32_t)d | ((uint32_t)(4 + d) << 4);
            const uint32_t lo = __byte_perm(sp[0], sp[8], sel), hi = __byte_perm(sp[16], sp[24], sel);
            stv1(c.X + a, __byte_perm(lo, hi, 0x5410));
        }
        // re-arm the other parity set for every token slot (a later step may run a larger batch)
        if (m < c.max_batch) stv1(c.Xc + a, kSentD);
    }
}
// same for 16-column blocks (stage D1): each fragment word gets a 16-bit half from this CTA
__device__ __forceinline__ void publish16(Ctx& c, size_t xoff, size_t tok_stride, int npb, int col0) {
    const int nown = 16 * npb, items = kMaxTok * npb * 32;
    for (int it = c.tid; it < items; it += kCT) {
        const int jd = it & 31, j = jd >> 2, d = jd & 3;
        const int r = it >> 5, m = r >= npb ? 1 : 0, blk = r - m * npb;
        const int c0 = col0 + blk * 16;
        const size_t a = xoff + (size_t)m * tok_stride + frag_word(c0 & ~31, j, d);
        const int bo = (c0 & 16) ? 2 : 0;
        if (m < c.M) {
            const uint32_t* sp = c.S.stage + m * nown + blk * 16 + j;
            stv16(reinterpret_cast<char*>(c.X + a) + bo, __byte_perm(sp[0], sp[8], (uint32_t)d | ((uint32_t)(4 + d) << 4)) & 0xFFFFu);
        }
        if (m < c.max_batch) stv16(reinterpret_cast<char*>(c.Xc + a) + bo, 0x8080u);
    }
}

// exchange of per-CTA statistics records [M][ncta][kStatW]: poll every CTA's record, then reduce in a fixed order:
// quantity q < nq is a sum of (hi, lo) pairs, the nmax quantities after are maxima. Result: S.redd[m * (nq + nmax) + q].
__device__ __forceinline__ void stats_exchange(Ctx& c, size_t soff, int nq, int nmax) {
    for (int m = 0; m < c.M; ++m)
        poll_copy_f(c.X + soff + (size_t)m * c.ncta * kStatW, c.S.stat + (size_t)m * c.ncta * kStatW, c.ncta * kStatW / 4, c.tid, c.abort_flag);
    csync(c, __LINE__);
    const int per = nq + nmax;
    if (c.warp < per * c.M) {
        const int m = c.warp >= per ? 1 : 0, qn = c.warp - m * per;
        const float* st = reinterpret_cast<const float*>(c.S.stat) + (size_t)m * c.ncta * kStatW;
        if (qn < nq) {
            double s = 0.0;
            for (int k = c.lane; k < c.ncta; k += 32) {
                const float2 hl = *reinterpret_cast<const float2*>(st + k * kStatW + 2 * qn);
                s += (double)hl.x + (double)hl.y;
            }
            s = warp_sum_d(s);
            if (c.lane == 0) c.S.redd[c.warp] = s;
        } else {
            float mx = 0.f;
            for (int k = c.lane; k < c.ncta; k += 32) mx = fmaxf(mx, st[k * kStatW + 2 * nq + (qn - nq)]);
            mx = warp_max_f(mx);
            if (c.lane == 0) c.S.redd[c.warp] = (double)mx;
        }
    }
    csync(c, __LINE__);
}

// x_hat = resid * rr[token] * ln_w -> x' for q, k, v of layer l -> digits -> exchange (inputs of stage A)   (:67-81)
// FROM_BLOCK: parameters of layer l come from the parameter block of layer l - 1 ("next layer" slices); else (layer 0,
// once per step) straight from global memory.
template <bool FROM_BLOCK>
__device__ __noinline__ void publish_qkv_inputs(Ctx& cref, const Params& P, int l, float resid, const float* rr) {
    Ctx c = cref;  // scalar-replaced local copy: the fields live in registers, not behind a pointer
    if (c.ownerC) {
        float lw, hq, hk, hv;
        int eq, ek, ev;
        if (FROM_BLOCK) {
            lw = pbf(c.pb, PB_lnN, c.oc, c.pdt); hq = pbf(c.pb, PB_hq, c.oc, c.pdt);
            hk = pbf(c.pb, PB_hk, c.oc, c.pdt); hv = pbf(c.pb, PB_hv, c.oc, c.pdt);
            eq = pbi(c.pb, 1, 0); ek = pbi(c.pb, 1, 1); ev = pbi(c.pb, 1, 2);
        } else {
            const LayerDev& Ly = P.layers[l];
            const int col = c.colC0 + c.oc;
            lw = ldp(Ly.ln_in, col, c.pdt); hq = ldp(Ly.q.h, col, c.pdt); hk = ldp(Ly.k.h, col, c.pdt); hv = ldp(Ly.v.h, col, c.pdt);
            eq = Ly.e_qkv[0]; ek = Ly.e_qkv[1]; ev = Ly.e_qkv[2];
        }
        const int col = c.colC0 + c.oc;
        const float xh = resid * rr[c.om] * lw;
        c.S.stage[(size_t)(c.om * 3 + 0) * c.nownC + c.oc] = quant_digits(xh * hq, eq, col);
        c.S.stage[(size_t)(c.om * 3 + 1) * c.nownC + c.oc] = quant_digits(xh * hk, ek, col);
        c.S.stage[(size_t)(c.om * 3 + 2) * c.nownC + c.oc] = quant_digits(xh * hv, ev, col);
    }
    csync(c, __LINE__);
    publish32(c, (size_t)l * P.per_layer + P.o_xA, (size_t)kMaxTok * c.H, (size_t)c.H, 3, c.nownC, c.colC0);
    cref.R = c.R;
    cref.tseq = c.tseq;
}

// x_hat = RMSNorm(post_attention_layernorm) of the updated stream -> x' for gate / up -> digits -> exchange (stage D1 inputs)
__device__ __noinline__ void publish_gu_inputs(Ctx& cref, const Params& P, int l, float resid, const float* rr) {
    Ctx c = cref;
    if (c.ownerC) {
        const int col = c.colC0 + c.oc;
        const float xh = resid * rr[c.om] * pbf(c.pb, PB_lnP, c.oc, c.pdt);
        c.S.stage[(size_t)(c.om * 2 + 0) * c.nownC + c.oc] = quant_digits(xh * pbf(c.pb, PB_hg, c.oc, c.pdt), pbi(c.pb, 0, 4), col);
        c.S.stage[(size_t)(c.om * 2 + 1) * c.nownC + c.oc] = quant_digits(xh * pbf(c.pb, PB_hu, c.oc, c.pdt), pbi(c.pb, 0, 5), col);
    }
    csync(c, __LINE__);
    publish32(c, (size_t)l * P.per_layer + P.o_xD1, (size_t)kMaxTok * c.H, (size_t)c.H, 2, c.nownC, c.colC0);
    cref.tseq = c.tseq;
}

// stages C / D2 share their shape: rows of o_proj / down_proj owned as 32-row blocks, then
// x <- x + LayerNorm(g*t) (:912 / :918) and the RMSNorm factor of the next BitLinear group (:67-81), all from ONE
// exchange of five per-CTA sums (sum u, sum u^2, sum r, sum r^2, sum r*u).
__device__ __noinline__ void residual_stage(Ctx& cref, const Params& P, int st, const uint32_t* xin0, size_t o_stat, float* resid_io,
                                            float* rr_out /*[M]*/, int ts) {
    Ctx c = cref;  // scalar-replaced local copy: the fields live in registers, not behind a pointer
    stage_core(c, st, 1, xin0, xin0, ts);
    float u = 0.f, resid = *resid_io;
    if (c.ownerC) {
        const StageTab& tb = c.S.tab[st];
        u = row_value(c.S.red, c.oc >> 4, 1 << tb.lgKG, c.oc & 15, c.om, (long long)c.S.q128[c.om], c.S.invs[c.om]) *
            pbf(c.pb, tb.goff[c.oc >> 4], c.oc & 15, c.pdt);
        c.S.u[c.om * 192 + c.oc] = u;
        c.S.u[c.om * 192 + 96 + c.oc] = resid;
    }
    csync(c, __LINE__);
    if (c.tid < 2 * kMaxTok) c.S.q128[c.tid] = 0ull;
    if (c.warp < kMaxTok) {  // this CTA's five sums per token
        const int m = c.warp;
        double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (m < c.M)
            for (int k = c.lane; k < c.nownC; k += 32) {
                const double uu = (double)c.S.u[m * 192 + k], r = (double)c.S.u[m * 192 + 96 + k];
                s[0] += uu; s[1] += uu * uu; s[2] += r; s[3] += r * r; s[4] += r * uu;
            }
#pragma unroll
        for (int i = 0; i < 5; ++i) s[i] = warp_sum_d(s[i]);
        if (c.lane == 0 && m < c.max_batch) {
            const int nc = 1;
            const size_t a = o_stat + ((size_t)m * c.ncta + c.cta) * kStatW;  // this CTA's record: 12 consecutive words
            if (m < c.M) {
#pragma unroll
                for (int i = 0; i < 5; ++i) put_hilo(c.X + a + 2 * i * nc, c.Xc + a + 2 * i * nc, nc, s[i]);
                stv1(c.X + a + 10 * nc, 0u); stv1(c.X + a + 11 * nc, 0u);
                stv1(c.Xc + a + 10 * nc, kSentF); stv1(c.Xc + a + 11 * nc, kSentF);
            } else {
                for (int i = 0; i < kStatW; ++i) stv1(c.Xc + a + i * nc, kSentF);
            }
        }
    }
    stats_exchange(c, o_stat, 5, 0);
    stamp(c, ts + 2);
    if (c.tid < c.M) {  // one thread per token: LayerNorm statistics and the RMSNorm factor of the updated stream
        const double* rd = c.S.redd + c.tid * 5;
        const double N = (double)c.H;
        float mean, rstd;
        ln_finish(rd[0], rd[1], P.inv_H, P.ln_eps, &mean, &rstd);
        const double mu = (double)mean, rs = (double)rstd;
        // sum over the full width of (r + (u - mu) rs)^2, with the SAME rounded mean / rstd the owners apply
        const double ssq = rd[3] + 2.0 * rs * (rd[4] - mu * rd[2]) + rs * rs * (rd[1] - 2.0 * mu * rd[0] + N * mu * mu);
        c.S.fscr[c.tid] = rsqrtf((float)(fmax(ssq, 0.0) * P.inv_H) + P.rms_eps);  // LlamaRMSNorm
        c.S.fscr[kMaxTok + c.tid] = mean;
        c.S.fscr[2 * kMaxTok + c.tid] = rstd;
    }
    csync(c, __LINE__);
    if (c.ownerC) *resid_io = resid + (u - c.S.fscr[kMaxTok + c.om]) * c.S.fscr[2 * kMaxTok + c.om];
    for (int m = 0; m < c.M; ++m) rr_out[m] = c.S.fscr[m];
    cref.R = c.R;
    cref.tseq = c.tseq;
}

// stage A: q, k, v = BitLinear(RMSNorm(x)) (:522-524) — publishes raw g*t and per-CTA LayerNorm partials
__device__ __noinline__ void stage_qkv(Ctx& cref, const Params& P, int l, const int* s_pos) {
    Ctx c = cref;  // scalar-replaced local copy: the fields live in registers, not behind a pointer
    const StageTab& tb = c.S.tab[ST_A];
    const int H = c.H, T = tb.T, p_lo = tb.p_lo;
    uint32_t* XL = c.X + (size_t)l * P.per_layer;
    uint32_t* XLc = c.Xc + (size_t)l * P.per_layer;
    if (c.tid < 2 * kMaxTok) c.S.invs[c.tid] = pow2d(pbi(c.pb, 0, min(p_lo + (c.tid >> 1), 2)) - 29);
    // attention CTAs: pull this layer's cached K/V rows of their (sequence, head) towards L2
    if (c.cta < P.heads * c.M) {
        const int am = c.cta >= P.heads ? 1 : 0, ah = c.cta - am * P.heads;
        const size_t base = ((((size_t)l * P.max_batch + am) * P.heads + ah) * P.max_seq) * kHeadDim;
        const int lines = s_pos[am] * 2;  // 256 B per row = 2 x 128 B lines
        for (int i = c.tid; i < lines; i += kCT) {
            prefetch_l2(reinterpret_cast<const char*>(P.kcache + base) + (size_t)i * 128);
            prefetch_l2(reinterpret_cast<const char*>(P.vcache + base) + (size_t)i * 128);
        }
    }
    stage_core(c, ST_A, tb.nsets, XL + P.o_xA + (size_t)p_lo * kMaxTok * H, XL + P.o_xA + (size_t)min(p_lo + 1, 2) * kMaxTok * H, 6);
    const int Rr = 16 * T;
    if (c.tid < Rr * kMaxTok) {
        const int em = c.tid >= Rr ? 1 : 0, er = c.tid - em * Rr, ti = er >> 4;
        const size_t a = P.o_qkv + (size_t)em * 3 * H + tb.aux[ti] + (er & 15);
        if (em < c.M) {
            const int ps = tb.pslot[ti];
            const float u = row_value(c.S.red, ti, 1 << tb.lgKG, er & 15, em, (long long)c.S.q128[ps * kMaxTok + em], c.S.invs[ps * kMaxTok + em]) *
                            pbf(c.pb, tb.goff[ti], er & 15, c.pdt);
            c.S.u[em * 192 + er] = u;
            stv1(XL + a, fbits(u));
        }
        if (em < c.max_batch) stv1(XLc + a, kSentF);
    }
    csync(c, __LINE__);
    if (c.tid < 2 * kMaxTok) c.S.q128[c.tid] = 0ull;
    if (c.warp < 3 * kMaxTok) {  // per-(token, projection) partial (sum, sum of squares) of this CTA's rows
        const int m = c.warp >= 3 ? 1 : 0, p = c.warp - 3 * m;
        double s = 0.0, q = 0.0;
        if (m < c.M)
            for (int r = c.lane; r < Rr; r += 32)
                if ((int)tb.pslot[r >> 4] + p_lo == p) {
                    const double v = (double)c.S.u[m * 192 + r];
                    s += v;
                    q += v * v;
                }
        s = warp_sum_d(s);
        q = warp_sum_d(q);
        if (c.lane == 0 && m < c.max_batch) {
            const int nc = 1;
            const size_t a = P.o_qst + (((size_t)m * 3 + p) * c.ncta + c.cta) * 4;  // [token][projection][cta][word 0..3]
            if (m < c.M) {
                put_hilo(XL + a, XLc + a, nc, s);
                put_hilo(XL + a + 2 * nc, XLc + a + 2 * nc, nc, q);
            } else {
                for (int i = 0; i < 4; ++i) stv1(XLc + a + i * nc, kSentF);
            }
        }
    }
    cref.R = c.R;
    cref.tseq = c.tseq;
}

// stage B: attention for one new token per sequence (:536-563): LayerNorm of q/k/v (bitnet.py:118) from the partials,
// RoPE (:176-181), cache append, fp32 online softmax over the cache, digits of o_proj's input
__device__ __noinline__ void stage_attention(Ctx& cref, const Params& P, int l, const int* s_pos) {
    Ctx c = cref;  // scalar-replaced local copy: the fields live in registers, not behind a pointer
    const int H = c.H, tid = c.tid, lane = c.lane, warp = c.warp, ncta = c.ncta;
    uint32_t* XL = c.X + (size_t)l * P.per_layer;
    uint32_t* XLc = c.Xc + (size_t)l * P.per_layer;
    if (c.cta >= P.heads * c.M) {
        if (c.cta < P.heads * P.max_batch && tid < 128) {  // a larger batch would publish here: keep the other set armed
            const int am = c.cta / P.heads, ah = c.cta - am * P.heads;
            stv1(XLc + P.o_xC + (size_t)am * H + frag_word(ah * kHeadDim + (tid >> 5) * 32, (tid & 31) >> 2, tid & 3), kSentD);
        }
        return;
    }
    const int am = c.cta / P.heads, ah = c.cta - am * P.heads;
    const int pos = s_pos[am], Tctx = pos + 1;
    float* raw = reinterpret_cast<float*>(c.S.red);  // [3][128] raw q, k, v of this head
    float* sq = raw + 384;                            // [128] rotated, scaled q
    float* sk = sq + 128;                             // [128] new k row (fp16-rounded)
    float* sv = sk + 128;                             // [128] new v row (fp16-rounded)
    float* part = sv + 128;                           // [16][132]: m, l, o[128] per warp
    poll_copy_f(XL + P.o_qst + (size_t)am * 3 * ncta * 4, c.S.stat, 3 * ncta, tid, c.abort_flag);
    if (tid < 96) {
        const int p = tid >> 5, i = tid & 31;
        const uint32_t* src = XL + P.o_qkv + ((size_t)am * 3 + p) * H + ah * kHeadDim + 4 * i;
        uint4 v = ldv4(src);
        int spins = 0;
        while (bad_f(v)) {
            if (spin_giveup(spins, c.abort_flag)) break;
            v = ldv4(src);
        }
        *reinterpret_cast<uint4*>(raw + p * 128 + 4 * i) = v;
    }
    csync(c, __LINE__);
    if (warp < 6) {  // (projection, sum | sumsq): fixed-order reduction over the CTAs
        const int p = warp >> 1, which = warp & 1;
        const float* st = reinterpret_cast<const float*>(c.S.stat) + (size_t)p * ncta * 4 + 2 * which;
        double s = 0.0;
        for (int k = lane; k < ncta; k += 32) {
            const float2 hl = *reinterpret_cast<const float2*>(st + k * 4);
            s += (double)hl.x + (double)hl.y;
        }
        s = warp_sum_d(s);
        if (lane == 0) c.S.redd[warp] = s;
    }
    csync(c, __LINE__);
    stamp(c, 17);
    if (tid < 3) ln_finish(c.S.redd[2 * tid], c.S.redd[2 * tid + 1], P.inv_H, P.ln_eps, &c.S.fscr[2 * tid], &c.S.fscr[2 * tid + 1]);
    csync(c, __LINE__);
    __half* kc = P.kcache + ((((size_t)l * P.max_batch + am) * P.heads + ah) * P.max_seq) * kHeadDim;
    __half* vc = P.vcache + ((((size_t)l * P.max_batch + am) * P.heads + ah) * P.max_seq) * kHeadDim;
    if (tid < kHeadDim) {
        const float mean[3] = {c.S.fscr[0], c.S.fscr[2], c.S.fscr[4]}, rstd[3] = {c.S.fscr[1], c.S.fscr[3], c.S.fscr[5]};
        const int d = tid, half = kHeadDim / 2, dp = d < half ? d + half : d - half, fi = d < half ? d : d - half;
        const float cs = P.rope_cos[(size_t)pos * half + fi], sn = P.rope_sin[(size_t)pos * half + fi];
        const float q0 = (raw[d] - mean[0]) * rstd[0], q1 = (raw[dp] - mean[0]) * rstd[0];
        const float k0 = (raw[128 + d] - mean[1]) * rstd[1], k1 = (raw[128 + dp] - mean[1]) * rstd[1];
        const float qr = d < half ? q0 * cs - q1 * sn : q0 * cs + q1 * sn;  // rotate_half: (-x2, x1), :168-181
        const float kr = d < half ? k0 * cs - k1 * sn : k0 * cs + k1 * sn;
        const float vv = (raw[256 + d] - mean[2]) * rstd[2];
        const __half kh = __float2half_rn(kr), vh = __float2half_rn(vv);
        kc[(size_t)pos * kHeadDim + d] = kh;
        vc[(size_t)pos * kHeadDim + d] = vh;
        sq[d] = qr * 0.08838834764831845f;  // 1 / sqrt(128), :546
        sk[d] = __half2float(kh);
        sv[d] = __half2float(vh);
    }
    csync(c, __LINE__);
    stamp(c, 18);
    {   // each warp: online softmax over positions warp, warp+16, ...; lane holds dims 4*lane .. 4*lane+3
        const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * lane);
        float mx = -INFINITY, lsum = 0.f;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = warp; j0 < Tctx; j0 += 4 * kCW) {
            uint2 kr[4], vr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * kCW;
                kr[u] = make_uint2(0u, 0u);
                vr[u] = make_uint2(0u, 0u);
                if (j < pos) {
                    kr[u] = *reinterpret_cast<const uint2*>(kc + (size_t)j * kHeadDim + 4 * lane);
                    vr[u] = *reinterpret_cast<const uint2*>(vc + (size_t)j * kHeadDim + 4 * lane);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * kCW;
                if (j < Tctx) {
                    float4 kf, vf;
                    if (j < pos) {
                        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&kr[u].x));
                        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&kr[u].y));
                        const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&vr[u].x));
                        const float2 d2 = __half22float2(*reinterpret_cast<const __half2*>(&vr[u].y));
                        kf = make_float4(a.x, a.y, b.x, b.y);
                        vf = make_float4(c2.x, c2.y, d2.x, d2.y);
                    } else {
                        kf = *reinterpret_cast<const float4*>(sk + 4 * lane);
                        vf = *reinterpret_cast<const float4*>(sv + 4 * lane);
                    }
                    float dot = qv.x * kf.x + qv.y * kf.y + qv.z * kf.z + qv.w * kf.w;
                    dot = warp_sum(dot);
                    const float mn = fmaxf(mx, dot);
                    const float sc = expf(mx - mn), pj = expf(dot - mn);
                    lsum = lsum * sc + pj;
                    o.x = o.x * sc + pj * vf.x; o.y = o.y * sc + pj * vf.y;
                    o.z = o.z * sc + pj * vf.z; o.w = o.w * sc + pj * vf.w;
                    mx = mn;
                }
            }
        }
        float* pw = part + warp * 132;
        if (lane == 0) { pw[0] = mx; pw[1] = lsum; }
        *reinterpret_cast<float4*>(pw + 4 + 4 * lane) = o;
    }
    csync(c, __LINE__);
    stamp(c, 19);
    if (tid < kHeadDim) {
        float mx = -INFINITY;
        for (int w = 0; w < kCW; ++w) mx = fmaxf(mx, part[w * 132]);
        float den = 0.f, num = 0.f;
        for (int w = 0; w < kCW; ++w) {
            const float m_w = part[w * 132];
            const float f = m_w == -INFINITY ? 0.f : expf(m_w - mx);
            den += part[w * 132 + 1] * f;
            num += part[w * 132 + 4 + tid] * f;
        }
        const int col = ah * kHeadDim + tid;
        c.S.stage[tid] = quant_digits((num / den) * pbf(c.pb, PB_hO, tid, c.pdt), pbi(c.pb, 0, 3), col);
    }
    csync(c, __LINE__);
    if (tid < 128) {
        const int blk = tid >> 5, jd = tid & 31, j = jd >> 2, d = jd & 3;
        const uint32_t* sp = c.S.stage + blk * 32 + j;
        const uint32_t sel = (uint32_t)d | ((uint32_t)(4 + d) << 4);
        const uint32_t word = __byte_perm(__byte_perm(sp[0], sp[8], sel), __byte_perm(sp[16], sp[24], sel), 0x5410);
        const size_t a = P.o_xC + (size_t)am * H + frag_word(ah * kHeadDim + blk * 32, j, d);
        stv1(XL + a, word);
        stv1(XLc + a, kSentD);
    }
    cref.R = c.R;
    cref.tseq = c.tseq;
}

// stage D1: gate, up (:257) — 16 gate rows + 16 up rows of the same columns live in the same CTA, so
// silu(LN(gate)) * LN(up) * input_factor(down) is finished by the owner after ONE exchange of sums and bounds
__device__ __noinline__ void stage_gate_up(Ctx& cref, const Params& P, int l, int d_b0, int d_b1) {
    Ctx c = cref;  // scalar-replaced local copy: the fields live in registers, not behind a pointer
    const int H = c.H, tid = c.tid, lane = c.lane, warp = c.warp;
    uint32_t* XL = c.X + (size_t)l * P.per_layer;
    uint32_t* XLc = c.Xc + (size_t)l * P.per_layer;
    const int npb = d_b1 - d_b0, T = 2 * npb, nown = 16 * npb, col0 = 16 * d_b0;
    if (tid < 2 * kMaxTok) c.S.invs[tid] = pow2d(pbi(c.pb, 0, 4 + (tid >> 1)) - 29);
    stage_core(c, ST_D1, 2, XL + P.o_xD1, XL + P.o_xD1 + (size_t)kMaxTok * H, 11);
    const int Rr = 16 * T;
    if (Rr > 0) {
        const int em = tid >= Rr ? 1 : 0, er = tid - em * Rr;
        if (em < c.M && tid < 2 * Rr) {
            const StageTab& tb = c.S.tab[ST_D1];
            const int ps = tb.pslot[er >> 4];
            c.S.u[em * 192 + er] = row_value(c.S.red, er >> 4, 1 << tb.lgKG, er & 15, em, (long long)c.S.q128[ps * kMaxTok + em], c.S.invs[ps * kMaxTok + em]) *
                                   pbf(c.pb, tb.goff[er >> 4], er & 15, c.pdt);
        }
    }
    csync(c, __LINE__);
    if (tid < 2 * kMaxTok) c.S.q128[tid] = 0ull;
    const bool ownerD = tid < nown * c.M;
    const int dm = (ownerD && tid >= nown) ? 1 : 0, dc = ownerD ? tid - dm * nown : 0;
    float gv = 0.f, uv = 0.f, hd = 0.f;
    if (ownerD) {
        gv = c.S.u[dm * 192 + dc];
        uv = c.S.u[dm * 192 + nown + dc];
        hd = pbf(c.pb, PB_hD, dc, c.pdt);
    }
    if (warp < kMaxTok) {
        const int m = warp;
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        float mx[3] = {0.f, 0.f, 0.f};
        if (m < c.M)
            for (int k = lane; k < nown; k += 32) {
                const float gf = c.S.u[m * 192 + k], uf = c.S.u[m * 192 + nown + k], hf = fabsf(pbf(c.pb, PB_hD, k, c.pdt));
                s[0] += (double)gf; s[1] += (double)gf * (double)gf; s[2] += (double)uf; s[3] += (double)uf * (double)uf;
                mx[0] = fmaxf(mx[0], fabsf(gf * uf) * hf); mx[1] = fmaxf(mx[1], fabsf(gf) * hf); mx[2] = fmaxf(mx[2], fabsf(uf) * hf);
            }
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = warp_sum_d(s[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) mx[i] = warp_max_f(mx[i]);
        if (lane == 0 && m < c.max_batch) {
            const int nc = 1;
            const size_t a = P.o_dst + ((size_t)m * c.ncta + c.cta) * kStatW;  // this CTA's record: 12 consecutive words
            if (m < c.M) {
#pragma unroll
                for (int i = 0; i < 4; ++i) put_hilo(XL + a + 2 * i * nc, XLc + a + 2 * i * nc, nc, s[i]);
#pragma unroll
                for (int i = 0; i < 3; ++i) { stv1(XL + a + (8 + i) * nc, fbits(mx[i])); stv1(XLc + a + (8 + i) * nc, kSentF); }
                stv1(XL + a + 11 * nc, 0u); stv1(XLc + a + 11 * nc, kSentF);
            } else {
                for (int i = 0; i < kStatW; ++i) stv1(XLc + a + i * nc, kSentF);
            }
        }
    }
    stats_exchange(c, (size_t)l * P.per_layer + P.o_dst, 4, 3);
    stamp(c, 13);
    if (tid < c.M) {  // per token: LayerNorm statistics of gate and up, power-of-two bound of down_proj's input
        const double* rd = c.S.redd + tid * 7;
        float mg, rg, mu, ru;
        ln_finish(rd[0], rd[1], P.inv_I, P.ln_eps, &mg, &rg);
        ln_finish(rd[2], rd[3], P.inv_I, P.ln_eps, &mu, &ru);
        const float amg = fabsf(mg), amu = fabsf(mu);
        // |silu(g^) u^ h| <= |g^||u^||h| <= rg ru (|g u h| + |mg||u h| + |mu||g h| + |mg mu||h|)
        const float bound = rg * ru * ((float)rd[4] + amg * (float)rd[6] + amu * (float)rd[5] + amg * amu * __int_as_float(pbi(c.pb, 0, 6))) * 1.0001f;
        int e = 0;
        if (bound > 0.f && bound < 3.0e38f) frexpf(bound, &e);
        float* f = c.S.fscr + tid * 8;
        f[0] = mg; f[1] = rg; f[2] = mu; f[3] = ru;
        reinterpret_cast<int*>(f)[4] = e;
        c.S.invs[tid] = pow2d(e - 29);  // scale of down_proj's input digits (pslot 0)
    }
    csync(c, __LINE__);
    if (ownerD) {
        const float* f = c.S.fscr + dm * 8;
        const float gh = (gv - f[0]) * f[1], uh = (uv - f[2]) * f[3];
        const float act = __fdiv_rn(gh, 1.0f + expf(-gh)) * uh;  // act_fn(gate) * up, :257
        c.S.stage[dm * nown + dc] = quant_digits(act * hd, reinterpret_cast<const int*>(f)[4], col0 + dc);
    }
    csync(c, __LINE__);
    publish16(c, (size_t)l * P.per_layer + P.o_xD2, (size_t)c.I, npb, col0);
    cref.R = c.R;
    cref.tseq = c.tseq;
}

// lm_head (:1610-1611) + greedy argmax (generation/utils.py:2540): kLmRows fp16 rows per ring chunk, one warp per chunk
__device__ __noinline__ void stage_lm_head(Ctx& cref, const Params& P, int v_b0, int v_b1, float* __restrict__ logits, const int* s_pos,
                                           unsigned long long step, unsigned long long* tr) {
    Ctx c = cref;  // scalar-replaced local copy: the fields live in registers, not behind a pointer
    const int H = c.H, M = c.M, tid = c.tid, lane = c.lane, warp = c.warp, ncta = c.ncta;
    uint32_t* Xt = c.X + (size_t)P.L * P.per_layer;
    uint32_t* Xtc = c.Xc + (size_t)P.L * P.per_layer;
    uint32_t* xs = c.S.dbuf;  // [M][H/2] half2 words
    for (int m = 0; m < M; ++m) poll_copy_f(Xt + P.o_xfin + (size_t)m * (H / 2), xs + (size_t)m * (H / 2), H / 8, tid, c.abort_flag);
    csync(c, __LINE__);
    float best[kMaxTok];
    int bidx[kMaxTok];
#pragma unroll
    for (int m = 0; m < kMaxTok; ++m) { best[m] = -INFINITY; bidx[m] = 0x7fffffff; }
    const int nch = H / 8;  // uint4 chunks per row
    // chunks of kLmRows (= 2) rows = 2 slots, starting on an even slot of an even-sized ring: never straddle the end.
    // Chunk j sits at slot (s0 + 2 (j mod hs)) mod NS, hs = NS / 2 <= 16. Warp w < hs owns slot pair w for the whole
    // stage (chunks w, w + hs, ...): a barrier is then always waited on by the same warp, one phase at a time (a second
    // warp sharing the slot could run two phases ahead and be fooled by the parity test).
    if (c.R.slot & 1u) c.R.slot = c.R.slot + 1u >= c.R.NS ? 0u : c.R.slot + 1u;  // (the producer skips the same slot)
    const uint32_t hs = c.R.NS >> 1;
    uint32_t sl = c.R.slot + 2u * (uint32_t)warp;
    if (sl >= c.R.NS) sl -= c.R.NS;
    uint32_t par = (c.R.phase >> sl) & 1u;
    for (int v0 = v_b0 + kLmRows * warp; (uint32_t)warp < hs && v0 < v_b1; v0 += kLmRows * (int)hs, par ^= 1u) {
        const int nr = min(kLmRows, v_b1 - v0);
        const uint32_t off = sl * (uint32_t)(2 * H);
        mbar_wait_b(&c.full[sl], par, c.abort_flag);
        float acc[kLmRows][kMaxTok];
#pragma unroll
        for (int r = 0; r < kLmRows; ++r)
#pragma unroll
            for (int m = 0; m < kMaxTok; ++m) acc[r][m] = 0.f;
        const uint4* wr = reinterpret_cast<const uint4*>(c.S.ring + off);
        for (int i = lane; i < nch; i += 32) {
            uint4 xv[kMaxTok];
#pragma unroll
            for (int m = 0; m < kMaxTok; ++m)
                xv[m] = m < M ? reinterpret_cast<const uint4*>(xs + (size_t)m * (H / 2))[i] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int r = 0; r < kLmRows; ++r) {
                if (r < nr) {
                    const uint4 wv = wr[r * nch + i];
                    const __half2* w2 = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                    for (int m = 0; m < kMaxTok; ++m) {
                        if (m < M) {
                            const __half2* x2 = reinterpret_cast<const __half2*>(&xv[m]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float2 a = __half22float2(w2[q]), b = __half22float2(x2[q]);
                                acc[r][m] += a.x * b.x + a.y * b.y;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane < kLmRows) mbar_arrive(&c.empty[sl + (uint32_t)lane]);
#pragma unroll
        for (int r = 0; r < kLmRows; ++r) {
            if (r < nr) {
                const int v = v0 + r;
#pragma unroll
                for (int m = 0; m < kMaxTok; ++m) {
                    if (m < M) {
                        const float rs = warp_sum(acc[r][m]);
                        if (lane == 0) logits[(size_t)m * P.V + v] = rs;
                        if (rs > best[m] || (rs == best[m] && v < bidx[m])) { best[m] = rs; bidx[m] = v; }
                    }
                }
            }
        }
    }
    float* sb = c.S.fscr;                                        // [kCW][kMaxTok] values
    int* si = reinterpret_cast<int*>(c.S.fscr + kCW * kMaxTok);  // [kCW][kMaxTok] indices
    if (lane == 0)
        for (int m = 0; m < kMaxTok; ++m) { sb[warp * kMaxTok + m] = best[m]; si[warp * kMaxTok + m] = bidx[m]; }
    csync(c, __LINE__);
    if (tid < kMaxTok && tid < c.max_batch) {
        const int m = tid;
        const size_t a = P.o_amax + ((size_t)m * ncta + c.cta) * 2;
        if (m < M) {
            float b = -INFINITY;
            int bi = 0x7fffffff;
            for (int w = 0; w < kCW; ++w) {
                const float x = sb[w * kMaxTok + m];
                const int xi = si[w * kMaxTok + m];
                if (x > b || (x == b && xi < bi)) { b = x; bi = xi; }
            }
            stv1(Xt + a, fbits(b));
            stv1(Xt + a + 1, (uint32_t)bi);
        }
        stv1(Xtc + a, kSentF);
        stv1(Xtc + a + 1, kSentF);
    }
    if (c.cta == 0) {
        poll_copy_f(Xt + P.o_amax, c.S.stat, M * ncta * 2 / 4, tid, c.abort_flag);  // ncta even: whole uint4s
        csync(c, __LINE__);
        if (tid < M) {
            const int m = tid;
            float b = -INFINITY;
            int bi = 0x7fffffff;
            for (int k = 0; k < ncta; ++k) {
                const float x = __uint_as_float(c.S.stat[(m * ncta + k) * 2]);
                const int xi = (int)c.S.stat[(m * ncta + k) * 2 + 1];
                if (x > b || (x == b && xi < bi)) { b = x; bi = xi; }
            }
            P.ids[m] = bi == 0x7fffffff ? 0 : bi;
            P.pos[m] = s_pos[m] + 1;
        }
        csync(c, __LINE__);
        if (tid == 0) {
            if (tr != nullptr) tr[1] = gtime();
            __threadfence();
            *P.step_counter = step + 1ull;
        }
    }
    cref.R = c.R;
    cref.tseq = c.tseq;
}

// TMA producer warp: walks the static schedule of this CTA's weight tiles, independent of the dependency chain.
// Per layer it first fetches the layer's PARAMETER BLOCK (every weight_scale / input_factor / norm-weight slice and
// quantiser exponent this CTA's compute warps will touch in the layer, see PB_* above) into one of two shared-memory
// buffers, so that no parameter load ever sits on the dependency chain.
__device__ __noinline__ void producer_loop(const Params& P, unsigned char* smem_raw, unsigned char* pbuf, int ring_bytes, uint64_t* s_full,
                                           uint64_t* s_empty, uint64_t* s_pfull, uint64_t* s_pempty, int lane, int cta, int M, int a_b0, int a_b1, int c_b0, int c_b1, int d_b0, int d_b1, int v_b0,
                                           int v_b1) {
    const int H = P.H, I = P.I, tH = H >> 4, L = P.L;
    const uint32_t tileH = 2u * (uint32_t)H, tileI = 2u * (uint32_t)I;  // bytes of a re-tiled 16-row tile
    const uint32_t slot_bytes = 2u * (uint32_t)H, nsD = (uint32_t)((I + H - 1) / H);
    Ring R;
    R.slot = 0; R.NS = (uint32_t)ring_bytes / slot_bytes; R.phase = 0xFFFFFFFFu;  // parity 1: a fresh barrier passes at once
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    auto take_empty = [&](uint32_t sl) -> bool {  // the previous occupant of the slot (if any) has been released
        if (!mbar_wait_b(&s_empty[sl], (R.phase >> sl) & 1u, P.abort_flag)) return false;
        R.phase ^= 1u << sl;
        return true;
    };
    auto issue = [&](const uint8_t* src, uint32_t bytes, uint32_t nslots) -> bool {
        uint32_t skip;
        const uint32_t sl = ring_take(R, nslots, skip);  // skipped slots: no barrier traffic on either side
        for (uint32_t k = 0; k < nslots; ++k)
            if (!take_empty(sl + k)) return false;
        if (lane == 0) {
            mbar_expect_tx(&s_full[sl], bytes);
            bulk_g2s_hint(smem_raw + (size_t)sl * slot_bytes, src, bytes, &s_full[sl], pol);
        }
        __syncwarp();
        return true;
    };
    const int es = (int)dsize(P.pdt);
    const int nownC = 32 * (c_b1 - c_b0), colC0 = 32 * c_b0, nownD = 16 * (d_b1 - d_b0), colD0 = 16 * d_b0, TA = 2 * (a_b1 - a_b0);
    const bool attn = cta < P.heads * M;
    for (int l = 0; l < L; ++l) {
        const LayerDev& Ly = P.layers[l];
        {   // ---- parameter block of layer l: lane i owns copy i
            unsigned char* pb = pbuf + (size_t)(l & 1) * kPBBytes;
            if (!mbar_wait_b(&s_pempty[l & 1], ((l >> 1) & 1) ^ 1, P.abort_flag)) return;
            const bool last = l + 1 == L;
            const LayerDev& Ln = P.layers[last ? l : l + 1];
            const char* src = nullptr;
            uint32_t bytes = 0, dst = 0;
            auto slice = [&](const void* base, int first, int count, int off_el) {
                src = static_cast<const char*>(base) + (size_t)first * es;
                bytes = (uint32_t)(count * es);
                dst = (uint32_t)(kPBScalarBytes + off_el * es);
            };
            switch (lane) {
                case 0: src = reinterpret_cast<const char*>(P.lscal + (size_t)l * 8); bytes = kPBScalarBytes; dst = 0; break;
                case 1: slice(Ly.o.g, colC0, nownC, PB_gO); break;
                case 2: slice(Ly.down.g, colC0, nownC, PB_gDn); break;
                case 3: slice(last ? P.final_norm : Ln.ln_in, colC0, nownC, PB_lnN); break;
                case 4: if (!last) slice(Ln.q.h, colC0, nownC, PB_hq); break;
                case 5: if (!last) slice(Ln.k.h, colC0, nownC, PB_hk); break;
                case 6: if (!last) slice(Ln.v.h, colC0, nownC, PB_hv); break;
                case 7: slice(Ly.ln_post, colC0, nownC, PB_lnP); break;
                case 8: slice(Ly.gate.h, colC0, nownC, PB_hg); break;
                case 9: slice(Ly.up.h, colC0, nownC, PB_hu); break;
                case 10: slice(Ly.gate.g, colD0, nownD, PB_gG); break;
                case 11: slice(Ly.up.g, colD0, nownD, PB_gU); break;
                case 12: slice(Ly.down.h, colD0, nownD, PB_hD); break;
                case 13: if (attn) slice(Ly.o.h, (cta % P.heads) * kHeadDim, kHeadDim, PB_hO); break;
                default: {
                    const int i = lane - 14;
                    if (i < TA) {
                        const int gt = 2 * a_b0 + i, prob = gt / tH, row0 = (gt - prob * tH) * 16;
                        slice(prob == 0 ? Ly.q.g : (prob == 1 ? Ly.k.g : Ly.v.g), row0, 16, PB_gA + 16 * i);
                    }
                }
            }
            uint32_t total = bytes;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            if (lane == 0) mbar_expect_tx(&s_pfull[l & 1], total);
            __syncwarp();
            if (bytes) imma::bulk_g2s(pb + dst, src, bytes, &s_pfull[l & 1]);
            __syncwarp();
        }
        for (int gt = 2 * a_b0; gt < 2 * a_b1; ++gt) {
            const int prob = gt / tH, lt = gt - prob * tH;
            const uint8_t* w = prob == 0 ? Ly.q.w : (prob == 1 ? Ly.k.w : Ly.v.w);
            if (!issue(w + (size_t)lt * tileH, tileH, 1u)) return;
        }
        for (int t = 2 * c_b0; t < 2 * c_b1; ++t) if (!issue(Ly.o.w + (size_t)t * tileH, tileH, 1u)) return;
        for (int pb = d_b0; pb < d_b1; ++pb) if (!issue(Ly.gate.w + (size_t)pb * tileH, tileH, 1u)) return;
        for (int pb = d_b0; pb < d_b1; ++pb) if (!issue(Ly.up.w + (size_t)pb * tileH, tileH, 1u)) return;
        for (int t = 2 * c_b0; t < 2 * c_b1; ++t) if (!issue(Ly.down.w + (size_t)t * tileI, tileI, nsD)) return;
    }
    if (R.slot & 1u) R.slot = R.slot + 1u >= R.NS ? 0u : R.slot + 1u;  // lm_head chunks start on an even slot (see stage_lm_head)
    for (int v = v_b0; v < v_b1; v += kLmRows) {  // lm_head rows are contiguous: kLmRows rows per chunk, one bulk copy
        const int nr = min(kLmRows, v_b1 - v);
        if (!issue(reinterpret_cast<const uint8_t*>(P.lm_head + (size_t)v * H), (uint32_t)(nr * 2 * H), (uint32_t)kLmRows)) return;
    }
}

}  // namespace

// ===============================================================================================================
// the kernel
// ===============================================================================================================
__global__ void __launch_bounds__(kThreads, 1)
step_kernel(const Params* __restrict__ Pg, const long long* __restrict__ ids_in, float* __restrict__ logits, int M,
            int ring_bytes, int dbuf_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kMaxSlots], s_empty[kMaxSlots], s_pfull[2], s_pempty[2];
    __shared__ Params P;
    __shared__ int s_tok[kMaxTok], s_pos[kMaxTok];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Pg);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&P);
        for (int i = tid; i < (int)(sizeof(Params) / 4); i += kThreads) dst[i] = src[i];
    }
    if (tid == 0) {
        s_deadline = gtime() + kStepBudgetNs;
        s_abort_info = Pg->trace;
        s_site = -1;
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_pfull[i], 1);
            mbar_init(&s_pempty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ncta = P.ncta, H = P.H, I = P.I, L = P.L;
    int a_b0, a_b1, c_b0, c_b1, d_b0, d_b1, v_b0, v_b1;
    block_range(3 * H / 32, cta, ncta, a_b0, a_b1);   // stage A: 32-row blocks of the [q;k;v] row space
    block_range(H / 32, cta, ncta, c_b0, c_b1);       // stages C, D2: 32-row blocks of o_proj / down_proj
    block_range(I / 16, cta, ncta, d_b0, d_b1);       // stage D1: pair blocks (16 gate rows + 16 up rows)
    block_range(P.V, cta, ncta, v_b0, v_b1);          // lm_head rows

    if (warp == kCW) {
        producer_loop(P, smem_raw, smem_raw + ring_bytes + dbuf_bytes, ring_bytes, s_full, s_empty, s_pfull, s_pempty, lane, cta,
                      M, a_b0, a_b1, c_b0, c_b1, d_b0, d_b1, v_b0, v_b1);
        return;
    }

    // ---- compute warps ----
    Ctx c;
    c.tid = tid; c.lane = lane; c.warp = warp; c.cta = cta; c.ncta = ncta; c.M = M; c.H = H; c.I = I; c.pdt = P.pdt;
    c.ring_bytes = ring_bytes; c.max_batch = P.max_batch;
    {
        unsigned char* p = smem_raw;
        c.S.ring = p; p += ring_bytes;
        c.S.dbuf = reinterpret_cast<uint32_t*>(p); p += dbuf_bytes;
        c.S.pbuf = p; p += 2 * kPBBytes;
        c.S.red = reinterpret_cast<int*>(p); p += kRedBytes;
        c.S.stat = reinterpret_cast<uint32_t*>(p); p += (size_t)M * ncta * kStatW * 4;
        c.S.u = reinterpret_cast<float*>(p); p += kMaxTok * 192 * 4;
        c.S.stage = reinterpret_cast<uint32_t*>(p); p += kMaxTok * 96 * 3 * 4;
        c.S.redd = reinterpret_cast<double*>(p); p += 32 * 8;
        c.S.q128 = reinterpret_cast<unsigned long long*>(p); p += 2 * kMaxTok * 8;
        c.S.invs = reinterpret_cast<double*>(p); p += 2 * kMaxTok * 8;
        c.S.tab = reinterpret_cast<StageTab*>(p); p += 4 * sizeof(StageTab);
        c.S.tslot = reinterpret_cast<uint32_t*>(p); p += kMaxTiles * 4;
        c.S.fscr = reinterpret_cast<float*>(p);
    }
    c.full = s_full; c.empty = s_empty; c.abort_flag = P.abort_flag;
    c.R.slot = 0; c.R.NS = (uint32_t)ring_bytes / (2u * (uint32_t)H); c.R.phase = 0u;
    const unsigned long long step = *P.step_counter;
    const int par = (int)(step & 1ull);
    c.X = P.xch[par];
    c.Xc = P.xch[par ^ 1];
    c.nownC = 32 * (c_b1 - c_b0); c.colC0 = 32 * c_b0;
    c.ownerC = c.nownC > 0 && tid < c.nownC * M;
    c.om = c.ownerC ? tid / c.nownC : 0;
    c.oc = c.ownerC ? tid - c.om * c.nownC : 0;
    const int who = !P.trace_on ? -1 : (cta == 0 ? 0 : (cta == ncta - 1 ? 1 : -1));
    unsigned long long* trace = who >= 0 ? P.trace + (size_t)who * (L + 2) * kTracePoints : nullptr;
    const bool tracer = who >= 0 && tid == 0;
    c.trl = nullptr;
    c.tseq = 0;
    if (tid < kMaxTok) {
        long long id = tid < M ? ids_in[tid] : 0;
        id = id < 0 ? 0 : (id >= P.V ? P.V - 1 : id);
        int ps = tid < M ? P.pos[tid] : 0;
        if (ps < 0 || ps >= P.max_seq) {  // decoding past the cache: refuse loudly (the host checks first), never overrun
            atomicExch(P.abort_flag, 2);
            ps = ps < 0 ? 0 : P.max_seq - 1;
        }
        s_tok[tid] = (int)id;
        s_pos[tid] = ps;
    }
    if (tid < 2 * kMaxTok) c.S.q128[tid] = 0ull;
    if (tid < 4) {  // the four layer-invariant stage descriptions of this CTA
        StageTab& tb = c.S.tab[tid];
        const int tH = H >> 4, npb = d_b1 - d_b0;
        tb.K = tid == ST_D2 ? I : H;
        tb.T = tid == ST_A ? 2 * (a_b1 - a_b0) : (tid == ST_D1 ? 2 * npb : 2 * (c_b1 - c_b0));
        tb.nsl = tid == ST_D2 ? (I + H - 1) / H : 1;
        plan_split(tb.T, tb.K >> 8, tb.lgKG, tb.tpg);
        tb.p_lo = (2 * a_b0) / tH;
        tb.nsets = tb.T > 0 ? (2 * a_b1 - 1) / tH - tb.p_lo + 1 : 1;
        for (int i = 0; i < tb.T; ++i) {
            int ps = 0, go = 0;
            if (tid == ST_A) {
                const int gt = 2 * a_b0 + i, prob = gt / tH;
                ps = prob - (2 * a_b0) / tH;
                go = PB_gA + 16 * i;
                tb.aux[i] = prob * H + (gt - prob * tH) * 16;
            }
            else if (tid == ST_C) { go = PB_gO + 16 * i; }
            else if (tid == ST_D1) { ps = i >= npb ? 1 : 0; go = i >= npb ? PB_gU + 16 * (i - npb) : PB_gG + 16 * i; }
            else { go = PB_gDn + 16 * i; }
            tb.pslot[i] = (unsigned char)ps;
            tb.goff[i] = (short)go;
        }
    }
    csync(c, __LINE__);
    if (tracer) trace[0] = gtime();

    float resid = 0.f;
    // ---- embedding -> residual stream, RMSNorm(input_layernorm of layer 0) -> digits of q/k/v inputs (:1202, :67-81)
    {
        float rr[kMaxTok];
        for (int m = 0; m < M; ++m) {
            const __half* erow = P.embed + (size_t)s_tok[m] * H;
            float part = 0.f;
            for (int i = tid; i < H / 8; i += kCT) {
                const uint4 raw = reinterpret_cast<const uint4*>(erow)[i];
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(h2[q]);
                    part += f.x * f.x + f.y * f.y;
                }
            }
            const double pd = warp_sum_d((double)part);
            if (lane == 0) c.S.redd[warp] = pd;
            csync(c, __LINE__);
            double tot = 0.0;
            for (int w = 0; w < kCW; ++w) tot += c.S.redd[w];
            rr[m] = rsqrtf((float)(tot / (double)H) + P.rms_eps);
            csync(c, __LINE__);
            if (c.ownerC && c.om == m) resid = __half2float(erow[c.colC0 + c.oc]);
        }
        publish_qkv_inputs<false>(c, P, 0, resid, rr);
    }

    for (int l = 0; l < L; ++l) {
        const size_t lbase = (size_t)l * P.per_layer;
        unsigned long long* tr = trace + (size_t)(1 + l) * kTracePoints;
        c.trl = who >= 0 ? tr : nullptr;
        c.pb = c.S.pbuf + (size_t)(l & 1) * kPBBytes;
        c.tseq = 0;
        if (tid == 0) s_site = l * 16 + 0;
        if (tracer) tr[0] = gtime();
        // the layer's parameter block has long arrived (the producer runs a layer ahead); the first readers are the
        // threads that set the stage scales, everybody else reads it after a CTA barrier
        if (tid < 2 * kMaxTok) mbar_wait_b(&s_pfull[l & 1], (l >> 1) & 1, c.abort_flag);
        stage_qkv(c, P, l, s_pos);
        if (tid == 0) s_site = l * 16 + 1;
        if (tracer) tr[1] = gtime();
        stage_attention(c, P, l, s_pos);
        if (tid == 0) s_site = l * 16 + 2;
        if (tracer) tr[2] = gtime();
        {   // stage C: o_proj (:580) + residual + post_attention_layernorm -> digits of gate / up inputs
            if (tid < kMaxTok) c.S.invs[tid] = pow2d(pbi(c.pb, 0, 3) - 29);
            float rr[kMaxTok];
            residual_stage(c, P, ST_C, c.X + lbase + P.o_xC, lbase + P.o_cst, &resid, rr, 8);
            publish_gu_inputs(c, P, l, resid, rr);
        }
        if (tid == 0) s_site = l * 16 + 3;
        if (tracer) tr[3] = gtime();
        stage_gate_up(c, P, l, d_b0, d_b1);
        if (tid == 0) s_site = l * 16 + 4;
        if (tracer) tr[4] = gtime();
        {   // stage D2: down_proj (:257) + residual (:918) + the next layer's input_layernorm (or the final norm, :1315)
            float rr[kMaxTok];
            residual_stage(c, P, ST_D2, c.X + lbase + P.o_xD2, lbase + P.o_d2st, &resid, rr, 14);
            if (l + 1 < L) {
                publish_qkv_inputs<true>(c, P, l + 1, resid, rr);
            } else {  // final RMSNorm -> fp16 x for lm_head, published as half2 words
                float* xf = reinterpret_cast<float*>(c.S.stage);
                if (c.ownerC) xf[c.om * 96 + c.oc] = resid * rr[c.om] * pbf(c.pb, PB_lnN, c.oc, c.pdt);
                csync(c, __LINE__);
                const int pairs = c.nownC / 2;
                uint32_t* Xt = c.X + (size_t)L * P.per_layer;
                uint32_t* Xtc = c.Xc + (size_t)L * P.per_layer;
                for (int it = tid; it < kMaxTok * pairs; it += kCT) {
                    const int m = it / pairs, pc = it - m * pairs;
                    const size_t a = P.o_xfin + (size_t)m * (H / 2) + (c.colC0 / 2) + pc;
                    if (m < M) {
                        const __half2 h2 = __floats2half2_rn(xf[m * 96 + 2 * pc], xf[m * 96 + 2 * pc + 1]);
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(&h2);
                        stv1(Xt + a, w == kSentF ? 0x7FFF7FFFu : w);
                    }
                    if (m < P.max_batch) stv1(Xtc + a, kSentF);
                }
            }
        }
        // every read of this layer's parameter block lies before the CTA barrier inside the publish above
        if (tid == 0) mbar_arrive(&s_pempty[l & 1]);
        if (tracer) tr[5] = gtime();
    }
    unsigned long long* tr = trace != nullptr ? trace + (size_t)(1 + L) * kTracePoints : nullptr;
    if (tid == 0) s_site = L * 16 + 5;
    if (tracer) tr[0] = gtime();
    stage_lm_head(c, P, v_b0, v_b1, logits, s_pos, step, tr);
}

}  // namespace persist

// ===================================================================================================================
// host side
// ===================================================================================================================
struct PersistState {
    persist::Params hp;              // host copy
    persist::Params* dp = nullptr;   // device copy
    persist::LayerDev* dlayers = nullptr;
    int* lscal = nullptr;
    uint8_t* wtiled = nullptr;       // all packed sign matrices, re-tiled (see row_pitch)
    uint32_t* xch[2] = {nullptr, nullptr};
    unsigned long long* step_counter = nullptr;
    int* abort_flag = nullptr;
    unsigned long long* trace = nullptr;
    int ncta = 0;
    int smem_limit = 0;
    int trace_words = 0;
};

namespace {

using namespace persist;

// [N][K/8] checkpoint layout -> tile-major fragment order (see row_pitch): one thread per 8 bytes of output
__global__ void retile_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int N, int Kb) {
    const size_t total = (size_t)N * Kb / 8;
    const int U = Kb / 32;
    for (size_t o8 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o8 < total; o8 += (size_t)gridDim.x * blockDim.x) {
        const size_t t = o8 / ((size_t)U * 64);
        const int r = (int)(o8 - t * (size_t)U * 64), u = r >> 6, q = r & 63, lane = q >> 1, half = q & 1, g = lane >> 2, t4 = lane & 3;
        const size_t row = t * 16 + g + 8 * half;
        reinterpret_cast<uint2*>(out)[o8] = *reinterpret_cast<const uint2*>(in + row * Kb + u * 32 + 8 * t4);
    }
}

__global__ void fill_words_kernel(uint32_t* p, size_t n, uint32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

float host_param(const void* p, size_t i, int dt) {
    if (dt == ONEBIT_F16) return __half2float(static_cast<const __half*>(p)[i]);
    if (dt == ONEBIT_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
    return static_cast<const float*>(p)[i];
}

int fetch(std::vector<float>& out, const void* dev, size_t n, int dt) {
    std::vector<unsigned char> raw(n * dtype_size(dt));
    ONEBIT_CUDA_TRY(cudaMemcpy(raw.data(), dev, raw.size(), cudaMemcpyDeviceToHost));
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = host_param(raw.data(), i, dt);
    return ONEBIT_OK;
}

int bound_exp(double b) {  // smallest e with b < 2^e (a little headroom for rounding in the norms)
    b *= 1.001;
    if (!(b > 0.0) || !std::isfinite(b)) return 0;
    int e = 0;
    std::frexp(b, &e);
    return e;
}

size_t round4(size_t w) { return (w + 3) & ~(size_t)3; }

void plan_smem(const Params& P, int M, int smem_limit, Geometry* g) {
    const int dbuf = (int)std::max<size_t>((size_t)2 * M * P.H * 4, (size_t)M * P.I * 4);
    const int fixed = kRedBytes + M * P.ncta * kStatW * 4 + kMaxTok * 192 * 4 + kMaxTok * 96 * 3 * 4 + 32 * 8 + 2 * kMaxTok * 8 * 2 +
                      4 * (int)sizeof(StageTab) + kMaxTiles * 4 + 64 * 4 + 128 + 2 * kPBBytes;
    g->dbuf_bytes = (dbuf + 127) & ~127;
    const int slot = 2 * P.H;  // one re-tiled 16-row tile of a K = H matrix = one fp16 lm_head row
    int ns = (smem_limit - 2048 - fixed - g->dbuf_bytes) / slot;
    ns = std::min(ns, kMaxSlots) & ~1;  // even: lm_head chunks are slot pairs
    g->ring_bytes = std::max(ns, 0) * slot;
    g->smem_bytes = g->ring_bytes + g->dbuf_bytes + fixed;
}

}  // namespace

bool persist_supported(const onebit_decoder_config& c) {
    static int env = -1;
    if (env < 0) {
        // Opt-in: measured on B200 (profiles/r02_persist_*), the single-kernel step is correct but slower than the
        // PDL-chained fused stages at batch 1 (2.75 ms vs 1.46 ms per LLaMA-7B step): every BitLinear stage costs two
        // all-CTA exchanges (data + LayerNorm statistics) instead of one kernel boundary.
        const char* e = getenv("ONEBIT_PERSIST");
        env = (e && e[0] == '1') ? 1 : 0;
    }
    if (!env) return false;
    const int tp = c.tp_size > 1 ? c.tp_size : 1;
    if (tp != 1) return false;
    if (c.max_batch > persist::kMaxTok) return false;
    if (c.hidden_size % 256 || c.intermediate_size % 256) return false;
    if (c.hidden_size != c.num_heads * persist::kHeadDim) return false;
    const int sms = num_sms();
    if (sms < 64 || (sms & 3)) return false;  // statistics rows of ncta words are polled as uint4
    if (c.num_heads * c.max_batch > sms) return false;
    // exact per-CTA tile counts against the kernel's limits and the weight ring this batch size leaves room for
    int dev = 0, smem_limit = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return false;
    persist::Params P;
    memset(&P, 0, sizeof(P));
    P.H = c.hidden_size; P.I = c.intermediate_size; P.ncta = sms;
    persist::Geometry g;
    plan_smem(P, c.max_batch, smem_limit, &g);
    const int ns = g.ring_bytes / (2 * P.H), nsD = (P.I + P.H - 1) / P.H;
    for (int cta = 0; cta < sms; ++cta) {
        auto tiles = [&](int nb, int per) { return (int)((long long)nb * (cta + 1) / sms - (long long)nb * cta / sms) * per; };
        const int tA = tiles(3 * P.H / 32, 2), tC = tiles(P.H / 32, 2), tD1 = tiles(P.I / 16, 2);
        if (tA > persist::kMaxTiles || tD1 > persist::kMaxTiles || tC * 16 > 96) return false;
        if (tA > ns || tD1 > ns || tC > ns || tC * nsD + nsD - 1 > ns) return false;
    }
    if (ns < 2 * persist::kLmRows) return false;
    return true;
}

int persist_create(PersistState** out, const onebit_decoder_config& cfg, const onebit_layer_params* layers, const void* embed,
                   const void* final_norm, const void* lm_head, const float* rope_cos, const float* rope_sin, __half* kcache,
                   __half* vcache, long long* ids, int* pos) {
    *out = nullptr;
    PersistState* S = new (std::nothrow) PersistState();
    ONEBIT_REQUIRE(S, "persist_create: out of host memory");
    Params& P = S->hp;
    memset(&P, 0, sizeof(P));
    const int H = cfg.hidden_size, I = cfg.intermediate_size, L = cfg.num_layers, dt = cfg.param_dtype;
    P.H = H; P.I = I; P.L = L; P.heads = cfg.num_heads; P.V = cfg.vocab_size; P.max_seq = cfg.max_seq_len;
    P.max_batch = cfg.max_batch; P.pdt = dt; P.ncta = num_sms(); P.rms_eps = cfg.rms_eps; P.ln_eps = cfg.ln_eps;
    P.inv_H = 1.0 / (double)H; P.inv_I = 1.0 / (double)I;
    S->ncta = P.ncta;
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    ONEBIT_CUDA_TRY(cudaDeviceGetAttribute(&S->smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // exact per-CTA tile bounds
    for (int c = 0; c < P.ncta; ++c) {
        auto rng = [&](int nb, int& b0, int& b1) { b0 = (int)((long long)nb * c / P.ncta); b1 = (int)((long long)nb * (c + 1) / P.ncta); };
        int b0, b1;
        rng(3 * H / 32, b0, b1);
        if (2 * (b1 - b0) > kMaxTiles) { delete S; return fail(ONEBIT_ERR_INVALID_ARGUMENT, "persist: too many q/k/v tiles per CTA"); }
        rng(I / 16, b0, b1);
        if (2 * (b1 - b0) > kMaxTiles) { delete S; return fail(ONEBIT_ERR_INVALID_ARGUMENT, "persist: too many gate/up tiles per CTA"); }
        rng(H / 32, b0, b1);
        if (32 * (b1 - b0) > 96) { delete S; return fail(ONEBIT_ERR_INVALID_ARGUMENT, "persist: too many o/down rows per CTA"); }
    }
    // ---- static quantiser bounds from the parameter vectors (host, once)
    std::vector<LayerDev> hl(L);
    const double sqH = std::sqrt((double)H);
    for (int l = 0; l < L; ++l) {
        const onebit_layer_params& lp = layers[l];
        LayerDev& d = hl[l];
        const onebit_bitlinear_params* src[7] = {&lp.q, &lp.k, &lp.v, &lp.o, &lp.gate, &lp.up, &lp.down};
        BLDev* dst[7] = {&d.q, &d.k, &d.v, &d.o, &d.gate, &d.up, &d.down};
        for (int i = 0; i < 7; ++i) {
            dst[i]->w = reinterpret_cast<const uint8_t*>(src[i]->weight);
            dst[i]->g = src[i]->weight_scale;
            dst[i]->h = src[i]->input_factor;
        }
        d.ln_in = lp.input_layernorm;
        d.ln_post = lp.post_attention_layernorm;
        std::vector<float> w_in, w_post, h;
        int rc = fetch(w_in, lp.input_layernorm, H, dt); if (rc) { delete S; return rc; }
        rc = fetch(w_post, lp.post_attention_layernorm, H, dt); if (rc) { delete S; return rc; }
        for (int p = 0; p < 3; ++p) {  // |RMSNorm(x)_k| <= sqrt(H) |w_k|
            rc = fetch(h, src[p]->input_factor, H, dt); if (rc) { delete S; return rc; }
            double mx = 0.0;
            for (int k = 0; k < H; ++k) mx = std::max(mx, std::fabs((double)w_in[k] * (double)h[k]));
            d.e_qkv[p] = bound_exp(sqH * mx);
        }
        {   // attention output: convex combination of LayerNorm outputs, |LN(.)_k| <= sqrt(H)
            rc = fetch(h, lp.o.input_factor, H, dt); if (rc) { delete S; return rc; }
            double mx = 0.0;
            for (int k = 0; k < H; ++k) mx = std::max(mx, std::fabs((double)h[k]));
            d.e_o = bound_exp(sqH * mx);
        }
        for (int p = 0; p < 2; ++p) {
            rc = fetch(h, src[4 + p]->input_factor, H, dt); if (rc) { delete S; return rc; }
            double mx = 0.0;
            for (int k = 0; k < H; ++k) mx = std::max(mx, std::fabs((double)w_post[k] * (double)h[k]));
            d.e_gu[p] = bound_exp(sqH * mx);
        }
        {
            rc = fetch(h, lp.down.input_factor, I, dt); if (rc) { delete S; return rc; }
            double mx = 0.0;
            for (int k = 0; k < I; ++k) mx = std::max(mx, std::fabs((double)h[k]));
            d.hmax_down = (float)mx;
        }
    }
    // ---- re-tiled copies of the packed sign matrices (1 bit per element, MMA-fragment order inside 16-row tiles)
    {
        const size_t per_layer_bytes = ((size_t)4 * H * H + (size_t)3 * H * I) / 8;
        cudaError_t e = cudaMalloc(&S->wtiled, per_layer_bytes * L);
        if (e != cudaSuccess) { persist_destroy(S); return fail(ONEBIT_ERR_CUDA, std::string("persist_create: cudaMalloc(weights): ") + cudaGetErrorString(e)); }
        uint8_t* wp = S->wtiled;
        for (int l = 0; l < L; ++l) {
            BLDev* bl[7] = {&hl[l].q, &hl[l].k, &hl[l].v, &hl[l].o, &hl[l].gate, &hl[l].up, &hl[l].down};
            const int Ns[7] = {H, H, H, H, I, I, H}, Ks[7] = {H, H, H, H, H, H, I};
            for (int i = 0; i < 7; ++i) {
                retile_kernel<<<592, 256>>>(bl[i]->w, wp, Ns[i], Ks[i] / 8);
                bl[i]->w = wp;
                wp += (size_t)Ns[i] * Ks[i] / 8;
            }
        }
        ONEBIT_CUDA_TRY(cudaGetLastError());
    }
    {
        std::vector<int> sc((size_t)(L + 1) * 8, 0);
        for (int l = 0; l < L; ++l) {
            int* w = &sc[(size_t)l * 8];
            w[0] = hl[l].e_qkv[0]; w[1] = hl[l].e_qkv[1]; w[2] = hl[l].e_qkv[2]; w[3] = hl[l].e_o;
            w[4] = hl[l].e_gu[0]; w[5] = hl[l].e_gu[1];
            memcpy(&w[6], &hl[l].hmax_down, 4);
        }
        if (cudaMalloc(&S->lscal, sc.size() * 4) != cudaSuccess) { persist_destroy(S); return fail(ONEBIT_ERR_CUDA, "persist_create: cudaMalloc failed"); }
        cudaMemcpy(S->lscal, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice);
        P.lscal = S->lscal;
    }
    // ---- exchange arenas
    const size_t B = kMaxTok, nc = P.ncta;
    size_t o = 0;
    P.o_xA = o; o += round4(3 * B * H);
    P.o_qkv = o; o += round4(B * 3 * H);
    P.o_qst = o; o += round4(B * 3 * nc * 4);
    P.o_xC = o; o += round4(B * H);
    P.o_cst = o; o += round4(B * nc * kStatW);
    P.o_xD1 = o; o += round4(2 * B * H);
    P.o_dst = o; o += round4(B * nc * kStatW);
    P.o_xD2 = o; o += round4(B * I);
    P.o_d2st = o; o += round4(B * nc * kStatW);
    P.per_layer = o;
    size_t t = 0;
    P.o_xfin = t; t += round4(B * (H / 2));
    P.o_amax = t; t += round4(B * nc * 2);
    P.total_words = P.per_layer * L + t;
    auto cleanup = [&]() { persist_destroy(S); };
    for (int s = 0; s < 2; ++s) {
        cudaError_t e = cudaMalloc(&S->xch[s], P.total_words * 4);
        if (e != cudaSuccess) { cleanup(); return fail(ONEBIT_ERR_CUDA, std::string("persist_create: cudaMalloc: ") + cudaGetErrorString(e)); }
        P.xch[s] = S->xch[s];
        // sentinels: digit regions 0x80808080, everything else 0xFFFFFFFF
        fill_words_kernel<<<256, 256>>>(S->xch[s], P.total_words, kSentF);
        for (int l = 0; l < L; ++l) {
            uint32_t* base = S->xch[s] + (size_t)l * P.per_layer;
            fill_words_kernel<<<64, 256>>>(base + P.o_xA, 3 * B * H, kSentD);
            fill_words_kernel<<<64, 256>>>(base + P.o_xC, B * H, kSentD);
            fill_words_kernel<<<64, 256>>>(base + P.o_xD1, 2 * B * H, kSentD);
            fill_words_kernel<<<64, 256>>>(base + P.o_xD2, B * I, kSentD);
        }
    }
    S->trace_words = kTracers * kTracePoints * (L + 2);
    if (cudaMalloc(&S->dlayers, sizeof(LayerDev) * L) != cudaSuccess || cudaMalloc(&S->dp, sizeof(Params)) != cudaSuccess ||
        cudaMalloc(&S->step_counter, 8) != cudaSuccess || cudaMalloc(&S->abort_flag, 4) != cudaSuccess ||
        cudaMalloc(&S->trace, sizeof(unsigned long long) * S->trace_words) != cudaSuccess) {
        cleanup();
        return fail(ONEBIT_ERR_CUDA, "persist_create: cudaMalloc failed");
    }
    cudaMemset(S->step_counter, 0, 8);
    cudaMemset(S->abort_flag, 0, 4);
    cudaMemset(S->trace, 0, sizeof(unsigned long long) * S->trace_words);
    cudaMemcpy(S->dlayers, hl.data(), sizeof(LayerDev) * L, cudaMemcpyHostToDevice);
    P.layers = S->dlayers;
    P.embed = static_cast<const __half*>(embed);
    P.final_norm = final_norm;
    P.lm_head = static_cast<const __half*>(lm_head);
    P.rope_cos = rope_cos; P.rope_sin = rope_sin;
    P.kcache = kcache; P.vcache = vcache;
    P.step_counter = S->step_counter; P.abort_flag = S->abort_flag; P.trace = S->trace;
    {
        const char* e = getenv("ONEBIT_PERSIST_TRACE");  // stage time stamps cost a few microseconds per layer: off by default
        P.trace_on = (e && e[0] == '1') ? 1 : 0;
    }
    P.ids = ids; P.pos = pos;
    cudaMemcpy(S->dp, &P, sizeof(Params), cudaMemcpyHostToDevice);
    ONEBIT_CUDA_TRY(cudaFuncSetAttribute(persist::step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S->smem_limit - 2048));
    {
        Geometry g;
        plan_smem(P, cfg.max_batch, S->smem_limit, &g);
        if (g.ring_bytes < 4 * H) { cleanup(); return fail(ONEBIT_ERR_INVALID_ARGUMENT, "persist_create: model too wide for the shared-memory ring"); }
    }
    ONEBIT_CUDA_TRY(cudaDeviceSynchronize());
    *out = S;
    return ONEBIT_OK;
}

int persist_step(PersistState* S, int batch, const long long* ids_in, float* logits, cudaStream_t s) {
    ONEBIT_REQUIRE(S && batch >= 1 && batch <= kMaxTok && batch <= S->hp.max_batch, "persist_step: bad batch");
    Geometry g;
    plan_smem(S->hp, batch, S->smem_limit, &g);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)S->ncta);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)g.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait on each other's data
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const Params* dp = S->dp;
    ONEBIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, persist::step_kernel, dp, ids_in, logits, batch, g.ring_bytes, g.dbuf_bytes));
    return ONEBIT_OK;
}

int persist_read_trace(PersistState* S, unsigned long long* out, int n) {
    ONEBIT_REQUIRE(S && out && n > 0, "persist_read_trace: bad arguments");
    const int m = std::min(n, S->trace_words);
    ONEBIT_CUDA_TRY(cudaMemcpy(out, S->trace, sizeof(unsigned long long) * m, cudaMemcpyDeviceToHost));
    return m;
}

int persist_abort_flag(PersistState* S, int* out) {
    ONEBIT_REQUIRE(S && out, "persist_abort_flag: bad arguments");
    ONEBIT_CUDA_TRY(cudaMemcpy(out, S->abort_flag, 4, cudaMemcpyDeviceToHost));
    return ONEBIT_OK;
}

void persist_destroy(PersistState* S) {
    if (!S) return;
    cudaFree(S->xch[0]);
    cudaFree(S->xch[1]);
    cudaFree(S->dlayers);
    cudaFree(S->lscal);
    cudaFree(S->wtiled);
    cudaFree(S->dp);
    cudaFree(S->step_counter);
    cudaFree(S->abort_flag);
    cudaFree(S->trace);
    delete S;
}

}  // namespace onebit
