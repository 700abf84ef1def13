// Large-batch (prefill) variant of the 1-bit linear layer on the 5th-generation tensor cores:
//     t[m][n] = g[n] * sum_k s(n,k) * h[k] * x[m][k]                      (bitnet.py:113-116)
// computed as D[n][m] = A[n][:] . B[m][:] with tcgen05.mma (kind::f16, fp32 accumulators in TMEM):
//   A (M side, 2 x 128 rows per CTA) = the sign matrix with h folded in: A[n][k] = bit ? -h[k] : h[k], built in
//       registers straight from the 1-bit layout and written once per K chunk to shared memory in the UMMA
//       SWIZZLE_128B K-major layout. Sign application costs ~1.1 op/weight and needs no shift:
//       (byte * (0x40008000 >> 2i)) & 0x80008000 puts bits 2i / 2i+1 of a weight byte on the sign positions
//       of an fp16 pair, and the same LOP3 XORs them onto the (h[k], h[k+1]) pair.
//   B (N side, 256 tokens per CTA) = x itself (fp16), TMA-loaded (cp.async.bulk.tensor, 128B swizzle), K-major as
//       stored. h is folded into A, so x needs no pre-multiplication pass and h*x is formed exactly (fp16 x fp16
//       products are exact in the fp32 accumulator).
// One expanded weight tile (256 x 64 fp16 = 32 KB from 2 KB of packed bits) is reused across the 256 tokens of
// the CTA, so the expansion (the reason tcgen05 is NOT used for decode, see imma_gemv.cuh) is amortised:
// 2 x 4 MMAs of 128x256x16 per chunk = 1024 tensor cycles vs ~170 ALU instructions per expander thread.
// B traffic: 32 KB per 1024 cycles per SM = 4.7 KB/clk chip-wide, under the ~6.3 KB/clk L2 limit (a 128-row CTA
// tile would need 9.5 KB/clk).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = sign expanders, then epilogue (tcgen05.ld -> * g -> coalesced fp32 stores).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace onebit {
namespace {

// Two tile configurations of the same kernel:
//   prefill   HALVES = 2 (256 weight rows per CTA), TM = 256 tokens, 3 stages  — one expanded tile serves 256 tokens;
//   decode    HALVES = 1 (128 weight rows per CTA), TM = 64 tokens, 2 stages, optional split-K over grid.z — batches of
//             5..64 sequences: many small CTAs (4 per SM, 53 KB each) so that all 148 SMs stream and expand weights.
// DENSE = true replaces the sign expanders by a second TMA stream of a dense fp16 matrix (lm_head for those batches).
constexpr int kChunkK = 64;     // K columns per pipeline stage (= one 128-byte swizzle row of fp16)
constexpr int kThreads = 320;    // warp 0 TMA, warp 1 MMA, warps 2-9 sign expanders (warps 2-5 also run the epilogue)
constexpr int kExpanders = 256;
constexpr int kMaxProblems = 3; // projections that share the activation tile in one launch (q/k/v, gate/up)
template <int HALVES, int TM>
struct Tile {
    static constexpr int kTileN = HALVES * 128;
    static constexpr int kStages = HALVES == 2 ? 3 : 2;  // decode: shallow pipeline, 4 CTAs per SM instead
    static constexpr int kHStage = HALVES == 2 ? 1 : 256;  // uint4 slots of shared memory for the CTA's slice of input_factor (decode)
    static constexpr int kABytes = kTileN * kChunkK * 2;
    static constexpr int kBBytes = TM * kChunkK * 2;
    static constexpr int kSlabChunks = 8;                                  // K chunks per weight slab
    static constexpr int kSlabBytes = kTileN * kSlabChunks * 8;            // [rows][64 B of packed signs]
    static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 2 * kSlabBytes + 1024;  // + alignment slack
    static constexpr int kTmemCols = HALVES * TM < 32 ? 32 : HALVES * TM;       // power of two for these configurations
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start >> 4 | LBO (unused for one swizzle atom along K) |
// SBO = 1024 B between 8-row groups | version 1 (sm_100) | layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = F16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N, bool bf16) {
    const uint32_t ab = bf16 ? ((1u << 7) | (1u << 10)) : 0u;  // A / B element format: 0 = F16, 1 = BF16
    return (1u << 4) | ab | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint2 ldg_nc_u2(const uint8_t* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

#ifdef ONEBIT_TC5_TRACE
__device__ unsigned long long g_tc5_trace[512];
__device__ __forceinline__ unsigned long long tc5_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TC5_STAMP(cond, slot) do { if ((cond) && (slot) < 512) g_tc5_trace[(slot)] = tc5_now(); } while (0)
#else
#define TC5_STAMP(cond, slot) do { } while (0)
#endif

struct Tc5Problem {
    const uint8_t* w;   // [N][K/8] packed signs (unused by the DENSE variant)
    const __half* h;    // [K] fp16 input_factor
    const void* g;      // [N] TP weight_scale or nullptr
    float* t;           // [ksplit][M][N] fp32 (split z writes its partial sums at t + z * M * N)
    int N, tile_begin;  // rows; first row tile (blockIdx.x) of this problem
    int ldt;            // leading dimension of t (>= N; padded K of a row-parallel consumer under tensor parallelism)
};
struct PrefillArgs {
    Tc5Problem p[kMaxProblems];
    int nprob, M, K, ksplit;
    int wtma;  // packed signs arrive through TMA slabs (K % 128 == 0); else the expanders load them themselves
    int bf16;  // activations and the input_factor copy are bfloat16 (same bit tricks: the sign is bit 15 of either format)
};

template <typename TP, int HALVES, int TM, bool DENSE>
__global__ void __launch_bounds__(kThreads, HALVES == 2 ? 1 : 3)  // decode tile: 3 CTAs per SM (<= 68 registers)
prefill_tc5_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap amap,
                   const __grid_constant__ CUtensorMap wmap1, const __grid_constant__ CUtensorMap wmap2,
                   const __grid_constant__ PrefillArgs A) {
    using TL = Tile<HALVES, TM>;
    constexpr int kTileN = TL::kTileN, kTileM = TM, kStages = TL::kStages, kABytes = TL::kABytes, kBBytes = TL::kBBytes;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                          // [stage][HALVES * 128 rows][128 B] swizzled
    unsigned char* sB = smem + kStages * kABytes;      // [stage][TM rows][128 B] swizzled (TMA)
    unsigned char* sW = sB + kStages * kBBytes;        // [2][HALVES * 128 rows][64 B] packed-sign slabs of 8 chunks (TMA)
    __shared__ __align__(8) uint64_t full_a[kStages], full_b[kStages], empty[kStages], tmem_full, wfull[2], wempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint4 h_stage[TL::kHStage];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool tr0 = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    TC5_STAMP(tr0 && threadIdx.x == 0, 0);
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kMaxProblems; ++i)
        if (i < A.nprob && (int)blockIdx.x >= A.p[i].tile_begin) pi = i;
    const Tc5Problem& P = A.p[pi];
    const int n0 = ((int)blockIdx.x - P.tile_begin) * kTileN, m0 = blockIdx.y * kTileM;
    const int nchunks_all = A.K / kChunkK;
    // this CTA's K slice; slices start on even chunks (16-byte aligned packed-sign columns for the TMA slabs)
    const int half_all = nchunks_all >> 1;
    const int c_begin = 2 * (int)(((long long)half_all * blockIdx.z) / A.ksplit);
    const int c_end = (int)blockIdx.z + 1 == A.ksplit ? nchunks_all : 2 * (int)(((long long)half_all * (blockIdx.z + 1)) / A.ksplit);
    const int nchunks = c_end - c_begin;
    float* tout = P.t + (size_t)blockIdx.z * A.M * P.ldt;
    const int m_valid = min(kTileM, A.M - m0);
    const int umma_n = max(16, (m_valid + 15) & ~15);  // tokens covered by the MMA (multiple of 16)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_a[s], DENSE ? 1 : kExpanders / 32);  // one arrival per expander warp
            mbar_init(&full_b[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tmem_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], kExpanders / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    }
    if (warp == 1) {  // TMEM: HALVES accumulators of 128 lanes x TM fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TL::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    TC5_STAMP(tr0 && threadIdx.x == 0, 1);

    if (warp == 0) {
        // ===== TMA producer: x tile [256 tokens][64 k] per stage =====
        if (lane == 0) {
            // packed signs: 64-byte-wide slabs (8 chunks) of the CTA's rows, two slabs ahead of the expanders. They come
            // through TMA because a global load in flight in an expander thread would be waited for by the MEMBAR that
            // fence.proxy.async implies: one HBM round trip per chunk (measured 0.75 us per chunk, tools/trace_tc5.py).
            const CUtensorMap* wm = pi == 0 ? &amap : (pi == 1 ? &wmap1 : &wmap2);
            const int nslabs = (nchunks + TL::kSlabChunks - 1) / TL::kSlabChunks;
            auto issue_slab = [&](int j) {
                if (DENSE || !A.wtma || j >= nslabs) return;
                const int b = j & 1;
                if (j >= 2) mbar_wait(&wempty[b], ((j >> 1) - 1) & 1);
                mbar_expect_tx(&wfull[b], TL::kSlabBytes);
                tma_load_2d(sW + b * TL::kSlabBytes, wm, (c_begin + j * TL::kSlabChunks) * 8, n0, &wfull[b]);
            };
            issue_slab(0);
            issue_slab(1);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kStages, it = c / kStages;
                if (c > 0 && (c % TL::kSlabChunks) == 0) issue_slab(c / TL::kSlabChunks + 1);
                if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
                mbar_expect_tx(&full_b[s], kBBytes);
                tma_load_2d(sB + s * kBBytes, &xmap, (c_begin + c) * kChunkK, m0, &full_b[s]);
                if (DENSE) {  // the A tile is a plain fp16 matrix: second TMA stream, same 128B-swizzled layout
                    mbar_expect_tx(&full_a[s], kABytes);
                    tma_load_2d(sA + s * kABytes, &amap, (c_begin + c) * kChunkK, n0, &full_a[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, umma_n, A.bf16 != 0);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kStages, ph = (c / kStages) & 1;
                mbar_wait(&full_a[s], ph);
                mbar_wait(&full_b[s], ph);
                TC5_STAMP(tr0 && c < 40, 18 + 4 * c);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(sA + s * kABytes), b_addr = smem_u32(sB + s * kBBytes);
#pragma unroll
                for (int half = 0; half < HALVES; ++half) {
#pragma unroll
                    for (int k = 0; k < kChunkK / 16; ++k) {
                        const uint64_t ad = umma_desc_sw128(a_addr + half * (128 * 128) + k * 32);
                        const uint64_t bd = umma_desc_sw128(b_addr + k * 32);
                        umma_f16(tmem_base + half * TM, ad, bd, idesc, (c > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);  // frees the stage when these MMAs have read it
                TC5_STAMP(tr0 && c < 40, 19 + 4 * c);
            }
            umma_commit(&tmem_full);
        }
    } else {
        // ===== sign expanders: 256 threads. Prefill tile (256 rows): thread e2 owns row e2, all 64 columns of a chunk.
        // Decode tile (128 rows): threads e2 and e2 + 128 share row e2 & 127, 32 columns each (a lone warp per scheduler
        // runs the ~100 dependent instructions of a row-chunk at ~6 cycles each: more threads, shorter chains) =====
        const int e2 = threadIdx.x - 64;
        const int e = HALVES == 2 ? e2 : (e2 & 127);           // row of the CTA tile
        const int q_lo = HALVES == 2 ? 0 : 4 * (e2 >> 7);      // first 16-byte group (8 columns each) of the chunk
        constexpr int kQ = HALVES == 2 ? 8 : 4;                // groups per thread and chunk
        const int Kb = A.K >> 3;
        const uint8_t* wrow0 = P.w + (size_t)min(n0 + e, P.N - 1) * Kb + (size_t)c_begin * 8;
        const bool h_staged = !DENSE && TL::kHStage > 1 && nchunks * 8 <= TL::kHStage;
        if (h_staged) {  // the CTA's K slice of input_factor: one cooperative copy instead of 8 L1/L2 round trips per chunk
            const uint4* hsrc = reinterpret_cast<const uint4*>(P.h + (size_t)c_begin * kChunkK);
            for (int i = e2; i < nchunks * 8; i += kExpanders) h_stage[i] = __ldg(hsrc + i);
            asm volatile("bar.sync 1, %0;" ::"n"(kExpanders) : "memory");
        }
        constexpr int kPF = 8;  // = TL::kSlabChunks: chunks per slab / per register refill
        uint2 q0[kPF];
#pragma unroll
        for (int i = 0; i < kPF; ++i) {
            q0[i] = make_uint2(0u, 0u);
            if (!DENSE && !A.wtma && i < nchunks) q0[i] = ldg_nc_u2(wrow0 + i * 8);  // legacy path (K % 128 != 0): own loads
        }
        for (int cb = 0; !DENSE && cb < nchunks; cb += kPF) {
          if (A.wtma) {  // this thread's 64 bytes (8 chunks) of rows e / e + 128 from the slab, then the slab is free again
            const int j = cb / kPF, b = j & 1;
            mbar_wait(&wfull[b], (j >> 1) & 1);
            const uint4* rp = reinterpret_cast<const uint4*>(sW + b * TL::kSlabBytes + e * 64);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = rp[i];
                q0[2 * i] = make_uint2(v.x, v.y);
                q0[2 * i + 1] = make_uint2(v.z, v.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&wempty[b]);
          }
#pragma unroll
          for (int ci = 0; ci < kPF; ++ci) {
            const int c = cb + ci;
            if (c >= nchunks) break;
            const int s = c % kStages, it = c / kStages;
            const uint2 w = q0[ci];
            if (!A.wtma && c + kPF < nchunks) q0[ci] = ldg_nc_u2(wrow0 + (c + kPF) * 8);  // refill with the chunk kPF ahead
            // 8 x 16 B of input_factor, the same for every row: from the staged slice (decode) or through L1 (prefill)
            const uint4* hp = h_staged ? h_stage + c * 8 : reinterpret_cast<const uint4*>(P.h + (size_t)(c_begin + c) * kChunkK);
            if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
            TC5_STAMP(tr0 && e2 == 0 && c < 40, 16 + 4 * c);
            unsigned char* base = sA + s * kABytes;
            {
                const int r = e;
                unsigned char* rowp = base + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
                for (int qi = 0; qi < kQ; ++qi) {  // 16-byte group q = columns 8q .. 8q+7 = byte q of the 64-bit word
                    const int q = q_lo + qi;
                    const uint4 hv = h_staged ? hp[q] : __ldg(hp + q);
                    // byte q of the row chunk: bit i = column 8q + i. An 8-bit value times (0x40008000 >> 2i) is two
                    // disjoint shifted copies (no carries): bit 2i lands on bit 15, bit 2i+1 on bit 31.
                    const uint32_t b8 = (((q < 4 ? w.x : w.y) >> (8 * (q & 3))) & 0xFFu);
                    uint4 o;
                    o.x = ((b8 * 0x40008000u) & 0x80008000u) ^ hv.x;
                    o.y = ((b8 * 0x10002000u) & 0x80008000u) ^ hv.y;
                    o.z = ((b8 * 0x04000800u) & 0x80008000u) ^ hv.z;
                    o.w = ((b8 * 0x01000200u) & 0x80008000u) ^ hv.w;
                    *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) << 4)) = o;  // SWIZZLE_128B: chunk ^= row % 8
                }
            }
            TC5_STAMP(tr0 && e2 == 0 && c < 40, 256 + 2 * c);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
            TC5_STAMP(tr0 && e2 == 0 && c < 40, 257 + 2 * c);
            __syncwarp();  // one arrival per warp: 256 arrivals on one shared-memory word per chunk serialise
            if (lane == 0) mbar_arrive(&full_a[s]);
            TC5_STAMP(tr0 && e2 == 0 && c < 40, 17 + 4 * c);
          }
        }
        // ===== epilogue: TMEM -> registers -> * g -> t[m][n] (lanes = consecutive n: coalesced) =====
        mbar_wait(&tmem_full, 0);
        TC5_STAMP(tr0 && threadIdx.x == 64, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;  // two warps per quarter: they split the halves (prefill) or the token columns (decode)
#pragma unroll 1
        for (int half = (HALVES == 2 ? grp : 0); half < (HALVES == 2 ? grp + 1 : 1); ++half) {
            const int n = n0 + half * 128 + quarter * 32 + lane;
            const float gs = (P.g != nullptr && n < P.N) ? to_f32(static_cast<const TP*>(P.g)[n]) : 1.f;
#pragma unroll 1
            for (int cb = (HALVES == 2 ? 0 : 32 * grp); cb < umma_n; cb += (HALVES == 2 ? 32 : 64)) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * TM + cb);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
                    "%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (n < P.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int m = m0 + cb + j;
                        if (m < A.M) tout[(size_t)m * P.ldt + n] = __uint_as_float(v[j]) * gs;
                    }
                }
            }
        }
    }
    TC5_STAMP(tr0 && threadIdx.x == 64, 3);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC5_STAMP(tr0 && threadIdx.x == 0, 4);
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TL::kTmemCols) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Decode tile with the A operand in TENSOR MEMORY ("TS" form of tcgen05.mma): the expanders write the sign-expanded
// (+-h) tile straight from registers into TMEM with tcgen05.st — no shared-memory copy of A, and above all no
// generic->async proxy fence per chunk (fence.proxy.async = MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, which with several CTAs per SM
// waits behind the other CTAs' TMA traffic: measured 0.64 us per chunk per SM with 3 CTAs against 0.35 us for a lone CTA).
// Layout: 128 weight rows <-> 128 TMEM lanes (a warp reaches the 32 lanes of its quarter, warp % 4), a 64-column K chunk
// <-> 32 TMEM columns (two fp16 per column, K-major); the MMA of K step k reads columns [8k, 8k + 8) of the stage.
// TMEM plan (128 columns per CTA): accumulator 64 | two A stages of 32.
constexpr int kTsStages = 2;
constexpr int kTsTM = 64;
constexpr int kTsSlabBytes = 128 * 8 * 8;
constexpr int kTsBBytes = kTsTM * kChunkK * 2;
constexpr int kTsSmemBytes = kTsStages * kTsBBytes + 2 * kTsSlabBytes + 1024;

__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

template <typename TP>
__global__ void __launch_bounds__(kThreads, 3)
tc5_decode_ts_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap0,
                     const __grid_constant__ CUtensorMap wmap1, const __grid_constant__ CUtensorMap wmap2,
                     const __grid_constant__ PrefillArgs A) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = smem;                              // [stage][64 tokens][128 B] swizzled (TMA)
    unsigned char* sW = sB + kTsStages * kTsBBytes;        // [2][128 rows][64 B] packed-sign slabs of 8 chunks (TMA)
    __shared__ __align__(8) uint64_t full_a[kTsStages], full_b[kTsStages], empty[kTsStages], tmem_full, wfull[2], wempty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ uint4 h_stage[256];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kMaxProblems; ++i)
        if (i < A.nprob && (int)blockIdx.x >= A.p[i].tile_begin) pi = i;
    const Tc5Problem& P = A.p[pi];
    const int n0 = ((int)blockIdx.x - P.tile_begin) * 128;
    const int nchunks_all = A.K / kChunkK, half_all = nchunks_all >> 1;
    const int c_begin = 2 * (int)(((long long)half_all * blockIdx.z) / A.ksplit);
    const int c_end = (int)blockIdx.z + 1 == A.ksplit ? nchunks_all : 2 * (int)(((long long)half_all * (blockIdx.z + 1)) / A.ksplit);
    const int nchunks = c_end - c_begin;
    float* tout = P.t + (size_t)blockIdx.z * A.M * P.ldt;
    const int umma_n = max(16, (min(kTsTM, A.M) + 15) & ~15);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTsStages; ++s) {
            mbar_init(&full_a[s], kExpanders / 32);
            mbar_init(&full_b[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&tmem_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], kExpanders / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s, tmem_a0 = tmem_base + 64u;  // accumulator: columns 0..63; A stages: 64.., 96..

    if (warp == 0) {
        if (lane == 0) {
            const CUtensorMap* wm = pi == 0 ? &wmap0 : (pi == 1 ? &wmap1 : &wmap2);
            const int nslabs = (nchunks + 7) / 8;
            auto issue_slab = [&](int j) {
                if (j >= nslabs) return;
                const int b = j & 1;
                if (j >= 2) mbar_wait(&wempty[b], ((j >> 1) - 1) & 1);
                mbar_expect_tx(&wfull[b], kTsSlabBytes);
                tma_load_2d(sW + b * kTsSlabBytes, wm, (c_begin + j * 8) * 8, n0, &wfull[b]);
            };
            issue_slab(0);
            issue_slab(1);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kTsStages, it = c / kTsStages;
                if (c > 0 && (c & 7) == 0) issue_slab(c / 8 + 1);
                if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
                mbar_expect_tx(&full_b[s], kTsBBytes);
                tma_load_2d(sB + s * kTsBBytes, &xmap, (c_begin + c) * kChunkK, 0, &full_b[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, umma_n, A.bf16 != 0);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % kTsStages, ph = (c / kTsStages) & 1;
                mbar_wait(&full_a[s], ph);
                mbar_wait(&full_b[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t b_addr = smem_u32(sB + s * kTsBBytes);
#pragma unroll
                for (int k = 0; k < kChunkK / 16; ++k)
                    umma_f16_ts(tmem_base, tmem_a0 + (uint32_t)(s * 32 + k * 8), umma_desc_sw128(b_addr + k * 32), idesc,
                                (c > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty[s]);
            }
            umma_commit(&tmem_full);
        }
    } else {
        // ===== expanders: warps 2..9; warp w reaches TMEM lanes (w % 4) * 32 .. + 31 = its weight rows; warps w and w + 4 share
        // the rows and split a chunk's 64 columns (32 each = 16 TMEM columns)
        const int quarter = warp & 3, qhalf = (warp - 2) >> 2;
        const int e2 = threadIdx.x - 64;
        const int row = quarter * 32 + lane;
        const bool h_staged = nchunks * 8 <= 256;
        if (h_staged) {
            const uint4* hsrc = reinterpret_cast<const uint4*>(P.h + (size_t)c_begin * kChunkK);
            for (int i = e2; i < nchunks * 8; i += kExpanders) h_stage[i] = __ldg(hsrc + i);
            asm volatile("bar.sync 1, %0;" ::"n"(kExpanders) : "memory");
        }
        uint2 q0[8];
        for (int cb = 0; cb < nchunks; cb += 8) {
            {
                const int j = cb >> 3, b = j & 1;
                mbar_wait(&wfull[b], (j >> 1) & 1);
                const uint4* rp = reinterpret_cast<const uint4*>(sW + b * kTsSlabBytes + row * 64);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 v = rp[i];
                    q0[2 * i] = make_uint2(v.x, v.y);
                    q0[2 * i + 1] = make_uint2(v.z, v.w);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&wempty[b]);
            }
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
                const int c = cb + ci;
                if (c >= nchunks) break;
                const int s = c % kTsStages, it = c / kTsStages;
                const uint32_t wbits = qhalf ? q0[ci].y : q0[ci].x;  // this thread's 32 columns of the chunk
                const uint4* hp = (h_staged ? h_stage + c * 8 : reinterpret_cast<const uint4*>(P.h + (size_t)(c_begin + c) * kChunkK)) + 4 * qhalf;
                uint32_t o[16];
#pragma unroll
                for (int qi = 0; qi < 4; ++qi) {
                    const uint4 hv = h_staged ? hp[qi] : __ldg(hp + qi);
                    const uint32_t b8 = (wbits >> (8 * qi)) & 0xFFu;
                    o[4 * qi + 0] = ((b8 * 0x40008000u) & 0x80008000u) ^ hv.x;
                    o[4 * qi + 1] = ((b8 * 0x10002000u) & 0x80008000u) ^ hv.y;
                    o[4 * qi + 2] = ((b8 * 0x04000800u) & 0x80008000u) ^ hv.z;
                    o[4 * qi + 3] = ((b8 * 0x01000200u) & 0x80008000u) ^ hv.w;
                }
                if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);  // the MMAs that read this TMEM stage have completed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_a0 + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * 32 + qhalf * 16);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                    "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]), "r"(o[9]),
                    "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15])
                    : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_a[s]);
            }
        }
        // ===== epilogue (warps 2..5 take token columns 0..31, warps 6..9 columns 32..63)
        mbar_wait(&tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int n = n0 + row;
        const float gs = (P.g != nullptr && n < P.N) ? to_f32(static_cast<const TP*>(P.g)[n]) : 1.f;
        const int cb0 = 32 * qhalf;
        if (cb0 < umma_n) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cb0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
                "%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n < P.N) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int m = cb0 + j;
                    if (m < A.M) tout[(size_t)m * P.ldt + n] = __uint_as_float(v[j]) * gs;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
}

// x (bf16 / fp32) -> fp16 scratch, so that the TMA / MMA B operand is always fp16
template <typename TX>
__global__ void to_half_kernel(const TX* __restrict__ x, __half* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2half_rn(to_f32(x[i]));
}

template <typename TX>
__global__ void to_bf16_kernel(const TX* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2bfloat16_rn(to_f32(x[i]));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_2d(CUtensorMap* map, const __half* base, int64_t rows, int64_t k, int box_rows, bool bf16 = false) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return fail(ONEBIT_ERR_CUDA, "tcgen05 path: cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(ONEBIT_ERR_CUDA, "tcgen05 path: cuTensorMapEncodeTiled failed, code " + std::to_string((int)cr));
    return ONEBIT_OK;
}

// packed signs [N][K/8] as a 2-D byte tensor: boxes of 64 bytes (8 K chunks) x box_rows rows, no swizzle
int encode_w(CUtensorMap* map, const uint8_t* base, int64_t rows, int64_t kb, int box_rows) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return fail(ONEBIT_ERR_CUDA, "tcgen05 path: cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t dims[2] = {(cuuint64_t)kb, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kb};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(ONEBIT_ERR_CUDA, "tcgen05 path: cuTensorMapEncodeTiled (weights) failed, code " + std::to_string((int)cr));
    return ONEBIT_OK;
}

template <typename TP, int HALVES, int TM, bool DENSE>
int launch_inst(const CUtensorMap& xmap, const CUtensorMap* wm, const PrefillArgs& a, dim3 grid, cudaStream_t s) {
    auto kern = prefill_tc5_kernel<TP, HALVES, TM, DENSE>;
    static bool configured[64] = {false};
    int dev = 0;
    ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tile<HALVES, TM>::kSmemBytes));
        ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[dev] = true;
    }
    kern<<<grid, kThreads, Tile<HALVES, TM>::kSmemBytes, s>>>(xmap, wm[0], wm[1], wm[2], a);
    ONEBIT_CUDA_TRY(cudaGetLastError());
    return ONEBIT_OK;
}

}  // namespace

bool prefill_tc5_supported(int64_t m, int64_t k, int64_t n) {
    return m >= 1 && k % kChunkK == 0 && k >= kChunkK && n >= 1 && m < (1ll << 31) && n < (1ll << 31);
}

size_t prefill_tc5_workspace_bytes(int64_t m, int64_t k, int act_dtype, int param_dtype) {
    size_t b = 0;
    if (act_dtype != ONEBIT_F16) b += (((size_t)m * k * 2) + 255) & ~(size_t)255;
    b += (((size_t)k * 2) + 255) & ~(size_t)255;  // input_factor copy in the activation-side 16-bit format
    return b + 256;
}

int launch_to_half(const void* src, __half* dst, int64_t n, int dtype, cudaStream_t s) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (dtype == ONEBIT_BF16) to_half_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src), dst, n);
    else if (dtype == ONEBIT_F32) to_half_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(src), dst, n);
    else return fail(ONEBIT_ERR_INVALID_ARGUMENT, "launch_to_half: source is already fp16");
    ONEBIT_CUDA_TRY(cudaGetLastError());
    return ONEBIT_OK;
}

int launch_prefill_tc5(const void* x, const int8_t* w, const void* g, const void* h, float* t, int64_t m, int64_t k,
                       int64_t n, int act_dtype, int param_dtype, bool scale_by_g, void* workspace, cudaStream_t s) {
    ONEBIT_REQUIRE(prefill_tc5_supported(m, k, n), "prefill_tc5: needs K % 64 == 0");
    char* ws = static_cast<char*>(workspace);
    // bfloat16 activations stay bfloat16 (tcgen05 kind::f16 takes BF16 operands; a conversion to fp16 would overflow beyond
    // 65504 and flush small values — ADVICE r01). fp32 activations are rounded to fp16: values beyond the fp16 range saturate
    // (documented in onebit_b200.h; the bit-plane GEMV of M <= 8 keeps 23 bits).
    const bool bf = act_dtype == ONEBIT_BF16;
    const __half* x16 = static_cast<const __half*>(x);
    if (act_dtype == ONEBIT_F32) {
        __half* buf = reinterpret_cast<__half*>(ws);
        ws += (((size_t)m * k * 2) + 255) & ~(size_t)255;
        const int rc = launch_to_half(x, buf, m * k, act_dtype, s);
        if (rc) return rc;
        x16 = buf;
    } else if (bf) {
        ws += (((size_t)m * k * 2) + 255) & ~(size_t)255;  // (slot reserved by prefill_tc5_workspace_bytes, unused)
    }
    const __half* h16 = static_cast<const __half*>(h);
    if (bf && param_dtype != ONEBIT_BF16) {  // input_factor as bfloat16 next to bfloat16 activations
        __nv_bfloat16* buf = reinterpret_cast<__nv_bfloat16*>(ws);
        const unsigned grid = (unsigned)((k + 255) / 256);
        if (param_dtype == ONEBIT_F16) to_bf16_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half*>(h), buf, k);
        else to_bf16_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(h), buf, k);
        ONEBIT_CUDA_TRY(cudaGetLastError());
        h16 = reinterpret_cast<const __half*>(buf);
    } else if (!bf && param_dtype != ONEBIT_F16) {
        __half* buf = reinterpret_cast<__half*>(ws);
        const int rc = launch_to_half(h, buf, k, param_dtype, s);
        if (rc) return rc;
        h16 = buf;
    }
    Tc5Launch L = {};
    L.bf16 = bf ? 1 : 0;
    L.x16 = x16; L.M = (int)m; L.K = (int)k; L.nprob = 1; L.ksplit = 1; L.param_dtype = param_dtype;
    L.p[0].w = w; L.p[0].h16 = h16; L.p[0].g = scale_by_g ? g : nullptr; L.p[0].t = t; L.p[0].N = (int)n;
    return launch_tc5(L, s);
}

// Shared launcher: `nprob` projections over the same fp16 activations [M][K]; each writes ksplit partial outputs
// [ksplit][M][N] (ksplit = 1: the final t). M <= 64 selects the decode tile configuration.
int launch_tc5(const Tc5Launch& L, cudaStream_t s) {
    ONEBIT_REQUIRE(L.nprob >= 1 && L.nprob <= kMaxProblems && L.M >= 1 && L.K % kChunkK == 0 && L.ksplit >= 1 &&
                       L.K / kChunkK >= L.ksplit, "launch_tc5: bad arguments");
    const bool small = L.M <= 64;
    const int tile_n = small ? 128 : 256, tile_m = small ? 64 : 256;
    ONEBIT_REQUIRE(small || L.ksplit == 1, "launch_tc5: split-K is built for the decode tile configuration only");
    CUtensorMap xmap;
    int rc = encode_2d(&xmap, L.x16, L.M, L.K, tile_m, L.bf16 != 0);
    if (rc) return rc;
    PrefillArgs a = {};
    a.nprob = L.nprob; a.M = L.M; a.K = L.K; a.ksplit = L.ksplit; a.bf16 = L.bf16;
    a.wtma = (L.K % 128 == 0) ? 1 : 0;
    ONEBIT_REQUIRE(L.ksplit == 1 || L.K / kChunkK / 2 >= L.ksplit, "launch_tc5: K too small for this split");
    CUtensorMap wm[3];
    int tiles = 0;
    for (int i = 0; i < L.nprob; ++i) {
        a.p[i].w = reinterpret_cast<const uint8_t*>(L.p[i].w); a.p[i].h = L.p[i].h16; a.p[i].g = L.p[i].g; a.p[i].t = L.p[i].t;
        a.p[i].N = L.p[i].N; a.p[i].tile_begin = tiles; a.p[i].ldt = L.p[i].ldt > 0 ? L.p[i].ldt : L.p[i].N;
        tiles += (L.p[i].N + tile_n - 1) / tile_n;
        if (!aligned16(L.p[i].w)) a.wtma = 0;
    }
    for (int i = 0; i < 3; ++i) {
        if (a.wtma && i < L.nprob) {
            rc = encode_w(&wm[i], a.p[i].w, a.p[i].N, L.K / 8, tile_n);
            if (rc) return rc;
        } else {
            wm[i] = xmap;  // unused slot
        }
    }
    dim3 grid((unsigned)tiles, (unsigned)((L.M + tile_m - 1) / tile_m), (unsigned)L.ksplit);
    ONEBIT_REQUIRE(grid.y <= 65535, "prefill_tc5: M too large (max 16.7M tokens)");
    static int use_ts = -1;  // decode tile: A operand in tensor memory (tc5_decode_ts_kernel); ONEBIT_TC5_TS=0 keeps the shared-memory form
    if (use_ts < 0) {
        const char* e = getenv("ONEBIT_TC5_TS");
        use_ts = (e && e[0] == '0') ? 0 : 1;
    }
    return dispatch_dtype(L.param_dtype, [&](auto pt) {
        using TP = decltype(pt);
        if (small && a.wtma && use_ts) {
            auto kern = tc5_decode_ts_kernel<TP>;
            static bool configured[64] = {false};
            int dev = 0;
            ONEBIT_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 0 && dev < 64 && !configured[dev]) {
                ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes));
                ONEBIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                configured[dev] = true;
            }
            kern<<<grid, kThreads, kTsSmemBytes, s>>>(xmap, wm[0], wm[1], wm[2], a);
            ONEBIT_CUDA_TRY(cudaGetLastError());
            return (int)ONEBIT_OK;
        }
        return small ? launch_inst<TP, 1, 64, false>(xmap, wm, a, grid, s) : launch_inst<TP, 2, 256, false>(xmap, wm, a, grid, s);
    });
}

// out[m][n] = sum_k W[n][k] * x[m][k]: dense fp16 weights through the same pipeline (lm_head for decode batches > 8)
int launch_dense_tc5(const __half* x16, const __half* w16, float* out, int64_t m, int64_t k, int64_t n, cudaStream_t s) {
    ONEBIT_REQUIRE(m >= 1 && m <= 64 && k % kChunkK == 0 && n >= 1, "launch_dense_tc5: needs 1 <= M <= 64, K % 64 == 0");
    CUtensorMap xmap, amap;
    int rc = encode_2d(&xmap, x16, m, k, 64);
    if (rc) return rc;
    rc = encode_2d(&amap, w16, n, k, 128);
    if (rc) return rc;
    PrefillArgs a = {};
    a.nprob = 1; a.M = (int)m; a.K = (int)k; a.ksplit = 1;
    a.p[0].t = out; a.p[0].N = (int)n; a.p[0].tile_begin = 0; a.p[0].ldt = (int)n;
    dim3 grid((unsigned)((n + 127) / 128), 1, 1);
    const CUtensorMap wm[3] = {amap, xmap, xmap};
    return launch_inst<__half, 1, 64, true>(xmap, wm, a, grid, s);
}

}  // namespace onebit

// Debug only (side builds with -DONEBIT_TC5_TRACE): %globaltimer stamps of CTA (0,0,0) of the LAST tcgen05 launch:
// [0] kernel start, [1] setup done, [2] accumulators complete, [3] epilogue done, [4] CTA end; per K chunk c < 40 at
// 16 + 4c: expander stage free / expander done / MMA inputs ready / MMA issued.
extern "C" __attribute__((visibility("default"))) int onebit_debug_tc5_trace(unsigned long long* out512) {
#ifdef ONEBIT_TC5_TRACE
    return cudaMemcpyFromSymbol(out512, onebit::g_tc5_trace, sizeof(unsigned long long) * 512) == cudaSuccess ? 0 : -2;
#else
    (void)out512;
    return -1;
#endif
}
