// Scoring on the 5th-generation tensor cores: U_tile x Q^T with split-TF32 (3 MMAs per product: hi*hi +
// hi*lo + lo*hi, fp32 accumulation in TMEM), TMA-staged 128-byte-swizzled operand tiles, and the known-item
// mask + per-user top-k fused into the TMEM epilogue -- the score matrix is never written to HBM.
//
// Stands in for the predict-all + sort + head of RecommenderBase.recommend (recommender_base.py:245-266)
// for the linear / sigmoid rank key  b_i + p.q  and the rbf rank key  2 p.q - |q|^2  (SURVEY.md 9.1 item 16).
//
// One CTA (192 threads) owns 128 users and walks over all item tiles of 128 items:
//   warp 0      TMA producer: per k-block of 32 floats loads {U_hi, U_lo, Q_hi, Q_lo} tiles (4 x 16 KB) into a
//               2-stage shared-memory ring, completion on `full` mbarriers;
//   warp 1      TMEM allocator + MMA issuer: per k-block 4 x 3 tcgen05.mma.kind::tf32 (M=128, N=128, K=8) into one of
//               four TMEM accumulators (all 512 columns), tcgen05.commit releases the stage / publishes the accumulator;
//   warps 2..5  epilogue: tcgen05.ld of the thread's row (one user per thread), key = alpha*acc + beta[item],
//               known-item mask by a cursor over the user's sorted list (ids prefetched one block of four ahead),
//               threshold test against the row's current k-th best; the survivors are PARKED in the row's small buffer
//               and a warp drains its buffers at a tile boundary (replace-root + sift-down in the row's k-entry min-heap in
//               shared memory) -- a row gains an entry only ~k ln(n/k) times per pass, but some row of a warp does in
//               almost every chunk, so inserting on the spot would make the whole warp pay the scan for each one.
//               The four warps are independent of each other (each keeps its own copy of the tile's beta values and
//               there is no barrier among them): they only meet in the 128 arrivals that free an accumulator, and the
//               MMAs run up to three tiles ahead, so a warp that drains does not hold up the others.
//
// Rows of up to 128 floats (TS = true): the user tile is the A operand IN TENSOR MEMORY for the whole life of the CTA --
// the epilogue threads split their own row of P into hi / lo and write it with tcgen05.st (columns 0..127 hi, 128..255
// lo; two accumulators in the other 256 columns), the MMAs are  tcgen05.mma [d], [a_tmem], b_desc .  Only the item
// tiles stream through shared memory (four stages of {Q_hi, Q_lo}): half the L2 -> SM traffic (the 128 KB user tile was
// re-read for every item tile) and half the operand reads of the tensor core from shared memory, which it shares with
// the epilogue's top-k lists.  With the shared memory that frees (ES = 2, k <= 55) EIGHT epilogue warps work on a tile: two
// threads per user row, each with the columns of one half of the tile, its own top-k list and its own cursor over the
// known-item list; they exchange only their thresholds (the k-th best of either half is a lower bound of the row's k-th
// best) and merge their two sorted lists at the end.  The epilogue is latency-bound with one warp per scheduler; two
// per scheduler overlap.
#include <cuda.h>

#include <cstdint>
#include <cstdlib>

#include "mfk_common.cuh"

namespace mfk {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 2, TC_KCAP = 64;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;        // 16 KB
constexpr int TC_STAGE_BYTES = 4 * TC_TILE_BYTES;       // U_hi, U_lo, Q_hi, Q_lo
constexpr int TC_THREADS = 192;
constexpr int TC_ACC = 4;                               // TMEM accumulators (4 x 128 columns = all of tensor memory): the MMAs run up to three tiles ahead of the epilogue
#ifndef MFK_TC_SOFT
#define MFK_TC_SOFT 10  // parked survivors in some row from which a shared drain is requested for the next tile boundary
#endif
constexpr int TC_CAP = 30;                              // parked survivors per row (a half chunk can add 16: a warp drains beyond TC_CAP - 16)
constexpr int TC_USER_CHUNK = 128 * 148 * 64;           // users per launch (whole waves of 148 CTAs; one launch for every shape the workspace is sized for)

struct TcParams {
    int32_t m;        // users in this chunk
    int32_t n_items;
    int32_t kblocks;  // padded factor count / 32
    int32_t k;
    int32_t kernel;
    float alpha;                 // key = alpha * (p.q) + beta[item]
    const float *beta;           // [n_pad]   b_i  (linear / sigmoid)  or  -|q|^2 (rbf)
    const float *unorm;          // [m_pad]   |p|^2 (rbf)
    const int32_t *users;        // [m]
    const float *bu;
    const int64_t *mask_ptr;     // [m + 1] (already offset to this chunk) or null
    const int32_t *mask_items;   // sorted ascending inside every row
    float mu, gamma, a, c, lo, hi;
    const float *P;              // TS: user factors [n_users][ldP] (the epilogue threads read their rows themselves)
    int32_t ldP, F;
    int32_t bound;
    int32_t debug;               // MFK_TC_DEBUG bits: 1 = skip epilogue scan, 2 = skip MMA issue (timing experiments only)
    float *out_scores;           // [m][k]
    int32_t *out_items;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a pipeline bug must trap, not hang the GPU.  The suspend-time hint parks the thread in the barrier unit until
// the phase completes (or the hint runs out): without it try_wait comes back at once and the loop spins -- ncu counted 1.2 G of
// the kernel's 4.2 G warp instructions in this loop, issue slots taken from the epilogue warps of the same scheduler.
constexpr uint32_t kMbarSuspendNs = 20000;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
        if (ok) return;
        if ((it & 0xfff) == 0xfff) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int32_t x, int32_t y,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// K-major operand tile, 128-byte swizzle, 8-row groups 1024 bytes apart (what the TMA box {32 floats, rows} writes)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);         // start address
    d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t f2key_tc(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f_tc(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// ---------------------------------------------------------------- operand split
// dst_hi / dst_lo [rows_pad][kp] : tf32-rounded value and the fp32 remainder; optional squared norms.
__global__ void k_split_rows(const float *__restrict__ src, int32_t ld, int32_t F, const int32_t *__restrict__ rows,
                             int32_t n_rows, int32_t rows_pad, int32_t kp, float *dst_hi, float *dst_lo, float *norm2) {
    const int lane = threadIdx.x & 31;
    int32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows_pad) return;
    const bool live = r < n_rows;
    const float *s = live ? src + (size_t)(rows ? rows[r] : r) * ld : nullptr;
    float acc = 0.f;
    for (int c = lane; c < kp; c += 32) {
        float x = (live && c < F) ? s[c] : 0.f;
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        float h = __uint_as_float(hb);
        dst_hi[(size_t)r * kp + c] = h;
        dst_lo[(size_t)r * kp + c] = x - h;
        acc = fmaf(x, x, acc);
    }
    acc = warp_sum(acc);
    if (norm2 && lane == 0) norm2[r] = acc;
}

__global__ void k_make_beta(const float *__restrict__ bi, const float *__restrict__ qnorm, int32_t n_items,
                            int32_t n_pad, int rbf, float *beta) {
    int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pad) return;
    beta[j] = j < n_items ? (rbf ? -qnorm[j] : bi[j]) : -INFINITY;  // padded columns can never be candidates
}

// ---------------------------------------------------------------- main kernel
template <bool TS, int ES>
__global__ void __launch_bounds__(64 + 128 * ES, 1)
k_score_tc(const __grid_constant__ CUtensorMap tm_uhi, const __grid_constant__ CUtensorMap tm_ulo,
           const __grid_constant__ CUtensorMap tm_qhi, const __grid_constant__ CUtensorMap tm_qlo, TcParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned operand ring (128B swizzle atoms), then the epilogue's state
    // (offset arithmetic on the shared array keeps the pointers in the shared address space -> LDS/STS, not generic)
    unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *stage_mem = base;
    // per-row top-k buffer: composite (sortable score key << 32 | ~item) -- larger = better, unique
    static_assert(ES == 1 || (ES == 2 && TS), "two epilogue sets need the shared memory the TS variant frees");
    // TS: stages of {Q_hi, Q_lo} (half the bytes; four of them, or two next to the second set's lists), two accumulators
    // behind the 256 columns of the A operand
    constexpr int NST = TS ? (ES == 2 ? 2 : 4) : TC_STAGES, STB = TS ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES;
    constexpr int NTHR = 64 + 128 * ES;
    constexpr int CAPV = ES == 2 ? 24 : TC_CAP;            // parked survivors per row and set
    constexpr int SOFTV = ES == 2 ? 5 : MFK_TC_SOFT;
    const int KR = ES == 2 ? p.k : TC_KCAP;                // rows of a top-k list
    unsigned long long *tk = reinterpret_cast<unsigned long long *>(base + NST * STB);  // [ES][KR][128]
    float *sbeta = reinterpret_cast<float *>(tk + ES * KR * TC_BM);                      // [4 ES epilogue warps][128]
    unsigned long long *sbuf = reinterpret_cast<unsigned long long *>(sbeta + 4 * ES * TC_BN);  // [ES][CAPV][128] parked survivors per row
    volatile float *s_thr = reinterpret_cast<volatile float *>(sbuf + ES * CAPV * TC_BM);       // [2][128] k-th best of either half (ES = 2)
    volatile int32_t *s_cnt = reinterpret_cast<volatile int32_t *>(s_thr + 2 * TC_BM);          // [128] entries of the second set's list
    uint64_t *bars = reinterpret_cast<uint64_t *>(const_cast<int32_t *>(s_cnt) + TC_BM);
    constexpr int NACC = TS ? 2 : TC_ACC;
    constexpr uint32_t ACC_COL0 = TS ? 2 * TC_BN : 0;
    uint64_t *full = bars, *empty = bars + NST, *acc_full = bars + 2 * NST, *acc_empty = acc_full + NACC;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int32_t m0 = blockIdx.x * TC_BM;
    const int32_t n_tiles = (p.n_items + TC_BN - 1) / TC_BN;
    const int32_t KB = p.kblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(acc_full + a, 1);
            mbar_init(acc_empty + a, 128 * ES);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: TC_ACC 128-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(TC_ACC * TC_BN))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = threadIdx.x; j < ES * KR * TC_BM; j += NTHR) tk[j] = 0ull;
    for (int j = threadIdx.x; j < 2 * TC_BM; j += NTHR) s_thr[j] = -INFINITY;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int32_t nt = 0; nt < n_tiles; ++nt) {
                for (int32_t kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    unsigned char *st = stage_mem + stage * STB;
                    mbar_expect_tx(full + stage, STB);
                    if (TS) {
                        tma_load_2d(st, &tm_qhi, kb * TC_BK, nt * TC_BN, full + stage);
                        tma_load_2d(st + TC_TILE_BYTES, &tm_qlo, kb * TC_BK, nt * TC_BN, full + stage);
                    } else {
                        tma_load_2d(st, &tm_uhi, kb * TC_BK, m0, full + stage);
                        tma_load_2d(st + TC_TILE_BYTES, &tm_ulo, kb * TC_BK, m0, full + stage);
                        tma_load_2d(st + 2 * TC_TILE_BYTES, &tm_qhi, kb * TC_BK, nt * TC_BN, full + stage);
                        tma_load_2d(st + 3 * TC_TILE_BYTES, &tm_qlo, kb * TC_BK, nt * TC_BN, full + stage);
                    }
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (TS) {  // the epilogue threads have written the user tile into tensor memory
            asm volatile("bar.sync 2, %0;" ::"n"(128 * ES + 32) : "memory");
            tc_fence_after();
        }
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t stage = 0, phase = 0;
            for (int32_t nt = 0; nt < n_tiles; ++nt) {
                const uint32_t acc = (uint32_t)nt % NACC;
                mbar_wait(acc_empty + acc, (((uint32_t)nt / NACC) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + ACC_COL0 + acc * TC_BN;
                for (int32_t kb = 0; kb < KB; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_mem + stage * STB);
                    const uint64_t d_uhi = umma_desc_sw128(sa), d_ulo = umma_desc_sw128(sa + TC_TILE_BYTES);
                    const uint64_t d_qhi = umma_desc_sw128(sa + (TS ? 0 : 2) * TC_TILE_BYTES), d_qlo = umma_desc_sw128(sa + (TS ? 1 : 3) * TC_TILE_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < TC_BK / 8; ++k4) {
                        const uint64_t adv = (uint64_t)((k4 * 32) >> 4);  // 8 tf32 = 32 bytes inside the swizzle atom
                        if (p.debug & 2) continue;
                        if (TS) {
                            const uint32_t a_hi = tmem_base + (uint32_t)(kb * TC_BK + k4 * 8), a_lo = a_hi + TC_BN;
                            umma_tf32_ts(tmem_d, a_hi, d_qhi + adv, idesc, (kb | k4) ? 1u : 0u);
                            umma_tf32_ts(tmem_d, a_hi, d_qlo + adv, idesc, 1u);
                            umma_tf32_ts(tmem_d, a_lo, d_qhi + adv, idesc, 1u);
                        } else {
                            umma_tf32(tmem_d, d_uhi + adv, d_qhi + adv, idesc, (kb | k4) ? 1u : 0u);
                            umma_tf32(tmem_d, d_uhi + adv, d_qlo + adv, idesc, 1u);
                            umma_tf32(tmem_d, d_ulo + adv, d_qhi + adv, idesc, 1u);
                        }
                    }
                    umma_commit(empty + stage);  // stage reusable once these MMAs have read it
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(acc_full + acc);  // accumulator complete
            }
        }
    } else {
        // ===== epilogue: thread <-> TMEM lane <-> user row =====
        const int quad = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) ..
        const int row = quad * 32 + lane;
        const int set = (warp - 2) >> 2;  // ES = 2: the half of every tile this thread scans
        unsigned long long *const tks = tk + set * KR * TC_BM;
        unsigned long long *const sbufs = sbuf + set * CAPV * TC_BM;
        const bool live = (m0 + row) < p.m;
        const int32_t k = p.k;
        float un_row = 0.f;  // TS: |p|^2 of the row (rbf)
        if (TS && set == 0) {
            // the thread's own user row, split into the tf32 part and the fp32 remainder, into tensor memory:
            // lane `row`, columns c (hi) and 128 + c (lo)
            const float *prow = p.P + (size_t)(live ? p.users[m0 + row] : 0) * p.ldP;
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
            for (int c0 = 0; c0 < p.kblocks * TC_BK; c0 += 32) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int c = c0 + 4 * c4;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live && c < p.F) x = __ldg(reinterpret_cast<const float4 *>(prow + c));  // (ld is a multiple of 4: inside the row)
                    const float xs[4] = {x.x, c + 1 < p.F ? x.y : 0.f, c + 2 < p.F ? x.z : 0.f, c + 3 < p.F ? x.w : 0.f};  // (columns beyond F do not count)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint32_t hb;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(xs[e]));
                        hi[4 * c4 + e] = hb;
                        lo[4 * c4 + e] = __float_as_uint(xs[e] - __uint_as_float(hb));
                        un_row = fmaf(xs[e], xs[e], un_row);
                    }
                }
                tmem_st32(trow + (uint32_t)c0, hi);
                tmem_st32(trow + (uint32_t)(TC_BN + c0), lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (TS) {
            tc_fence_before();
            asm volatile("bar.sync 2, %0;" ::"n"(128 * ES + 32) : "memory");  // -> the MMA issuer
        }
        // known-item mask: the row's sorted list is consumed in item order; the next four ids sit in registers
        int64_t mcur = 0, me = 0;
        if (live && p.mask_ptr) {
            mcur = p.mask_ptr[m0 + row];
            me = p.mask_ptr[m0 + row + 1];
        }
        // (mk: the block in use;  nk: the block after it, requested when mk was taken over -- its load latency is off the path)
        int32_t mk0 = INT32_MAX, mk1 = INT32_MAX, mk2 = INT32_MAX, mk3 = INT32_MAX;
        int32_t nk0 = INT32_MAX, nk1 = INT32_MAX, nk2 = INT32_MAX, nk3 = INT32_MAX;
        auto mask_load_next = [&](int64_t at) {
            nk0 = (at + 0 < me) ? __ldg(p.mask_items + at + 0) : INT32_MAX;
            nk1 = (at + 1 < me) ? __ldg(p.mask_items + at + 1) : INT32_MAX;
            nk2 = (at + 2 < me) ? __ldg(p.mask_items + at + 2) : INT32_MAX;
            nk3 = (at + 3 < me) ? __ldg(p.mask_items + at + 3) : INT32_MAX;
        };
        auto mask_refill = [&]() {  // mcur: list position of the first id of the block that becomes current
            mk0 = nk0, mk1 = nk1, mk2 = nk2, mk3 = nk3;
            mask_load_next(mcur + 4);
        };
        mask_load_next(mcur);
        mask_refill();
        float thr = -INFINITY;  // candidates must beat it: the k-th best of this thread's list (-inf while the list is not full)
        float pthr = -INFINITY;  // ... or of the other half's (ES = 2)
        int count = 0;
        int nbuf = 0;  // survivors parked in this lane's buffer: (score bits << 32 | column), in item order
        auto drain = [&]() {
            int mx = (p.debug & 4) ? 0 : nbuf;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            for (int r = 0; r < mx; ++r) {
                if (r >= nbuf) continue;
                const unsigned long long ent = sbufs[r * TC_BM + row];
                const float sc = __uint_as_float((uint32_t)(ent >> 32));
                if (!(sc > thr)) continue;  // the threshold has risen since the survivor was parked
                const uint32_t col = (uint32_t)(ent & 0xffffffffull);
                // the row's k best as a binary MIN-HEAP in shared memory (entry i of the row at tks[i * 128 + row]): the list fills
                // in arrival order and is heapified once when it is full; from then on a survivor replaces the root and sifts
                // down -- at most log2(k) dependent steps of two loads and a 64-bit compare, instead of a scan of all k entries
                const unsigned long long ck = ((unsigned long long)f2key_tc(sc) << 32) | (unsigned long long)(0xffffffffu - col);
                auto sift_down = [&](int i, unsigned long long val) {
                    for (;;) {
                        const int l = 2 * i + 1;
                        if (l >= k) break;
                        const unsigned long long hl = tks[l * TC_BM + row];
                        const unsigned long long hr = (l + 1 < k) ? tks[(l + 1) * TC_BM + row] : ~0ull;
                        const bool right = hr < hl;
                        const unsigned long long hc = right ? hr : hl;
                        if (!(hc < val)) break;
                        tks[i * TC_BM + row] = hc;
                        i = right ? l + 1 : l;
                    }
                    tks[i * TC_BM + row] = val;
                };
                bool full = false;
                if (count < k) {
                    tks[count * TC_BM + row] = ck;
                    if (++count == k) {
                        for (int s0 = k / 2 - 1; s0 >= 0; --s0) sift_down(s0, tks[s0 * TC_BM + row]);
                        full = true;
                    }
                } else {
                    sift_down(0, ck);  // (ck beats the root: its score is above the threshold, which is at least the root's)
                    full = true;
                }
                if (full) {
                    const float othr = key2f_tc((uint32_t)(tks[row] >> 32));
                    if (ES == 2) s_thr[set * TC_BM + row] = othr;
                    thr = fmaxf(othr, pthr);
                }
            }
            nbuf = 0;
        };
        // the warp's own copy of the tile's 128 beta values (a lane brings four of them; the next tile's are requested a
        // tile ahead); padded columns carry -inf: never candidates
        float *wbeta = sbeta + (warp - 2) * TC_BN;
        float4 bnext = *reinterpret_cast<const float4 *>(p.beta + 4 * lane);
        for (int32_t nt = 0; nt < n_tiles; ++nt) {
            const uint32_t acc = (uint32_t)nt % NACC;
            __syncwarp();  // (every lane is done with the previous tile's values)
            *reinterpret_cast<float4 *>(wbeta + 4 * lane) = bnext;
            if (nt + 1 < n_tiles) bnext = *reinterpret_cast<const float4 *>(p.beta + (size_t)(nt + 1) * TC_BN + 4 * lane);
            __syncwarp();
            if (ES == 2) {  // the other half's k-th best is a lower bound of the row's, too (monotone: a stale value is still valid)
                pthr = fmaxf(pthr, s_thr[(set ^ 1) * TC_BM + row]);
                thr = fmaxf(thr, pthr);
            }
            // a tile boundary is where a warp drains (one round per parked survivor of its fullest row)
            if (__ballot_sync(0xffffffffu, nbuf > SOFTV) != 0u) drain();
            mbar_wait(acc_full + acc, ((uint32_t)nt / NACC) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ACC_COL0 + acc * TC_BN;
#pragma unroll 1
            for (int ch = set * (TC_BN / 32 / ES); ch < (set + 1) * (TC_BN / 32 / ES); ++ch) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
                    "%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr + (uint32_t)(ch * 32))
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p.debug & 1) continue;  // (warp-uniform; rows beyond the chunk are masked below)
                // pass 1 (branch-free): scores and the bitmask of those above the row's threshold
                const float4 *bt = reinterpret_cast<const float4 *>(wbeta + ch * 32);
                uint32_t cand = 0u;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b4 = bt[j4];
                    float s0 = fmaf(p.alpha, __uint_as_float(v[4 * j4 + 0]), b4.x);
                    float s1 = fmaf(p.alpha, __uint_as_float(v[4 * j4 + 1]), b4.y);
                    float s2 = fmaf(p.alpha, __uint_as_float(v[4 * j4 + 2]), b4.z);
                    float s3 = fmaf(p.alpha, __uint_as_float(v[4 * j4 + 3]), b4.w);
                    v[4 * j4 + 0] = __float_as_uint(s0);
                    v[4 * j4 + 1] = __float_as_uint(s1);
                    v[4 * j4 + 2] = __float_as_uint(s2);
                    v[4 * j4 + 3] = __float_as_uint(s3);
                    cand |= (s0 > thr ? 1u : 0u) << (4 * j4 + 0);
                    cand |= (s1 > thr ? 1u : 0u) << (4 * j4 + 1);
                    cand |= (s2 > thr ? 1u : 0u) << (4 * j4 + 2);
                    cand |= (s3 > thr ? 1u : 0u) << (4 * j4 + 3);
                }
                if (!live) cand = 0u;
                // known items of this chunk's column range are no candidates (the cursor only moves forward)
                const int32_t col0 = nt * TC_BN + ch * 32;
                while (mk0 < col0 + 32) {
                    if (mk0 >= col0) cand &= ~(1u << (mk0 - col0));
                    mk0 = mk1; mk1 = mk2; mk2 = mk3; mk3 = INT32_MAX;
                    ++mcur;
                    if (mk0 == INT32_MAX && mcur < me) mask_refill();
                }
                // pass 2, decoupled: a row gains a top-k entry only ~k ln(n/k) times over the whole pass, but SOME lane of the
                // warp has a survivor in almost every chunk -- handled on the spot the warp would pay the replace-minimum scan
                // in lockstep for every single one.  Survivors are therefore parked in the lane's own buffer (predicated
                // stores, no divergence) and the buffers are drained together once some lane's is about to overflow: the
                // warp then runs max-fill rounds instead of sum-of-fills.  (The threshold only rises at a drain; stale
                // thresholds let a few more candidates through, they are re-checked when drained.)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t hc = (p.debug & 8) ? 0u : (cand >> (16 * half)) & 0xffffu;
                    if (__ballot_sync(0xffffffffu, hc != 0u) == 0u) continue;  // (warp-uniform)
                    if (__ballot_sync(0xffffffffu, nbuf > CAPV - 16) != 0u) drain();  // (rare: the drains at tile boundaries come first)
                    // slot of survivor j = nbuf + survivors before it: no dependent chain through the counter; groups of four
                    // columns without a survivor in any lane are skipped
                    const uint32_t wany = __reduce_or_sync(0xffffffffu, hc);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (((wany >> (4 * g)) & 0xfu) == 0u) continue;  // (warp-uniform)
#pragma unroll
                        for (int j = 4 * g; j < 4 * g + 4; ++j) {
                            if ((hc >> j) & 1u) {
                                const int slot = nbuf + __popc(hc & ((1u << j) - 1u));
                                sbufs[slot * TC_BM + row] = ((unsigned long long)v[16 * half + j] << 32) | (unsigned long long)(uint32_t)(col0 + 16 * half + j);
                            }
                        }
                    }
                    nbuf += __popc(hc);
                }
            }
            if (nt + 1 == n_tiles) drain();
            tc_fence_before();
            mbar_arrive(acc_empty + acc);
        }
        // ---- winners -> predictions: insertion-sort the row's buffer (descending), then emit
        if (live) {
            for (int i = 1; i < count; ++i) {
                const unsigned long long x = tks[i * TC_BM + row];
                int j = i - 1;
                while (j >= 0 && tks[j * TC_BM + row] < x) {
                    tks[(j + 1) * TC_BM + row] = tks[j * TC_BM + row];
                    --j;
                }
                tks[(j + 1) * TC_BM + row] = x;
            }
        }
        int countB = 0, ia = 0, ib = 0;  // ES = 2: the two sorted lists of a row are merged by the first set's thread
        if (ES == 2) {
            if (set == 1) s_cnt[row] = count;
            asm volatile("bar.sync %0, 64;" ::"r"(3 + quad) : "memory");  // the two warps of this quarter of the rows
            if (set == 0) countB = s_cnt[row];
        }
        if (live && set == 0) {
            const unsigned long long *tkb = tk + KR * TC_BM;
            const int32_t user = p.users[m0 + row];
            const float ub = (p.kernel == MFK_KERNEL_RBF) ? 0.f : p.bu[user];
            const float un = (p.kernel == MFK_KERNEL_RBF) ? (TS ? un_row : p.unorm[m0 + row]) : 0.f;
            for (int j = 0; j < k; ++j) {
                float score = -INFINITY;
                int32_t item = -1;
                if (j < count + countB) {
                    unsigned long long ck;
                    if (ES == 2) {  // the larger head of the two lists (keys are unique)
                        const unsigned long long ka = ia < count ? tks[ia * TC_BM + row] : 0ull;
                        const unsigned long long kb2 = ib < countB ? tkb[ib * TC_BM + row] : 0ull;
                        if (ka > kb2) ck = ka, ++ia;
                        else ck = kb2, ++ib;
                    } else {
                        ck = tks[j * TC_BM + row];
                    }
                    item = (int32_t)(0xffffffffu - (uint32_t)(ck & 0xffffffffull));
                    const float kv = key2f_tc((uint32_t)(ck >> 32));
                    if (p.kernel == MFK_KERNEL_LINEAR) score = p.mu + ub + kv;
                    else if (p.kernel == MFK_KERNEL_SIGMOID) score = p.a + p.c * (1.0f / (1.0f + expf(-(p.mu + ub + kv))));
                    else score = p.a + p.c * expf(-p.gamma * fmaxf(un - kv, 0.f));  // |p-q|^2 = |p|^2 - (2p.q - |q|^2)
                    if (p.bound) score = score > p.hi ? p.hi : (score < p.lo ? p.lo : score);
                }
                p.out_scores[(size_t)(m0 + row) * k + j] = score;
                p.out_items[(size_t)(m0 + row) * k + j] = item;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(TC_ACC * TC_BN)) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// [rows][kp] fp32 row-major -> tiles of {32 floats (128 B), 128 rows}, 128-byte swizzle
static int make_tmap(CUtensorMap *tm, const float *base, int64_t rows, int32_t kp) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return MFK_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kp * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return MFK_ERR_CUDA;
    }
    return MFK_OK;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

size_t score_tc_workspace_bytes(int64_t m, int32_t n_items, int32_t n_factors) {
    const int64_t kp = round_up(n_factors, TC_BK);
    const int64_t mc = round_up(m < TC_USER_CHUNK ? (m > 0 ? m : 1) : TC_USER_CHUNK, TC_BM);
    const int64_t np = round_up(n_items, TC_BN);
    return (size_t)(2 * mc * kp + 2 * np * kp + 2 * np + mc + 64) * sizeof(float);
}

int score_tc(int kernel, const int32_t *d_users, int64_t m, const float *d_P, const float *d_Q, const float *d_bu,
             const float *d_bi, int32_t n_items, int32_t n_factors, int32_t ld, float mu, float gamma, float lo,
             float hi, const int64_t *d_mask_ptr, const int32_t *d_mask_items, int32_t k, int bound, float *d_scores,
             int32_t *d_items, void *d_ws, cudaStream_t st) {
    const int32_t kp = (int32_t)round_up(n_factors, TC_BK);
    const int64_t mc_max = round_up(m < TC_USER_CHUNK ? m : TC_USER_CHUNK, TC_BM);
    const int64_t np = round_up(n_items, TC_BN);
    float *ws = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(d_ws) + 255) & ~(uintptr_t)255);
    float *uhi = ws, *ulo = uhi + mc_max * kp, *qhi = ulo + mc_max * kp, *qlo = qhi + np * kp;
    float *qnorm = qlo + np * kp, *beta = qnorm + np, *unorm = beta + np;
    const bool rbf = kernel == MFK_KERNEL_RBF;

    k_split_rows<<<(unsigned)((np * 32 + 255) / 256), 256, 0, st>>>(d_Q, ld, n_factors, nullptr, n_items, (int32_t)np, kp,
                                                                   qhi, qlo, qnorm);
    MFK_LAUNCH_CHECK();
    k_make_beta<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(d_bi, qnorm, n_items, (int32_t)np, rbf ? 1 : 0, beta);
    MFK_LAUNCH_CHECK();
    CUtensorMap tm_qhi, tm_qlo;
    int rc = make_tmap(&tm_qhi, qhi, np, kp);
    if (rc == MFK_OK) rc = make_tmap(&tm_qlo, qlo, np, kp);
    if (rc) return rc;

    const size_t smem = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + (size_t)TC_KCAP * TC_BM * 8 + 4 * TC_BN * 4 +
                        (size_t)TC_CAP * TC_BM * 8 + 3 * TC_BM * 4 + 192;
    // rows of up to 128 floats: the user tile lives in tensor memory (MFK_SCORE_TS=0 keeps it in shared memory)
    const char *ts_env = getenv("MFK_SCORE_TS");
    const bool ts = kp <= TC_BN && !(ts_env && ts_env[0] == '0');
    // ... and with k small enough for two lists per row next to two stages, eight epilogue warps (MFK_SCORE_ES=1: four)
    const size_t smem2 = 1024 + (size_t)TC_STAGES * (TC_STAGE_BYTES / 2) + 2 * (size_t)k * TC_BM * 8 + 8 * TC_BN * 4 +
                         2 * (size_t)24 * TC_BM * 8 + 3 * TC_BM * 4 + 192;
    const char *es_env = getenv("MFK_SCORE_ES");
    DeviceProps props;
    rc = device_props(&props);
    if (rc) return rc;
    const bool es2 = ts && smem2 <= props.smem_optin && !(es_env && es_env[0] == '1');
    MFK_CUDA(cudaFuncSetAttribute(k_score_tc<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MFK_CUDA(cudaFuncSetAttribute(k_score_tc<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MFK_CUDA(cudaFuncSetAttribute(k_score_tc<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    for (int64_t u0 = 0; u0 < m; u0 += TC_USER_CHUNK) {
        const int64_t mt = (m - u0 < TC_USER_CHUNK) ? (m - u0) : TC_USER_CHUNK;
        const int64_t mp = round_up(mt, TC_BM);
        CUtensorMap tm_uhi = tm_qhi, tm_ulo = tm_qlo;  // (TS: not used)
        if (!ts) {
            k_split_rows<<<(unsigned)((mp * 32 + 255) / 256), 256, 0, st>>>(d_P, ld, n_factors, d_users + u0, (int32_t)mt,
                                                                           (int32_t)mp, kp, uhi, ulo, unorm);
            MFK_LAUNCH_CHECK();
            rc = make_tmap(&tm_uhi, uhi, mp, kp);
            if (rc == MFK_OK) rc = make_tmap(&tm_ulo, ulo, mp, kp);
            if (rc) return rc;
        }
        TcParams p;
        p.m = (int32_t)mt;
        p.n_items = n_items;
        p.kblocks = kp / TC_BK;
        p.k = k;
        p.kernel = kernel;
        p.alpha = rbf ? 2.0f : 1.0f;
        p.beta = beta;
        p.unorm = unorm;
        p.users = d_users + u0;
        p.bu = d_bu;
        p.mask_ptr = d_mask_ptr ? d_mask_ptr + u0 : nullptr;
        p.mask_items = d_mask_items;
        p.mu = mu;
        p.gamma = gamma;
        p.a = lo;
        p.c = hi - lo;
        p.lo = lo;
        p.hi = hi;
        p.bound = bound;
        p.P = d_P;
        p.ldP = ld;
        p.F = n_factors;
        {
            const char *e = getenv("MFK_TC_DEBUG");
            p.debug = e ? atoi(e) : 0;
        }
        p.out_scores = d_scores + (size_t)u0 * k;
        p.out_items = d_items + (size_t)u0 * k;
        if (es2) k_score_tc<true, 2><<<(unsigned)(mp / TC_BM), 64 + 128 * 2, smem2, st>>>(tm_uhi, tm_ulo, tm_qhi, tm_qlo, p);
        else if (ts) k_score_tc<true, 1><<<(unsigned)(mp / TC_BM), TC_THREADS, smem, st>>>(tm_uhi, tm_ulo, tm_qhi, tm_qlo, p);
        else k_score_tc<false, 1><<<(unsigned)(mp / TC_BM), TC_THREADS, smem, st>>>(tm_uhi, tm_ulo, tm_qhi, tm_qlo, p);
        MFK_LAUNCH_CHECK();
    }
    return MFK_OK;
}

}  // namespace mfk
