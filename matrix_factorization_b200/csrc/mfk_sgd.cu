// Stratified conflict-free SGD epoch (sm_100a, SIMT, HBM/L2-bound).
//
// Stands in for the rating loop of kernel_matrix_factorization.py:374-425 with the update
// rules of kernels.py:108-180 (linear), :183-262 (sigmoid), :265-327 (rbf), and for
// baseline_model.py:255-266 (bias-only SGD).
//
// Execution model ("ring DSGD"): W worker warps, all co-resident (cooperative launch).
// Worker w owns the item stripe dealt to it for the whole epoch -- its Q rows and item biases
// live in the warp's private slice of shared memory (loaded once, written back once) -- and
// walks its rating list in step order.  At step s it owns user stripe (w + s) mod W; stripes
// move around the ring w+1 -> w, so a worker may enter step s only when its neighbour w+1 has
// finished every step < s.  Hand-off is a monotone progress flag: shared memory inside a CTA,
// a release-store / relaxed-load global flag across CTAs.
//
// Per rating: one warp, 128-bit coalesced accesses, shuffle-reduced dot product, bias / factor
// / kernel-gradient updates fused in registers.  User rows (and the 16-byte chunk holding the
// user bias) are prefetched D ratings ahead with cp.async.cg (L2-coherent, no L1) into a
// per-warp shared-memory ring, as far as the neighbour's progress has released them.
//
// Hot-item chains (consecutive ratings of one item) are the epoch's critical path: the item row
// stays in registers, and for the linear kernel four chained ratings are resolved at once --
// all ten dot products (p_j.q, p_j.p_l) are reduced together and the four errors follow by a
// scalar forward substitution, which is algebraically the sequential update rule
//     q_{j+1} = (1-lr*reg) q_j - lr*err_j p_j,   err_j = mu + b_u + b_i + p_j.q_j - r_j
// (kernels.py:145-178) with one reduction latency per four ratings instead of four.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "mfk_common.cuh"
#include "mfk_plan.h"

#ifndef MFK_RING_PROFILE
#define MFK_RING_PROFILE 0
#endif
#if MFK_RING_PROFILE
#define PROF_T0() long long prof_t_ = clock64()
#define PROF_ADD(slot) do { long long n_ = clock64(); prof[slot] += n_ - prof_t_; prof_t_ = n_; } while (0)
#else
#define PROF_T0() do { } while (0)
#define PROF_ADD(slot) do { } while (0)
#endif

namespace mfk {

constexpr int kRingDepthMax = 8;
constexpr uint32_t kDefaultSleepNs = 256;  // cap of the poll back-off while waiting for a neighbour

struct RingView {
    const int4 *rec;  // per rating {user, slot, rating bits, ctrl}; ctrl = step | kCtrl* flags
    const int64_t *wbeg;
    const int32_t *witems;
    const int32_t *cbeg = nullptr;  // flat plans: [W][R + 1] first list position of every (worker, step) cell
    int32_t *flags;
    int32_t W, k, max_slots;
    int32_t R, slack;  // steps per epoch (= stripes = slack * W); step s needs the neighbour's step s - slack
    int32_t depth;  // prefetch ring depth D (2..kRingDepthMax)
    unsigned long long watchdog_ns;  // trap if a wait sees no progress for this long (0 = never)
    uint32_t max_sleep_ns;           // cap of the poll back-off
    int32_t *uver;                   // [n_users] per-user version counters (dataflow schedule), zero at epoch start
    long long *stats;                // [W][4]: total cycles, cycles blocked in hand-off waits, 4-chains, singles
    long long *prof;                 // [W][8]: phase cycle counters (only written by MFK_RING_PROFILE builds)
};

struct SgdParams {
    float *P, *Q, *bu, *bi;
    int32_t F;   // n_factors rounded up to a multiple of 4 (columns that exist in memory)
    int32_t ld;  // row stride in floats
    int32_t n_users;
    float mu, lr, reg, gamma, a, c;
    int32_t upd_user, upd_item;
    int32_t base;  // flag base of this epoch
};

// progress hand-off state of one worker warp
struct Ring {
    volatile int32_t *sflags;  // shared, one per warp of the CTA
    int32_t *gpub;             // global flag this warp publishes (warp 0 only) or nullptr
    const int32_t *gpoll;      // global flag this warp polls (last warp only) or nullptr
    int32_t warp, base, pub, rel, pending;
    uint32_t max_ns;  // cap of the poll back-off
    int32_t ahead;  // slack - 1: how many steps beyond the neighbour's completed count may be entered
    int32_t last_rel;
    unsigned long long t0, watchdog_ns;
    long long wait_cycles;
    bool dirty;

    __device__ __forceinline__ void publish(int32_t t) {
        pub = t;
        if (gpub) {
            // every lane orders its own row stores before the flag (MEMBAR.ALL.GPU, no L1 flush)
            st_release_gpu_i(gpub, base + t);
        } else {
            sflags[warp] = base + t;
        }
    }
    __device__ __forceinline__ int32_t load_flag() const {
        return (gpoll ? ld_strong_i(gpoll) : sflags[warp + 1]) - base;
    }
    // Non-blocking look at the neighbour: consume the flag value requested last time, request a new
    // one (the load's latency overlaps the rating being processed).
    __device__ __forceinline__ void refresh_async() {
        int32_t f = pending - base;
        if (f > rel) rel = f;
        pending = gpoll ? ld_strong_i(gpoll) : sflags[warp + 1];
    }
    // Wait until the neighbour has completed every step < target, forwarding (publishing, capped at
    // `cap`) the progress observed meanwhile.  The wait loop is a handful of instructions and backs off with nanosleep so that
    // waiting warps leave the issue slots to the working warps of their SM sub-partition.
    __device__ __forceinline__ void wait_for(int32_t target, int32_t cap) {
        long long c0 = clock64();
        unsigned ns = 64, it = 0;
        for (;;) {
            int32_t f = load_flag();
            if (f > rel) {
                rel = f;
                int32_t t = min(cap, rel + 1 + ahead);
                if (t > pub) publish(t);
                if (rel >= target) break;
                ns = 64;
                continue;
            }
            __nanosleep(ns);
            if (ns < max_ns) ns <<= 1;
            if (((++it) & 0x3ffu) == 0 && watchdog_ns) {  // a ring without progress for seconds is a bug: trap
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0 || rel != last_rel) { t0 = now; last_rel = rel; }
                else if (now - t0 > watchdog_ns) __trap();
            }
        }
        wait_cycles += clock64() - c0;
    }
    // may step s be entered (its stripe was released by the neighbour's step s - slack)?
    __device__ __forceinline__ bool open(int32_t s) const { return s <= rel + ahead; }
    // Block until the neighbour has completed every step <= s - slack.
    __device__ __forceinline__ void advance_to(int32_t s) {
        if (dirty) {
            if (!gpub) asm volatile("fence.acq_rel.cta;" ::: "memory");  // CTA-scope release; gpu scope rides on st.release
            dirty = false;
        }
        int32_t t = min(s, rel + 1 + ahead);
        if (t > pub) publish(t);
        if (open(s)) return;
        wait_for(s - ahead, s);
    }
    // After the last rating: keep forwarding until the whole ring has drained (pub == R, the step count).
    __device__ __forceinline__ void finish(int32_t W) {
        if (dirty) {
            if (!gpub) asm volatile("fence.acq_rel.cta;" ::: "memory");
            dirty = false;
        }
        int32_t t = min(W, rel + 1 + ahead);
        if (t > pub) publish(t);
        if (pub >= W) return;
        wait_for(W - 1 - ahead, W);  // rel >= W-1-ahead  =>  pub == W
    }
};

template <int NV>
struct Row {
    float4 v[NV];
};

template <int NV>
__device__ __forceinline__ void load_row_strong(Row<NV> &x, const float *row, int lane, int F) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        int c = 4 * lane + 128 * j;
        x.v[j] = (c < F) ? ld_strong_f4(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int NV>
__device__ __forceinline__ void load_row(Row<NV> &x, const float *row, int lane, int F) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        int c = 4 * lane + 128 * j;
        x.v[j] = (c < F) ? *reinterpret_cast<const float4 *>(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int NV>
__device__ __forceinline__ void store_row(const Row<NV> &x, float *row, int lane, int F) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        int c = 4 * lane + 128 * j;
        if (c < F) *reinterpret_cast<float4 *>(row + c) = x.v[j];
    }
}

// dst is a 32-bit shared-window address (computed once per warp: the generic->shared conversion is not free)
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const float *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most n of this thread's cp.async groups are pending (n is warp-uniform, 0..7)
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}

// One SGD step on registers, in three parts: the per-lane partial of the dot product (squared distance for rbf),
// the scalar part (error, bias updates, gradient factor) and the row updates.  p, q, ub, ib are updated in place
// (subject to the flags).
template <int KERNEL, int NV>
__device__ __forceinline__ float sgd_partial(const Row<NV> &p, const Row<NV> &q) {
    float acc = 0.f;
    if (KERNEL == MFK_KERNEL_RBF) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float dx = p.v[j].x - q.v[j].x, dy = p.v[j].y - q.v[j].y;
            float dz = p.v[j].z - q.v[j].z, dw = p.v[j].w - q.v[j].w;
            acc = fmaf(dx, dx, acc);
            acc = fmaf(dy, dy, acc);
            acc = fmaf(dz, dz, acc);
            acc = fmaf(dw, dw, acc);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            acc = fmaf(p.v[j].x, q.v[j].x, acc);
            acc = fmaf(p.v[j].y, q.v[j].y, acc);
            acc = fmaf(p.v[j].z, q.v[j].z, acc);
            acc = fmaf(p.v[j].w, q.v[j].w, acc);
        }
    }
    return acc;
}

// returns gp:  p -= lr*(gp*q' + reg*p) with q' = q (linear/sigmoid) or (q - p) (rbf)
template <int KERNEL>
__device__ __forceinline__ float sgd_scalar(float acc, float &ub, float &ib, float r, const SgdParams &prm) {
    const float lr = prm.lr, reg = prm.reg;
    if (KERNEL == MFK_KERNEL_LINEAR) {
        float err = (prm.mu + ib + ub + acc) - r;  // kernels.py:145-153
        if (prm.upd_user) ub -= lr * (err + reg * ub);
        if (prm.upd_item) ib -= lr * (err + reg * ib);
        return err;
    } else if (KERNEL == MFK_KERNEL_SIGMOID) {
        float x = prm.mu + ub + ib + acc;  // kernels.py:224-234
        float ex = expf(-x);
        float s = 1.0f / (1.0f + ex);
        float err = (prm.a + prm.c * s) - r;
        float D = (s * s) * ex;  // sigma^2 * e^-x, no factor c
        if (prm.upd_user) ub -= lr * (err * D + reg * ub);
        if (prm.upd_item) ib -= lr * (err * D + reg * ib);
        return err * D;
    } else {
        float E = expf(-prm.gamma * acc);  // kernels.py:301-309
        float err = (prm.a + prm.c * E) - r;
        float D = 2.0f * E * prm.gamma;  // no factor c
        return err * D;
    }
}

template <int KERNEL, int NV>
__device__ __forceinline__ void sgd_apply(Row<NV> &p, Row<NV> &q, float gp, const SgdParams &prm) {
    const float lr = prm.lr, reg = prm.reg;
    // The decay 1 - lr*reg is applied as  x - (lr*reg) x  (one FMA): the constant 1 - lr*reg rounded to fp32 is off by up to
    // 3e-8 relative, a BIAS that a row rated 100 000 times would compound to 4e-3 -- the FMA form only rounds the result.
    const float eps = lr * reg;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        float4 pv = p.v[j], qv = q.v[j];
        float4 pn, qn;
        if (KERNEL == MFK_KERNEL_RBF) {
            // p -= lr*(err*D*(q-p) + reg*p);  q -= lr*(err*D*(p-q) + reg*q)
            pn.x = pv.x - lr * (gp * (qv.x - pv.x) + reg * pv.x);
            pn.y = pv.y - lr * (gp * (qv.y - pv.y) + reg * pv.y);
            pn.z = pv.z - lr * (gp * (qv.z - pv.z) + reg * pv.z);
            pn.w = pv.w - lr * (gp * (qv.w - pv.w) + reg * pv.w);
            qn.x = qv.x - lr * (gp * (pv.x - qv.x) + reg * qv.x);
            qn.y = qv.y - lr * (gp * (pv.y - qv.y) + reg * qv.y);
            qn.z = qv.z - lr * (gp * (pv.z - qv.z) + reg * qv.z);
            qn.w = qv.w - lr * (gp * (pv.w - qv.w) + reg * qv.w);
        } else {
            // p -= lr*(g*q + reg*p) == decay*p - (lr*g)*q   (both sides use the OLD p, q)
            float lg = lr * gp;
            pn.x = fmaf(-lg, qv.x, fmaf(-eps, pv.x, pv.x));
            pn.y = fmaf(-lg, qv.y, fmaf(-eps, pv.y, pv.y));
            pn.z = fmaf(-lg, qv.z, fmaf(-eps, pv.z, pv.z));
            pn.w = fmaf(-lg, qv.w, fmaf(-eps, pv.w, pv.w));
            qn.x = fmaf(-lg, pv.x, fmaf(-eps, qv.x, qv.x));
            qn.y = fmaf(-lg, pv.y, fmaf(-eps, qv.y, qv.y));
            qn.z = fmaf(-lg, pv.z, fmaf(-eps, qv.z, qv.z));
            qn.w = fmaf(-lg, pv.w, fmaf(-eps, qv.w, qv.w));
        }
        if (prm.upd_user) p.v[j] = pn;
        if (prm.upd_item) q.v[j] = qn;
    }
}

template <int KERNEL, int NV>
__device__ __forceinline__ void sgd_step(Row<NV> &p, Row<NV> &q, float &ub, float &ib, float r,
                                         const SgdParams &prm) {
    const float acc = warp_sum(sgd_partial<KERNEL, NV>(p, q));
    const float gp = sgd_scalar<KERNEL>(acc, ub, ib, r, prm);
    sgd_apply<KERNEL, NV>(p, q, gp, prm);
}

// Sum each of four per-lane partials over the warp and return the four totals in every lane: recursive halving
// (2 + 1 + 3 shuffles) and four broadcasts -- the latency of one butterfly reduction for four values.
__device__ __forceinline__ void reduce4(float (&v)[4], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    const float a0 = (b4 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, b4 ? v[0] : v[2], 16);
    const float a1 = (b4 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, b4 ? v[1] : v[3], 16);
    float b = (b3 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, b3 ? a0 : a1, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 4);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    v[0] = __shfl_sync(0xffffffffu, b, 0);   // (bit4, bit3) = (0, 0)
    v[1] = __shfl_sync(0xffffffffu, b, 8);   // (0, 1)
    v[2] = __shfl_sync(0xffffffffu, b, 16);  // (1, 0)
    v[3] = __shfl_sync(0xffffffffu, b, 24);  // (1, 1)
}

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}
// x = ca*x - cb*y  (elementwise)
// (1 - ea) x - cb y, the decay as x - ea x (see sgd_apply)
__device__ __forceinline__ float4 axmby(float ea, const float4 &x, float cb, const float4 &y) {
    float4 o;
    o.x = fmaf(-cb, y.x, fmaf(-ea, x.x, x.x));
    o.y = fmaf(-cb, y.y, fmaf(-ea, x.y, x.y));
    o.z = fmaf(-cb, y.z, fmaf(-ea, x.z, x.z));
    o.w = fmaf(-cb, y.w, fmaf(-ea, x.w, x.w));
    return o;
}

// Sum each of ten per-lane partials over the warp and return all ten totals in every lane.
// Recursive halving: at each butterfly stage a lane keeps one half of its values and sends the other
// half to its partner, so 5+3+2+1+1 = 12 shuffles replace the 50 of ten separate butterflies; ten
// independent index-shuffles then broadcast the totals (value i ends up in lane kQuadLane[i]).
__device__ __forceinline__ void reduce10(float (&v)[10], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float a[5], b[3], c[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        float send = b4 ? v[i] : v[i + 5];
        float keep = b4 ? v[i + 5] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    {
        float s0 = b3 ? a[0] : a[3], k0 = b3 ? a[3] : a[0];
        float s1 = b3 ? a[1] : a[4], k1 = b3 ? a[4] : a[1];
        float s2 = b3 ? a[2] : 0.f, k2 = b3 ? 0.f : a[2];
        b[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
        b[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
        b[2] = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
    }
    {
        float s0 = b2 ? b[0] : b[2], k0 = b2 ? b[2] : b[0];
        float s1 = b2 ? b[1] : 0.f, k1 = b2 ? 0.f : b[1];
        c[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
        c[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
    }
    float d = (b1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b1 ? c[0] : c[1], 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    // value i: i<5 -> bit4 = 0, a-index i;  a-index 0,1,2 -> bit3 = 0 (b-index same), 3,4 -> bit3 = 1 (b 0,1);
    // b-index 0,1 -> bit2 = 0 (c-index same), 2 -> bit2 = 1 (c 0);  c-index = bit1.
    constexpr int kQuadLane[10] = {0, 2, 4, 8, 10, 16, 18, 20, 24, 26};
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = __shfl_sync(0xffffffffu, d, kQuadLane[i]);
}

// NV = number of float4 per lane (row of up to 128*NV floats); NV == 0 is the bias-only model.
template <int NV>
constexpr int ring_max_threads() {
    return NV >= 8 ? 256 : 512;  // >= 128 registers per thread: the chain code must not spill or rematerialise
}

constexpr int kFlowBatch = 4;        // finished ratings published per memory fence (dataflow schedule)
constexpr uint32_t kFlowAhead = 24;  // look for more ready records when fewer than this many are known ahead

// FLOW = false: ring schedule -- workers hand whole user stripes around (Ring above).
// FLOW = true:  dataflow schedule -- a record carries the number of earlier ratings of its user (`need`); it may be
//   applied once uver[user] == need, and applying it publishes need + 1.  Workers walk their lists in order, so
//   the emitted order (a linear extension of the per-worker and per-user orders) is reproduced exactly while
//   every rating only waits for the one rating it really depends on.
template <int KERNEL, int NV, bool QSMEM, bool FLOW>
__global__ void __launch_bounds__(ring_max_threads<NV>(), 1) k_sgd_ring(RingView rv, SgdParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *sflags = reinterpret_cast<int32_t *>(smem_raw);  // [32]
    float *swarp_all = reinterpret_cast<float *>(smem_raw + 128);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t w = blockIdx.x * rv.k + warp;
    constexpr int NVR = NV > 0 ? NV : 1;
    constexpr bool has_bias = (KERNEL != MFK_KERNEL_RBF);
    constexpr bool kQuads = (KERNEL == MFK_KERNEL_LINEAR) && (NV == 1 || NV == 2);
    constexpr bool kPar = QSMEM && NV == 1;  // groups of independent ratings are interleaved (item rows in shared memory)
    constexpr uint32_t RS = 128u * NV + 4u;  // ring slot: a full-width row (zero beyond F) + the 16-byte bias chunk
    const int D = rv.depth;                  // power of two, >= 2
    const uint32_t dmask = (uint32_t)D - 1u;
    const int ld = prm.ld, F = prm.F;
    const int lc = 4 * lane;

    if (lane == 0) sflags[warp] = prm.base;
    // per-warp slice of shared memory:
    //   [Q stripe: max_slots rows of ld floats + max_slots item biases]   (QSMEM only)
    //   [prefetch ring: D slots of RS floats]   [record window: 64 x int4]
    const size_t q_floats = QSMEM ? (((size_t)rv.max_slots * (size_t)(ld + 1) + 3) & ~(size_t)3) : 0;
    const size_t per_warp = q_floats + (size_t)D * RS + 256;
    float *sq = swarp_all + (size_t)warp * per_warp;
    float *sbi = sq + (size_t)rv.max_slots * ld;
    float *sring = sq + q_floats;
    int4 *srec = reinterpret_cast<int4 *>(sring + (size_t)D * RS);
    if (QSMEM) {
        for (int s = 0; s < rv.max_slots; ++s) {
            int32_t it = rv.witems[(int64_t)s * rv.W + w];
            if (it < 0) continue;
            if (NV > 0) {
                Row<NVR> t;
                load_row<NVR>(t, prm.Q + (size_t)it * ld, lane, F);
                store_row<NVR>(t, sq + (size_t)s * ld, lane, F);
            }
            if (lane == 0) sbi[s] = has_bias ? prm.bi[it] : 0.f;
        }
    }
    for (int d = 0; d < D; ++d) {  // zero the ring once: columns beyond F are never written again
#pragma unroll
        for (int v = 0; v < NV; ++v) *reinterpret_cast<float4 *>(sring + d * RS + lc + 128 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane == 0) *reinterpret_cast<float4 *>(sring + d * RS + 128 * NV) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    Ring ring;
    ring.sflags = sflags;
    ring.warp = warp;
    ring.base = prm.base;
    ring.pub = 0;
    ring.rel = 0;
    ring.pending = prm.base;
    ring.dirty = false;
    ring.last_rel = -1;
    ring.t0 = 0;
    ring.watchdog_ns = rv.watchdog_ns;
    ring.wait_cycles = 0;
    ring.ahead = rv.slack - 1;
    ring.max_ns = rv.max_sleep_ns;
    const long long clk_start = clock64();
    int32_t n_quads = 0, n_singles = 0, n_par = 0;
#if MFK_RING_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // 0 step hand-off, 1 prefetch issue, 2 item switch, 3 cp wait,
                                                  // 4 quad math, 5 quad update/store, 6 single, 7 window slide
#endif
    ring.gpub = (warp == 0) ? rv.flags + w : nullptr;
    {
        int32_t nb = (w + 1 == rv.W) ? 0 : w + 1;
        ring.gpoll = (warp == rv.k - 1) ? rv.flags + nb : nullptr;
    }

    const int64_t beg = rv.wbeg[w];
    const uint32_t n_list = (uint32_t)(rv.wbeg[w + 1] - beg);
    const int4 *recs = rv.rec + beg;
    int32_t cur_slot = -1, cur_item = -1;  // slot / item id whose q, ib are live in registers
    Row<NVR> q;
    float ib = 0.f;
#pragma unroll
    for (int j = 0; j < NVR; ++j) q.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto flush_q = [&]() {
        if (cur_slot < 0 || !prm.upd_item) return;
        if (QSMEM) {
            if (NV > 0) store_row<NVR>(q, sq + (size_t)cur_slot * ld, lane, F);
            if (lane == 0) sbi[cur_slot] = ib;
        } else {
            if (NV > 0) store_row<NVR>(q, prm.Q + (size_t)cur_item * ld, lane, F);
            if (has_bias && lane == 0) prm.bi[cur_item] = ib;
        }
    };
    auto load_q = [&](int32_t slot) {
        flush_q();
        cur_slot = slot;
        if (QSMEM) {
            if (NV > 0) load_row<NVR>(q, sq + (size_t)slot * ld, lane, F);
            ib = sbi[slot];
        } else {
            cur_item = rv.witems[(int64_t)slot * rv.W + w];
            if (NV > 0) load_row<NVR>(q, prm.Q + (size_t)cur_item * ld, lane, F);
            ib = has_bias ? prm.bi[cur_item] : 0.f;
        }
    };

    // ---- record window [w0, w0+64) in shared memory, the following batch staged in registers
    const int4 rec_none = make_int4(0, 0, 0, 0xffff);
    uint32_t w0 = 0;
    srec[lane] = (lane < n_list) ? __ldcs(recs + lane) : rec_none;
    srec[32 + lane] = (32u + lane < n_list) ? __ldcs(recs + 32 + lane) : rec_none;
    int4 next_rec = (64u + lane < n_list) ? __ldcs(recs + 64 + lane) : rec_none;
    __syncwarp();

    // ---- prefetch ring: index x lives in slot x & (D-1); one cp.async group per index
    uint32_t pf = 0;  // next index to issue
    auto slot_of = [&](uint32_t x) { return sring + (x & dmask) * RS; };
    const uint32_t sring_s = (uint32_t)__cvta_generic_to_shared(sring);
    auto slot_s = [&](uint32_t x) { return sring_s + (x & dmask) * (RS * 4u); };  // byte address in the shared window
    auto issue_row = [&](uint32_t x, int32_t u, int32_t ctl) {  // the row only; closes the index's group
        if (NV > 0 && !(ctl & kCtrlDup)) {
            const uint32_t slot = slot_s(x) + 4u * (uint32_t)lc;
            const float *row = prm.P + (size_t)u * ld + lc;
#pragma unroll
            for (int j = 0; j < NVR; ++j)
                if (lc + 128 * j < F) cp_async16(slot + 512u * j, row + 128 * j);
        }
        cp_async_commit();
    };
    auto issue = [&](uint32_t x, int32_t u, int32_t ctl) {
        if (has_bias && lane == 0 && !(ctl & kCtrlDup)) cp_async16(slot_s(x) + 512u * NV, prm.bu + (u & ~3));
        issue_row(x, u, ctl);
    };

    // constants of the update rule
    const float aq = prm.upd_item ? 1.0f - prm.lr * prm.reg : 1.0f, lq = prm.upd_item ? prm.lr : 0.f;
    const float ap = prm.upd_user ? 1.0f - prm.lr * prm.reg : 1.0f, lp = prm.upd_user ? prm.lr : 0.f;
    const float aq2 = aq * aq, aq3 = aq2 * aq;
    const float eq = prm.upd_item ? prm.lr * prm.reg : 0.f, ep = prm.upd_user ? prm.lr * prm.reg : 0.f;  // decays as x - e x

    // ---- dataflow state: records [0, rdy_end) are known ready; one poll of the next <= 32 records may be in flight
    uint32_t rdy_end = (FLOW && prm.upd_user) ? 0u : n_list;  // users that are only read never block anybody
    bool f_polling = false;
    int32_t f_ver = 0, f_need = 0;  // lane j: version seen / needed by record rdy_end + j
    int32_t pend_u = 0, pend_v = 0, n_pend = 0;  // lane j < n_pend: version pend_v of user pend_u awaits publication
    auto flow_poll = [&](uint32_t w0_) {
        const uint32_t x = rdy_end + (uint32_t)lane;
        f_need = INT32_MAX;
        f_ver = 0;
        if (x < n_list && x < w0_ + 64u) {
            const int4 rp = srec[x & 63u];
            f_need = rp.w & kNeedMask;
            f_ver = (rp.w & kCtrlOwn) ? INT32_MAX : ld_strong_i(rv.uver + rp.x);
        }
        f_polling = true;
    };
    auto flow_consume = [&]() {
        const unsigned m = __ballot_sync(0xffffffffu, f_ver >= f_need);
        rdy_end += (m == 0xffffffffu) ? 32u : (uint32_t)(__ffs((int)~m) - 1);
        f_polling = false;
    };
    // publish the versions of the ratings applied since the last flush: every lane fences its own row stores first
    auto flow_flush = [&]() {
        if (n_pend == 0) return;
        asm volatile("fence.release.gpu;" ::: "memory");  // MEMBAR.ALL.GPU, no L1 invalidation
        __syncwarp();
        // max, not store: one flush may carry several versions of the same user (its ratings in this cell)
        if (lane < n_pend) asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" ::"l"(rv.uver + pend_u), "r"(pend_v) : "memory");
        n_pend = 0;
    };
    auto flow_done = [&](int32_t user, int32_t ctl) {
        if (lane == n_pend) {
            pend_u = user;
            pend_v = (ctl & kNeedMask) + 1;
        }
        ++n_pend;
    };

    uint32_t k = 0;
    bool first = true;
    while (k < n_list) {
        PROF_T0();
        const int4 rc = srec[k & 63u];
        const int32_t ctrl = rc.w;
        if constexpr (FLOW) {
            if (f_polling) flow_consume();
            // a 4-chain is entered only with all four ratings ready: whether ratings are resolved as a chain or one
            // by one must not depend on timing (the two differ in rounding)
            const uint32_t head_end =
                k + ((kQuads && (ctrl & kCtrlQuad) && D >= 4) ? 4u : ((kPar && D >= 4) ? (uint32_t)(rc.y >> 24) + 1u : 1u));
            if (rdy_end < head_end) {
                // the head record waits for another worker: publish what we owe, then poll with back-off
                flow_flush();
                const long long c0 = clock64();
                unsigned ns = 32, it = 0;
                for (;;) {
                    flow_poll(w0);
                    flow_consume();
                    if (rdy_end >= head_end) break;
                    __nanosleep(ns);
                    if (ns < rv.max_sleep_ns) ns <<= 1;
                    if (((++it) & 0x3ffu) == 0 && ring.watchdog_ns) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        if (ring.t0 == 0) ring.t0 = now;
                        else if (now - ring.t0 > ring.watchdog_ns) {
                            if (lane == 0) {
                                const int4 rb = srec[k & 63u];
                                printf("mfk flow watchdog: worker %d record %u of %u (user %d, needs version %d, sees %d)\n", w, k,
                                       n_list, rb.x, rb.w & kNeedMask, ld_strong_i(rv.uver + rb.x));
                            }
                            __trap();
                        }
                    }
                }
                ring.t0 = 0;
                ring.wait_cycles += clock64() - c0;
            }
            if (rdy_end < n_list && rdy_end < k + kFlowAhead && rdy_end < w0 + 64u) flow_poll(w0);
        }
        if constexpr (!FLOW) {
            if (first || (ctrl & kCtrlNewStep)) ring.advance_to(ctrl & 0xffff);
        }
        PROF_ADD(0);
        // issue as far as the ring depth, the record window and the neighbour's progress allow
        {
            const uint32_t lim = min(min(min(n_list, k + (uint32_t)D), w0 + 64u), rdy_end);
            if (pf + 4u <= lim) {  // common case in a chain: four at once, one progress check
                const int4 a0 = srec[pf & 63u], a1 = srec[(pf + 1u) & 63u], a2 = srec[(pf + 2u) & 63u],
                           a3 = srec[(pf + 3u) & 63u];
                if (FLOW || ring.open(a3.w & 0xffff)) {
                    if (has_bias && lane < 4) {  // the four bias chunks with one instruction (lane j -> index pf+j)
                        const int32_t uj = lane == 0 ? a0.x : (lane == 1 ? a1.x : (lane == 2 ? a2.x : a3.x));
                        const int32_t cj = lane == 0 ? a0.w : (lane == 1 ? a1.w : (lane == 2 ? a2.w : a3.w));
                        if (!(cj & kCtrlDup)) cp_async16(slot_s(pf + (uint32_t)lane) + 512u * NV, prm.bu + (uj & ~3));
                    }
                    issue_row(pf, a0.x, a0.w);
                    issue_row(pf + 1u, a1.x, a1.w);
                    issue_row(pf + 2u, a2.x, a2.w);
                    issue_row(pf + 3u, a3.x, a3.w);
                    pf += 4u;
                }
            }
            while (pf < lim) {
                const int4 rp = srec[pf & 63u];
                if (!FLOW && !ring.open(rp.w & 0xffff)) {
                    ring.refresh_async();  // non-blocking look at the neighbour; otherwise retry next rating
                    if (!ring.open(rp.w & 0xffff)) break;
                }
                issue(pf, rp.x, rp.w);
                ++pf;
            }
        }
        PROF_ADD(1);
        const int32_t slot_k = rc.y & 0xffffff;
        const uint32_t grp = (kPar && D >= 4) ? (uint32_t)(rc.y >> 24) + 1u : 1u;  // independent ratings starting here
        if (grp == 1u && (first || (ctrl & kCtrlNewItem) || (kPar && cur_slot != slot_k))) load_q(slot_k);
        first = false;
        PROF_ADD(2);
        const int32_t u = rc.x;

        if (kQuads && (ctrl & kCtrlQuad) && pf >= k + 4u) {
            // ---- exact 4-chain: same item, same step, four distinct users, all four rows in flight
            const int4 r1 = srec[(k + 1u) & 63u], r2 = srec[(k + 2u) & 63u], r3 = srec[(k + 3u) & 63u];
            cp_async_wait_dyn((int)(pf - 4u - k));
            __syncwarp();  // lane 0's bias-chunk copies become visible to every lane
            PROF_ADD(3);
            const float *s0 = slot_of(k), *s1 = slot_of(k + 1u), *s2 = slot_of(k + 2u), *s3 = slot_of(k + 3u);
            Row<NVR> p0, p1, p2, p3;
#pragma unroll
            for (int v = 0; v < NVR; ++v) {
                p0.v[v] = *reinterpret_cast<const float4 *>(s0 + lc + 128 * v);
                p1.v[v] = *reinterpret_cast<const float4 *>(s1 + lc + 128 * v);
                p2.v[v] = *reinterpret_cast<const float4 *>(s2 + lc + 128 * v);
                p3.v[v] = *reinterpret_cast<const float4 *>(s3 + lc + 128 * v);
            }
            const float ub0 = s0[128 * NV + (u & 3)], ub1 = s1[128 * NV + (r1.x & 3)];
            const float ub2 = s2[128 * NV + (r2.x & 3)], ub3 = s3[128 * NV + (r3.x & 3)];
            // ten reductions at once: t_j = p_j.q (0..3), g_jl = p_j.p_l for l < j (10,20,21,30,31,32)
            float v10[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) v10[i] = 0.f;
#pragma unroll
            for (int v = 0; v < NVR; ++v) {
                const float4 qv = q.v[v];
                v10[0] = dot4(p0.v[v], qv, v10[0]);
                v10[1] = dot4(p1.v[v], qv, v10[1]);
                v10[2] = dot4(p2.v[v], qv, v10[2]);
                v10[3] = dot4(p3.v[v], qv, v10[3]);
                v10[4] = dot4(p1.v[v], p0.v[v], v10[4]);
                v10[5] = dot4(p2.v[v], p0.v[v], v10[5]);
                v10[6] = dot4(p2.v[v], p1.v[v], v10[6]);
                v10[7] = dot4(p3.v[v], p0.v[v], v10[7]);
                v10[8] = dot4(p3.v[v], p1.v[v], v10[8]);
                v10[9] = dot4(p3.v[v], p2.v[v], v10[9]);
            }
            reduce10(v10, lane);
            // forward substitution (item bias rides along as an extra all-ones factor)
            const float e0 = (prm.mu + ub0 - __int_as_float(rc.z)) + (v10[0] + ib);
            const float e1 = (prm.mu + ub1 - __int_as_float(r1.z)) + aq * (v10[1] + ib) - lq * (e0 * (v10[4] + 1.f));
            const float e2 = (prm.mu + ub2 - __int_as_float(r2.z)) + aq2 * (v10[2] + ib) -
                             lq * (aq * e0 * (v10[5] + 1.f) + e1 * (v10[6] + 1.f));
            const float e3 = (prm.mu + ub3 - __int_as_float(r3.z)) + aq3 * (v10[3] + ib) -
                             lq * (aq2 * e0 * (v10[7] + 1.f) + aq * e1 * (v10[8] + 1.f) + e2 * (v10[9] + 1.f));
            PROF_ADD(4);
            // sequential update with the errors known:  p_j' uses q_j,  q_{j+1} uses the OLD p_j
            float *g0 = prm.P + (size_t)u * ld, *g1 = prm.P + (size_t)r1.x * ld;
            float *g2 = prm.P + (size_t)r2.x * ld, *g3 = prm.P + (size_t)r3.x * ld;
#pragma unroll
            for (int v = 0; v < NVR; ++v) {
                const int c = lc + 128 * v;
                const bool in = c < F;
                float4 qv = q.v[v];
                float4 n0 = axmby(ep, p0.v[v], lp * e0, qv);
                qv = axmby(eq, qv, lq * e0, p0.v[v]);
                float4 n1 = axmby(ep, p1.v[v], lp * e1, qv);
                qv = axmby(eq, qv, lq * e1, p1.v[v]);
                float4 n2 = axmby(ep, p2.v[v], lp * e2, qv);
                qv = axmby(eq, qv, lq * e2, p2.v[v]);
                float4 n3 = axmby(ep, p3.v[v], lp * e3, qv);
                qv = axmby(eq, qv, lq * e3, p3.v[v]);
                q.v[v] = qv;
                if (prm.upd_user && in) {
                    *reinterpret_cast<float4 *>(g0 + c) = n0;
                    *reinterpret_cast<float4 *>(g1 + c) = n1;
                    *reinterpret_cast<float4 *>(g2 + c) = n2;
                    *reinterpret_cast<float4 *>(g3 + c) = n3;
                }
            }
            if (prm.upd_user) {
                prm.bu[u] = fmaf(-lp, e0, fmaf(-ep, ub0, ub0));
                prm.bu[r1.x] = fmaf(-lp, e1, fmaf(-ep, ub1, ub1));
                prm.bu[r2.x] = fmaf(-lp, e2, fmaf(-ep, ub2, ub2));
                prm.bu[r3.x] = fmaf(-lp, e3, fmaf(-ep, ub3, ub3));
                ring.dirty = true;
                if constexpr (FLOW) {
                    flow_done(u, ctrl);
                    flow_done(r1.x, r1.w);
                    flow_done(r2.x, r2.w);
                    flow_done(r3.x, r3.w);
                    flow_flush();
                }
            }
            ib = fmaf(-lq, e0, fmaf(-eq, ib, ib));
            ib = fmaf(-lq, e1, fmaf(-eq, ib, ib));
            ib = fmaf(-lq, e2, fmaf(-eq, ib, ib));
            ib = fmaf(-lq, e3, fmaf(-eq, ib, ib));
            k += 4u;
            ++n_quads;
            PROF_ADD(5);
        } else if (kPar && grp > 1u && pf >= k + grp) {
            // ---- a group of 2..4 independent ratings (distinct users, distinct items): loads, reductions and stores
            //      are interleaved, so the group costs about the latency of one rating
            const int4 r1 = srec[(k + 1u) & 63u], r2 = srec[(k + 2u) & 63u], r3 = srec[(k + 3u) & 63u];
            flush_q();
            cur_slot = -1;  // no item row is live in registers after this group
            cp_async_wait_dyn((int)(pf - grp - k));
            __syncwarp();
            PROF_ADD(3);
            const int32_t us[4] = {u, r1.x, r2.x, r3.x};
            const int32_t sl[4] = {slot_k, r1.y & 0xffffff, r2.y & 0xffffff, r3.y & 0xffffff};
            const float rr[4] = {__int_as_float(rc.z), __int_as_float(r1.z), __int_as_float(r2.z), __int_as_float(r3.z)};
            Row<NVR> pp[4];
            float ubs[4], ibs[4], v4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v4[j] = 0.f;
                ubs[j] = 0.f;
                ibs[j] = 0.f;
                pp[j].v[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < (int)grp) {
                    const float *sj = slot_of(k + (uint32_t)j);
                    pp[j].v[0] = *reinterpret_cast<const float4 *>(sj + lc);
                    Row<NVR> qj;
                    load_row<NVR>(qj, sq + (size_t)sl[j] * ld, lane, F);
                    if (has_bias) ubs[j] = sj[128 * NV + (us[j] & 3)];
                    ibs[j] = sbi[sl[j]];
                    v4[j] = sgd_partial<KERNEL, NVR>(pp[j], qj);
                }
            }
            reduce4(v4, lane);
            PROF_ADD(4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < (int)grp) {
                    Row<NVR> qj;
                    load_row<NVR>(qj, sq + (size_t)sl[j] * ld, lane, F);
                    const float gp = sgd_scalar<KERNEL>(v4[j], ubs[j], ibs[j], rr[j], prm);
                    sgd_apply<KERNEL, NVR>(pp[j], qj, gp, prm);
                    if (prm.upd_user) {
                        store_row<NVR>(pp[j], prm.P + (size_t)us[j] * ld, lane, F);
                        if (has_bias) prm.bu[us[j]] = ubs[j];
                    }
                    if (prm.upd_item) {
                        store_row<NVR>(qj, sq + (size_t)sl[j] * ld, lane, F);
                        if (lane == 0) sbi[sl[j]] = ibs[j];
                    }
                }
            }
            if (prm.upd_user) {
                ring.dirty = true;
                if constexpr (FLOW) {
                    flow_done(u, ctrl);
                    flow_done(r1.x, r1.w);
                    if (grp > 2u) flow_done(r2.x, r2.w);
                    if (grp > 3u) flow_done(r3.x, r3.w);
                    flow_flush();
                }
            }
            k += grp;
            ++n_par;
            PROF_ADD(5);
        } else {
            // ---- single rating
            if (kPar && grp > 1u && cur_slot != slot_k) load_q(slot_k);  // (a group that could not be taken as one)
            Row<NVR> p;
            float ub = 0.f;
            if (ctrl & kCtrlDup) {  // recently updated (or last bias chunk): read directly, after our own stores
                if (NV > 0) load_row_strong<NVR>(p, prm.P + (size_t)u * ld, lane, F);
                if (has_bias) ub = ld_strong_f(prm.bu + u);
            } else {
                cp_async_wait_dyn((int)(pf - 1u - k));
                if (has_bias) __syncwarp();
                PROF_ADD(3);
                const float *sl = slot_of(k);
#pragma unroll
                for (int v = 0; v < NVR; ++v)
                    if (NV > 0) p.v[v] = *reinterpret_cast<const float4 *>(sl + lc + 128 * v);
                if (has_bias) ub = sl[128 * NV + (u & 3)];
            }
            const float r = __int_as_float(rc.z);
            if (NV > 0) {
                sgd_step<KERNEL, NVR>(p, q, ub, ib, r, prm);
            } else {
                // baseline_model.py:259-266: err = r - pred;  b += lr*(err - reg*b)
                float err = r - (prm.mu + ub + ib);
                if (prm.upd_user) ub += prm.lr * (err - prm.reg * ub);
                if (prm.upd_item) ib += prm.lr * (err - prm.reg * ib);
            }
            if (prm.upd_user) {
                if (NV > 0) store_row<NVR>(p, prm.P + (size_t)u * ld, lane, F);
                if (has_bias) prm.bu[u] = ub;  // every lane stores the same value (own program order)
                ring.dirty = true;
                if constexpr (FLOW) {
                    flow_done(u, ctrl);
                    if (n_pend >= kFlowBatch) flow_flush();
                }
            }
            k += 1u;
            ++n_singles;
            PROF_ADD(6);
        }
        if (k >= w0 + 32u) {  // slide the record window: the half [w0, w0+32) is dead
            srec[(w0 + lane) & 63u] = next_rec;
            w0 += 32u;
            const uint32_t nx = w0 + 64u + lane;
            next_rec = (nx < n_list) ? __ldcs(recs + nx) : rec_none;
            __syncwarp();
        }
        PROF_ADD(7);
    }
    flush_q();
    const long long clk_work = clock64();
    if constexpr (FLOW) flow_flush();
    else ring.finish(rv.R);
    if (rv.stats && lane == 0) {
        rv.stats[4 * (int64_t)w + 0] = clk_work - clk_start;
        rv.stats[4 * (int64_t)w + 1] = ring.wait_cycles;
        rv.stats[4 * (int64_t)w + 2] = n_quads;
        rv.stats[4 * (int64_t)w + 3] = n_singles;
#if MFK_RING_PROFILE
        for (int j = 0; j < 7; ++j) rv.prof[8 * (int64_t)w + j] = prof[j];
#endif
        rv.prof[8 * (int64_t)w + 7] = n_par;  // groups of four independent ratings
    }

    if (QSMEM && prm.upd_item) {
        __syncwarp();
        for (int s = 0; s < rv.max_slots; ++s) {
            int32_t it = rv.witems[(int64_t)s * rv.W + w];
            if (it < 0) continue;
            if (NV > 0) {
                Row<NVR> t;
                load_row<NVR>(t, sq + (size_t)s * ld, lane, F);
                store_row<NVR>(t, prm.Q + (size_t)it * ld, lane, F);
            }
            if (has_bias && lane == 0) prm.bi[it] = sbi[s];
        }
    }
}

#include "mfk_sgd_hot.inc"
#include "mfk_sgd_hot_pipe.inc"
#include "mfk_sgd_batch.inc"
#include "mfk_sgd_flat.inc"

template <int KERNEL, int NV, bool QSMEM>
static int launch_ring(const mfk_plan *plan, const SgdParams &prm, int depth, size_t smem, cudaStream_t st) {
    RingView rv;
    rv.rec = plan->rec;
    rv.wbeg = plan->wbeg;
    rv.witems = plan->witems;
    rv.flags = plan->flags;
    rv.W = plan->W;
    rv.R = plan->R;
    rv.slack = plan->slack;
    rv.uver = plan->uver;
    rv.k = plan->warps_per_cta;
    rv.max_slots = plan->max_slots;
    rv.depth = depth;
    rv.stats = plan->stats;
    rv.prof = plan->stats + 4 * (size_t)plan->W;
    {
        static const long long wd_ms = [] {
            const char *e = getenv("MFK_RING_WATCHDOG_MS");
            return e ? atoll(e) : 10000ll;
        }();
        rv.watchdog_ns = wd_ms > 0 ? (unsigned long long)wd_ms * 1000000ull : 0ull;
        const char *e = getenv("MFK_RING_SLEEP_NS");
        rv.max_sleep_ns = e ? (uint32_t)atoi(e) : kDefaultSleepNs;
    }
    auto kern = plan->flow ? k_sgd_ring<KERNEL, NV, QSMEM, true> : k_sgd_ring<KERNEL, NV, QSMEM, false>;
    if (plan->flow) MFK_CUDA(cudaMemsetAsync(plan->uver, 0, sizeof(int32_t) * (size_t)plan->n_users, st));
    if (plan->warps_per_cta * 32 > ring_max_threads<NV>()) {
        set_error("sgd ring: plan has %d warps per CTA but rows of %d floats allow at most %d; rebuild the plan with "
                  "mfk_plan_opts.n_factors set", plan->warps_per_cta, prm.F, ring_max_threads<NV>() / 32);
        return MFK_ERR_UNSUPPORTED;
    }
    MFK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MFK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, plan->warps_per_cta * 32, smem));
    DeviceProps props;
    int rc = device_props(&props);
    if (rc) return rc;
    if ((int64_t)per_sm * props.sm_count < plan->n_ctas) {
        set_error("sgd ring: %d CTAs of %d warps (%zu B smem) cannot be co-resident (%d per SM x %d SMs)",
                  plan->n_ctas, plan->warps_per_cta, smem, per_sm, props.sm_count);
        return MFK_ERR_UNSUPPORTED;
    }
    SgdParams prm_copy = prm;
    void *args[] = {(void *)&rv, (void *)&prm_copy};
    // cooperative launch: the ring spins on its neighbours, so every CTA must be resident
    MFK_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(plan->n_ctas), dim3(plan->warps_per_cta * 32), args,
                                         smem, st));
    return MFK_OK;
}

// Shared memory plan: Q stripe in smem if it fits next to a ring of depth >= 4, else Q stays in global/L2.
template <int KERNEL, int NV>
static int launch_ring_q(const mfk_plan *plan, const SgdParams &prm, cudaStream_t st) {
    DeviceProps props;
    int rc = device_props(&props);
    if (rc) return rc;
    const size_t budget = props.smem_optin > 2048 ? props.smem_optin - 1024 : 0;
    const size_t k = (size_t)plan->warps_per_cta;
    const size_t q_floats = ((size_t)plan->max_slots * (size_t)(prm.ld + 1) + 3) & ~(size_t)3;
    const size_t slot_floats = 128 * (size_t)NV + 4;  // full-width ring slots (zero beyond F)
    auto bytes = [&](bool qsmem, int d) {
        return 128 + 4 * k * ((qsmem ? q_floats : 0) + (size_t)d * slot_floats + 256 /* record window */);
    };
    for (int d = kRingDepthMax; d >= 4; d >>= 1)
        if (bytes(true, d) <= budget) return launch_ring<KERNEL, NV, true>(plan, prm, d, bytes(true, d), st);
    for (int d = kRingDepthMax; d >= 2; d >>= 1)
        if (bytes(false, d) <= budget) return launch_ring<KERNEL, NV, false>(plan, prm, d, bytes(false, d), st);
    set_error("sgd ring: no shared-memory configuration fits (%d warps/CTA, ld=%d)", plan->warps_per_cta, prm.ld);
    return MFK_ERR_UNSUPPORTED;
}

template <int KERNEL>
static int launch_ring_nv(const mfk_plan *plan, const SgdParams &prm, cudaStream_t st) {
    int nv = (prm.F + 127) / 128;
    if (nv <= 1) return launch_ring_q<KERNEL, 1>(plan, prm, st);
    if (nv == 2) return launch_ring_q<KERNEL, 2>(plan, prm, st);
    if (nv <= 4) return launch_ring_q<KERNEL, 4>(plan, prm, st);
    return launch_ring_q<KERNEL, 8>(plan, prm, st);
}

// MFK_HOT_PIPE=0 selects the unpipelined hot kernel (diagnostics)
static bool use_hot_pipe() {
    static const bool on = [] {
        const char *e = getenv("MFK_HOT_PIPE");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// MFK_HOT_ENGINE=legacy selects the round-1 hot kernels (A/B diagnostics)
static bool use_batch_engine() {
    static const bool on = [] {
        const char *e = getenv("MFK_HOT_ENGINE");
        return !(e && strcmp(e, "legacy") == 0);
    }();
    return on;
}
// the round-1 kernels rescale by a^-k: keep them away from small or negative a = 1 - lr*reg
static bool legacy_hot_ok(const SgdParams &prm) {
    const float a = 1.0f - prm.lr * prm.reg;
    return a >= 0.5f && a <= 1.0f;
}

// parameters for a role-swapped sub-plan (its "users" are items and vice versa)
static SgdParams swap_roles(const SgdParams &prm, const mfk_plan *sub) {
    SgdParams s = prm;
    s.P = prm.Q;
    s.Q = prm.P;
    s.bu = prm.bi;
    s.bi = prm.bu;
    s.upd_user = prm.upd_item;
    s.upd_item = prm.upd_user;
    s.n_users = sub->n_users;
    return s;
}

static int32_t next_base(mfk_plan *plan, cudaStream_t st, int *rc) {
    *rc = MFK_OK;
    int64_t base = plan->epoch * (int64_t)(plan->R + 1);
    if (base > (int64_t)1 << 30) {  // keep the monotone flags inside int32
        cudaError_t e = cudaMemsetAsync(plan->flags, 0, sizeof(int32_t) * (size_t)(plan->W + 32), st);
        if (e != cudaSuccess) {
            set_error("cudaMemsetAsync(flags) failed: %s", cudaGetErrorString(e));
            *rc = MFK_ERR_CUDA;
        }
        plan->epoch = 0;
        base = 0;
    }
    plan->epoch += 1;
    return (int32_t)base;
}

}  // namespace mfk

using namespace mfk;

extern "C" int mfk_kmf_sgd_epoch(mfk_plan *plan, int kernel, float *d_P, float *d_Q, float *d_bu, float *d_bi,
                                 int32_t n_factors, int32_t ld, float global_mean, float lr, float reg,
                                 float gamma, float min_rating, float max_rating, int update_user_params,
                                 int update_item_params, void *stream) {
    MFK_REQUIRE(plan != nullptr, "mfk_kmf_sgd_epoch: plan is NULL");
    MFK_REQUIRE(kernel >= 0 && kernel <= 2, "mfk_kmf_sgd_epoch: bad kernel %d", kernel);
    MFK_REQUIRE(d_P && d_Q && d_bu && d_bi, "mfk_kmf_sgd_epoch: null parameter array");
    MFK_REQUIRE(n_factors >= 1, "mfk_kmf_sgd_epoch: n_factors must be >= 1");
    MFK_REQUIRE(ld >= n_factors && ld % 4 == 0, "mfk_kmf_sgd_epoch: ld=%d must be >= n_factors and a multiple of 4", ld);
    MFK_REQUIRE((((uintptr_t)d_P | (uintptr_t)d_Q | (uintptr_t)d_bu) & 15) == 0,
                "mfk_kmf_sgd_epoch: P/Q/bu must be 16-byte aligned");
    if (n_factors > MFK_MAX_FACTORS) {
        set_error("mfk_kmf_sgd_epoch: n_factors=%d > %d unsupported", n_factors, MFK_MAX_FACTORS);
        return MFK_ERR_UNSUPPORTED;
    }
    if (plan->n_total == 0) return MFK_OK;
    cudaStream_t st = as_stream(stream);
    SgdParams prm;
    prm.P = d_P;
    prm.Q = d_Q;
    prm.bu = d_bu;
    prm.bi = d_bi;
    prm.F = (n_factors + 3) & ~3;
    prm.ld = ld;
    prm.n_users = plan->n_users;
    prm.mu = global_mean;
    prm.lr = lr;
    prm.reg = reg;
    prm.gamma = gamma;
    prm.a = min_rating;
    prm.c = max_rating - min_rating;
    prm.upd_user = update_user_params ? 1 : 0;
    prm.upd_item = update_item_params ? 1 : 0;
    int rc;
    // The two hot phases side by side (the plan kept their rows disjoint): the hot-user launch goes to a second stream between
    // a fork and a join event, the hot-item launch stays on the caller's stream.  Both are cooperative launches whose workers
    // add up to the SM count at most, so they are resident together; neither waits for the other.
    const bool fork = plan->hot_parallel && plan->hot && plan->hot->n > 0 && plan->hot_users && plan->hot_users->n > 0 &&
                      (plan->phases & 3u) == 3u;
    cudaStream_t st_main = st;
    if (fork) {
        if (!plan->aux_stream) {
            MFK_CUDA(cudaStreamCreateWithFlags(&plan->aux_stream, cudaStreamNonBlocking));
            MFK_CUDA(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
            MFK_CUDA(cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming));
        }
        MFK_CUDA(cudaEventRecord(plan->ev_fork, st_main));
        MFK_CUDA(cudaStreamWaitEvent(plan->aux_stream, plan->ev_fork, 0));
    }
    // one hot sub-plan on a stream: exact mini-batches for the linear kernel (k_sgd_batch); the other kernels walk the same
    // sub-plan with the ring kernel, one warp per id
    auto launch_hot_phase = [&](mfk_plan *sub, SgdParams hp, cudaStream_t s) -> int {
        int rc2 = MFK_OK;
        hp.base = next_base(sub, s, &rc2);
        if (rc2) return rc2;
        if (batch_engine_ok(kernel, hp) && use_batch_engine()) return launch_batch(sub, hp, s);
        if (kernel == MFK_KERNEL_LINEAR && prm.F <= 128 && sub->max_slots == 1 && legacy_hot_ok(hp))
            return use_hot_pipe() ? launch_hot_pipe<1>(sub, hp, s) : launch_hot<1>(sub, hp, s);
        if (kernel == MFK_KERNEL_LINEAR && prm.F <= 256 && sub->max_slots == 1 && legacy_hot_ok(hp)) return launch_hot<2>(sub, hp, s);
        if (kernel == MFK_KERNEL_LINEAR) return launch_ring_nv<MFK_KERNEL_LINEAR>(sub, hp, s);
        if (kernel == MFK_KERNEL_SIGMOID) return launch_ring_nv<MFK_KERNEL_SIGMOID>(sub, hp, s);
        return launch_ring_nv<MFK_KERNEL_RBF>(sub, hp, s);
    };
    const bool run_items = plan->hot && plan->hot->n > 0 && (plan->phases & 1u);
    const bool run_users = plan->hot_users && plan->hot_users->n > 0 && (plan->phases & 2u);
    if (fork) {  // (hot users first: their launch goes to the second stream and the hot items follow at once on the caller's)
        rc = launch_hot_phase(plan->hot_users, swap_roles(prm, plan->hot_users), plan->aux_stream);
        if (rc) return rc;
        MFK_CUDA(cudaEventRecord(plan->ev_join, plan->aux_stream));
    }
    if (run_items) {  // the most-rated items, one CTA per worker
        rc = launch_hot_phase(plan->hot, prm, st_main);
        if (rc) return rc;
    }
    if (run_users && !fork) {
        // then the most active users: the update rules are symmetric in (p_u, b_u) <-> (q_i, b_i), so the same
        // kernels run on the role-swapped sub-plan with the parameter arrays exchanged
        rc = launch_hot_phase(plan->hot_users, swap_roles(prm, plan->hot_users), st_main);
        if (rc) return rc;
    }
    st = st_main;
    if (fork) MFK_CUDA(cudaStreamWaitEvent(st_main, plan->ev_join, 0));
    if (plan->n == 0 || !(plan->phases & 4u)) return MFK_OK;
    prm.base = next_base(plan, st, &rc);
    if (rc) return rc;
    if (plan->flat) {  // CTA workers, batches of independent ratings
        if (kernel == MFK_KERNEL_LINEAR) return launch_flat<MFK_KERNEL_LINEAR>(plan, prm, st);
        if (kernel == MFK_KERNEL_SIGMOID) return launch_flat<MFK_KERNEL_SIGMOID>(plan, prm, st);
        return launch_flat<MFK_KERNEL_RBF>(plan, prm, st);
    }
    if (kernel == MFK_KERNEL_LINEAR) return launch_ring_nv<MFK_KERNEL_LINEAR>(plan, prm, st);
    if (kernel == MFK_KERNEL_SIGMOID) return launch_ring_nv<MFK_KERNEL_SIGMOID>(plan, prm, st);
    return launch_ring_nv<MFK_KERNEL_RBF>(plan, prm, st);
}

extern "C" int mfk_bias_sgd_epoch(mfk_plan *plan, float *d_bu, float *d_bi, float global_mean, float lr, float reg,
                                  int update_user_params, int update_item_params, void *stream) {
    MFK_REQUIRE(plan != nullptr, "mfk_bias_sgd_epoch: plan is NULL");
    MFK_REQUIRE(d_bu && d_bi, "mfk_bias_sgd_epoch: null parameter array");
    MFK_REQUIRE(((uintptr_t)d_bu & 15) == 0, "mfk_bias_sgd_epoch: bu must be 16-byte aligned");
    if (plan->n_total == 0) return MFK_OK;
    cudaStream_t st = as_stream(stream);
    SgdParams prm;
    prm.P = nullptr;
    prm.Q = nullptr;
    prm.bu = d_bu;
    prm.bi = d_bi;
    prm.F = 0;
    prm.ld = 0;
    prm.n_users = plan->n_users;
    prm.mu = global_mean;
    prm.lr = lr;
    prm.reg = reg;
    prm.gamma = 0.f;
    prm.a = 0.f;
    prm.c = 0.f;
    prm.upd_user = update_user_params ? 1 : 0;
    prm.upd_item = update_item_params ? 1 : 0;
    int rc;
    if (plan->hot && plan->hot->n > 0 && (plan->phases & 1u)) {
        SgdParams hp = prm;
        hp.base = next_base(plan->hot, st, &rc);
        if (rc) return rc;
        rc = launch_ring_q<MFK_KERNEL_LINEAR, 0>(plan->hot, hp, st);
        if (rc) return rc;
    }
    if (plan->hot_users && plan->hot_users->n > 0 && (plan->phases & 2u)) {
        SgdParams hp = swap_roles(prm, plan->hot_users);
        hp.base = next_base(plan->hot_users, st, &rc);
        if (rc) return rc;
        rc = launch_ring_q<MFK_KERNEL_LINEAR, 0>(plan->hot_users, hp, st);
        if (rc) return rc;
    }
    if (plan->n == 0 || !(plan->phases & 4u)) return MFK_OK;
    MFK_REQUIRE(!plan->flat, "mfk_bias_sgd_epoch: the plan was built for factor rows (n_factors > 0); build it with n_factors = 0");
    prm.base = next_base(plan, st, &rc);
    if (rc) return rc;
    return launch_ring_q<MFK_KERNEL_LINEAR, 0>(plan, prm, st);
}

// diagnostics of profile builds: clock stamps of one iteration of the batch engine (not part of include/mfk.h)
extern "C" int mfk_debug_trace(long long *h_out, int n) {
#if MFK_BATCH_PROFILE
    if (n > 256) n = 256;
    MFK_CUDA(cudaMemcpyFromSymbol(h_out, g_bt_trace, sizeof(long long) * (size_t)n));
    return MFK_OK;
#else
    (void)h_out;
    (void)n;
    return MFK_ERR_UNSUPPORTED;
#endif
}
