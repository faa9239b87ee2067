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
// finished every step < s.  Hand-off is a monotone progress flag: shared memory inside a
// CTA, a release/relaxed global flag across CTAs.  User rows are read and written through
// L2 (strong loads), one warp per rating, 128-bit coalesced accesses, shuffle-reduced dot
// product, bias/factor/kernel-gradient updates fused in registers.  Consecutive ratings of
// one item keep q in registers (the hot-item chain never leaves the SM).
#include <algorithm>

#include "mfk_common.cuh"
#include "mfk_plan.h"

namespace mfk {

struct RingView {
    const int32_t *su, *si, *sslot, *sstep;
    const float *sr;
    const int64_t *wbeg;
    const int32_t *witems;
    int32_t *flags;
    int32_t W, k, max_slots;
};

struct SgdParams {
    float *P, *Q, *bu, *bi;
    int32_t F;   // n_factors rounded up to a multiple of 4 (columns that exist in memory)
    int32_t ld;  // row stride in floats
    float mu, lr, reg, gamma, a, c;
    int32_t upd_user, upd_item;
    int32_t base;  // flag base of this epoch
};

// progress hand-off state of one worker warp
struct Ring {
    volatile int32_t *sflags;  // shared, one per warp of the CTA
    int32_t *gpub;             // global flag this warp publishes (warp 0 only) or nullptr
    const int32_t *gpoll;      // global flag this warp polls (last warp only) or nullptr
    int32_t warp, base, pub, rel;
    uint32_t spins;
    unsigned long long t0;
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
    __device__ __forceinline__ int32_t poll() {
        int32_t f = gpoll ? ld_strong_i(gpoll) : sflags[warp + 1];
        f -= base;
        // watchdog: a ring that makes no progress for ~4 s is a bug -- trap instead of hanging the GPU
        if (((++spins) & 0xfffu) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (f > rel || t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
        return f > rel ? f : rel;
    }
    // Block until every step < s of the neighbour is complete; forwards progress meanwhile.
    __device__ __forceinline__ void advance_to(int32_t s) {
        if (dirty) {
            if (!gpub) __threadfence_block();  // release at CTA scope; gpu scope rides on st.release
            dirty = false;
        }
        for (;;) {
            int32_t t = min(s, rel + 1);
            if (t > pub) publish(t);
            if (rel >= s) break;
            rel = poll();
        }
    }
    // After the last rating: keep forwarding until the whole ring has drained (pub == W).
    __device__ __forceinline__ void finish(int32_t W) {
        if (dirty) {
            if (!gpub) __threadfence_block();
            dirty = false;
        }
        for (;;) {
            int32_t t = min(W, rel + 1);
            if (t > pub) publish(t);
            if (pub >= W) break;
            rel = poll();
        }
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

// One SGD step on registers.  p, q, ub, ib are updated in place (subject to the flags).
template <int KERNEL, int NV>
__device__ __forceinline__ void sgd_step(Row<NV> &p, Row<NV> &q, float &ub, float &ib, float r,
                                         const SgdParams &prm) {
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
    acc = warp_sum(acc);

    const float lr = prm.lr, reg = prm.reg;
    float gp, gq;  // p -= lr*(gp*q' + reg*p) with q' = q (linear/sigmoid) or (q - p) (rbf)
    if (KERNEL == MFK_KERNEL_LINEAR) {
        float err = (prm.mu + ib + ub + acc) - r;  // kernels.py:145-153
        if (prm.upd_user) ub -= lr * (err + reg * ub);
        if (prm.upd_item) ib -= lr * (err + reg * ib);
        gp = err;
        gq = err;
    } else if (KERNEL == MFK_KERNEL_SIGMOID) {
        float x = prm.mu + ub + ib + acc;  // kernels.py:224-234
        float ex = expf(-x);
        float s = 1.0f / (1.0f + ex);
        float err = (prm.a + prm.c * s) - r;
        float D = (s * s) * ex;  // sigma^2 * e^-x, no factor c
        if (prm.upd_user) ub -= lr * (err * D + reg * ub);
        if (prm.upd_item) ib -= lr * (err * D + reg * ib);
        gp = err * D;
        gq = err * D;
    } else {
        float E = expf(-prm.gamma * acc);  // kernels.py:301-309
        float err = (prm.a + prm.c * E) - r;
        float D = 2.0f * E * prm.gamma;  // no factor c
        gp = err * D;
        gq = err * D;
    }
    const float decay = 1.0f - lr * reg;
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
            qn.x = qv.x - lr * (gq * (pv.x - qv.x) + reg * qv.x);
            qn.y = qv.y - lr * (gq * (pv.y - qv.y) + reg * qv.y);
            qn.z = qv.z - lr * (gq * (pv.z - qv.z) + reg * qv.z);
            qn.w = qv.w - lr * (gq * (pv.w - qv.w) + reg * qv.w);
        } else {
            // p -= lr*(g*q + reg*p) == decay*p - (lr*g)*q   (both sides use the OLD p, q)
            float lg = lr * gp;
            pn.x = fmaf(-lg, qv.x, decay * pv.x);
            pn.y = fmaf(-lg, qv.y, decay * pv.y);
            pn.z = fmaf(-lg, qv.z, decay * pv.z);
            pn.w = fmaf(-lg, qv.w, decay * pv.w);
            qn.x = fmaf(-lg, pv.x, decay * qv.x);
            qn.y = fmaf(-lg, pv.y, decay * qv.y);
            qn.z = fmaf(-lg, pv.z, decay * qv.z);
            qn.w = fmaf(-lg, pv.w, decay * qv.w);
        }
        if (prm.upd_user) p.v[j] = pn;
        if (prm.upd_item) q.v[j] = qn;
    }
}

// NV = number of float4 per lane (row of up to 128*NV floats); NV == 0 is the bias-only model.
template <int NV>
constexpr int ring_max_threads() {
    return NV >= 8 ? 256 : (NV >= 4 ? 512 : 1024);  // keeps the row registers out of local memory
}

template <int KERNEL, int NV, bool QSMEM>
__global__ void __launch_bounds__(ring_max_threads<NV>(), 1) k_sgd_ring(RingView rv, SgdParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *sflags = reinterpret_cast<int32_t *>(smem_raw);  // [32]
    float *sq_all = reinterpret_cast<float *>(smem_raw + 128);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t w = blockIdx.x * rv.k + warp;
    constexpr int NVR = NV > 0 ? NV : 1;
    const bool has_bias = (KERNEL != MFK_KERNEL_RBF);

    if (lane == 0) sflags[warp] = prm.base;
    // per-warp slice of shared memory: max_slots rows of ld floats, then max_slots item biases
    float *sq = nullptr, *sbi = nullptr;
    if (QSMEM) {
        size_t per_warp = (size_t)rv.max_slots * (size_t)(prm.ld + 1);
        per_warp = (per_warp + 3) & ~(size_t)3;
        sq = sq_all + (size_t)warp * per_warp;
        sbi = sq + (size_t)rv.max_slots * prm.ld;
        for (int s = 0; s < rv.max_slots; ++s) {
            int32_t it = rv.witems[(int64_t)s * rv.W + w];
            if (it < 0) continue;
            if (NV > 0) {
                Row<NVR> t;
                load_row<NVR>(t, prm.Q + (size_t)it * prm.ld, lane, prm.F);
                store_row<NVR>(t, sq + (size_t)s * prm.ld, lane, prm.F);
            }
            if (lane == 0) sbi[s] = has_bias ? prm.bi[it] : 0.f;
        }
    }
    __syncthreads();

    Ring ring;
    ring.sflags = sflags;
    ring.warp = warp;
    ring.base = prm.base;
    ring.pub = 0;
    ring.rel = 0;
    ring.dirty = false;
    ring.spins = 0;
    ring.t0 = 0;
    ring.gpub = (warp == 0) ? rv.flags + w : nullptr;
    {
        int32_t nb = (w + 1 == rv.W) ? 0 : w + 1;
        ring.gpoll = (warp == rv.k - 1) ? rv.flags + nb : nullptr;
    }

    const int64_t beg = rv.wbeg[w], end = rv.wbeg[w + 1];
    int32_t cur_step = -1;
    int32_t cur_item = -1;  // slot (QSMEM) or item id whose q / ib are live in registers
    Row<NVR> q;
    float ib = 0.f;
#pragma unroll
    for (int j = 0; j < NVR; ++j) q.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto flush_q = [&]() {
        if (cur_item < 0 || !prm.upd_item) return;
        if (QSMEM) {
            if (NV > 0) store_row<NVR>(q, sq + (size_t)cur_item * prm.ld, lane, prm.F);
            if (lane == 0) sbi[cur_item] = ib;
        } else {
            if (NV > 0) store_row<NVR>(q, prm.Q + (size_t)cur_item * prm.ld, lane, prm.F);
            if (has_bias && lane == 0) prm.bi[cur_item] = ib;
        }
    };

    Row<NVR> pn;  // prefetched / forwarded user row of the next rating
    float ubn = 0.f;
    bool have_next = false;

    for (int64_t b0 = beg; b0 < end; b0 += 32) {
        const int64_t kk = b0 + lane;
        const bool valid = kk < end;
        const int32_t ru = valid ? ld_stream_i(rv.su + kk) : 0;
        const int32_t ri = valid ? ld_stream_i((QSMEM ? rv.sslot : rv.si) + kk) : 0;
        const float rr = valid ? ld_stream_f(rv.sr + kk) : 0.f;
        const int32_t rs = valid ? ld_stream_i(rv.sstep + kk) : 0;
        const int cnt = (int)min((int64_t)32, end - b0);
        for (int j = 0; j < cnt; ++j) {
            const int32_t u = __shfl_sync(0xffffffffu, ru, j);
            const int32_t it = __shfl_sync(0xffffffffu, ri, j);
            const float r = __shfl_sync(0xffffffffu, rr, j);
            const int32_t s = __shfl_sync(0xffffffffu, rs, j);
            if (s != cur_step) {
                ring.advance_to(s);
                cur_step = s;
            }
            // ---- operands
            Row<NVR> p;
            float ub = 0.f;
            if (have_next) {
                p = pn;
                ub = ubn;
            } else {
                if (NV > 0) load_row_strong<NVR>(p, prm.P + (size_t)u * prm.ld, lane, prm.F);
                if (has_bias) ub = ld_strong_f(prm.bu + u);
            }
            if (it != cur_item) {
                flush_q();
                cur_item = it;
                if (QSMEM) {
                    if (NV > 0) load_row<NVR>(q, sq + (size_t)it * prm.ld, lane, prm.F);
                    ib = sbi[it];
                } else {
                    if (NV > 0) load_row<NVR>(q, prm.Q + (size_t)it * prm.ld, lane, prm.F);
                    ib = has_bias ? prm.bi[it] : 0.f;
                }
            }
            // ---- prefetch the next rating's user row if its stripe is already released
            have_next = false;
            int32_t u1 = -1;
            if (j + 1 < cnt) {
                const int32_t s1 = __shfl_sync(0xffffffffu, rs, j + 1);
                u1 = __shfl_sync(0xffffffffu, ru, j + 1);
                if (s1 <= ring.rel) {
                    have_next = true;
                    if (u1 != u) {
                        if (NV > 0) load_row_strong<NVR>(pn, prm.P + (size_t)u1 * prm.ld, lane, prm.F);
                        if (has_bias) ubn = ld_strong_f(prm.bu + u1);
                    }
                }
            }
            // ---- update
            if (NV > 0) {
                sgd_step<KERNEL, NVR>(p, q, ub, ib, r, prm);
            } else {
                // baseline_model.py:259-266: err = r - pred;  b += lr*(err - reg*b)
                float err = r - (prm.mu + ub + ib);
                if (prm.upd_user) ub += prm.lr * (err - prm.reg * ub);
                if (prm.upd_item) ib += prm.lr * (err - prm.reg * ib);
            }
            if (prm.upd_user) {
                if (NV > 0) store_row<NVR>(p, prm.P + (size_t)u * prm.ld, lane, prm.F);
                if (has_bias) prm.bu[u] = ub;  // every lane stores the same value (own program order)
                ring.dirty = true;
            }
            if (have_next && u1 == u) {  // same user again: forward the fresh row in registers
                pn = p;
                ubn = ub;
            }
        }
    }
    flush_q();
    ring.finish(rv.W);

    if (QSMEM && prm.upd_item) {
        __syncwarp();
        for (int s = 0; s < rv.max_slots; ++s) {
            int32_t it = rv.witems[(int64_t)s * rv.W + w];
            if (it < 0) continue;
            if (NV > 0) {
                Row<NVR> t;
                load_row<NVR>(t, sq + (size_t)s * prm.ld, lane, prm.F);
                store_row<NVR>(t, prm.Q + (size_t)it * prm.ld, lane, prm.F);
            }
            if (has_bias && lane == 0) prm.bi[it] = sbi[s];
        }
    }
}

template <int KERNEL, int NV, bool QSMEM>
static int launch_ring(const mfk_plan *plan, const SgdParams &prm, size_t smem, cudaStream_t st) {
    RingView rv;
    rv.su = plan->su;
    rv.si = plan->si;
    rv.sslot = plan->sslot;
    rv.sstep = plan->sstep;
    rv.sr = plan->sr;
    rv.wbeg = plan->wbeg;
    rv.witems = plan->witems;
    rv.flags = plan->flags;
    rv.W = plan->W;
    rv.k = plan->warps_per_cta;
    rv.max_slots = plan->max_slots;
    auto kern = k_sgd_ring<KERNEL, NV, QSMEM>;
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
        set_error("sgd ring: %d CTAs of %d warps cannot be co-resident (%d per SM x %d SMs)", plan->n_ctas,
                  plan->warps_per_cta, per_sm, props.sm_count);
        return MFK_ERR_UNSUPPORTED;
    }
    SgdParams prm_copy = prm;
    void *args[] = {(void *)&rv, (void *)&prm_copy};
    // cooperative launch: the ring spins on its neighbours, so every CTA must be resident
    MFK_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(plan->n_ctas), dim3(plan->warps_per_cta * 32), args,
                                         smem, st));
    return MFK_OK;
}

template <int KERNEL, int NV>
static int launch_ring_q(const mfk_plan *plan, const SgdParams &prm, cudaStream_t st) {
    DeviceProps props;
    int rc = device_props(&props);
    if (rc) return rc;
    size_t per_warp = ((size_t)plan->max_slots * (size_t)(prm.ld + 1) + 3) & ~(size_t)3;
    size_t smem_q = 128 + per_warp * sizeof(float) * (size_t)plan->warps_per_cta;
    if (smem_q + 1024 <= props.smem_optin) return launch_ring<KERNEL, NV, true>(plan, prm, smem_q, st);
    return launch_ring<KERNEL, NV, false>(plan, prm, 128, st);
}

template <int KERNEL>
static int launch_ring_nv(const mfk_plan *plan, const SgdParams &prm, cudaStream_t st) {
    int nv = (prm.F + 127) / 128;
    if (nv <= 1) return launch_ring_q<KERNEL, 1>(plan, prm, st);
    if (nv == 2) return launch_ring_q<KERNEL, 2>(plan, prm, st);
    if (nv <= 4) return launch_ring_q<KERNEL, 4>(plan, prm, st);
    return launch_ring_q<KERNEL, 8>(plan, prm, st);
}

static int32_t next_base(mfk_plan *plan, cudaStream_t st, int *rc) {
    *rc = MFK_OK;
    int64_t base = plan->epoch * (int64_t)(plan->W + 1);
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
    MFK_REQUIRE((((uintptr_t)d_P | (uintptr_t)d_Q) & 15) == 0, "mfk_kmf_sgd_epoch: P/Q must be 16-byte aligned");
    if (n_factors > MFK_MAX_FACTORS) {
        set_error("mfk_kmf_sgd_epoch: n_factors=%d > %d unsupported", n_factors, MFK_MAX_FACTORS);
        return MFK_ERR_UNSUPPORTED;
    }
    if (plan->n == 0) return MFK_OK;
    cudaStream_t st = as_stream(stream);
    SgdParams prm;
    prm.P = d_P;
    prm.Q = d_Q;
    prm.bu = d_bu;
    prm.bi = d_bi;
    prm.F = (n_factors + 3) & ~3;
    prm.ld = ld;
    prm.mu = global_mean;
    prm.lr = lr;
    prm.reg = reg;
    prm.gamma = gamma;
    prm.a = min_rating;
    prm.c = max_rating - min_rating;
    prm.upd_user = update_user_params ? 1 : 0;
    prm.upd_item = update_item_params ? 1 : 0;
    int rc;
    prm.base = next_base(plan, st, &rc);
    if (rc) return rc;
    if (kernel == MFK_KERNEL_LINEAR) return launch_ring_nv<MFK_KERNEL_LINEAR>(plan, prm, st);
    if (kernel == MFK_KERNEL_SIGMOID) return launch_ring_nv<MFK_KERNEL_SIGMOID>(plan, prm, st);
    return launch_ring_nv<MFK_KERNEL_RBF>(plan, prm, st);
}

extern "C" int mfk_bias_sgd_epoch(mfk_plan *plan, float *d_bu, float *d_bi, float global_mean, float lr, float reg,
                                  int update_user_params, int update_item_params, void *stream) {
    MFK_REQUIRE(plan != nullptr, "mfk_bias_sgd_epoch: plan is NULL");
    MFK_REQUIRE(d_bu && d_bi, "mfk_bias_sgd_epoch: null parameter array");
    if (plan->n == 0) return MFK_OK;
    cudaStream_t st = as_stream(stream);
    SgdParams prm;
    prm.P = nullptr;
    prm.Q = nullptr;
    prm.bu = d_bu;
    prm.bi = d_bi;
    prm.F = 0;
    prm.ld = 0;
    prm.mu = global_mean;
    prm.lr = lr;
    prm.reg = reg;
    prm.gamma = 0.f;
    prm.a = 0.f;
    prm.c = 0.f;
    prm.upd_user = update_user_params ? 1 : 0;
    prm.upd_item = update_item_params ? 1 : 0;
    int rc;
    prm.base = next_base(plan, st, &rc);
    if (rc) return rc;
    return launch_ring_q<MFK_KERNEL_LINEAR, 0>(plan, prm, st);
}
