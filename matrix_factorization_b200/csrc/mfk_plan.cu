// Stratified conflict-free SGD plan, built on the GPU.
//
// Replaces `np.random.shuffle(X)` (kernel_matrix_factorization.py:371, baseline_model.py:252)
// by a static DSGD schedule:
//   * items are dealt to W worker warps by descending degree, snake order (nnz balance);
//   * users are dealt to W stripes the same way;
//   * a rating (u, i) belongs to worker w = worker(i) and step s = (stripe(u) - w) mod W;
//   * ratings are sorted by (w, s, slot(i)) so each worker walks one contiguous list and
//     consecutive ratings of one item form a register-resident chain.
// At step s worker w is the only one touching stripe (w + s) mod W and its own items, so no
// two ratings in flight share a user or an item; the step-major linearisation is a valid
// sequential order (mfk_plan_order).
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <vector>

#include <mutex>

#include "mfk_common.cuh"
#include "mfk_plan.h"

// ---- retaining device-memory pool (see mfk_common.cuh)
namespace mfk {
namespace {
struct PoolState {
    bool tried = false, ok = false;
    cudaMemPool_t pool = nullptr;
    cudaStream_t stream = nullptr;
};
PoolState g_pools[64];
std::mutex g_pool_mu;

PoolState *pool_state() {
    static const bool enabled = [] {
        const char *e = getenv("MFK_POOL");
        return !(e && atoi(e) == 0);
    }();
    if (!enabled) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    PoolState &st = g_pools[dev];
    if (!st.tried) {
        st.tried = true;
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&st.pool, &props) == cudaSuccess &&
            cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking) == cudaSuccess) {
            unsigned long long keep = ~0ull;  // never give cached blocks back on a synchronisation
            cudaMemPoolSetAttribute(st.pool, cudaMemPoolAttrReleaseThreshold, &keep);
            st.ok = true;
        } else {
            cudaGetLastError();  // (clear; fall back to cudaMalloc)
        }
    }
    return st.ok ? &st : nullptr;
}
}  // namespace

cudaError_t pool_malloc(void **p, size_t bytes) {
    PoolState *st = pool_state();
    if (!st) return cudaMalloc(p, bytes);
    cudaError_t e = cudaMallocFromPoolAsync(p, bytes, st->pool, st->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st->stream);  // usable from any stream from here on
    return e;
}

cudaError_t pool_free(void *p) {
    if (!p) return cudaSuccess;
    PoolState *st = pool_state();
    if (!st) return cudaFree(p);
    cudaError_t e = cudaDeviceSynchronize();  // what cudaFree does implicitly: nothing in flight may still use the block
    cudaError_t e2 = cudaFreeAsync(p, st->stream);
    return e != cudaSuccess ? e : e2;
}
}  // namespace mfk

namespace mfk {

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int device_props(DeviceProps *out) {
    static thread_local DeviceProps cache[16];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice failed: %s (is a CUDA device visible?)", cudaGetErrorString(e));
        return MFK_ERR_NO_DEVICE;
    }
    if (dev < 16 && cache[dev].device == dev) {
        *out = cache[dev];
        return MFK_OK;
    }
    DeviceProps p;
    p.device = dev;
    int major = 0, minor = 0, optin = 0;
    MFK_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
    MFK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MFK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    MFK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    p.cc = major * 10 + minor;
    p.smem_optin = (size_t)optin;
    if (p.cc != 100) {
        set_error("libmfk_b200 is built for sm_100a only; device %d is sm_%d", dev, p.cc);
        return MFK_ERR_NO_DEVICE;
    }
    if (dev < 16) cache[dev] = p;
    *out = p;
    return MFK_OK;
}

// ------------------------------------------------------------------ kernels
__global__ void k_degrees(const int32_t *__restrict__ u, const int32_t *__restrict__ i, int64_t n,
                          int32_t n_users, int32_t n_items, int32_t *deg_u, int32_t *deg_i,
                          int32_t *bad) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        int32_t uu = u[k], ii = i[k];
        if ((uint32_t)uu >= (uint32_t)n_users || (uint32_t)ii >= (uint32_t)n_items) {
            atomicAdd(bad, 1);
            continue;
        }
        atomicAdd(deg_u + uu, 1);
        atomicAdd(deg_i + ii, 1);
    }
}

__global__ void k_iota(int32_t *a, int32_t n) {
    int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) a[k] = k;
}

// rank r (descending degree) -> snake-dealt bin and round.  With n_ctas > 0 the snake position is
// additionally spread CTA-major (position x -> warp x / n_ctas of CTA x % n_ctas): the heaviest
// workers -- the epoch's critical path -- then sit on different SMs (and, as warps 0, 1, 2, ... of
// their CTA, on different SM sub-partitions) instead of sharing the issue slots of one SM.
__global__ void k_deal(const int32_t *__restrict__ sorted_ids, int32_t n_ids, int32_t W, int32_t n_ctas,
                       int32_t warps_per_cta, int32_t *bin_of, int32_t *round_of,
                       int32_t *table /* [rounds][W] or null */) {
    int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_ids) return;
    int32_t id = sorted_ids[r];
    int32_t round = r / W, pos = r % W;
    int32_t x = (round & 1) ? (W - 1 - pos) : pos;
    // (+1: warp 0 of a CTA is the one that publishes through global memory -- keep the heaviest off it)
    int32_t bin = n_ctas > 0 ? (x % n_ctas) * warps_per_cta + (x / n_ctas + 1) % warps_per_cta : x;
    bin_of[id] = bin;
    if (round_of) round_of[id] = round;
    if (table) table[(int64_t)round * W + bin] = id;
}

__global__ void k_keys(const int32_t *__restrict__ u, const int32_t *__restrict__ i, int64_t n,
                       const int32_t *__restrict__ ustripe, const int32_t *__restrict__ iworker,
                       const int32_t *__restrict__ islot, const uint8_t *__restrict__ layer /* nullable */, int32_t R,
                       int32_t slack, uint64_t *keys, int32_t *idx) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        int32_t ii = i[k];
        int32_t w = iworker[ii];
        int32_t s = ustripe[u[k]] - slack * w;  // worker w meets stripe (slack * w + s) mod R at step s
        if (s < 0) s += R;
        keys[k] = ((uint64_t)w << kPlanWorkerShift) | ((uint64_t)s << kPlanStepShift) |
                  ((uint64_t)(layer ? layer[k] : 0) << kPlanLayerShift) | (uint64_t)islot[ii];
        idx[k] = (int32_t)k;
    }
}

__global__ void k_gather(const uint64_t *__restrict__ keys, const int32_t *__restrict__ idx, int64_t n,
                         const int32_t *__restrict__ u, const int32_t *__restrict__ i,
                         const float *__restrict__ r, int32_t *su, int32_t *si, int32_t *sslot,
                         float *sr, int32_t *sstep) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        uint64_t key = keys[k];
        int32_t j = idx[k];
        su[k] = u[j];
        si[k] = i[j];
        sr[k] = r[j];
        sslot[k] = (int32_t)(key & ((1ull << kPlanLayerShift) - 1));
        sstep[k] = (int32_t)((key >> kPlanStepShift) & ((1ull << (kPlanWorkerShift - kPlanStepShift)) - 1));
    }
}

// Order inside a (worker, step) block.  After a first sort by (worker, step, slot), the c ratings of one item in
// one block split into a chain part -- the first 4 * (c / 4), which stay together and are resolved as exact
// 4-chains -- and a remainder of up to three, which goes to layers 1..3: layer l holds at most one rating per item,
// so consecutive records of a layer are ratings of distinct items that the kernel can process side by side.
__global__ void k_layers(const uint64_t *__restrict__ keys_sorted, const int32_t *__restrict__ idx_sorted, int64_t n,
                         uint8_t *layer /* by original index */) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const uint64_t key = keys_sorted[k];
        if (k > 0 && keys_sorted[k - 1] == key) continue;  // not the head of its run
        int64_t len = 1;
        while (k + len < n && keys_sorted[k + len] == key) ++len;
        const int64_t chain = len & ~(int64_t)3;
        for (int64_t j = 0; j < len; ++j) layer[idx_sorted[k + j]] = j < chain ? 0 : (uint8_t)(1 + j - chain);
    }
}

// rank[original index] = position of the rating inside its run of equal keys (keys sorted; capped at `cap`: the
// caller checks the maximum against the key layout)
__global__ void k_run_rank(const uint64_t *__restrict__ keys_sorted, const int32_t *__restrict__ idx_sorted, int64_t n,
                           int32_t *rank, int32_t *max_rank) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const uint64_t key = keys_sorted[k];
        if (k > 0 && keys_sorted[k - 1] == key) continue;  // not the head of its run
        int64_t len = 1;
        while (k + len < n && keys_sorted[k + len] == key) ++len;
        for (int64_t j = 0; j < len; ++j) rank[idx_sorted[k + j]] = (int32_t)j;
        atomicMax(max_rank, (int32_t)(len - 1));
    }
}

// flat plans: keys of the three sorting passes
//   pass 0: [worker | step | slot]  -> rank among the item's ratings in the cell
//   pass 1: [worker | step | user]  -> rank among the user's ratings in the cell
//   pass 2: [worker | step | rank_u | rank_i | slot]  -> final order
__global__ void k_flat_keys(const int32_t *__restrict__ u, const int32_t *__restrict__ i, int64_t n,
                            const int32_t *__restrict__ ustripe, const int32_t *__restrict__ iworker,
                            const int32_t *__restrict__ islot, const int32_t *__restrict__ rank_u,
                            const int32_t *__restrict__ rank_i, int32_t R, int32_t slack, int pass, uint64_t *keys, int32_t *idx) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const int32_t ii = i[k], uu = u[k];
        const int32_t w = iworker[ii];
        int32_t s = ustripe[uu] - slack * w;  // worker w meets stripe (slack * w + s) mod R at step s
        if (s < 0) s += R;
        const uint64_t cell = ((uint64_t)w << kFlatWorkerShift) | ((uint64_t)s << kFlatStepShift);
        uint64_t low;
        if (pass == 0) low = (uint64_t)islot[ii];
        else if (pass == 1) low = (uint64_t)(uint32_t)uu;
        else low = ((uint64_t)rank_u[k] << kFlatRankUShift) | ((uint64_t)rank_i[k] << kFlatRankIShift) | (uint64_t)islot[ii];
        keys[k] = cell | low;
        idx[k] = (int32_t)k;
    }
}

// flat plans: gather the rating arrays in final order (the records are completed by k_flat_link / k_flat_records)
__global__ void k_flat_gather(const uint64_t *__restrict__ keys, const int32_t *__restrict__ idx, int64_t n,
                              const int32_t *__restrict__ u, const int32_t *__restrict__ i, const float *__restrict__ r,
                              int32_t *su, int32_t *si, int32_t *sslot, float *sr, int32_t *sstep) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const uint64_t key = keys[k];
        const int32_t j = idx[k];
        su[k] = u[j];
        si[k] = i[j];
        sr[k] = r[j];
        sslot[k] = (int32_t)(key & (kFlatMaxSlots - 1));
        sstep[k] = (int32_t)((key >> kFlatStepShift) & 0xffff);
    }
}

// flat plans: keys [cell | id] over the FINAL positions (id = user or slot); a stable sort then lists every id's
// ratings inside a cell in list order
__global__ void k_flat_link_keys(const uint64_t *__restrict__ final_keys, const int32_t *__restrict__ ids, int64_t n,
                                 uint64_t *keys, int32_t *pos) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        keys[k] = ((final_keys[k] >> kFlatStepShift) << kFlatStepShift) | (uint64_t)(uint32_t)ids[k];
        pos[k] = (int32_t)k;
    }
}
// distance (in list positions) to the previous / next rating of the same id in the cell; 0 = none
__global__ void k_flat_link(const uint64_t *__restrict__ keys_sorted, const int32_t *__restrict__ pos_sorted, int64_t n,
                            int32_t *dprev, int32_t *dnext /* nullable */, int32_t *max_dist) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        const int32_t p = pos_sorted[k];
        int32_t d = 0;
        if (k > 0 && keys_sorted[k - 1] == keys_sorted[k]) d = p - pos_sorted[k - 1];
        dprev[p] = d;
        if (dnext) dnext[p] = (k + 1 < n && keys_sorted[k + 1] == keys_sorted[k]) ? pos_sorted[k + 1] - p : 0;
        if (d > 0) atomicMax(max_dist, d);
    }
}
// the records k_sgd_flat streams: {user, rating bits, slot | ordinal of the rating among the worker's ratings of the item << 12,
//                                  distance to the user's previous rating | distance to the user's next rating << 16}
__global__ void k_flat_records(const int32_t *__restrict__ su, const float *__restrict__ sr, const int32_t *__restrict__ sslot,
                               const int32_t *__restrict__ du_prev, const int32_t *__restrict__ du_next,
                               const int32_t *__restrict__ i_ord, int64_t n, int4 *rec) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride)
        rec[k] = make_int4(su[k], __float_as_int(sr[k]), sslot[k] | (i_ord[k] << 12), du_prev[k] | (du_next[k] << 16));
}
// keys [worker | slot] over the final positions: a stable sort lists every item's ratings in the worker's list order
__global__ void k_flat_item_keys(const uint64_t *__restrict__ final_keys, const int32_t *__restrict__ sslot, int64_t n,
                                 uint64_t *keys, int32_t *pos) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        keys[k] = ((final_keys[k] >> kFlatWorkerShift) << kFlatWorkerShift) | (uint64_t)(uint32_t)sslot[k];
        pos[k] = (int32_t)k;
    }
}
// cbeg[w * (R + 1) + s] = first list position (absolute) of cell (w, s); the entry s = R closes the worker's list
__global__ void k_flat_cells(const uint64_t *__restrict__ keys, int64_t n, int32_t W, int32_t R, int32_t *cbeg) {
    int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * (R + 1)) return;
    const uint64_t w = (uint64_t)(t / (R + 1)), s = (uint64_t)(t % (R + 1));
    const uint64_t target = (w << 16) | s;  // first key with (worker, step) >= (w, s); s == R rolls over to the next worker
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((keys[mid] >> kFlatStepShift) < target) lo = mid + 1;
        else hi = mid;
    }
    cbeg[t] = (int32_t)lo;
}

__global__ void k_flat_worker_bounds(const uint64_t *__restrict__ keys, int64_t n, int32_t W, int64_t *wbeg) {
    int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > W) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> kFlatWorkerShift) < (int64_t)w) lo = mid + 1;
        else hi = mid;
    }
    wbeg[w] = lo;
}

// Schedule records streamed by the SGD kernel: {user, slot | (group - 1) << 24, rating bits, ctrl} with
//   kCtrlDup     user occurs among the previous 15 records of this worker (the kernel prefetches user rows up to
//                8 ratings ahead and must not prefetch a row it is about to rewrite), or belongs to the
//                last partial 16-byte chunk of the bias array: such rows are read directly;
//   kCtrlQuad    records k..k+3 have the same (worker, step, slot) key and none of them is a dup, so
//                the four users are distinct and the kernel may resolve them as one exact 4-chain;
//   kCtrlNewStep / kCtrlNewItem   block and item boundaries inside a worker's list;
//   kCtrlOwn     the user occurs among the previous 15 records of this worker: the user's previous rating in
//                the emitted order is then that record (a user meets a worker in one step only);
//   group        g in 1..4: records k..k+g-1 lie in one (worker, step) block outside the chain parts, have
//                pairwise distinct items and users and none is a dup -- independent ratings the kernel interleaves.
// need (nullable): per record, the number of earlier ratings of its user (dataflow schedule); it replaces
// the step in the low bits of ctrl.
__global__ void k_build_records(const uint64_t *__restrict__ keys, const int32_t *__restrict__ su,
                                const float *__restrict__ sr, const int32_t *__restrict__ need, int64_t n,
                                int32_t n_users, int4 *rec) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t slot_mask = (1ull << kPlanLayerShift) - 1;
    const uint64_t low_mask = (1ull << kPlanStepShift) - 1;
    for (; k < n; k += stride) {
        const uint64_t key = keys[k];
        auto is_own = [&](int64_t x) {
            const int32_t u = su[x];
            const uint64_t wk = keys[x] >> kPlanWorkerShift;
            bool d = false;
            for (int j = 1; j <= 15 && x - j >= 0; ++j) d |= (su[x - j] == u && (keys[x - j] >> kPlanWorkerShift) == wk);
            return d;
        };
        // users of the last partial 16-byte bias chunk are read directly too (no over-read of bu)
        auto is_dup = [&](int64_t x) { return (su[x] | 3) >= n_users || is_own(x); };
        int32_t ctrl = need ? (need[k] & kNeedMask) : (int32_t)((key >> kPlanStepShift) & 0xffff);
        if (is_own(k)) ctrl |= kCtrlOwn;
        const bool dup = is_dup(k);
        if (dup) ctrl |= kCtrlDup;
        if (k == 0 || (keys[k - 1] >> kPlanStepShift) != (key >> kPlanStepShift)) ctrl |= kCtrlNewStep;
        if (k == 0 || (keys[k - 1] >> kPlanWorkerShift) != (key >> kPlanWorkerShift) ||
            (keys[k - 1] & slot_mask) != (key & slot_mask))
            ctrl |= kCtrlNewItem;
        if (!dup && k + 3 < n && keys[k + 1] == key && keys[k + 2] == key && keys[k + 3] == key &&
            !is_dup(k + 1) && !is_dup(k + 2) && !is_dup(k + 3))
            ctrl |= kCtrlQuad;
        int g = 1;
        if (!dup && ((key & low_mask) >> kPlanLayerShift) != 0) {
            const uint64_t blk = key >> kPlanStepShift;
            uint64_t slots[4] = {key & slot_mask, 0, 0, 0};
            while (g < 4 && k + g < n) {
                const uint64_t kg = keys[k + g];
                if ((kg >> kPlanStepShift) != blk || ((kg & low_mask) >> kPlanLayerShift) == 0 || is_dup(k + g)) break;
                const uint64_t sg = kg & slot_mask;
                bool clash = false;
                for (int j = 0; j < g; ++j) clash |= (slots[j] == sg);
                if (clash) break;
                slots[g] = sg;
                ++g;
            }
        }
        rec[k] = make_int4(su[k], (int32_t)(key & slot_mask) | ((g - 1) << 24), __float_as_int(sr[k]), ctrl);
    }
}

// wbeg[w] = first sorted position whose worker >= w  (w in 0..W)
__global__ void k_worker_bounds(const uint64_t *__restrict__ keys, int64_t n, int32_t W, int64_t *wbeg) {
    int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > W) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> kPlanWorkerShift) < (int64_t)w) lo = mid + 1;
        else hi = mid;
    }
    wbeg[w] = lo;
}

// dataflow schedule: key (user, step) of every list position
__global__ void k_user_step_keys(const int32_t *__restrict__ su, const int32_t *__restrict__ sstep, int64_t n,
                                 uint64_t *keys, int32_t *pos) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        keys[k] = ((uint64_t)(uint32_t)su[k] << 16) | (uint64_t)(uint32_t)sstep[k];
        pos[k] = (int32_t)k;
    }
}
// j-th entry of the (user, step)-sorted list is the (j - ustart[user])-th rating of its user
__global__ void k_user_ranks(const uint64_t *__restrict__ keys_sorted, const int32_t *__restrict__ pos_sorted,
                             const int32_t *__restrict__ ustart, int64_t n, int32_t *need) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < n; j += stride) need[pos_sorted[j]] = (int32_t)(j - (int64_t)ustart[keys_sorted[j] >> 16]);
}

__global__ void k_pos_iota(int32_t *a, int64_t n) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) a[k] = (int32_t)k;
}

__global__ void k_order_out(const int32_t *__restrict__ pos, const int32_t *__restrict__ sidx, int64_t n,
                            int64_t *order) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) order[k] = (int64_t)sidx[pos[k]];
}

__global__ void k_assignment(const int64_t *__restrict__ wbeg, int32_t W, const int32_t *__restrict__ sidx,
                             const int32_t *__restrict__ sstep, int64_t n, int32_t *worker, int32_t *step) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        int32_t lo = 0, hi = W;  // largest w with wbeg[w] <= k
        while (hi - lo > 1) {
            int32_t mid = (lo + hi) >> 1;
            if (wbeg[mid] <= k) lo = mid;
            else hi = mid;
        }
        int32_t j = sidx[k];
        worker[j] = lo;
        step[j] = sstep[k];
    }
}

static int grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148 * 32));
}

static int bits_for(uint64_t v) {
    int b = 1;
    while ((v >> b) != 0) ++b;
    return b;
}

// sort ids 0..n_ids-1 by descending degree (stable: ties keep id order)
static int sort_by_degree(const int32_t *d_deg, int32_t n_ids, int32_t *d_sorted_ids, cudaStream_t st) {
    int32_t *keys_out = nullptr, *ids_in = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    MFK_CUDA(pool_malloc(&keys_out, sizeof(int32_t) * (size_t)n_ids));
    MFK_CUDA(pool_malloc(&ids_in, sizeof(int32_t) * (size_t)n_ids));
    k_iota<<<(n_ids + 255) / 256, 256, 0, st>>>(ids_in, n_ids);
    MFK_LAUNCH_CHECK();
    MFK_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_deg, keys_out, ids_in,
                                                       d_sorted_ids, n_ids, 0, 32, st));
    MFK_CUDA(pool_malloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    MFK_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, d_deg, keys_out, ids_in,
                                                       d_sorted_ids, n_ids, 0, 32, st));
    MFK_CUDA(cudaStreamSynchronize(st));
    pool_free(tmp);
    pool_free(keys_out);
    pool_free(ids_in);
    return MFK_OK;
}

void choose_workers(int64_t n, int32_t n_users, int32_t n_items, int sm_count, const mfk_plan_opts *opts,
                    int32_t *n_ctas, int32_t *warps_per_cta) {
    int32_t req_w = opts ? opts->n_workers : 0;
    int32_t req_k = opts ? opts->warps_per_cta : 0;
    int32_t cap = std::max(1, std::min(n_users, n_items));  // more workers than users/items is useless
    int32_t W;
    if (req_w > 0) {
        W = req_w;
    } else {
        // ~3 ratings per (worker, step) block keeps the ring hand-off amortised while all SMs
        // get work:  W ~ sqrt(n / 3), bounded by one full wave of 32-warp CTAs.
        double w = std::sqrt((double)std::max<int64_t>(n, 1) / 3.0);
        W = (int32_t)std::max(1.0, std::min(w, (double)sm_count * 16));
        W = std::min(W, std::max(1, cap / 2));
    }
    int32_t k = req_k > 0 ? req_k : 0;
    if (k == 0) {
        if (W >= sm_count * 4) k = std::min(16, std::max(4, (W + sm_count - 1) / sm_count));
        else k = std::min(8, W);
    }
    int32_t kmax = 16;  // 512-thread CTAs leave 128 registers per thread for the 4-chain path
    if (opts && opts->n_factors > 512) kmax = 8;
    k = std::max(1, std::min(kmax, k));
    int32_t ctas = std::max(1, (W + k - 1) / k);
    if (req_w == 0 && ctas > sm_count) ctas = sm_count;
    *n_ctas = ctas;
    *warps_per_cta = k;
}

}  // namespace mfk

using namespace mfk;

extern "C" const char *mfk_last_error(void) { return g_err; }
extern "C" int mfk_abi_version(void) { return MFK_ABI_VERSION; }

extern "C" int mfk_device_query(int device, int *sm_count, int *cc, size_t *smem_optin) {
    int major = 0, minor = 0, sms = 0, optin = 0;
    MFK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    MFK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    MFK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    MFK_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (sm_count) *sm_count = sms;
    if (cc) *cc = major * 10 + minor;
    if (smem_optin) *smem_optin = (size_t)optin;
    return MFK_OK;
}

// The C-ABI entry points of the plan (create / destroy / info / order / assignment / stats).
#include "mfk_plan_api.inc"
