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
#include <vector>

#include "mfk_common.cuh"
#include "mfk_plan.h"

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
                       const int32_t *__restrict__ islot, int32_t W, uint64_t *keys, int32_t *idx) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        int32_t ii = i[k];
        int32_t w = iworker[ii];
        int32_t s = ustripe[u[k]] - w;
        if (s < 0) s += W;
        keys[k] = ((uint64_t)w << kPlanWorkerShift) | ((uint64_t)s << kPlanStepShift) | (uint64_t)islot[ii];
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
        sslot[k] = (int32_t)(key & ((1ull << kPlanStepShift) - 1));
        sstep[k] = (int32_t)((key >> kPlanStepShift) & ((1ull << (kPlanWorkerShift - kPlanStepShift)) - 1));
    }
}

// Schedule records streamed by the SGD kernel: {user, slot, rating bits, ctrl} with
//   kCtrlDup     user occurs among the previous 15 records (the kernel prefetches user rows up to 8
//                ratings ahead and must not prefetch a row it is about to rewrite), or belongs to the
//                last partial 16-byte chunk of the bias array: such rows are read directly;
//   kCtrlQuad    records k..k+3 have the same (worker, step, slot) key and none of them is a dup, so
//                the four users are distinct and the kernel may resolve them as one exact 4-chain;
//   kCtrlNewStep / kCtrlNewItem   block and item boundaries inside a worker's list.
__global__ void k_build_records(const uint64_t *__restrict__ keys, const int32_t *__restrict__ su,
                                const float *__restrict__ sr, int64_t n, int32_t n_users, int4 *rec) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t slot_mask = (1ull << kPlanStepShift) - 1;
    for (; k < n; k += stride) {
        const uint64_t key = keys[k];
        auto is_dup = [&](int64_t x) {
            int32_t u = su[x];
            // users of the last partial 16-byte bias chunk are read directly too (no over-read of bu)
            bool d = (u | 3) >= n_users;
            for (int j = 1; j <= 15 && x - j >= 0; ++j) d |= (su[x - j] == u);
            return d;
        };
        int32_t ctrl = (int32_t)((key >> kPlanStepShift) & 0xffff);
        const bool dup = is_dup(k);
        if (dup) ctrl |= kCtrlDup;
        if (k == 0 || (keys[k - 1] >> kPlanStepShift) != (key >> kPlanStepShift)) ctrl |= kCtrlNewStep;
        if (k == 0 || (keys[k - 1] >> kPlanWorkerShift) != (key >> kPlanWorkerShift) ||
            (keys[k - 1] & slot_mask) != (key & slot_mask))
            ctrl |= kCtrlNewItem;
        if (!dup && k + 3 < n && keys[k + 1] == key && keys[k + 2] == key && keys[k + 3] == key &&
            !is_dup(k + 1) && !is_dup(k + 2) && !is_dup(k + 3))
            ctrl |= kCtrlQuad;
        rec[k] = make_int4(su[k], (int32_t)(key & slot_mask), __float_as_int(sr[k]), ctrl);
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
    MFK_CUDA(cudaMalloc(&keys_out, sizeof(int32_t) * (size_t)n_ids));
    MFK_CUDA(cudaMalloc(&ids_in, sizeof(int32_t) * (size_t)n_ids));
    k_iota<<<(n_ids + 255) / 256, 256, 0, st>>>(ids_in, n_ids);
    MFK_LAUNCH_CHECK();
    MFK_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d_deg, keys_out, ids_in,
                                                       d_sorted_ids, n_ids, 0, 32, st));
    MFK_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    MFK_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, d_deg, keys_out, ids_in,
                                                       d_sorted_ids, n_ids, 0, 32, st));
    MFK_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    cudaFree(keys_out);
    cudaFree(ids_in);
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

static void plan_free(mfk_plan *p) {
    if (!p) return;
    void *ptrs[] = {p->su, p->si, p->sslot, p->sr, p->sstep, p->rec, p->sidx, p->wbeg, p->witems,
                    p->iworker, p->islot, p->ustripe, p->flags, p->stats};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    delete p;
}

extern "C" int mfk_plan_destroy(mfk_plan *plan) {
    plan_free(plan);
    return MFK_OK;
}

extern "C" int mfk_plan_create(mfk_plan **out, const int32_t *d_u, const int32_t *d_i, const float *d_r,
                               int64_t n, int32_t n_users, int32_t n_items, const mfk_plan_opts *opts,
                               void *stream) {
    MFK_REQUIRE(out != nullptr, "mfk_plan_create: out is NULL");
    *out = nullptr;
    MFK_REQUIRE(n >= 0 && n < (int64_t)INT32_MAX, "mfk_plan_create: n=%lld out of range", (long long)n);
    MFK_REQUIRE(n_users > 0 && n_items > 0, "mfk_plan_create: n_users/n_items must be positive");
    MFK_REQUIRE(n == 0 || (d_u && d_i && d_r), "mfk_plan_create: null rating arrays");
    DeviceProps props;
    int rc = device_props(&props);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);

    mfk_plan *p = new mfk_plan();
    p->n = n;
    p->n_users = n_users;
    p->n_items = n_items;
    choose_workers(n, n_users, n_items, props.sm_count, opts, &p->n_ctas, &p->warps_per_cta);
    p->W = p->n_ctas * p->warps_per_cta;
    const int32_t W = p->W;
    if (W >= (1 << (kPlanWorkerShift - kPlanStepShift))) {
        plan_free(p);
        set_error("mfk_plan_create: %d workers exceed the key layout", W);
        return MFK_ERR_UNSUPPORTED;
    }
    p->max_slots = (n_items + W - 1) / W;
    if ((uint64_t)p->max_slots >= (1ull << kPlanStepShift)) {
        plan_free(p);
        set_error("mfk_plan_create: %d items per worker exceed the key layout", p->max_slots);
        return MFK_ERR_UNSUPPORTED;
    }

#define PLAN_CUDA(call)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            plan_free(p);                                                                     \
            for (void *q : scratch) cudaFree(q);                                              \
            return MFK_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)
    std::vector<void *> scratch;
    auto dalloc = [&](void **ptr, size_t bytes) { return cudaMalloc(ptr, bytes ? bytes : 16); };

    size_t nn = (size_t)std::max<int64_t>(n, 1);
    PLAN_CUDA(dalloc((void **)&p->su, nn * 4));
    PLAN_CUDA(dalloc((void **)&p->si, nn * 4));
    PLAN_CUDA(dalloc((void **)&p->sslot, nn * 4));
    PLAN_CUDA(dalloc((void **)&p->sr, nn * 4));
    PLAN_CUDA(dalloc((void **)&p->sstep, nn * 4));
    PLAN_CUDA(dalloc((void **)&p->rec, nn * 16));
    PLAN_CUDA(dalloc((void **)&p->sidx, nn * 4));
    PLAN_CUDA(dalloc((void **)&p->wbeg, sizeof(int64_t) * (size_t)(W + 1)));
    PLAN_CUDA(dalloc((void **)&p->witems, sizeof(int32_t) * (size_t)p->max_slots * W));
    PLAN_CUDA(dalloc((void **)&p->iworker, sizeof(int32_t) * (size_t)n_items));
    PLAN_CUDA(dalloc((void **)&p->islot, sizeof(int32_t) * (size_t)n_items));
    PLAN_CUDA(dalloc((void **)&p->ustripe, sizeof(int32_t) * (size_t)n_users));
    PLAN_CUDA(dalloc((void **)&p->flags, sizeof(int32_t) * (size_t)(W + 32)));
    PLAN_CUDA(cudaMemsetAsync(p->flags, 0, sizeof(int32_t) * (size_t)(W + 32), st));
    PLAN_CUDA(dalloc((void **)&p->stats, sizeof(long long) * 12 * (size_t)W));
    PLAN_CUDA(cudaMemsetAsync(p->stats, 0, sizeof(long long) * 12 * (size_t)W, st));
    PLAN_CUDA(cudaMemsetAsync(p->witems, 0xff, sizeof(int32_t) * (size_t)p->max_slots * W, st));

    int32_t *deg_u = nullptr, *deg_i = nullptr, *sorted_u = nullptr, *sorted_i = nullptr, *bad = nullptr;
    PLAN_CUDA(dalloc((void **)&deg_u, sizeof(int32_t) * (size_t)n_users)); scratch.push_back(deg_u);
    PLAN_CUDA(dalloc((void **)&deg_i, sizeof(int32_t) * (size_t)n_items)); scratch.push_back(deg_i);
    PLAN_CUDA(dalloc((void **)&sorted_u, sizeof(int32_t) * (size_t)n_users)); scratch.push_back(sorted_u);
    PLAN_CUDA(dalloc((void **)&sorted_i, sizeof(int32_t) * (size_t)n_items)); scratch.push_back(sorted_i);
    PLAN_CUDA(dalloc((void **)&bad, sizeof(int32_t))); scratch.push_back(bad);
    PLAN_CUDA(cudaMemsetAsync(deg_u, 0, sizeof(int32_t) * (size_t)n_users, st));
    PLAN_CUDA(cudaMemsetAsync(deg_i, 0, sizeof(int32_t) * (size_t)n_items, st));
    PLAN_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
    if (n > 0) {
        k_degrees<<<grid_for(n), 256, 0, st>>>(d_u, d_i, n, n_users, n_items, deg_u, deg_i, bad);
        PLAN_CUDA(cudaGetLastError());
    }
    int32_t h_bad = 0;
    PLAN_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaStreamSynchronize(st));
    if (h_bad != 0) {
        plan_free(p);
        for (void *q : scratch) cudaFree(q);
        set_error("mfk_plan_create: %d ratings have ids outside [0,n_users) x [0,n_items)", h_bad);
        return MFK_ERR_ARG;
    }
    rc = sort_by_degree(deg_u, n_users, sorted_u, st);
    if (rc == MFK_OK) rc = sort_by_degree(deg_i, n_items, sorted_i, st);
    if (rc != MFK_OK) {
        plan_free(p);
        for (void *q : scratch) cudaFree(q);
        return rc;
    }
    k_deal<<<(n_users + 255) / 256, 256, 0, st>>>(sorted_u, n_users, W, 0, 0, p->ustripe, nullptr, nullptr);
    PLAN_CUDA(cudaGetLastError());
    k_deal<<<(n_items + 255) / 256, 256, 0, st>>>(sorted_i, n_items, W, p->n_ctas, p->warps_per_cta, p->iworker,
                                                    p->islot, p->witems);
    PLAN_CUDA(cudaGetLastError());
    {   // degree extremes for the info struct
        int32_t top_u = 0, top_i = 0, hu = 0, hi = 0;
        PLAN_CUDA(cudaMemcpyAsync(&top_u, sorted_u, 4, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaMemcpyAsync(&top_i, sorted_i, 4, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaStreamSynchronize(st));
        PLAN_CUDA(cudaMemcpyAsync(&hu, deg_u + top_u, 4, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaMemcpyAsync(&hi, deg_i + top_i, 4, cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaStreamSynchronize(st));
        p->max_user_degree = hu;
        p->max_item_degree = hi;
    }

    if (n > 0) {
        uint64_t *keys_a = nullptr, *keys_b = nullptr;
        int32_t *idx_a = nullptr;
        void *tmp = nullptr;
        size_t tmp_bytes = 0;
        PLAN_CUDA(dalloc((void **)&keys_a, nn * 8)); scratch.push_back(keys_a);
        PLAN_CUDA(dalloc((void **)&keys_b, nn * 8)); scratch.push_back(keys_b);
        PLAN_CUDA(dalloc((void **)&idx_a, nn * 4)); scratch.push_back(idx_a);
        k_keys<<<grid_for(n), 256, 0, st>>>(d_u, d_i, n, p->ustripe, p->iworker, p->islot, W, keys_a, idx_a);
        PLAN_CUDA(cudaGetLastError());
        int end_bit = std::min(64, kPlanWorkerShift + bits_for((uint64_t)W));
        PLAN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, idx_a, p->sidx, (int)n, 0,
                                                  end_bit, st));
        PLAN_CUDA(dalloc(&tmp, tmp_bytes)); scratch.push_back(tmp);
        PLAN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, idx_a, p->sidx, (int)n, 0,
                                                  end_bit, st));
        k_gather<<<grid_for(n), 256, 0, st>>>(keys_b, p->sidx, n, d_u, d_i, d_r, p->su, p->si, p->sslot,
                                               p->sr, p->sstep);
        PLAN_CUDA(cudaGetLastError());
        k_build_records<<<grid_for(n), 256, 0, st>>>(keys_b, p->su, p->sr, n, n_users, p->rec);
        PLAN_CUDA(cudaGetLastError());
        k_worker_bounds<<<(W + 1 + 255) / 256, 256, 0, st>>>(keys_b, n, W, p->wbeg);
        PLAN_CUDA(cudaGetLastError());
    } else {
        PLAN_CUDA(cudaMemsetAsync(p->wbeg, 0, sizeof(int64_t) * (size_t)(W + 1), st));
    }
    {
        std::vector<int64_t> h_wbeg((size_t)W + 1);
        PLAN_CUDA(cudaMemcpyAsync(h_wbeg.data(), p->wbeg, sizeof(int64_t) * (size_t)(W + 1),
                                  cudaMemcpyDeviceToHost, st));
        PLAN_CUDA(cudaStreamSynchronize(st));
        int64_t mx = 0;
        for (int32_t w = 0; w < W; ++w) mx = std::max(mx, h_wbeg[w + 1] - h_wbeg[w]);
        p->max_worker_ratings = mx;
    }
    for (void *q : scratch) cudaFree(q);
    scratch.clear();
#undef PLAN_CUDA
    p->epoch = 0;
    *out = p;
    return MFK_OK;
}

extern "C" int mfk_plan_get_info(const mfk_plan *plan, mfk_plan_info *info) {
    MFK_REQUIRE(plan && info, "mfk_plan_get_info: null argument");
    info->n = plan->n;
    info->n_users = plan->n_users;
    info->n_items = plan->n_items;
    info->n_workers = plan->W;
    info->n_ctas = plan->n_ctas;
    info->warps_per_cta = plan->warps_per_cta;
    info->max_items_per_worker = plan->max_slots;
    info->max_worker_ratings = plan->max_worker_ratings;
    info->max_item_degree = plan->max_item_degree;
    info->max_user_degree = plan->max_user_degree;
    return MFK_OK;
}

extern "C" int mfk_plan_order(const mfk_plan *plan, int64_t *d_order, void *stream) {
    MFK_REQUIRE(plan && (d_order || plan->n == 0), "mfk_plan_order: null argument");
    if (plan->n == 0) return MFK_OK;
    cudaStream_t st = as_stream(stream);
    const int64_t n = plan->n;
    int32_t *pos_in = nullptr, *pos_out = nullptr, *step_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    MFK_CUDA(cudaMalloc(&pos_in, (size_t)n * 4));
    MFK_CUDA(cudaMalloc(&pos_out, (size_t)n * 4));
    MFK_CUDA(cudaMalloc(&step_out, (size_t)n * 4));
    k_pos_iota<<<grid_for(n), 256, 0, st>>>(pos_in, n);
    MFK_LAUNCH_CHECK();
    int end_bit = bits_for((uint64_t)plan->W);
    // stable sort by step: positions are worker-major already, so the result is
    // (step, worker, in-block position) -- every step is one conflict-free wave.
    MFK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, plan->sstep, step_out, pos_in, pos_out, (int)n, 0,
                                             end_bit, st));
    MFK_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    MFK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, plan->sstep, step_out, pos_in, pos_out, (int)n, 0,
                                             end_bit, st));
    k_order_out<<<grid_for(n), 256, 0, st>>>(pos_out, plan->sidx, n, d_order);
    MFK_LAUNCH_CHECK();
    MFK_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    cudaFree(pos_in);
    cudaFree(pos_out);
    cudaFree(step_out);
    return MFK_OK;
}

extern "C" int mfk_plan_stats(const mfk_plan *plan, int64_t *d_stats, void *stream) {
    MFK_REQUIRE(plan && d_stats, "mfk_plan_stats: null argument");
    MFK_CUDA(cudaMemcpyAsync(d_stats, plan->stats, sizeof(long long) * 12 * (size_t)plan->W, cudaMemcpyDeviceToDevice,
                             as_stream(stream)));
    return MFK_OK;
}

extern "C" int mfk_plan_assignment(const mfk_plan *plan, int32_t *d_worker, int32_t *d_step, void *stream) {
    MFK_REQUIRE(plan && ((d_worker && d_step) || plan->n == 0), "mfk_plan_assignment: null argument");
    if (plan->n == 0) return MFK_OK;
    k_assignment<<<grid_for(plan->n), 256, 0, as_stream(stream)>>>(plan->wbeg, plan->W, plan->sidx, plan->sstep,
                                                                  plan->n, d_worker, d_step);
    MFK_LAUNCH_CHECK();
    return MFK_OK;
}
