// Internal layout of the opaque mfk_plan handle (see mfk_plan.cu for how it is built).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mfk {
// sort key: [worker | step | layer | slot]
constexpr int kPlanWorkerShift = 40;
constexpr int kPlanStepShift = 24;
constexpr int kPlanLayerShift = 22;  // inside the low 24 bits: [layer (2) | slot (22)]
// ctrl word of a schedule record: flags in the high bits; the low bits hold the step (ring schedule, 16 bits) or
// the number of earlier ratings of the same user in the emitted order (dataflow schedule, 25 bits)
constexpr int32_t kCtrlDup = 1 << 30;      // this user occurs among the previous 15 records (do not prefetch its row)
constexpr int32_t kCtrlQuad = 1 << 29;     // records k..k+3 share worker, step and item, none is kCtrlDup
constexpr int32_t kCtrlNewStep = 1 << 28;  // first record of a (worker, step) block
constexpr int32_t kCtrlNewItem = 1 << 27;  // item differs from the previous record of this worker
constexpr int32_t kCtrlOwn = 1 << 26;      // the user's previous rating is one of this worker's previous 15 records
constexpr int32_t kNeedMask = (1 << 25) - 1;
// "flat" plans (CTA workers, batches of independent ratings, k_sgd_flat): sort key
//   [worker (16) | step (16) | rank of the rating among its user's ratings in the cell (10) |
//    rank among its item's ratings in the cell (10) | slot (12)]
// -- neighbours in the list are ratings of different users and items wherever the cell allows it
constexpr int kFlatWorkerShift = 48, kFlatStepShift = 32, kFlatRankUShift = 22, kFlatRankIShift = 12;
constexpr int32_t kFlatMaxSlots = 1 << 12, kFlatMaxRank = 1 << 10;
constexpr int32_t kFlatDefaultSlack = 2;  // stripes per worker of a flat plan when mfk_plan_opts.stripe_slack is 0
// shared memory of k_sgd_flat<*, NV, NBUF>: nbuf row buffers of C rows with their records and user biases, the
// worker's item rows and biases, control words (C = 64 for rows of up to 128 floats, 32 up to 256)
inline size_t flat_smem_bytes(int nv, int max_slots, int nbuf) {
    const size_t fw = 128 * (size_t)nv, c = nv == 1 ? 64 : 32;
    return 4 * ((size_t)nbuf * c * (fw + 5) + (size_t)max_slots * (fw + 2) + 8 + 64);
}
// row buffers a flat plan can afford (3 preferred, 2 minimum, 0 = the worker's item rows do not fit: no flat plan)
inline int flat_row_buffers(int n_factors, int max_slots, size_t smem_optin) {
    const int f4 = (n_factors + 3) & ~3;
    if (n_factors < 1 || f4 > 256 || max_slots >= kFlatMaxSlots) return 0;
    const int nv = f4 <= 128 ? 1 : 2;
    for (int nbuf = 3; nbuf >= 2; --nbuf)
        if (flat_smem_bytes(nv, max_slots, nbuf) + 1024 <= smem_optin) return nbuf;
    return 0;
}
constexpr int32_t kHotSlotsMax = 32;        // hot ids (items / users) per worker of the batch engine at most
constexpr int32_t kDefaultSlack = 1;       // stripes per worker when mfk_plan_opts.stripe_slack is 0
}  // namespace mfk

struct mfk_plan {
    int64_t n = 0;
    int32_t n_users = 0, n_items = 0;
    int32_t W = 0;  // worker warps
    int32_t slack = 1;  // stripes per worker: step s of worker w needs step s - slack of worker w + 1
    int32_t R = 0;      // user stripes == steps per epoch == slack * W
    int32_t n_ctas = 0, warps_per_cta = 0;
    int32_t max_slots = 0;  // items per worker (rows of witems)
    int64_t max_worker_ratings = 0, max_item_degree = 0, max_user_degree = 0;
    // ratings sorted by (worker, step, slot); all device arrays of length n
    int32_t *su = nullptr;     // user id
    int32_t *si = nullptr;     // item id
    int32_t *sslot = nullptr;  // item's slot inside its worker (index into the worker's smem stripe)
    float *sr = nullptr;       // rating
    int32_t *sstep = nullptr;  // step
    int4 *rec = nullptr;       // {user, slot, rating bits, ctrl}: what the SGD kernel streams
    int32_t *sidx = nullptr;   // index into the arrays given to mfk_plan_create
    int64_t *wbeg = nullptr;   // [W+1] list bounds per worker
    int32_t *cbeg = nullptr;   // flat plans: [W][R + 1] first list position of every (worker, step) cell
    int32_t *witems = nullptr; // [max_slots][W] item id owned by (slot, worker) or -1
    int32_t *iworker = nullptr, *islot = nullptr;  // per item
    int32_t *ustripe = nullptr;                    // per user
    int32_t *flags = nullptr;  // [W] ring progress flags (monotone across epochs)
    int32_t flat = 0;          // 1: flat plan -- workers are CTAs, records are grouped into conflict-free batches (k_sgd_flat)
    int32_t flow = 0;          // 1: dataflow schedule (records carry the user's version, uver is the live counter)
    int32_t *uver = nullptr;   // [n_users] ratings of each user applied so far in this epoch (dataflow schedule)
    // hot/cold split (optional): `hot` is a plan over the ratings of the most-rated items (one item per
    // worker, worker = one CTA); this plan then covers the remaining ratings.  n_total counts both.
    mfk_plan *hot = nullptr;
    // `hot_users`: likewise for the most active users among the remaining ratings, built with the roles of users
    // and items exchanged (swapped = 1: su holds item ids, si user ids, the workers own users)
    mfk_plan *hot_users = nullptr;
    int32_t n_hot_users = 0, swapped = 0;
    // hot_parallel = 1: no rating of a hot item by a hot user is in `hot` (they are part of this plan), the two hot phases touch
    // disjoint rows and run side by side -- `hot_users` on aux_stream between the two events
    int32_t hot_parallel = 0;
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    uint32_t phases = 7;  // diagnostics: bit 0 hot items, bit 1 hot users, bit 2 the rest (mfk_plan_set_phases)
    int64_t n_total = 0;
    int32_t n_hot_items = 0;
    long long *stats = nullptr;  // [W][4] per-worker counters of the last SGD epoch (diagnostics)
    int64_t epoch = 0;         // epochs run so far (flag base = epoch * (R + 1))
};
