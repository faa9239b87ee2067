/*
 * mfk.h -- C ABI of libmfk_b200.so: the B200 (sm_100a) core that replaces the reference's
 * module-level numba functions (the de-facto operator boundary).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ / torch types.
 *   - every function returns an int status (MFK_OK == 0); mfk_last_error() returns a
 *     thread-local message for the last non-zero status.
 *   - d_* pointers are DEVICE pointers, h_* pointers are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls
 *     taking a stream are asynchronous w.r.t. the host unless stated otherwise.
 *   - factor matrices are fp32 row-major with row stride `ld` floats (ld >= n_factors,
 *     ld % 4 == 0, base pointer 16-byte aligned, padding columns must be zero);
 *     ids are int32, ratings fp32.
 *   - plans / csr handles own their device memory (created once per fit, like an FFT
 *     plan); the per-epoch and per-query calls allocate nothing.
 *
 * Each entry point cites the reference function it stands in for (paths under the
 * reference repository root).
 */
#ifndef MFK_H_
#define MFK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFK_ABI_VERSION 1

#define MFK_OK 0
#define MFK_ERR_ARG 1         /* bad argument (null pointer, bad size, bad enum) */
#define MFK_ERR_CUDA 2        /* a CUDA runtime call failed; see mfk_last_error() */
#define MFK_ERR_UNSUPPORTED 3 /* valid request this build cannot serve (e.g. n_factors too large) */
#define MFK_ERR_NO_DEVICE 4   /* no sm_100 device / kernel image not loadable */

/* matrix_factorization/kernels.py: kernel_linear / kernel_sigmoid / kernel_rbf */
#define MFK_KERNEL_LINEAR 0
#define MFK_KERNEL_SIGMOID 1
#define MFK_KERNEL_RBF 2

#define MFK_MAX_FACTORS 1024

const char *mfk_last_error(void);
int mfk_abi_version(void);
/* sm count, compute capability (major*10+minor) and opt-in shared memory of `device`. */
int mfk_device_query(int device, int *sm_count, int *cc, size_t *smem_optin);

/* ------------------------------------------------------------------------------------
 * Stratified conflict-free SGD plan.
 *
 * Replaces the per-epoch `np.random.shuffle(X)` + sequential rating loop of
 * kernel_matrix_factorization.py:369-425 and baseline_model.py:250-266 by a static DSGD
 * schedule: items are dealt to W worker warps (nnz-balanced), users to R = c * W stripes; at
 * step s worker w owns user stripe (c * w + s) mod R and may enter it once worker w + 1 has
 * finished its step s - c, so no two ratings in flight share a user or an item.  The plan holds the ratings re-ordered per worker (step-major, item-minor).
 * ---------------------------------------------------------------------------------- */
typedef struct mfk_plan mfk_plan;

typedef struct {
    int32_t n_workers;     /* 0 = choose from n / n_users / n_items and the device */
    int32_t warps_per_cta; /* 0 = choose */
    int32_t n_factors;     /* hint: n_factors the plan will run with (caps warps_per_cta: rows
                              wider than 256 / 512 floats need 512- / 256-thread CTAs); 0 = unknown */
    uint32_t hot_min_degree; /* items rated at least this often (at most 32 per SM) are split off into a "hot"
                              sub-plan whose chains a whole CTA resolves as exact mini-batches (linear kernel);
                              0 = default (4096), 0xffffffff = never split */
    uint32_t stripe_slack; /* user stripes per worker (steps per epoch = stripe_slack * n_workers): step s of a
                              worker needs step s - stripe_slack of its ring neighbour, so a worker may run
                              stripe_slack - 1 steps ahead of the hand-off; 0 = default */
    uint32_t schedule;     /* 0 = default: 3 where it applies, else 2.
                              1 = dataflow ring -- worker warps, every rating waits for exactly the previous rating
                              of its user (per-user version counters in global memory);
                              2 = ring -- worker warps hand whole user stripes around in lockstep;
                              3 = flat -- one CTA per worker, the worker's item rows in shared memory, stripes
                              handed around the ring of CTAs; inside a (worker, step) cell every rating waits for
                              the previous rating of its user and of its item only (shared-memory progress
                              counters).  Needs n_factors in 1..256 and the item rows of a worker to fit into shared
                              memory (n_workers = CTAs, at most one per SM; 0 = choose), otherwise the plan falls
                              back to 2.  The default only picks it with n_workers / warps_per_cta left 0 */
    uint32_t no_hot_users; /* 1 = split off hot items only (by default the most active users among the remaining
                              ratings get the same treatment, with the roles of users and items exchanged) */
} mfk_plan_opts;

typedef struct {
    int64_t n;
    int32_t n_users, n_items;
    int32_t n_workers;     /* W: worker warps */
    int32_t n_ctas, warps_per_cta;
    int32_t max_items_per_worker;
    int64_t max_worker_ratings; /* longest worker list (load balance / critical path) */
    int64_t max_item_degree, max_user_degree;
    int32_t n_hot_items;   /* items in the hot sub-plan (0 = no split) */
    int32_t n_steps;       /* R: user stripes == steps per epoch (stripe_slack * n_workers) */
    int64_t n_hot_ratings;
    int32_t n_hot_users;   /* users in the hot-user sub-plan */
    int32_t flat;          /* 1: the main plan is a flat plan (schedule 3) */
    int64_t n_hot_user_ratings;
    int32_t n_hot_workers;      /* CTAs of the hot-item sub-plan (each owns up to hot_max_slots items) */
    int32_t n_hot_user_workers; /* CTAs of the hot-user sub-plan */
    int32_t hot_max_slots;
    int32_t hot_parallel;        /* 1: the hot-item and the hot-user phase touch disjoint rows and run side by side (two streams) */
} mfk_plan_info;

/* d_u/d_i/d_r: the n ratings as internal ids (0..n_users-1 / 0..n_items-1).  Synchronises
 * `stream` before returning (the auto-sizing reads degree statistics back). */
int mfk_plan_create(mfk_plan **out, const int32_t *d_u, const int32_t *d_i, const float *d_r,
                    int64_t n, int32_t n_users, int32_t n_items, const mfk_plan_opts *opts,
                    void *stream);
int mfk_plan_destroy(mfk_plan *plan);
int mfk_plan_get_info(const mfk_plan *plan, mfk_plan_info *info);
/* d_order[n] (int64, device): indices into the arrays given to mfk_plan_create, in an order
 * whose SEQUENTIAL replay is equivalent to one parallel epoch of this plan (step-major,
 * worker-minor; SURVEY.md 8c "same-order protocol"). */
int mfk_plan_order(const mfk_plan *plan, int64_t *d_order, void *stream);
/* per original rating: owning worker and step (int32, device) -- used by the tests to prove
 * that each (step) wave is conflict-free. */
int mfk_plan_assignment(const mfk_plan *plan, int32_t *d_worker, int32_t *d_step, void *stream);

/* Diagnostics of the last SGD epoch run on the plan: d_stats is int64[12 * (n_workers + n_hot_workers +
 * n_hot_user_workers)] (device; the blocks of the hot-item and the hot-user sub-plan follow the main block, same layout):
 * first [n_workers][4] = {SM cycles from start to last rating, cycles blocked in ring hand-off
 * waits, 4-rating chains resolved at once, ratings processed singly}, then [n_workers][8] phase
 * cycle counters that only builds with -DMFK_RING_PROFILE=1 fill in. */
int mfk_plan_stats(const mfk_plan *plan, int64_t *d_stats, void *stream);
/* Diagnostics: which phases of a split plan the SGD epoch entry points run (bit 0 hot items, bit 1 hot users,
 * bit 2 the rest; default 7).  Lets a benchmark time the three kernels of an epoch one by one. */
int mfk_plan_set_phases(mfk_plan *plan, uint32_t mask);

/* ------------------------------------------------------------------------------------
 * KernelMF.  One epoch of kernel_matrix_factorization.py:374-425 (the rating loop of _sgd,
 * update rules kernels.py:108-327) over the plan's schedule, in place.
 * ---------------------------------------------------------------------------------- */
int mfk_kmf_sgd_epoch(mfk_plan *plan, int kernel, float *d_P, float *d_Q, float *d_bu,
                      float *d_bi, int32_t n_factors, int32_t ld, float global_mean, float lr,
                      float reg, float gamma, float min_rating, float max_rating,
                      int update_user_params, int update_item_params, void *stream);

/* Sum of squared errors of UNCLIPPED predictions over n ratings
 * (kernel_matrix_factorization.py:240-317, _calculate_rmse = sqrt(sse / n)).
 * d_ws: workspace of mfk_sse_workspace_bytes() bytes; d_sse: one double on the device. */
size_t mfk_sse_workspace_bytes(void);
int mfk_kmf_sse(int kernel, const int32_t *d_u, const int32_t *d_i, const float *d_r, int64_t n,
                const float *d_P, const float *d_Q, const float *d_bu, const float *d_bi,
                int32_t n_factors, int32_t ld, float global_mean, float gamma, float min_rating,
                float max_rating, void *d_ws, double *d_sse, void *stream);
/* Same, over the ratings stored in the plan (no need to keep the COO arrays around). */
int mfk_kmf_sse_plan(const mfk_plan *plan, int kernel, const float *d_P, const float *d_Q,
                     const float *d_bu, const float *d_bi, int32_t n_factors, int32_t ld,
                     float global_mean, float gamma, float min_rating, float max_rating,
                     void *d_ws, double *d_sse, void *stream);

/* kernel_matrix_factorization.py:448-541 (_predict).  id -1 = unknown (bias 0, zero vector).
 * d_pred[n] fp32, d_possible[n] uint8 (1 iff both ids known). */
int mfk_kmf_predict(int kernel, const int32_t *d_u, const int32_t *d_i, int64_t n,
                    const float *d_P, const float *d_Q, const float *d_bu, const float *d_bi,
                    int32_t n_factors, int32_t ld, float global_mean, float gamma,
                    float min_rating, float max_rating, int bound_ratings, float *d_pred,
                    uint8_t *d_possible, void *stream);

/* ------------------------------------------------------------------------------------
 * BaselineModel.
 * ---------------------------------------------------------------------------------- */
/* One epoch of baseline_model.py:255-266 (bias SGD rating loop) over the plan's schedule. */
int mfk_bias_sgd_epoch(mfk_plan *plan, float *d_bu, float *d_bi, float global_mean, float lr,
                       float reg, int update_user_params, int update_item_params, void *stream);

/* CSR (by user) + CSC (by item) layout of the ratings, built once on the GPU. */
typedef struct mfk_csr mfk_csr;
int mfk_csr_create(mfk_csr **out, const int32_t *d_u, const int32_t *d_i, const float *d_r,
                   int64_t n, int32_t n_users, int32_t n_items, void *stream);
int mfk_csr_destroy(mfk_csr *csr);
/* Copies the layout into caller-provided DEVICE buffers (any may be NULL to skip):
 * row_ptr[n_users+1] int64, col[n] int32, val[n] fp32 sorted by (user, item);
 * col_ptr[n_items+1] int64, row[n] int32, cval[n] fp32 sorted by (item, user). */
int mfk_csr_export(const mfk_csr *csr, int64_t *d_row_ptr, int32_t *d_col, float *d_val,
                   int64_t *d_col_ptr, int32_t *d_row, float *d_cval, void *stream);
/* One epoch of baseline_model.py:326-348 (_als): user pass from zeros, then item pass with the
 * new user biases; two segmented reductions. */
int mfk_bias_als_epoch(const mfk_csr *csr, float *d_bu, float *d_bi, float global_mean, float reg,
                       void *stream);
/* baseline_model.py:183-212 (_calculate_rmse) as SSE, and :365-417 (_predict). */
int mfk_bias_sse(const int32_t *d_u, const int32_t *d_i, const float *d_r, int64_t n,
                 const float *d_bu, const float *d_bi, float global_mean, void *d_ws,
                 double *d_sse, void *stream);
int mfk_bias_predict(const int32_t *d_u, const int32_t *d_i, int64_t n, const float *d_bu,
                     const float *d_bi, float global_mean, float min_rating, float max_rating,
                     int bound_ratings, float *d_pred, uint8_t *d_possible, void *stream);

/* ------------------------------------------------------------------------------------
 * Scoring: recommender_base.py:214-271 (recommend) batched over users.
 * For each of the m users in d_users: score every item with the model's kernel (UNBOUNDED),
 * drop the user's masked items, keep the top k (descending score, ties by lower item id),
 * then clip if bound_ratings.  Masked items of user j are d_mask_items[d_mask_ptr[j] ..
 * d_mask_ptr[j+1]) (internal item ids, sorted ascending inside each row); d_mask_ptr may be NULL for no mask.
 * For k <= 64 the contraction runs on the tensor cores (tcgen05, split-TF32, TMA) with mask and top-k
 * fused into the epilogue; larger k (or MFK_SCORE_SIMT=1) use the fp32 SIMT path.
 * Output rows are padded with item -1 / score -inf when fewer than k candidates remain.
 * d_ws: workspace of mfk_score_workspace_bytes(m, n_items, n_factors, k) bytes.
 * ---------------------------------------------------------------------------------- */
size_t mfk_score_workspace_bytes(int64_t m, int32_t n_items, int32_t n_factors, int32_t k);
int mfk_score_topk(int kernel, const int32_t *d_users, int64_t m, const float *d_P,
                   const float *d_Q, const float *d_bu, const float *d_bi, int32_t n_items,
                   int32_t n_factors, int32_t ld, float global_mean, float gamma,
                   float min_rating, float max_rating, const int64_t *d_mask_ptr,
                   const int32_t *d_mask_items, int32_t k, int bound_ratings, float *d_scores,
                   int32_t *d_items, void *d_ws, void *stream);

/* Merge step of the item-sharded recommend: d_scores_in / d_items_in are [m][c] candidate lists (the
 * all-gathered per-shard top-k lists: UNBOUNDED scores, global item ids, item < 0 = padding); keeps
 * the best k per user (score descending, ties by lower item id) and clips if bound_ratings. */
int mfk_topk_merge(const float *d_scores_in, const int32_t *d_items_in, int64_t m, int32_t c, int32_t k,
                   int bound_ratings, float min_rating, float max_rating, float *d_scores, int32_t *d_items,
                   void *stream);

/* ------------------------------------------------------------------------------------
 * Preprocessing on the GPU -- replaces the id-mapping part of RecommenderBase._preprocess_data for integer raw ids
 * (recommender_base.py:125-164): first-appearance internal ids on the shuffled rows, duplicate check.
 *
 * mfk_first_appearance: shuffled[k] = d_raw[d_perm[k]] (d_perm NULL: identity); d_internal[k] = rank of shuffled[k] among
 * the distinct ids ordered by first occurrence (what `{id: index}` built by a loop over the shuffled rows gives,
 * :133-140); d_unique[r] = the id of rank r (capacity n); *h_n_unique = number of distinct ids.  Allocates its scratch
 * internally and synchronises the stream.
 * mfk_has_duplicate_pairs: *h_flag = 1 if some (d_u[k], d_i[k]) occurs more than once (:127-128). */
int mfk_first_appearance(const int64_t *d_raw, const int64_t *d_perm, int64_t n, int32_t *d_internal, int64_t *d_unique,
                         int32_t *h_n_unique, void *stream);
int mfk_has_duplicate_pairs(const int32_t *d_u, const int32_t *d_i, int64_t n, int64_t n_items, int32_t *h_flag,
                            void *stream);

/* ------------------------------------------------------------------------------------
 * Host-buffer entry points (what a cgo / JNI / ctypes binding without torch would call).
 * All pointers are HOST pointers; the call allocates device memory, copies in, runs, copies
 * back and synchronises before returning.
 * ---------------------------------------------------------------------------------- */
/* kernel_matrix_factorization.py:320-445 (_sgd): n_epochs epochs, parameters updated in
 * place, h_train_rmse[n_epochs] filled.  h_order (nullable, int64[n]) receives the plan's
 * replay order. */
int mfk_kmf_sgd_host(int kernel, const int32_t *h_u, const int32_t *h_i, const float *h_r,
                     int64_t n, int32_t n_users, int32_t n_items, float *h_P, float *h_Q,
                     float *h_bu, float *h_bi, int32_t n_factors, int32_t ld, float global_mean,
                     int32_t n_epochs, float lr, float reg, float gamma, float min_rating,
                     float max_rating, int update_user_params, int update_item_params,
                     const mfk_plan_opts *opts, double *h_train_rmse, int64_t *h_order);
/* baseline_model.py:215-280 (_sgd) and :283-362 (_als). */
int mfk_bias_sgd_host(const int32_t *h_u, const int32_t *h_i, const float *h_r, int64_t n,
                      int32_t n_users, int32_t n_items, float *h_bu, float *h_bi,
                      float global_mean, int32_t n_epochs, float lr, float reg,
                      int update_user_params, int update_item_params, const mfk_plan_opts *opts,
                      double *h_train_rmse, int64_t *h_order);
int mfk_bias_als_host(const int32_t *h_u, const int32_t *h_i, const float *h_r, int64_t n,
                      int32_t n_users, int32_t n_items, float *h_bu, float *h_bi,
                      float global_mean, int32_t n_epochs, float reg, double *h_train_rmse);

#ifdef __cplusplus
}
#endif
#endif /* MFK_H_ */
