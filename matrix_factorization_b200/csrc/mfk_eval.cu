// Read-only passes: sum of squared errors (train_rmse) and predict.
//
//   mfk_kmf_sse / mfk_kmf_sse_plan : kernel_matrix_factorization.py:240-317 (_calculate_rmse)
//   mfk_kmf_predict                : kernel_matrix_factorization.py:448-541 (_predict)
//   mfk_bias_sse / mfk_bias_predict: baseline_model.py:183-212, :365-417
//
// One warp per rating, 128-bit coalesced row loads, shuffle-reduced dot product; the SSE is
// accumulated in double per warp, reduced per block, and summed in a fixed order by a second
// tiny kernel (deterministic).
#include "mfk_common.cuh"
#include "mfk_plan.h"

namespace mfk {

constexpr int kEvalThreads = 256;
constexpr int kEvalMaxBlocks = 148 * 8;

struct EvalParams {
    const float *P, *Q, *bu, *bi;
    int32_t F, ld;
    float mu, gamma, a, c;
};

// dot (linear/sigmoid) or squared distance (rbf) of two rows; -1 ids give the zero vector
template <int KERNEL>
__device__ __forceinline__ float row_reduce(const EvalParams &e, int32_t u, int32_t i, int lane) {
    float acc = 0.f;
    const float *p = e.P + (size_t)(u < 0 ? 0 : u) * e.ld;
    const float *q = e.Q + (size_t)(i < 0 ? 0 : i) * e.ld;
    for (int c = 4 * lane; c < e.F; c += 128) {
        float4 pv = (u < 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4 *>(p + c));
        float4 qv = (i < 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4 *>(q + c));
        if (KERNEL == MFK_KERNEL_RBF) {
            float dx = pv.x - qv.x, dy = pv.y - qv.y, dz = pv.z - qv.z, dw = pv.w - qv.w;
            acc = fmaf(dx, dx, acc);
            acc = fmaf(dy, dy, acc);
            acc = fmaf(dz, dz, acc);
            acc = fmaf(dw, dw, acc);
        } else {
            acc = fmaf(pv.x, qv.x, acc);
            acc = fmaf(pv.y, qv.y, acc);
            acc = fmaf(pv.z, qv.z, acc);
            acc = fmaf(pv.w, qv.w, acc);
        }
    }
    return warp_sum(acc);
}

template <int KERNEL, bool BIAS_ONLY>
__global__ void __launch_bounds__(kEvalThreads) k_sse(const int32_t *__restrict__ u, const int32_t *__restrict__ i,
                                                      const float *__restrict__ r, int64_t n, EvalParams e,
                                                      double *partial) {
    __shared__ double s_part[kEvalThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (kEvalThreads / 32);
    double acc = 0.0;
    // each warp takes 32 consecutive ratings at a time: one coalesced record load, then 32 gathers
    for (int64_t b0 = ((int64_t)blockIdx.x * (kEvalThreads / 32) + warp) * 32; b0 < n; b0 += warps_total * 32) {
        int64_t k = b0 + lane;
        int32_t ru = k < n ? u[k] : 0, ri = k < n ? i[k] : 0;
        float rr = k < n ? r[k] : 0.f;
        int cnt = (int)min((int64_t)32, n - b0);
        if (BIAS_ONLY) {
            if (k < n) {
                float err = rr - (e.mu + e.bu[ru] + e.bi[ri]);
                acc += (double)err * (double)err;
            }
        } else {
            for (int j = 0; j < cnt; ++j) {
                int32_t uu = __shfl_sync(0xffffffffu, ru, j), ii = __shfl_sync(0xffffffffu, ri, j);
                float red = row_reduce<KERNEL>(e, uu, ii, lane);
                if (lane == j) {
                    float ub = (KERNEL == MFK_KERNEL_RBF) ? 0.f : e.bu[uu];
                    float ib = (KERNEL == MFK_KERNEL_RBF) ? 0.f : e.bi[ii];
                    float err = rr - kmf_predict_from_dot(KERNEL, e.mu, ub, ib, red, e.gamma, e.a, e.c);
                    acc += (double)err * (double)err;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kEvalThreads / 32; ++w) t += s_part[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void k_sse_final(const double *partial, int nblocks, double *out) {
    // single warp, fixed order
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) acc += partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) *out = acc;
}

template <int KERNEL, bool BIAS_ONLY>
__global__ void __launch_bounds__(kEvalThreads) k_predict(const int32_t *__restrict__ u,
                                                          const int32_t *__restrict__ i, int64_t n, EvalParams e,
                                                          float lo, float hi, int bound, float *pred,
                                                          uint8_t *possible) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (kEvalThreads / 32);
    for (int64_t b0 = ((int64_t)blockIdx.x * (kEvalThreads / 32) + warp) * 32; b0 < n; b0 += warps_total * 32) {
        int64_t k = b0 + lane;
        int32_t ru = k < n ? u[k] : -1, ri = k < n ? i[k] : -1;
        int cnt = (int)min((int64_t)32, n - b0);
        float v = 0.f;
        if (BIAS_ONLY) {
            v = e.mu;  // baseline_model.py:399-405
            if (ru >= 0) v += e.bu[ru];
            if (ri >= 0) v += e.bi[ri];
        } else {
            for (int j = 0; j < cnt; ++j) {
                int32_t uu = __shfl_sync(0xffffffffu, ru, j), ii = __shfl_sync(0xffffffffu, ri, j);
                float red = row_reduce<KERNEL>(e, uu, ii, lane);
                if (lane == j) {
                    float ub = (KERNEL == MFK_KERNEL_RBF || uu < 0) ? 0.f : e.bu[uu];
                    float ib = (KERNEL == MFK_KERNEL_RBF || ii < 0) ? 0.f : e.bi[ii];
                    v = kmf_predict_from_dot(KERNEL, e.mu, ub, ib, red, e.gamma, e.a, e.c);
                }
            }
        }
        if (k < n) {
            if (bound) v = v > hi ? hi : (v < lo ? lo : v);  // kmf:532-536
            pred[k] = v;
            possible[k] = (uint8_t)((ru >= 0) && (ri >= 0));
        }
    }
}

static int eval_blocks(int64_t n) {
    int64_t per_block = (kEvalThreads / 32) * 32;
    int64_t b = (n + per_block - 1) / per_block;
    return (int)max((int64_t)1, min(b, (int64_t)kEvalMaxBlocks));
}

static int check_factors(const char *who, const float *P, const float *Q, const float *bu, const float *bi,
                         int32_t F, int32_t ld) {
    MFK_REQUIRE(P && Q && bu && bi, "%s: null parameter array", who);
    MFK_REQUIRE(F >= 1 && ld >= F && ld % 4 == 0, "%s: need 1 <= n_factors <= ld and ld %% 4 == 0 (F=%d ld=%d)", who,
                F, ld);
    MFK_REQUIRE((((uintptr_t)P | (uintptr_t)Q) & 15) == 0, "%s: P/Q must be 16-byte aligned", who);
    return MFK_OK;
}

struct RatingSeg {
    const int32_t *u, *i;
    const float *r;
    int64_t n;
};

// Up to kMaxSegs rating segments (the phases of a split plan); the partial sums go to consecutive workspace
// slots and are added in a fixed order.
constexpr int kMaxSegs = 3;
static int run_sse(int kernel, bool bias_only, const RatingSeg *segs, int n_segs, const EvalParams &e, void *ws,
                   double *out, cudaStream_t st) {
    MFK_REQUIRE(ws && out, "sse: null workspace/output");
    MFK_REQUIRE(n_segs <= kMaxSegs, "sse: too many segments");
    double *partial = reinterpret_cast<double *>(ws);
    int64_t n_all = 0;
    for (int seg = 0; seg < n_segs; ++seg) n_all += segs[seg].n;
    if (n_all == 0) {
        MFK_CUDA(cudaMemsetAsync(out, 0, sizeof(double), st));
        return MFK_OK;
    }
    int total_blocks = 0;
    for (int seg = 0; seg < n_segs; ++seg) {
        const int32_t *su = segs[seg].u, *si = segs[seg].i;
        const float *sr = segs[seg].r;
        const int64_t sn = segs[seg].n;
        if (sn == 0) continue;
        MFK_REQUIRE(su && si && sr, "sse: null rating arrays");
        const int blocks = eval_blocks(sn);
        double *pp = partial + total_blocks;
        if (bias_only) k_sse<MFK_KERNEL_LINEAR, true><<<blocks, kEvalThreads, 0, st>>>(su, si, sr, sn, e, pp);
        else if (kernel == MFK_KERNEL_LINEAR) k_sse<MFK_KERNEL_LINEAR, false><<<blocks, kEvalThreads, 0, st>>>(su, si, sr, sn, e, pp);
        else if (kernel == MFK_KERNEL_SIGMOID) k_sse<MFK_KERNEL_SIGMOID, false><<<blocks, kEvalThreads, 0, st>>>(su, si, sr, sn, e, pp);
        else k_sse<MFK_KERNEL_RBF, false><<<blocks, kEvalThreads, 0, st>>>(su, si, sr, sn, e, pp);
        MFK_LAUNCH_CHECK();
        total_blocks += blocks;
    }
    k_sse_final<<<1, 32, 0, st>>>(partial, total_blocks, out);
    MFK_LAUNCH_CHECK();
    return MFK_OK;
}

}  // namespace mfk

using namespace mfk;

extern "C" size_t mfk_sse_workspace_bytes(void) { return sizeof(double) * kMaxSegs * (size_t)kEvalMaxBlocks; }

extern "C" int mfk_kmf_sse(int kernel, const int32_t *d_u, const int32_t *d_i, const float *d_r, int64_t n,
                           const float *d_P, const float *d_Q, const float *d_bu, const float *d_bi,
                           int32_t n_factors, int32_t ld, float global_mean, float gamma, float min_rating,
                           float max_rating, void *d_ws, double *d_sse, void *stream) {
    MFK_REQUIRE(kernel >= 0 && kernel <= 2, "mfk_kmf_sse: bad kernel %d", kernel);
    int rc = check_factors("mfk_kmf_sse", d_P, d_Q, d_bu, d_bi, n_factors, ld);
    if (rc) return rc;
    EvalParams e{d_P, d_Q, d_bu, d_bi, (n_factors + 3) & ~3, ld, global_mean, gamma, min_rating,
                 max_rating - min_rating};
    const RatingSeg seg{d_u, d_i, d_r, n};
    return run_sse(kernel, false, &seg, 1, e, d_ws, d_sse, as_stream(stream));
}

// the rating segments of a plan: cold part, hot items, hot users (whose arrays hold the roles exchanged)
static int plan_segments(const mfk_plan *plan, RatingSeg *segs) {
    int k = 0;
    segs[k++] = RatingSeg{plan->su, plan->si, plan->sr, plan->n};
    if (plan->hot) segs[k++] = RatingSeg{plan->hot->su, plan->hot->si, plan->hot->sr, plan->hot->n};
    if (plan->hot_users) segs[k++] = RatingSeg{plan->hot_users->si, plan->hot_users->su, plan->hot_users->sr, plan->hot_users->n};
    return k;
}

extern "C" int mfk_kmf_sse_plan(const mfk_plan *plan, int kernel, const float *d_P, const float *d_Q,
                                const float *d_bu, const float *d_bi, int32_t n_factors, int32_t ld,
                                float global_mean, float gamma, float min_rating, float max_rating, void *d_ws,
                                double *d_sse, void *stream) {
    MFK_REQUIRE(plan != nullptr, "mfk_kmf_sse_plan: plan is NULL");
    MFK_REQUIRE(kernel >= 0 && kernel <= 2, "mfk_kmf_sse_plan: bad kernel %d", kernel);
    int rc = check_factors("mfk_kmf_sse_plan", d_P, d_Q, d_bu, d_bi, n_factors, ld);
    if (rc) return rc;
    EvalParams e{d_P, d_Q, d_bu, d_bi, (n_factors + 3) & ~3, ld, global_mean, gamma, min_rating,
                 max_rating - min_rating};
    RatingSeg segs[kMaxSegs];
    const int n_segs = plan_segments(plan, segs);
    return run_sse(kernel, false, segs, n_segs, e, d_ws, d_sse, as_stream(stream));
}

extern "C" int mfk_kmf_predict(int kernel, const int32_t *d_u, const int32_t *d_i, int64_t n, const float *d_P,
                               const float *d_Q, const float *d_bu, const float *d_bi, int32_t n_factors,
                               int32_t ld, float global_mean, float gamma, float min_rating, float max_rating,
                               int bound_ratings, float *d_pred, uint8_t *d_possible, void *stream) {
    MFK_REQUIRE(kernel >= 0 && kernel <= 2, "mfk_kmf_predict: bad kernel %d", kernel);
    int rc = check_factors("mfk_kmf_predict", d_P, d_Q, d_bu, d_bi, n_factors, ld);
    if (rc) return rc;
    if (n == 0) return MFK_OK;
    MFK_REQUIRE(d_u && d_i && d_pred && d_possible, "mfk_kmf_predict: null array");
    EvalParams e{d_P, d_Q, d_bu, d_bi, (n_factors + 3) & ~3, ld, global_mean, gamma, min_rating,
                 max_rating - min_rating};
    cudaStream_t st = as_stream(stream);
    int blocks = eval_blocks(n);
    if (kernel == MFK_KERNEL_LINEAR)
        k_predict<MFK_KERNEL_LINEAR, false><<<blocks, kEvalThreads, 0, st>>>(d_u, d_i, n, e, min_rating, max_rating, bound_ratings, d_pred, d_possible);
    else if (kernel == MFK_KERNEL_SIGMOID)
        k_predict<MFK_KERNEL_SIGMOID, false><<<blocks, kEvalThreads, 0, st>>>(d_u, d_i, n, e, min_rating, max_rating, bound_ratings, d_pred, d_possible);
    else
        k_predict<MFK_KERNEL_RBF, false><<<blocks, kEvalThreads, 0, st>>>(d_u, d_i, n, e, min_rating, max_rating, bound_ratings, d_pred, d_possible);
    MFK_LAUNCH_CHECK();
    return MFK_OK;
}

extern "C" int mfk_bias_sse(const int32_t *d_u, const int32_t *d_i, const float *d_r, int64_t n, const float *d_bu,
                            const float *d_bi, float global_mean, void *d_ws, double *d_sse, void *stream) {
    MFK_REQUIRE(d_bu && d_bi, "mfk_bias_sse: null bias array");
    EvalParams e{nullptr, nullptr, d_bu, d_bi, 0, 0, global_mean, 0.f, 0.f, 0.f};
    const RatingSeg seg{d_u, d_i, d_r, n};
    return run_sse(0, true, &seg, 1, e, d_ws, d_sse, as_stream(stream));
}

extern "C" int mfk_bias_predict(const int32_t *d_u, const int32_t *d_i, int64_t n, const float *d_bu,
                                const float *d_bi, float global_mean, float min_rating, float max_rating,
                                int bound_ratings, float *d_pred, uint8_t *d_possible, void *stream) {
    MFK_REQUIRE(d_bu && d_bi, "mfk_bias_predict: null bias array");
    if (n == 0) return MFK_OK;
    MFK_REQUIRE(d_u && d_i && d_pred && d_possible, "mfk_bias_predict: null array");
    EvalParams e{nullptr, nullptr, d_bu, d_bi, 0, 0, global_mean, 0.f, 0.f, 0.f};
    k_predict<MFK_KERNEL_LINEAR, true><<<eval_blocks(n), kEvalThreads, 0, as_stream(stream)>>>(
        d_u, d_i, n, e, min_rating, max_rating, bound_ratings, d_pred, d_possible);
    MFK_LAUNCH_CHECK();
    return MFK_OK;
}
