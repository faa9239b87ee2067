// Host-buffer entry points: the call a binding without torch (cgo / JNI / ctypes) makes.
// They stand in for the reference's njit drivers one-for-one:
//   mfk_kmf_sgd_host  <- kernel_matrix_factorization.py:320-445 (_sgd)
//   mfk_bias_sgd_host <- baseline_model.py:215-280 (_sgd)
//   mfk_bias_als_host <- baseline_model.py:283-362 (_als)
// Each allocates device memory, copies in, runs n_epochs epochs (+ the per-epoch RMSE pass),
// copies the parameters back and synchronises before returning.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "mfk_common.cuh"

namespace mfk {

struct DevBufs {
    std::vector<void *> ptrs;
    cudaStream_t st = nullptr;
    ~DevBufs() {
        for (void *p : ptrs)
            if (p) pool_free(p);
        if (st) cudaStreamDestroy(st);
    }
    template <typename T>
    cudaError_t alloc(T **out, size_t count) {
        void *p = nullptr;
        cudaError_t e = pool_malloc(&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = reinterpret_cast<T *>(p);
        return e;
    }
};

// MFK_HOST_TIMING=1: wall-clock of the phases of a host call on stderr (diagnostics)
struct PhaseTimer {
    bool on = getenv("MFK_HOST_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char *what, cudaStream_t st) {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[mfk host] %-10s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

template <typename T>
static cudaError_t up(T *dst, const T *src, size_t count, cudaStream_t st) {
    return cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st);
}
template <typename T>
static cudaError_t down(T *dst, const T *src, size_t count, cudaStream_t st) {
    return cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, st);
}

}  // namespace mfk

using namespace mfk;

extern "C" int mfk_kmf_sgd_host(int kernel, const int32_t *h_u, const int32_t *h_i, const float *h_r, int64_t n,
                                int32_t n_users, int32_t n_items, float *h_P, float *h_Q, float *h_bu, float *h_bi,
                                int32_t n_factors, int32_t ld, float global_mean, int32_t n_epochs, float lr,
                                float reg, float gamma, float min_rating, float max_rating, int update_user_params,
                                int update_item_params, const mfk_plan_opts *opts, double *h_train_rmse,
                                int64_t *h_order) {
    MFK_REQUIRE(kernel >= 0 && kernel <= 2, "mfk_kmf_sgd_host: bad kernel %d", kernel);
    MFK_REQUIRE(n >= 0 && n_users > 0 && n_items > 0 && n_epochs >= 0, "mfk_kmf_sgd_host: bad sizes");
    MFK_REQUIRE(n == 0 || (h_u && h_i && h_r), "mfk_kmf_sgd_host: null rating arrays");
    MFK_REQUIRE(h_P && h_Q && h_bu && h_bi, "mfk_kmf_sgd_host: null parameter array");
    MFK_REQUIRE(n_factors >= 1 && ld >= n_factors && ld % 4 == 0, "mfk_kmf_sgd_host: bad n_factors/ld");
    MFK_REQUIRE(n_epochs == 0 || h_train_rmse, "mfk_kmf_sgd_host: null train_rmse");
    DevBufs d;
    MFK_CUDA(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
    int32_t *du, *di;
    float *dr, *dP, *dQ, *dbu, *dbi;
    double *dsse;
    void *dws;
    int64_t *dorder = nullptr;
    const size_t pn = (size_t)n_users * ld, qn = (size_t)n_items * ld;
    MFK_CUDA(d.alloc(&du, (size_t)n));
    MFK_CUDA(d.alloc(&di, (size_t)n));
    MFK_CUDA(d.alloc(&dr, (size_t)n));
    MFK_CUDA(d.alloc(&dP, pn));
    MFK_CUDA(d.alloc(&dQ, qn));
    MFK_CUDA(d.alloc(&dbu, (size_t)n_users));
    MFK_CUDA(d.alloc(&dbi, (size_t)n_items));
    MFK_CUDA(d.alloc(&dsse, (size_t)(n_epochs > 0 ? n_epochs : 1)));
    MFK_CUDA(d.alloc((unsigned char **)&dws, mfk_sse_workspace_bytes()));
    PhaseTimer tm;
    tm.lap("alloc", d.st);
    if (n > 0) {
        MFK_CUDA(up(du, h_u, (size_t)n, d.st));
        MFK_CUDA(up(di, h_i, (size_t)n, d.st));
        MFK_CUDA(up(dr, h_r, (size_t)n, d.st));
    }
    MFK_CUDA(up(dP, h_P, pn, d.st));
    MFK_CUDA(up(dQ, h_Q, qn, d.st));
    MFK_CUDA(up(dbu, h_bu, (size_t)n_users, d.st));
    MFK_CUDA(up(dbi, h_bi, (size_t)n_items, d.st));
    tm.lap("h2d", d.st);
    mfk_plan *plan = nullptr;
    int rc = mfk_plan_create(&plan, du, di, dr, n, n_users, n_items, opts, d.st);
    if (rc) return rc;
    tm.lap("plan", d.st);
    if (h_order && n > 0) {
        rc = d.alloc(&dorder, (size_t)n) == cudaSuccess ? MFK_OK : MFK_ERR_CUDA;
        if (rc == MFK_OK) rc = mfk_plan_order(plan, dorder, d.st);
        if (rc == MFK_OK && down(h_order, dorder, (size_t)n, d.st) != cudaSuccess) rc = MFK_ERR_CUDA;
    }
    for (int e = 0; rc == MFK_OK && e < n_epochs; ++e) {
        rc = mfk_kmf_sgd_epoch(plan, kernel, dP, dQ, dbu, dbi, n_factors, ld, global_mean, lr, reg, gamma, min_rating,
                               max_rating, update_user_params, update_item_params, d.st);
        if (rc == MFK_OK)
            rc = mfk_kmf_sse_plan(plan, kernel, dP, dQ, dbu, dbi, n_factors, ld, global_mean, gamma, min_rating,
                                  max_rating, dws, dsse + e, d.st);
    }
    if (rc == MFK_OK) {
        cudaError_t ce = cudaSuccess;
        if (ce == cudaSuccess) ce = down(h_P, dP, pn, d.st);
        if (ce == cudaSuccess) ce = down(h_Q, dQ, qn, d.st);
        if (ce == cudaSuccess) ce = down(h_bu, dbu, (size_t)n_users, d.st);
        if (ce == cudaSuccess) ce = down(h_bi, dbi, (size_t)n_items, d.st);
        if (ce == cudaSuccess && n_epochs > 0) ce = down(h_train_rmse, dsse, (size_t)n_epochs, d.st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(d.st);
        if (ce != cudaSuccess) {
            set_error("mfk_kmf_sgd_host: copy-back failed: %s", cudaGetErrorString(ce));
            rc = MFK_ERR_CUDA;
        }
    }
    tm.lap("epochs+d2h", d.st);
    mfk_plan_destroy(plan);
    tm.lap("plan free", d.st);
    if (rc == MFK_OK)
        for (int e = 0; e < n_epochs; ++e) h_train_rmse[e] = n > 0 ? std::sqrt(h_train_rmse[e] / (double)n) : NAN;
    return rc;
}

extern "C" int mfk_bias_sgd_host(const int32_t *h_u, const int32_t *h_i, const float *h_r, int64_t n, int32_t n_users,
                                 int32_t n_items, float *h_bu, float *h_bi, float global_mean, int32_t n_epochs,
                                 float lr, float reg, int update_user_params, int update_item_params,
                                 const mfk_plan_opts *opts, double *h_train_rmse, int64_t *h_order) {
    MFK_REQUIRE(n >= 0 && n_users > 0 && n_items > 0 && n_epochs >= 0, "mfk_bias_sgd_host: bad sizes");
    MFK_REQUIRE(n == 0 || (h_u && h_i && h_r), "mfk_bias_sgd_host: null rating arrays");
    MFK_REQUIRE(h_bu && h_bi, "mfk_bias_sgd_host: null parameter array");
    MFK_REQUIRE(n_epochs == 0 || h_train_rmse, "mfk_bias_sgd_host: null train_rmse");
    DevBufs d;
    MFK_CUDA(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
    int32_t *du, *di;
    float *dr, *dbu, *dbi;
    double *dsse;
    void *dws;
    int64_t *dorder = nullptr;
    MFK_CUDA(d.alloc(&du, (size_t)n));
    MFK_CUDA(d.alloc(&di, (size_t)n));
    MFK_CUDA(d.alloc(&dr, (size_t)n));
    MFK_CUDA(d.alloc(&dbu, (size_t)n_users));
    MFK_CUDA(d.alloc(&dbi, (size_t)n_items));
    MFK_CUDA(d.alloc(&dsse, (size_t)(n_epochs > 0 ? n_epochs : 1)));
    MFK_CUDA(d.alloc((unsigned char **)&dws, mfk_sse_workspace_bytes()));
    if (n > 0) {
        MFK_CUDA(up(du, h_u, (size_t)n, d.st));
        MFK_CUDA(up(di, h_i, (size_t)n, d.st));
        MFK_CUDA(up(dr, h_r, (size_t)n, d.st));
    }
    MFK_CUDA(up(dbu, h_bu, (size_t)n_users, d.st));
    MFK_CUDA(up(dbi, h_bi, (size_t)n_items, d.st));
    mfk_plan *plan = nullptr;
    int rc = mfk_plan_create(&plan, du, di, dr, n, n_users, n_items, opts, d.st);
    if (rc) return rc;
    if (h_order && n > 0) {
        rc = d.alloc(&dorder, (size_t)n) == cudaSuccess ? MFK_OK : MFK_ERR_CUDA;
        if (rc == MFK_OK) rc = mfk_plan_order(plan, dorder, d.st);
        if (rc == MFK_OK && down(h_order, dorder, (size_t)n, d.st) != cudaSuccess) rc = MFK_ERR_CUDA;
    }
    for (int e = 0; rc == MFK_OK && e < n_epochs; ++e) {
        rc = mfk_bias_sgd_epoch(plan, dbu, dbi, global_mean, lr, reg, update_user_params, update_item_params, d.st);
        if (rc == MFK_OK) rc = mfk_bias_sse(du, di, dr, n, dbu, dbi, global_mean, dws, dsse + e, d.st);
    }
    if (rc == MFK_OK) {
        cudaError_t ce = down(h_bu, dbu, (size_t)n_users, d.st);
        if (ce == cudaSuccess) ce = down(h_bi, dbi, (size_t)n_items, d.st);
        if (ce == cudaSuccess && n_epochs > 0) ce = down(h_train_rmse, dsse, (size_t)n_epochs, d.st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(d.st);
        if (ce != cudaSuccess) {
            set_error("mfk_bias_sgd_host: copy-back failed: %s", cudaGetErrorString(ce));
            rc = MFK_ERR_CUDA;
        }
    }
    mfk_plan_destroy(plan);
    if (rc == MFK_OK)
        for (int e = 0; e < n_epochs; ++e) h_train_rmse[e] = n > 0 ? std::sqrt(h_train_rmse[e] / (double)n) : NAN;
    return rc;
}

extern "C" int mfk_bias_als_host(const int32_t *h_u, const int32_t *h_i, const float *h_r, int64_t n, int32_t n_users,
                                 int32_t n_items, float *h_bu, float *h_bi, float global_mean, int32_t n_epochs,
                                 float reg, double *h_train_rmse) {
    MFK_REQUIRE(n >= 0 && n_users > 0 && n_items > 0 && n_epochs >= 0, "mfk_bias_als_host: bad sizes");
    MFK_REQUIRE(n == 0 || (h_u && h_i && h_r), "mfk_bias_als_host: null rating arrays");
    MFK_REQUIRE(h_bu && h_bi, "mfk_bias_als_host: null parameter array");
    MFK_REQUIRE(n_epochs == 0 || h_train_rmse, "mfk_bias_als_host: null train_rmse");
    DevBufs d;
    MFK_CUDA(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
    int32_t *du, *di;
    float *dr, *dbu, *dbi;
    double *dsse;
    void *dws;
    MFK_CUDA(d.alloc(&du, (size_t)n));
    MFK_CUDA(d.alloc(&di, (size_t)n));
    MFK_CUDA(d.alloc(&dr, (size_t)n));
    MFK_CUDA(d.alloc(&dbu, (size_t)n_users));
    MFK_CUDA(d.alloc(&dbi, (size_t)n_items));
    MFK_CUDA(d.alloc(&dsse, (size_t)(n_epochs > 0 ? n_epochs : 1)));
    MFK_CUDA(d.alloc((unsigned char **)&dws, mfk_sse_workspace_bytes()));
    if (n > 0) {
        MFK_CUDA(up(du, h_u, (size_t)n, d.st));
        MFK_CUDA(up(di, h_i, (size_t)n, d.st));
        MFK_CUDA(up(dr, h_r, (size_t)n, d.st));
    }
    MFK_CUDA(up(dbu, h_bu, (size_t)n_users, d.st));
    MFK_CUDA(up(dbi, h_bi, (size_t)n_items, d.st));
    mfk_csr *csr = nullptr;
    int rc = mfk_csr_create(&csr, du, di, dr, n, n_users, n_items, d.st);
    if (rc) return rc;
    for (int e = 0; rc == MFK_OK && e < n_epochs; ++e) {
        rc = mfk_bias_als_epoch(csr, dbu, dbi, global_mean, reg, d.st);
        if (rc == MFK_OK) rc = mfk_bias_sse(du, di, dr, n, dbu, dbi, global_mean, dws, dsse + e, d.st);
    }
    if (rc == MFK_OK) {
        cudaError_t ce = down(h_bu, dbu, (size_t)n_users, d.st);
        if (ce == cudaSuccess) ce = down(h_bi, dbi, (size_t)n_items, d.st);
        if (ce == cudaSuccess && n_epochs > 0) ce = down(h_train_rmse, dsse, (size_t)n_epochs, d.st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(d.st);
        if (ce != cudaSuccess) {
            set_error("mfk_bias_als_host: copy-back failed: %s", cudaGetErrorString(ce));
            rc = MFK_ERR_CUDA;
        }
    }
    mfk_csr_destroy(csr);
    if (rc == MFK_OK)
        for (int e = 0; e < n_epochs; ++e) h_train_rmse[e] = n > 0 ? std::sqrt(h_train_rmse[e] / (double)n) : NAN;
    return rc;
}
