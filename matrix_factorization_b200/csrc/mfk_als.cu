// CSR/CSC build on the GPU and the ALS bias solver as two segmented reductions.
//
//   mfk_csr_create      : layout built once from the internal ids of utils.py / recommender_base.py
//   mfk_bias_als_epoch  : baseline_model.py:326-348 (_als epoch body)
//       b_u = sum_{i in R(u)} (r - mu - b_i) / (reg + n_u)          (all users, from zero)
//       b_i = sum_{u in R(i)} (r - mu - b_u) / (reg + n_i)          (with the NEW b_u)
// Sums are order-independent; they are accumulated in double in a fixed (sorted) order, so the
// result is deterministic and within fp32 rounding of the reference's fp64 scatter-add loop.
#include <cub/cub.cuh>

#include "mfk_common.cuh"

struct mfk_csr {
    int64_t n = 0;
    int32_t n_users = 0, n_items = 0;
    int64_t *row_ptr = nullptr, *col_ptr = nullptr;
    int32_t *col = nullptr, *row = nullptr;
    float *val = nullptr, *cval = nullptr;
};

namespace mfk {

__global__ void k_pack_keys(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int64_t n, uint64_t *keys,
                            int32_t *idx) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        keys[k] = ((uint64_t)(uint32_t)a[k] << 32) | (uint64_t)(uint32_t)b[k];
        idx[k] = (int32_t)k;
    }
}

__global__ void k_unpack(const uint64_t *__restrict__ keys, const int32_t *__restrict__ idx,
                         const float *__restrict__ r, int64_t n, int32_t *minor, float *val) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        minor[k] = (int32_t)(keys[k] & 0xffffffffu);
        val[k] = r[idx[k]];
    }
}

// ptr[j] = first position whose major id >= j  (j in 0..n_major)
__global__ void k_major_bounds(const uint64_t *__restrict__ keys, int64_t n, int32_t n_major, int64_t *ptr) {
    int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n_major) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> 32) < (int64_t)j) lo = mid + 1;
        else hi = mid;
    }
    ptr[j] = lo;
}

// One warp per segment: out[j] = sum(val - mu - other[minor]) / (reg + len).
// Long segments (hot items) are strided by the whole warp.
__global__ void __launch_bounds__(256) k_als_pass(const int64_t *__restrict__ ptr, const int32_t *__restrict__ minor,
                                                  const float *__restrict__ val, const float *__restrict__ other,
                                                  int32_t n_major, float mu, float reg, float *out) {
    const int lane = threadIdx.x & 31;
    int32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t stride = (gridDim.x * blockDim.x) >> 5;
    for (; j < n_major; j += stride) {
        const int64_t b = ptr[j], e = ptr[j + 1];
        double acc = 0.0;
        for (int64_t k = b + lane; k < e; k += 32) acc += (double)val[k] - (double)mu - (double)other[minor[k]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[j] = (float)(acc / ((double)reg + (double)(e - b)));
    }
}

static int grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    return (int)max((int64_t)1, min(b, (int64_t)148 * 32));
}

static int build_half(const int32_t *major, const int32_t *minor_in, const float *r, int64_t n, int32_t n_major,
                      int64_t *ptr, int32_t *minor_out, float *val_out, cudaStream_t st) {
    uint64_t *ka = nullptr, *kb = nullptr;
    int32_t *ia = nullptr, *ib = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    size_t nn = (size_t)max((int64_t)1, n);
    MFK_CUDA(cudaMalloc(&ka, nn * 8));
    MFK_CUDA(cudaMalloc(&kb, nn * 8));
    MFK_CUDA(cudaMalloc(&ia, nn * 4));
    MFK_CUDA(cudaMalloc(&ib, nn * 4));
    if (n > 0) {
        k_pack_keys<<<grid_for(n), 256, 0, st>>>(major, minor_in, n, ka, ia);
        MFK_LAUNCH_CHECK();
        MFK_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ka, kb, ia, ib, (int)n, 0, 64, st));
        MFK_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
        MFK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ka, kb, ia, ib, (int)n, 0, 64, st));
        k_unpack<<<grid_for(n), 256, 0, st>>>(kb, ib, r, n, minor_out, val_out);
        MFK_LAUNCH_CHECK();
    }
    k_major_bounds<<<(n_major + 1 + 255) / 256, 256, 0, st>>>(kb, n, n_major, ptr);
    MFK_LAUNCH_CHECK();
    MFK_CUDA(cudaStreamSynchronize(st));
    cudaFree(tmp);
    cudaFree(ka);
    cudaFree(kb);
    cudaFree(ia);
    cudaFree(ib);
    return MFK_OK;
}

}  // namespace mfk

using namespace mfk;

extern "C" int mfk_csr_destroy(mfk_csr *c) {
    if (!c) return MFK_OK;
    void *ptrs[] = {c->row_ptr, c->col_ptr, c->col, c->row, c->val, c->cval};
    for (void *q : ptrs)
        if (q) cudaFree(q);
    delete c;
    return MFK_OK;
}

extern "C" int mfk_csr_create(mfk_csr **out, const int32_t *d_u, const int32_t *d_i, const float *d_r, int64_t n,
                              int32_t n_users, int32_t n_items, void *stream) {
    MFK_REQUIRE(out != nullptr, "mfk_csr_create: out is NULL");
    *out = nullptr;
    MFK_REQUIRE(n >= 0 && n < (int64_t)INT32_MAX, "mfk_csr_create: n out of range");
    MFK_REQUIRE(n_users > 0 && n_items > 0, "mfk_csr_create: n_users/n_items must be positive");
    MFK_REQUIRE(n == 0 || (d_u && d_i && d_r), "mfk_csr_create: null rating arrays");
    DeviceProps props;
    int rc = device_props(&props);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    mfk_csr *c = new mfk_csr();
    c->n = n;
    c->n_users = n_users;
    c->n_items = n_items;
    size_t nn = (size_t)max((int64_t)1, n);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&c->row_ptr, sizeof(int64_t) * (size_t)(n_users + 1));
    if (e == cudaSuccess) e = cudaMalloc(&c->col_ptr, sizeof(int64_t) * (size_t)(n_items + 1));
    if (e == cudaSuccess) e = cudaMalloc(&c->col, nn * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->row, nn * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->val, nn * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->cval, nn * 4);
    if (e != cudaSuccess) {
        set_error("mfk_csr_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        mfk_csr_destroy(c);
        return MFK_ERR_CUDA;
    }
    rc = build_half(d_u, d_i, d_r, n, n_users, c->row_ptr, c->col, c->val, st);
    if (rc == MFK_OK) rc = build_half(d_i, d_u, d_r, n, n_items, c->col_ptr, c->row, c->cval, st);
    if (rc != MFK_OK) {
        mfk_csr_destroy(c);
        return rc;
    }
    *out = c;
    return MFK_OK;
}

extern "C" int mfk_csr_export(const mfk_csr *csr, int64_t *d_row_ptr, int32_t *d_col, float *d_val,
                              int64_t *d_col_ptr, int32_t *d_row, float *d_cval, void *stream) {
    MFK_REQUIRE(csr != nullptr, "mfk_csr_export: csr is NULL");
    cudaStream_t st = as_stream(stream);
    const size_t n = (size_t)csr->n;
    if (d_row_ptr) MFK_CUDA(cudaMemcpyAsync(d_row_ptr, csr->row_ptr, 8 * (size_t)(csr->n_users + 1), cudaMemcpyDeviceToDevice, st));
    if (d_col_ptr) MFK_CUDA(cudaMemcpyAsync(d_col_ptr, csr->col_ptr, 8 * (size_t)(csr->n_items + 1), cudaMemcpyDeviceToDevice, st));
    if (n) {
        if (d_col) MFK_CUDA(cudaMemcpyAsync(d_col, csr->col, 4 * n, cudaMemcpyDeviceToDevice, st));
        if (d_val) MFK_CUDA(cudaMemcpyAsync(d_val, csr->val, 4 * n, cudaMemcpyDeviceToDevice, st));
        if (d_row) MFK_CUDA(cudaMemcpyAsync(d_row, csr->row, 4 * n, cudaMemcpyDeviceToDevice, st));
        if (d_cval) MFK_CUDA(cudaMemcpyAsync(d_cval, csr->cval, 4 * n, cudaMemcpyDeviceToDevice, st));
    }
    return MFK_OK;
}

extern "C" int mfk_bias_als_epoch(const mfk_csr *csr, float *d_bu, float *d_bi, float global_mean, float reg,
                                  void *stream) {
    MFK_REQUIRE(csr && d_bu && d_bi, "mfk_bias_als_epoch: null argument");
    cudaStream_t st = as_stream(stream);
    // user pass reads the current item biases; item pass reads the user biases just written
    k_als_pass<<<grid_for((int64_t)csr->n_users * 32), 256, 0, st>>>(csr->row_ptr, csr->col, csr->val, d_bi,
                                                                     csr->n_users, global_mean, reg, d_bu);
    MFK_LAUNCH_CHECK();
    k_als_pass<<<grid_for((int64_t)csr->n_items * 32), 256, 0, st>>>(csr->col_ptr, csr->row, csr->cval, d_bu,
                                                                     csr->n_items, global_mean, reg, d_bi);
    MFK_LAUNCH_CHECK();
    return MFK_OK;
}
