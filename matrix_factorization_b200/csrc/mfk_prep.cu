// GPU side of RecommenderBase._preprocess_data (recommender_base.py:97-173) for integer raw ids: the row shuffle is
// applied from a permutation the HOST drew (the reference consumes numpy's global RNG there, and so does the caller),
// ids become internal ids in FIRST-APPEARANCE order of the shuffled rows (:133-140), and duplicate (user, item) pairs are
// detected (:125-128).  Bit-exact with the host path by construction: "first appearance" is resolved with stable
// radix sorts, no hashing order is involved.
//
//   mfk_first_appearance:  shuffled[k] = raw[perm[k]];  internal[k] = rank of shuffled[k] among the distinct ids ordered
//                          by the position of their first occurrence;  unique[r] = the id with rank r.
//   mfk_has_duplicate_pairs:  does any (u, i) occur twice?
#include <cub/cub.cuh>

#include <vector>

#include "mfk_common.cuh"

namespace mfk {
namespace {

struct Scratch {  // device allocations of one call, freed on every exit path
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t get(T **out, size_t count) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, sizeof(T) * (count ? count : 1));
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = reinterpret_cast<T *>(p);
        return e;
    }
};

__global__ void k_gather_keys(const int64_t *__restrict__ raw, const int64_t *__restrict__ perm, int64_t n, int64_t *keys,
                              int32_t *pos) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        keys[k] = raw[perm ? perm[k] : k];
        pos[k] = (int32_t)k;
    }
}
// heads of the runs of equal keys (sorted), as 0/1 flags
__global__ void k_heads(const int64_t *__restrict__ ks, int64_t n, int32_t *head) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        head[k] = (k == 0 || ks[k] != ks[k - 1]) ? 1 : 0;
}
// per run: its key and the position of its first occurrence (the sort was stable, positions ascend inside a run)
__global__ void k_run_firsts(const int64_t *__restrict__ ks, const int32_t *__restrict__ pos, const int32_t *__restrict__ head,
                             const int32_t *__restrict__ seg /* inclusive scan of head */, int64_t n, int64_t *run_key,
                             int32_t *run_first, int32_t *run_id) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        if (head[k]) {
            const int32_t s = seg[k] - 1;
            run_key[s] = ks[k];
            run_first[s] = pos[k];
            run_id[s] = s;
        }
}
__global__ void k_rank_of_run(const int32_t *__restrict__ run_sorted, const int64_t *__restrict__ run_key, int32_t n_runs,
                              int32_t *rank, int64_t *unique_out) {
    for (int32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_runs; t += gridDim.x * blockDim.x) {
        const int32_t s = run_sorted[t];
        rank[s] = t;
        unique_out[t] = run_key[s];
    }
}
__global__ void k_scatter_rank(const int32_t *__restrict__ pos, const int32_t *__restrict__ seg, const int32_t *__restrict__ rank,
                               int64_t n, int32_t *internal) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        internal[pos[k]] = rank[seg[k] - 1];
}
__global__ void k_pair_keys(const int32_t *__restrict__ u, const int32_t *__restrict__ i, int64_t n, int64_t n_items, int64_t *keys) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        keys[k] = (int64_t)u[k] * n_items + (int64_t)i[k];
}
__global__ void k_any_adjacent_equal(const int64_t *__restrict__ ks, int64_t n, int32_t *flag) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1; k < n; k += (int64_t)gridDim.x * blockDim.x)
        if (ks[k] == ks[k - 1]) *flag = 1;
}
inline int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 16); }

}  // namespace
}  // namespace mfk

using namespace mfk;

#define PREP_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return MFK_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

extern "C" int mfk_first_appearance(const int64_t *d_raw, const int64_t *d_perm, int64_t n, int32_t *d_internal,
                                    int64_t *d_unique, int32_t *h_n_unique, void *stream) {
    MFK_REQUIRE(h_n_unique != nullptr, "mfk_first_appearance: h_n_unique is NULL");
    *h_n_unique = 0;
    if (n == 0) return MFK_OK;
    MFK_REQUIRE(d_raw && d_internal && d_unique, "mfk_first_appearance: null array");
    MFK_REQUIRE(n < ((int64_t)1 << 31), "mfk_first_appearance: n = %lld exceeds 2^31 - 1", (long long)n);
    cudaStream_t st = as_stream(stream);
    Scratch sc;
    int64_t *keys = nullptr, *ks = nullptr, *run_key = nullptr;
    int32_t *pos = nullptr, *pos_s = nullptr, *head = nullptr, *seg = nullptr, *run_first = nullptr, *run_id = nullptr,
            *run_first_s = nullptr, *run_sorted = nullptr, *rank = nullptr;
    PREP_CUDA(sc.get(&keys, (size_t)n));
    PREP_CUDA(sc.get(&ks, (size_t)n));
    PREP_CUDA(sc.get(&pos, (size_t)n));
    PREP_CUDA(sc.get(&pos_s, (size_t)n));
    PREP_CUDA(sc.get(&head, (size_t)n));
    PREP_CUDA(sc.get(&seg, (size_t)n));
    const int g = grid_for(n);
    k_gather_keys<<<g, 256, 0, st>>>(d_raw, d_perm, n, keys, pos);
    PREP_CUDA(cudaGetLastError());
    // stable sort by id: inside a run of equal ids the positions stay ascending
    size_t tmp_bytes = 0, tb2 = 0;
    PREP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, ks, pos, pos_s, (int)n, 0, 64, st));
    PREP_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, head, seg, (int)n, st));
    tmp_bytes = std::max(tmp_bytes, tb2);
    void *tmp = nullptr;
    PREP_CUDA(sc.get((char **)&tmp, tmp_bytes));
    PREP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, ks, pos, pos_s, (int)n, 0, 64, st));
    k_heads<<<g, 256, 0, st>>>(ks, n, head);
    PREP_CUDA(cudaGetLastError());
    PREP_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, head, seg, (int)n, st));
    int32_t n_runs = 0;
    PREP_CUDA(cudaMemcpyAsync(&n_runs, seg + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PREP_CUDA(cudaStreamSynchronize(st));
    PREP_CUDA(sc.get(&run_key, (size_t)n_runs));
    PREP_CUDA(sc.get(&run_first, (size_t)n_runs));
    PREP_CUDA(sc.get(&run_id, (size_t)n_runs));
    PREP_CUDA(sc.get(&run_first_s, (size_t)n_runs));
    PREP_CUDA(sc.get(&run_sorted, (size_t)n_runs));
    PREP_CUDA(sc.get(&rank, (size_t)n_runs));
    k_run_firsts<<<g, 256, 0, st>>>(ks, pos_s, head, seg, n, run_key, run_first, run_id);
    PREP_CUDA(cudaGetLastError());
    // the distinct ids ordered by their first position = the order in which the reference's dict meets them
    size_t tb3 = 0;
    PREP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb3, run_first, run_first_s, run_id, run_sorted, n_runs, 0, 32, st));
    void *tmp3 = nullptr;
    PREP_CUDA(sc.get((char **)&tmp3, tb3));
    PREP_CUDA(cub::DeviceRadixSort::SortPairs(tmp3, tb3, run_first, run_first_s, run_id, run_sorted, n_runs, 0, 32, st));
    k_rank_of_run<<<grid_for(n_runs), 256, 0, st>>>(run_sorted, run_key, n_runs, rank, d_unique);
    PREP_CUDA(cudaGetLastError());
    k_scatter_rank<<<g, 256, 0, st>>>(pos_s, seg, rank, n, d_internal);
    PREP_CUDA(cudaGetLastError());
    PREP_CUDA(cudaStreamSynchronize(st));
    *h_n_unique = n_runs;
    return MFK_OK;
}

extern "C" int mfk_has_duplicate_pairs(const int32_t *d_u, const int32_t *d_i, int64_t n, int64_t n_items, int32_t *h_flag,
                                       void *stream) {
    MFK_REQUIRE(h_flag != nullptr, "mfk_has_duplicate_pairs: h_flag is NULL");
    *h_flag = 0;
    if (n < 2) return MFK_OK;
    MFK_REQUIRE(d_u && d_i && n_items > 0, "mfk_has_duplicate_pairs: null array or n_items <= 0");
    MFK_REQUIRE(n < ((int64_t)1 << 31), "mfk_has_duplicate_pairs: n = %lld exceeds 2^31 - 1", (long long)n);
    cudaStream_t st = as_stream(stream);
    Scratch sc;
    int64_t *keys = nullptr, *ks = nullptr;
    int32_t *flag = nullptr;
    PREP_CUDA(sc.get(&keys, (size_t)n));
    PREP_CUDA(sc.get(&ks, (size_t)n));
    PREP_CUDA(sc.get(&flag, 1));
    PREP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
    k_pair_keys<<<grid_for(n), 256, 0, st>>>(d_u, d_i, n, n_items, keys);
    PREP_CUDA(cudaGetLastError());
    size_t tb = 0;
    PREP_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, ks, (int)n, 0, 64, st));
    void *tmp = nullptr;
    PREP_CUDA(sc.get((char **)&tmp, tb));
    PREP_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, ks, (int)n, 0, 64, st));
    k_any_adjacent_equal<<<grid_for(n), 256, 0, st>>>(ks, n, flag);
    PREP_CUDA(cudaGetLastError());
    PREP_CUDA(cudaMemcpyAsync(h_flag, flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PREP_CUDA(cudaStreamSynchronize(st));
    return MFK_OK;
}
