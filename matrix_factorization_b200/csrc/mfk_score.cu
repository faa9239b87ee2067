// Batched scoring + known-item mask + top-k  (recommender_base.py:214-271, recommend()).
//
// Semantics kept from the reference: candidates are filtered BEFORE ranking, ranking uses the
// UNBOUNDED prediction, clipping happens AFTER selection (recommender_base.py:248-266).
// Rank keys are monotone in the prediction:
//     linear / sigmoid : key = b_i + p.q          rbf : key = -|p - q|^2
// so sigma / exp are applied only to the k winners.
//
// Round-1 implementation: a tiled fp32 SIMT contraction U_tile x Q^T writes the keys of one
// user tile to a scratch matrix, the mask entries are punched out, and one CTA per user does an
// exact radix-select top-k (ties broken by lower item id) followed by a bitonic sort of the
// winners.  (The tcgen05 split-TF32 contraction with the top-k fused into its epilogue is the
// planned replacement of stage 1; the selection semantics stay as they are here.)
#include <cfloat>
#include <cstdlib>

#include "mfk_common.cuh"

namespace mfk {

constexpr int kTileM = 64, kTileN = 64, kTileK = 16;
constexpr int kScoreUserTile = 2048;  // users per scratch tile
constexpr int kMaxTopK = 1024;
constexpr int kSelThreads = 256;

__device__ __forceinline__ uint32_t f2key(float f) {  // larger float <-> larger uint; masked = 0
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

struct ScoreParams {
    const int32_t *users;
    const float *P, *Q, *bi;
    int32_t n_items, F, ld;
    int64_t m;  // users in this tile
};

// keys[row][item] for a 64x64 tile; 256 threads, 4x4 micro-tile each.
template <int KERNEL>
__global__ void __launch_bounds__(256) k_score_tile(ScoreParams sp, uint32_t *keys) {
    __shared__ float As[kTileK][kTileM + 4];
    __shared__ float Bs[kTileK][kTileN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads
    const int64_t row0 = (int64_t)blockIdx.y * kTileM;
    const int32_t col0 = blockIdx.x * kTileN;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    // loader mapping: 64 rows x 16 k-values = 1024 elements, 4 per thread (one float4 along k)
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const int64_t arow = row0 + lrow;
    const int32_t auser = (arow < sp.m) ? sp.users[arow] : -1;
    const int32_t bitem = col0 + lrow;
    for (int k0 = 0; k0 < sp.F; k0 += kTileK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        if (auser >= 0 && k0 + lk < sp.F) av = __ldg(reinterpret_cast<const float4 *>(sp.P + (size_t)auser * sp.ld + k0 + lk));
        if (bitem < sp.n_items && k0 + lk < sp.F) bv = __ldg(reinterpret_cast<const float4 *>(sp.Q + (size_t)bitem * sp.ld + k0 + lk));
        As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
        Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTileK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                a[x] = As[kk][ty * 4 + x];
                b[x] = Bs[kk][tx * 4 + x];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    if (KERNEL == MFK_KERNEL_RBF) {
                        float d = a[x] - b[y];
                        acc[x][y] = fmaf(d, d, acc[x][y]);
                    } else {
                        acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        int64_t row = row0 + ty * 4 + x;
        if (row >= sp.m) continue;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
            int32_t col = col0 + tx * 4 + y;
            if (col >= sp.n_items) continue;
            float key = (KERNEL == MFK_KERNEL_RBF) ? -acc[x][y] : (sp.bi[col] + acc[x][y]);
            keys[(size_t)row * sp.n_items + col] = f2key(key);
        }
    }
}

__global__ void k_mask(const int64_t *__restrict__ mask_ptr, const int32_t *__restrict__ mask_items, int64_t user0,
                       int64_t m, int32_t n_items, uint32_t *keys) {
    // one warp per user row of the tile
    const int lane = threadIdx.x & 31;
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= m) return;
    const int64_t b = mask_ptr[user0 + row], e = mask_ptr[user0 + row + 1];
    for (int64_t k = b + lane; k < e; k += 32) {
        int32_t it = mask_items[k];
        if ((uint32_t)it < (uint32_t)n_items) keys[(size_t)row * n_items + it] = 0u;
    }
}

// One CTA per user: exact top-k of keys[0..n_items) (0 == masked), ties -> lower item id.
__global__ void __launch_bounds__(kSelThreads) k_select(const uint32_t *__restrict__ keys_all, int32_t n_items,
                                                        int32_t k, int kernel, const int32_t *__restrict__ users,
                                                        const float *__restrict__ bu, float mu, float gamma, float a,
                                                        float c, int bound, float lo, float hi, float *out_scores,
                                                        int32_t *out_items) {
    extern __shared__ unsigned long long s_sel[];  // [kpow2] composite keys
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_prefix, s_need, s_count, s_tiebase;
    __shared__ uint32_t s_scan[kSelThreads / 32];
    const int tid = threadIdx.x;
    const uint32_t *keys = keys_all + (size_t)blockIdx.x * n_items;
    int kpow2 = 1;
    while (kpow2 < k) kpow2 <<= 1;

    // ---- radix select of the k-th largest key (over non-masked entries)
    if (tid == 0) {
        s_prefix = 0;
        s_need = (uint32_t)k;
    }
    __syncthreads();
    uint32_t n_valid_total = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        s_hist[tid] = 0;  // kSelThreads == 256
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int32_t j = tid; j < n_items; j += kSelThreads) {
            uint32_t v = keys[j];
            if (v != 0u && (v & pmask) == prefix) atomicAdd(&s_hist[(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = s_need, run = 0;
            if (pass == 0) {
                uint32_t tot = 0;
                for (int b = 0; b < 256; ++b) tot += s_hist[b];
                s_count = tot;  // number of valid candidates
                if (need > tot) need = tot;
            }
            int b = 255;
            if (need > 0) {
                for (; b > 0; --b) {
                    if (run + s_hist[b] >= need) break;
                    run += s_hist[b];
                }
            }
            s_prefix = prefix | ((uint32_t)b << shift);
            s_need = need - run;  // how many are still needed inside bucket b
        }
        __syncthreads();
    }
    n_valid_total = s_count;
    const uint32_t kth = s_prefix;        // value of the k-th largest key
    const uint32_t need_ties = s_need;    // how many entries == kth to take (lowest item ids first)
    const uint32_t k_eff = min((uint32_t)k, n_valid_total);
    __syncthreads();
    if (tid == 0) {
        s_count = 0;
        s_tiebase = 0;
    }
    for (int j = tid; j < kpow2; j += kSelThreads) s_sel[j] = 0ull;
    __syncthreads();
    // ---- gather: everything > kth (any order), then the first need_ties entries == kth in id order
    for (int32_t j0 = 0; j0 < n_items; j0 += kSelThreads) {
        int32_t j = j0 + tid;
        uint32_t v = (j < n_items) ? keys[j] : 0u;
        if (k_eff > 0 && v > kth && v != 0u) {
            uint32_t pos = atomicAdd(&s_count, 1u);
            if (pos < (uint32_t)kpow2) s_sel[pos] = ((unsigned long long)v << 32) | (unsigned long long)(0xffffffffu - (uint32_t)j);
        }
        // ordered compaction of ties
        uint32_t is_tie = (k_eff > 0 && v == kth && v != 0u) ? 1u : 0u;
        uint32_t ball = __ballot_sync(0xffffffffu, is_tie);
        uint32_t wpre = __popc(ball & ((1u << (tid & 31)) - 1u));
        if ((tid & 31) == 0) s_scan[tid >> 5] = __popc(ball);
        __syncthreads();
        uint32_t base = s_tiebase;
        for (int w = 0; w < (tid >> 5); ++w) base += s_scan[w];
        if (is_tie && base + wpre < need_ties) {
            uint32_t pos = atomicAdd(&s_count, 1u);
            if (pos < (uint32_t)kpow2) s_sel[pos] = ((unsigned long long)v << 32) | (unsigned long long)(0xffffffffu - (uint32_t)j);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kSelThreads / 32; ++w) t += s_scan[w];
            s_tiebase += t;
        }
        __syncthreads();
    }
    // ---- bitonic sort (descending) of the composite keys
    for (int size = 2; size <= kpow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < kpow2 / 2; t += kSelThreads) {
                int lo_i = 2 * t - (t & (stride - 1));
                int hi_i = lo_i + stride;
                bool desc = ((lo_i & size) == 0);
                unsigned long long x = s_sel[lo_i], y = s_sel[hi_i];
                if ((x < y) == desc) {
                    s_sel[lo_i] = y;
                    s_sel[hi_i] = x;
                }
            }
            __syncthreads();
        }
    }
    // ---- winners -> predictions
    const int32_t user = users[blockIdx.x];
    const float ub = (kernel == MFK_KERNEL_RBF) ? 0.f : bu[user];
    for (int j = tid; j < k; j += kSelThreads) {
        float score = -INFINITY;
        int32_t item = -1;
        if ((uint32_t)j < k_eff) {
            unsigned long long ck = s_sel[j];
            item = (int32_t)(0xffffffffu - (uint32_t)(ck & 0xffffffffull));
            float key = key2f((uint32_t)(ck >> 32));
            if (kernel == MFK_KERNEL_LINEAR) score = mu + ub + key;                  // key = b_i + p.q
            else if (kernel == MFK_KERNEL_SIGMOID) score = a + c * (1.0f / (1.0f + expf(-(mu + ub + key))));
            else score = a + c * expf(gamma * key);                                  // key = -|p-q|^2
            if (bound) score = score > hi ? hi : (score < lo ? lo : score);
        }
        out_scores[(size_t)blockIdx.x * k + j] = score;
        out_items[(size_t)blockIdx.x * k + j] = item;
    }
}

// Merge of per-shard top-k lists (the all-gather step of the item-sharded recommend): one CTA per user
// sorts its c candidates (score desc, ties -> lower item id; item < 0 = padding) and keeps the best k.
__global__ void __launch_bounds__(kSelThreads) k_topk_merge(const float *__restrict__ in_scores,
                                                            const int32_t *__restrict__ in_items, int32_t c,
                                                            int32_t k, int bound, float lo, float hi,
                                                            float *out_scores, int32_t *out_items) {
    extern __shared__ unsigned long long s_sel[];
    const int tid = threadIdx.x;
    int cpow2 = 1;
    while (cpow2 < c) cpow2 <<= 1;
    const float *sc = in_scores + (size_t)blockIdx.x * c;
    const int32_t *it = in_items + (size_t)blockIdx.x * c;
    for (int j = tid; j < cpow2; j += kSelThreads) {
        unsigned long long key = 0ull;
        if (j < c && it[j] >= 0) key = ((unsigned long long)f2key(sc[j]) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)it[j]);
        s_sel[j] = key;
    }
    __syncthreads();
    for (int size = 2; size <= cpow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < cpow2 / 2; t += kSelThreads) {
                int lo_i = 2 * t - (t & (stride - 1));
                int hi_i = lo_i + stride;
                bool desc = ((lo_i & size) == 0);
                unsigned long long x = s_sel[lo_i], y = s_sel[hi_i];
                if ((x < y) == desc) {
                    s_sel[lo_i] = y;
                    s_sel[hi_i] = x;
                }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < k; j += kSelThreads) {
        float score = -INFINITY;
        int32_t item = -1;
        unsigned long long ck = j < cpow2 ? s_sel[j] : 0ull;
        if (ck != 0ull) {
            item = (int32_t)(0xffffffffu - (uint32_t)(ck & 0xffffffffull));
            score = key2f((uint32_t)(ck >> 32));
            if (bound) score = score > hi ? hi : (score < lo ? lo : score);
        }
        out_scores[(size_t)blockIdx.x * k + j] = score;
        out_items[(size_t)blockIdx.x * k + j] = item;
    }
}

}  // namespace mfk

using namespace mfk;

extern "C" int mfk_topk_merge(const float *d_scores_in, const int32_t *d_items_in, int64_t m, int32_t c, int32_t k,
                              int bound_ratings, float min_rating, float max_rating, float *d_scores,
                              int32_t *d_items, void *stream) {
    MFK_REQUIRE(m >= 0 && c >= 1 && c <= 8192 && k >= 1 && k <= c, "mfk_topk_merge: bad sizes (m=%lld c=%d k=%d)",
                (long long)m, c, k);
    if (m == 0) return MFK_OK;
    MFK_REQUIRE(d_scores_in && d_items_in && d_scores && d_items, "mfk_topk_merge: null array");
    int cpow2 = 1;
    while (cpow2 < c) cpow2 <<= 1;
    size_t smem = sizeof(unsigned long long) * (size_t)cpow2;
    if (smem > 48 * 1024) MFK_CUDA(cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_topk_merge<<<(unsigned)m, kSelThreads, smem, as_stream(stream)>>>(d_scores_in, d_items_in, c, k, bound_ratings,
                                                                      min_rating, max_rating, d_scores, d_items);
    MFK_LAUNCH_CHECK();
    return MFK_OK;
}

namespace mfk {
size_t score_tc_workspace_bytes(int64_t m, int32_t n_items, int32_t n_factors);
int score_tc(int kernel, const int32_t *d_users, int64_t m, const float *d_P, const float *d_Q, const float *d_bu,
             const float *d_bi, int32_t n_items, int32_t n_factors, int32_t ld, float mu, float gamma, float lo,
             float hi, const int64_t *d_mask_ptr, const int32_t *d_mask_items, int32_t k, int bound, float *d_scores,
             int32_t *d_items, void *d_ws, cudaStream_t st);
// The tensor-core path serves k <= 64; MFK_SCORE_SIMT=1 forces the fp32 SIMT path (used by the tests to compare).
static bool use_tensor_path(int32_t k) {
    static const bool force_simt = [] {
        const char *e = getenv("MFK_SCORE_SIMT");
        return e && e[0] == '1';
    }();
    return !force_simt && k <= 64;
}
}  // namespace mfk

extern "C" size_t mfk_score_workspace_bytes(int64_t m, int32_t n_items, int32_t n_factors, int32_t k) {
    (void)k;
    int64_t tile = m < kScoreUserTile ? m : kScoreUserTile;
    if (tile < 1) tile = 1;
    size_t simt = (size_t)tile * (size_t)(n_items > 0 ? n_items : 1) * sizeof(uint32_t);
    size_t tc = mfk::score_tc_workspace_bytes(m, n_items > 0 ? n_items : 1, n_factors > 0 ? n_factors : 1) + 512;
    return simt > tc ? simt : tc;
}

extern "C" int mfk_score_topk(int kernel, const int32_t *d_users, int64_t m, const float *d_P, const float *d_Q,
                              const float *d_bu, const float *d_bi, int32_t n_items, int32_t n_factors, int32_t ld,
                              float global_mean, float gamma, float min_rating, float max_rating,
                              const int64_t *d_mask_ptr, const int32_t *d_mask_items, int32_t k, int bound_ratings,
                              float *d_scores, int32_t *d_items, void *d_ws, void *stream) {
    MFK_REQUIRE(kernel >= 0 && kernel <= 2, "mfk_score_topk: bad kernel %d", kernel);
    MFK_REQUIRE(m >= 0 && n_items > 0, "mfk_score_topk: bad sizes");
    MFK_REQUIRE(k >= 1 && k <= kMaxTopK, "mfk_score_topk: k=%d must be in [1,%d]", k, kMaxTopK);
    MFK_REQUIRE(d_P && d_Q && d_bu && d_bi, "mfk_score_topk: null parameter array");
    MFK_REQUIRE(n_factors >= 1 && ld >= n_factors && ld % 4 == 0, "mfk_score_topk: bad n_factors/ld");
    MFK_REQUIRE((((uintptr_t)d_P | (uintptr_t)d_Q) & 15) == 0, "mfk_score_topk: P/Q must be 16-byte aligned");
    if (m == 0) return MFK_OK;
    MFK_REQUIRE(d_users && d_scores && d_items && d_ws, "mfk_score_topk: null array");
    MFK_REQUIRE(d_mask_ptr == nullptr || d_mask_items != nullptr, "mfk_score_topk: mask_ptr without mask_items");
    cudaStream_t st = as_stream(stream);
    if (use_tensor_path(k))
        return score_tc(kernel, d_users, m, d_P, d_Q, d_bu, d_bi, n_items, n_factors, ld, global_mean, gamma, min_rating,
                        max_rating, d_mask_ptr, d_mask_items, k, bound_ratings, d_scores, d_items, d_ws, st);
    uint32_t *keys = reinterpret_cast<uint32_t *>(d_ws);
    int kpow2 = 1;
    while (kpow2 < k) kpow2 <<= 1;
    size_t sel_smem = sizeof(unsigned long long) * (size_t)kpow2;
    const int F4 = (n_factors + 3) & ~3;
    for (int64_t u0 = 0; u0 < m; u0 += kScoreUserTile) {
        int64_t mt = m - u0 < kScoreUserTile ? m - u0 : kScoreUserTile;
        ScoreParams sp{d_users + u0, d_P, d_Q, d_bi, n_items, F4, ld, mt};
        dim3 grid((n_items + kTileN - 1) / kTileN, (unsigned)((mt + kTileM - 1) / kTileM));
        if (kernel == MFK_KERNEL_RBF) k_score_tile<MFK_KERNEL_RBF><<<grid, 256, 0, st>>>(sp, keys);
        else k_score_tile<MFK_KERNEL_LINEAR><<<grid, 256, 0, st>>>(sp, keys);
        MFK_LAUNCH_CHECK();
        if (d_mask_ptr) {
            k_mask<<<(unsigned)((mt * 32 + 255) / 256), 256, 0, st>>>(d_mask_ptr, d_mask_items, u0, mt, n_items, keys);
            MFK_LAUNCH_CHECK();
        }
        k_select<<<(unsigned)mt, kSelThreads, sel_smem, st>>>(keys, n_items, k, kernel, d_users + u0, d_bu, global_mean,
                                                             gamma, min_rating, max_rating - min_rating, bound_ratings,
                                                             min_rating, max_rating, d_scores + (size_t)u0 * k,
                                                             d_items + (size_t)u0 * k);
        MFK_LAUNCH_CHECK();
    }
    return MFK_OK;
}
