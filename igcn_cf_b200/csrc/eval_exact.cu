// Full-ranking evaluation, exact fp32 path: score tile (CUDA-core FMA chain, ascending d) ->
// threshold filter -> lazy seen-item mask -> per-user running top-k kept in shared memory.
// The B x I score matrix is never written to HBM.
//
// Replaces torch.mm (reference model.py:122), the -inf index_put (trainer.py:149-161) and
// torch.topk (trainer.py:163).  It defines the score summation order that the tensor-core path's
// re-scoring step reproduces bit for bit, and it is the fallback for users whose tensor-core
// candidate bound does not verify.  igcn_hits is the membership test of calculate_metrics
// (trainer.py:111-115).
#include <float.h>

#include "common.cuh"

namespace igcn {

constexpr int BM = 64;      // users per CTA
constexpr int BN = 128;     // items per tile
constexpr int CAP = 256;    // candidate slots per user (K <= CAP - BN)
constexpr int EX_THREADS = 256;

__device__ __forceinline__ uint32_t float_order(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// larger key == better candidate: higher score first, then lower item id
__device__ __forceinline__ uint64_t cand_key(float score, int32_t item) {
    return ((uint64_t)float_order(score) << 32) | (uint32_t)(0x7fffffff - item);
}

__device__ __forceinline__ bool item_masked(int64_t u, int32_t j, const int64_t *__restrict__ mptr,
                                            const int32_t *__restrict__ mitems, const uint32_t *__restrict__ banned) {
    if (banned && ((__ldg(banned + (j >> 5)) >> (j & 31)) & 1u)) return true;
    if (!mptr) return false;
    int64_t lo = __ldg(mptr + u), hi = __ldg(mptr + u + 1);
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(mitems + mid) < j) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(mitems + lo) == j;
}

// One warp sorts the first `n` (<= CAP) keys of a row descending (bitonic, padded with 0).
__device__ void warp_sort_desc(uint64_t *keys, int n, int lane) {
    int n_pad = 32;
    while (n_pad < n) n_pad <<= 1;
    for (int s = n + lane; s < n_pad; s += 32) keys[s] = 0ULL;
    __syncwarp();
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < n_pad / 2; t += 32) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncwarp();
        }
    }
}

struct ExactArgs {
    const float *rep;
    const int64_t *user_ids;
    int64_t n_eval, item_row0, n_items;
    int D;
    const int64_t *mask_ptr;
    const int32_t *mask_items;
    int64_t item_lo, item_hi;
    const uint32_t *banned;
    int k;
    int32_t *out_items;
    float *out_scores;
    const int32_t *out_rows;     // optional: output row of entry b (default b)
    const int32_t *n_eval_dev;   // optional: device-side entry count (<= n_eval)
    // item-range splitting for SHORT user lists (the tensor-core path's fallback is typically a handful of users:
    // one CTA scanning the whole catalogue for them would be a multi-millisecond serial tail)
    int n_item_splits;           // blockIdx.y range; used only while the entry count is <= split_cap
    int64_t split_cap;
    uint64_t *split_keys;        // [split_cap, n_item_splits, k] partial top-k keys, merged by exact_merge_kernel
};

// Dimensions of the item tile held in shared memory at a time: 64, or 32 for the wide representations of the sibling
// models (IMCGAE 192, NGCF 256 columns) so that the user tile [BM][D+4] still fits beside the candidate keys.
__host__ __device__ inline int exact_kc(int D) { return D > 128 ? 32 : (D < 64 ? D : 64); }

// smem: Us[BM][D+4] | Is[BN][exact_kc(D)+4] | keys[BM][CAP] (u64) | cnt[BM] | thr[BM]
__global__ void __launch_bounds__(EX_THREADS, 1) score_topk_exact_kernel(const __grid_constant__ ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.D, DP = D + 4;
    float *Us = reinterpret_cast<float *>(smem_raw);
    float *Is = Us + BM * DP;
    uint64_t *keys = reinterpret_cast<uint64_t *>(Is + BN * (exact_kc(D) + 4));
    int *cnt = reinterpret_cast<int *>(keys + BM * CAP);
    float *thr = reinterpret_cast<float *>(cnt + BM);
    __shared__ int need_compact;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;      // 16 item lanes x 16 user groups of 4
    const int64_t ubase = (int64_t)blockIdx.x * BM;
    const int d4 = D >> 2;
    const int64_t n_eval = a.n_eval_dev ? min(a.n_eval, (int64_t)*a.n_eval_dev) : a.n_eval;
    if (ubase >= n_eval) return;
    // two launches share this kernel when item splitting is on: gridDim.y > 1 is the split pass (runs only while
    // the entry count fits split_cap), gridDim.y == 1 the plain pass (runs only when it does not)
    const bool split_on = a.n_item_splits > 1 && n_eval <= a.split_cap;
    if (a.n_item_splits > 1 && split_on != (gridDim.y > 1)) return;
    const int n_splits = split_on ? a.n_item_splits : 1;
    const int split = blockIdx.y;

    for (int idx = tid; idx < BM * d4; idx += EX_THREADS) {
        const int r = idx / d4, c = idx % d4;
        float4 v = f4zero();
        if (ubase + r < n_eval) v = ld4(a.rep + __ldg(a.user_ids + ubase + r) * D + c * 4);
        st4(Us + r * DP + c * 4, v);
    }
    if (tid < BM) { cnt[tid] = 0; thr[tid] = -INFINITY; }
    if (tid == 0) need_compact = 0;
    __syncthreads();

    const int64_t j_first = max((int64_t)0, a.item_lo) / BN * BN;
    const int64_t j_end = min(a.n_items, a.item_hi);
    const int64_t n_tiles = (max(j_end, j_first) - j_first + BN - 1) / BN, per = (n_tiles + n_splits - 1) / n_splits;
    const int64_t j_begin = j_first + split * per * BN, j_stop = min(j_end, j_begin + per * BN);
    const int KC = exact_kc(D), KCP = KC + 4;     // item tile holds at most 64 dims at a time
    for (int64_t j0 = j_begin; j0 < j_stop; j0 += BN) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int kb = 0; kb < D; kb += KC) {
            const int kc4 = min(KC, D - kb) >> 2;
            if (kb) __syncthreads();
            for (int idx = tid; idx < BN * kc4; idx += EX_THREADS) {
                const int r = idx / kc4, c = idx % kc4;
                float4 v = f4zero();
                if (j0 + r < a.n_items) v = ld4(a.rep + (a.item_row0 + j0 + r) * D + kb + c * 4);
                st4(Is + r * KCP + c * 4, v);
            }
            __syncthreads();
            for (int c = 0; c < kc4; ++c) {
                float4 u[4], it[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) u[i] = ld4(Us + (ty * 4 + i) * DP + kb + c * 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) it[j] = ld4(Is + (tx + 16 * j) * KCP + c * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float s = acc[i][j];
                        s = fmaf(u[i].x, it[j].x, s); s = fmaf(u[i].y, it[j].y, s);
                        s = fmaf(u[i].z, it[j].z, s); s = fmaf(u[i].w, it[j].w, s);
                        acc[i][j] = s;
                    }
            }
        }
        // threshold filter (thr is stale-but-valid: never above the true k-th best so far)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            const float t = thr[r];
            const int64_t urow = ubase + r;
            if (urow >= n_eval) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t item = j0 + tx + 16 * j;
                const float s = acc[i][j];
                if (s > t && item >= a.item_lo && item < j_end) {
                    if (!item_masked(__ldg(a.user_ids + urow), (int32_t)item, a.mask_ptr, a.mask_items, a.banned)) {
                        const int slot = atomicAdd(cnt + r, 1);
                        keys[r * CAP + slot] = cand_key(s, (int32_t)item);
                        if (slot + 1 > CAP - BN) need_compact = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (need_compact) {
            for (int r = wid; r < BM; r += EX_THREADS / 32) {
                const int n = cnt[r];
                if (n > CAP - BN) {
                    warp_sort_desc(keys + r * CAP, n, lane);
                    if (lane == 0) {
                        cnt[r] = a.k;   // n > CAP - BN >= k
                        thr[r] = order_float((uint32_t)(keys[r * CAP + a.k - 1] >> 32));
                    }
                }
            }
            __syncthreads();
            if (tid == 0) need_compact = 0;
        }
        __syncthreads();
    }

    for (int r = wid; r < BM; r += EX_THREADS / 32) {
        const int64_t urow = ubase + r;
        if (urow >= n_eval) continue;
        const int n = cnt[r];
        const int64_t orow = a.out_rows ? (int64_t)__ldg(a.out_rows + urow) : urow;
        warp_sort_desc(keys + r * CAP, n, lane);
        if (n_splits > 1) {       // partial result of this item range; exact_merge_kernel picks the final top-k
            for (int q = lane; q < a.k; q += 32)
                a.split_keys[((size_t)urow * n_splits + split) * a.k + q] = q < n ? keys[r * CAP + q] : 0ULL;
            continue;
        }
        for (int q = lane; q < a.k; q += 32) {
            int32_t item = -1;
            float sc = -INFINITY;
            if (q < n) {
                const uint64_t key = keys[r * CAP + q];
                item = 0x7fffffff - (int32_t)(key & 0xffffffffu);
                sc = order_float((uint32_t)(key >> 32));
            }
            a.out_items[orow * a.k + q] = item;
            a.out_scores[orow * a.k + q] = sc;
        }
    }
}

// One warp per entry: merge the per-split top-k lists (keys are totally ordered: score, then lower item id).
__global__ void __launch_bounds__(256) exact_merge_kernel(const __grid_constant__ ExactArgs a) {
    extern __shared__ uint64_t merge_keys[];             // [8 warps][pow2 >= n_item_splits * k]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t n_eval = a.n_eval_dev ? min(a.n_eval, (int64_t)*a.n_eval_dev) : a.n_eval;
    if (n_eval > a.split_cap) return;                     // the scoring kernel ran unsplit and wrote the final lists
    const int64_t b = (int64_t)blockIdx.x * 8 + wid;
    if (b >= n_eval) return;
    const int n = a.n_item_splits * a.k;
    int n_pad = 32;
    while (n_pad < n) n_pad <<= 1;
    uint64_t *keys = merge_keys + (size_t)wid * n_pad;
    for (int e = lane; e < n; e += 32) keys[e] = a.split_keys[(size_t)b * n + e];
    __syncwarp();
    warp_sort_desc(keys, n, lane);
    const int64_t orow = a.out_rows ? (int64_t)__ldg(a.out_rows + b) : b;
    for (int q = lane; q < a.k; q += 32) {
        const uint64_t key = keys[q];
        const bool ok = key != 0ULL;
        a.out_items[orow * a.k + q] = ok ? 0x7fffffff - (int32_t)(key & 0xffffffffu) : -1;
        a.out_scores[orow * a.k + q] = ok ? order_float((uint32_t)(key >> 32)) : -INFINITY;
    }
}

// Dense score block for the reference's predict() API (model.py:118-123): out[b][j] = <rep[user_ids[b]], rep[item_row0 + j]>
// with the same fp32 FMA chain (ascending d) as the ranking kernels.  The evaluation path never materialises
// this matrix; predict() exists for callers of the reference API that want raw scores.
__global__ void __launch_bounds__(256) predict_scores_kernel(const float *__restrict__ rep, const int64_t *__restrict__ user_ids,
                                                             int64_t item_row0, int64_t n_items, int D, float *__restrict__ out) {
    extern __shared__ __align__(16) float urow[];
    const int64_t b = blockIdx.y;
    const float *u = rep + __ldg(user_ids + b) * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) urow[d] = u[d];
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_items) return;
    const float *it = rep + (item_row0 + j) * D;
    float s = 0.f;
    for (int d = 0; d < D; d += 4) {
        const float4 x = ld4(urow + d), y = ld4(it + d);
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    out[b * n_items + j] = s;
}

__global__ void hits_kernel(const int32_t *__restrict__ rec, int64_t n, int k, const int64_t *__restrict__ ptr,
                            const int32_t *__restrict__ items, float *hit) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * k) return;
    const int64_t u = idx / k;
    const int32_t j = rec[idx];
    int64_t lo = ptr[u], hi = ptr[u + 1];
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (items[mid] < j) lo = mid + 1; else hi = mid;
    }
    hit[idx] = (j >= 0 && lo < end && items[lo] == j) ? 1.f : 0.f;
}

// numpy's float32 add.reduce over a contiguous row of n <= 128 elements (numpy/core/src/umath/loops_utils.h,
// pairwise_sum): fewer than 8 elements are added left to right; otherwise eight accumulators take elements j, j + 8,
// ..., are combined as ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), and the n % 8 leftovers are added one by
// one; the result is added to the zero-initialised output.  Verified against np.sum on this image for n = 1..50.
__device__ __forceinline__ float numpy_row_sum(const float *a, int n) {
    float res;
    if (n < 8) {
        res = 0.f;
        for (int i = 0; i < n; ++i) res += a[i];
    } else {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
    }
    return 0.f + res;
}

constexpr int kMetricsMaxK = 32, kMetricsMaxTopks = 8;
struct MetricsArgs {
    int32_t topks[kMetricsMaxTopks];
    int n_topks, k;
};

// One thread per user: membership of its k recommended items in its eval list (the double loop of
// calculate_metrics, trainer.py:111-115), then per cut-off the hit count and the DCG = sum_j hit_j / log2(j + 2)
// with the reference's fp32 table and numpy's summation order, so that the per-user values -- and with them every
// mean the host takes -- carry the reference's bits.  Only 2 x n_topks x U floats leave the device.
__global__ void __launch_bounds__(256) user_metrics_kernel(const int32_t *__restrict__ rec, int64_t n, const int64_t *__restrict__ ptr,
                                                           const int32_t *__restrict__ items, const float *__restrict__ log2_tab,
                                                           const __grid_constant__ MetricsArgs ma, float *hit_num, float *dcg) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int k = ma.k;
    float gain[kMetricsMaxK];
    const int64_t lo0 = ptr[u], end = ptr[u + 1];
    for (int j = 0; j < k; ++j) {
        const int32_t it = rec[u * k + j];
        int64_t lo = lo0, hi = end;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (items[mid] < it) lo = mid + 1; else hi = mid;
        }
        const float hit = (it >= 0 && lo < end && items[lo] == it) ? 1.f : 0.f;
        gain[j] = hit;
    }
    for (int t = 0; t < ma.n_topks; ++t) {
        const int kk = ma.topks[t];
        float cnt = 0.f;
        float a[kMetricsMaxK];
        for (int j = 0; j < kk; ++j) {
            cnt += gain[j];                                  // exact: small integers
            a[j] = __fdiv_rn(gain[j], log2_tab[j]);          // hit / denominator, IEEE division like numpy's
        }
        hit_num[(int64_t)t * n + u] = cnt;
        dcg[(int64_t)t * n + u] = numpy_row_sum(a, kk);
    }
}

}  // namespace igcn

using namespace igcn;

extern "C" int igcn_user_metrics(const int32_t *rec, int64_t n_users, int32_t k, const int64_t *eval_ptr,
                                 const int32_t *eval_items, const float *log2_table, const int32_t *topks_host,
                                 int32_t n_topks, float *hit_num, float *dcg, void *stream) {
    IGCN_CHECK_ARG(rec && eval_ptr && eval_items && log2_table && topks_host && hit_num && dcg, "null pointer");
    IGCN_CHECK_ARG(k > 0 && k <= kMetricsMaxK, "k must be in [1, 32]");
    IGCN_CHECK_ARG(n_topks > 0 && n_topks <= kMetricsMaxTopks, "at most 8 cut-offs");
    MetricsArgs ma{};
    for (int t = 0; t < n_topks; ++t) {
        IGCN_CHECK_ARG(topks_host[t] > 0 && topks_host[t] <= k, "cut-off outside [1, k]");
        ma.topks[t] = topks_host[t];
    }
    ma.n_topks = n_topks; ma.k = k;
    if (n_users <= 0) return 0;
    user_metrics_kernel<<<(unsigned)((n_users + 255) / 256), 256, 0, as_stream(stream)>>>(rec, n_users, eval_ptr, eval_items, log2_table,
                                                                                           ma, hit_num, dcg);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_score_topk_exact(const float *rep, const int64_t *user_ids, int64_t n_eval, int64_t item_row0,
                                     int64_t n_items, int32_t D, const int64_t *mask_ptr, const int32_t *mask_items,
                                     int64_t item_lo, int64_t item_hi, const uint32_t *banned_bits, int32_t k,
                                     int32_t *out_items, float *out_scores, const int32_t *out_rows,
                                     const int32_t *n_eval_dev, int32_t n_item_splits, uint64_t *split_keys,
                                     int64_t split_cap, void *stream) {
    IGCN_CHECK_ARG(rep && user_ids && out_items && out_scores, "null pointer");
    IGCN_CHECK_ARG(D > 0 && D <= 256 && !(D & 3), "embedding size unsupported (need D % 4 == 0, D <= 256)");
    IGCN_CHECK_ARG(k > 0 && k <= CAP - BN, "k must be in [1, 128]");
    IGCN_CHECK_ARG(!mask_ptr || mask_items, "mask_ptr without mask_items");
    IGCN_CHECK_ARG(n_items > 0 && n_items < 0x7fffffff, "n_items out of range");
    if (n_eval <= 0) return 0;
    IGCN_CHECK_ARG(n_item_splits >= 1 && n_item_splits <= 64, "n_item_splits must be in [1, 64]");
    IGCN_CHECK_ARG(n_item_splits == 1 || (split_keys && split_cap > 0), "item splitting needs the split_keys scratch");
    ExactArgs a{rep, user_ids, n_eval, item_row0, n_items, D, mask_ptr, mask_items, item_lo, item_hi, banned_bits, k,
                out_items, out_scores, out_rows, n_eval_dev, n_item_splits, split_cap, split_keys};
    const size_t smem = ((size_t)BM * (D + 4) + (size_t)BN * (exact_kc(D) + 4)) * sizeof(float) + (size_t)BM * CAP * sizeof(uint64_t) + BM * 8;
    cudaError_t e = cudaFuncSetAttribute(score_topk_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("igcn_score_topk_exact: %s", cudaGetErrorString(e)); return (int)e; }
    const int64_t blocks = (n_eval + BM - 1) / BM;
    if (n_item_splits > 1) {
        // split pass over at most split_cap entries; both passes look at the device-side count and one of them exits
        const int64_t sblocks = (min(n_eval, split_cap) + BM - 1) / BM;
        score_topk_exact_kernel<<<dim3((unsigned)sblocks, (unsigned)n_item_splits), EX_THREADS, smem, as_stream(stream)>>>(a);
    }
    if (n_item_splits == 1 || n_eval > split_cap)
        score_topk_exact_kernel<<<dim3((unsigned)blocks, 1u), EX_THREADS, smem, as_stream(stream)>>>(a);
    if (n_item_splits > 1) {
        int n_pad = 32;
        while (n_pad < n_item_splits * k) n_pad <<= 1;
        const size_t msmem = (size_t)8 * n_pad * sizeof(uint64_t);
        e = cudaFuncSetAttribute(exact_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
        if (e != cudaSuccess) { set_error("igcn_score_topk_exact: %s", cudaGetErrorString(e)); return (int)e; }
        const int64_t mblocks = (min(n_eval, split_cap) + 7) / 8;
        exact_merge_kernel<<<(unsigned)mblocks, 256, msmem, as_stream(stream)>>>(a);
    }
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_predict_scores(const float *rep, const int64_t *user_ids, int64_t n_eval, int64_t item_row0,
                                   int64_t n_items, int32_t D, float *out, void *stream) {
    IGCN_CHECK_ARG(rep && user_ids && out, "null pointer");
    IGCN_CHECK_ARG(D > 0 && D <= 1024 && !(D & 3), "embedding size unsupported (need D % 4 == 0)");
    IGCN_CHECK_ARG(n_eval <= 65535, "at most 65535 users per call");
    if (n_eval <= 0 || n_items <= 0) return 0;
    predict_scores_kernel<<<dim3((unsigned)((n_items + 255) / 256), (unsigned)n_eval), 256, (size_t)D * sizeof(float), as_stream(stream)>>>(
        rep, user_ids, item_row0, n_items, D, out);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_hits(const int32_t *rec, int64_t n_users, int32_t k, const int64_t *eval_ptr,
                         const int32_t *eval_items, float *hit, void *stream) {
    IGCN_CHECK_ARG(rec && eval_ptr && eval_items && hit, "null pointer");
    IGCN_CHECK_ARG(k > 0, "k must be positive");
    const int64_t total = n_users * k;
    if (total <= 0) return 0;
    hits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(rec, n_users, k, eval_ptr, eval_items, hit);
    IGCN_CHECK_LAUNCH();
    return 0;
}
