// The BPR trainer step on B200: sampler, fused gather-dot-softplus forward, deterministic
// segmented gradient (sort + one owner per touched row, no float atomics), loss finalisation
// and Adam.  Reference call sites: dataset.py:119-131, model.py:108-116 / 293-299,
// trainer.py:231-248 / 294-320.  All of it is small gather/scatter work (3*B rows of 256 B).
#include "common.cuh"

namespace igcn {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ sampler
__global__ void sample_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int64_t col_offset,
                              int64_t n_users, int64_t n_items, int64_t B, uint64_t key, const uint64_t *step_dev, int64_t *out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    if (step_dev) key = mix64(key ^ mix64(*step_dev + 0x51ed2701ULL));
    uint64_t ctr = mix64(key ^ ((uint64_t)t * 0x9e3779b97f4a7c15ULL));
    auto next = [&]() { ctr += 0x9e3779b97f4a7c15ULL; return mix64(ctr); };
    // multiply-shift maps 64 uniform bits to [0, n) with bias < n / 2^64
    auto below = [&](uint64_t n) { return (int64_t)__umul64hi(next(), n); };
    int64_t u, lo, hi;
    do {
        u = below((uint64_t)n_users);
        lo = rowptr[u]; hi = rowptr[u + 1];
    } while (hi == lo);
    const int64_t pos = (int64_t)col[lo + below((uint64_t)(hi - lo))] - col_offset;
    int64_t neg;
    for (;;) {
        neg = below((uint64_t)n_items);
        const int32_t keyc = (int32_t)(neg + col_offset);
        int64_t a = lo, b = hi;
        while (a < b) {
            const int64_t mid = (a + b) >> 1;
            if (col[mid] < keyc) a = mid + 1; else b = mid;
        }
        if (a == hi || col[a] != keyc) break;
    }
    out[t * 3 + 0] = u; out[t * 3 + 1] = pos; out[t * 3 + 2] = neg;
}

// ------------------------------------------------------------------ forward
// Sum over the LANES lanes of a group, smallest stride first: the result is a balanced binary tree over the lane
// values in lane order -- ((l0 + l1) + (l2 + l3)) + ... -- so an aligned block of 2^j lanes holds ITS sum after j
// steps.  That is what makes the column-sharded step (engine.TrainStep, shard 'dims') bit-identical to one GPU: a rank
// that owns D / R columns computes exactly the subtree of its lanes (the other lanes contribute exact zeros) and
// bpr_combine_kernel adds the R subtree sums in the same tree order.
template <int LANES>
__device__ __forceinline__ float group_sum(float v, uint32_t gmask) {
#pragma unroll
    for (int o = 1; o < LANES; o <<= 1) v += __shfl_xor_sync(gmask, v, o, LANES);
    return v;
}

template <int LANES>
__device__ __forceinline__ uint32_t gmask_of() {
    if (LANES == 32) return 0xffffffffu;
    const uint32_t base = (LANES == 16) ? 0xffffu : 0xffu;
    return base << ((threadIdx.x & 31) & ~(LANES - 1));
}

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

struct BprDots { float ps, ns, qa, qb, qc; };

// <u * w, p>, <u * w, n> and the three squared norms of triple i over the D columns of T (group of LANES lanes)
template <int LANES>
__device__ __forceinline__ BprDots bpr_dots(const float *__restrict__ T, const float *__restrict__ L2T, const float *__restrict__ w,
                                            const int64_t *__restrict__ tri, int64_t i, int64_t off, int D, int lane, uint32_t gmask) {
    const bool active = lane * 4 < D;
    const int64_t u = tri[i * 3], p = tri[i * 3 + 1] + off, n = tri[i * 3 + 2] + off;
    float4 xu = f4zero(), xp = f4zero(), xn = f4zero(), ww = make_float4(1.f, 1.f, 1.f, 1.f);
    if (active) {
        xu = ld4(T + u * D + lane * 4); xp = ld4(T + p * D + lane * 4); xn = ld4(T + n * D + lane * 4);
        if (w) ww = ld4(w + lane * 4);
    }
    const float4 uw = make_float4(xu.x * ww.x, xu.y * ww.y, xu.z * ww.z, xu.w * ww.w);
    BprDots d;
    d.ps = group_sum<LANES>(dot4(uw, xp), gmask);
    d.ns = group_sum<LANES>(dot4(uw, xn), gmask);
    d.qa = d.qb = d.qc = 0.f;
    if (L2T) {
        float4 a = xu, b = xp, c = xn;
        if (L2T != T && active) { a = ld4(L2T + u * D + lane * 4); b = ld4(L2T + p * D + lane * 4); c = ld4(L2T + n * D + lane * 4); }
        d.qa = group_sum<LANES>(dot4(a, a), gmask);
        d.qb = group_sum<LANES>(dot4(b, b), gmask);
        d.qc = group_sum<LANES>(dot4(c, c), gmask);
    }
    return d;
}

__device__ __forceinline__ void bpr_finish(const BprDots &d, bool has_l2, int64_t i, float *sp, float *sig, float *l2) {
    const float x = d.ns - d.ps;
    sp[i] = x > 20.f ? x : log1pf(expf(x));      // F.softplus (beta 1, threshold 20)
    sig[i] = 1.f / (1.f + expf(-x));
    if (l2) l2[i] = has_l2 ? (d.qa + d.qb) + d.qc : 0.f;
}

template <int LANES>
__global__ void __launch_bounds__(kThreads) bpr_fwd_kernel(const float *__restrict__ T, const float *__restrict__ L2T,
                                                           const float *__restrict__ w, const int64_t *__restrict__ tri,
                                                           int64_t B, int64_t off, int D, float *sp, float *sig, float *l2) {
    const int lane = threadIdx.x % LANES;
    const uint32_t gmask = gmask_of<LANES>();
    const int64_t i = (int64_t)blockIdx.x * (kThreads / LANES) + threadIdx.x / LANES;
    if (i >= B) return;
    const BprDots d = bpr_dots<LANES>(T, L2T, w, tri, i, off, D, lane, gmask);
    if (lane == 0) bpr_finish(d, L2T != nullptr, i, sp, sig, l2);
}

// ---- column-sharded step: every rank holds D / R columns of the tables.  bpr_partial_kernel writes this rank's five
// partial sums of triple i into record i of EVERY rank's exchange buffer (peer stores over NVLink, 20 bytes per triple
// and peer); after the device barrier bpr_combine_kernel adds the R partials in group_sum's tree order.
// Buffer layout (floats): [2 step parities][R ranks][B triples][8]: slots 0..4 = ps, ns, qa, qb, qc of the main triples,
// 5..6 = ps, ns of the auxiliary triples.  The parity (igcn_step_state.step & 1) double-buffers consecutive steps.
struct PartArgs {
    float *peer[IGCN_MAX_PEERS];
    int n_peers, rank, slot0;
    const igcn_step_state *state;
    int64_t cap;                       // B capacity of the buffer (records per rank and parity)
};

template <int LANES>
__global__ void __launch_bounds__(kThreads) bpr_partial_kernel(const float *__restrict__ T, const float *__restrict__ L2T,
                                                               const float *__restrict__ w, const int64_t *__restrict__ tri,
                                                               int64_t B, int64_t off, int D, const __grid_constant__ PartArgs pa) {
    const int lane = threadIdx.x % LANES;
    const uint32_t gmask = gmask_of<LANES>();
    const int64_t i = (int64_t)blockIdx.x * (kThreads / LANES) + threadIdx.x / LANES;
    if (i >= B) return;
    const BprDots d = bpr_dots<LANES>(T, L2T, w, tri, i, off, D, lane, gmask);
    if (lane != 0) return;
    const int64_t parity = (int64_t)(pa.state->step & 1ULL);
    const int64_t rec = ((parity * pa.n_peers + pa.rank) * pa.cap + i) * 8 + pa.slot0;
    for (int p = 0; p < pa.n_peers; ++p) {
        float *dst = pa.peer[p] + rec;
        dst[0] = d.ps; dst[1] = d.ns;
        if (L2T) { dst[2] = d.qa; dst[3] = d.qb; dst[4] = d.qc; }
    }
}

__device__ __forceinline__ float tree_sum(const float *v, int n) {      // n = 2, 4 or 8 partials, group_sum's order
    float t[IGCN_MAX_PEERS];
    for (int r = 0; r < n; ++r) t[r] = v[r];
    for (int o = 1; o < n; o <<= 1)
        for (int r = 0; r < n; r += 2 * o) t[r] = t[r] + t[r + o];
    return t[0];
}

__global__ void __launch_bounds__(kThreads) bpr_combine_kernel(const float *__restrict__ parts, int64_t B, int64_t cap, int n_peers,
                                                               int slot0, int has_l2, const igcn_step_state *__restrict__ state,
                                                               float *sp, float *sig, float *l2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int64_t parity = (int64_t)(state->step & 1ULL);
    float v[5][IGCN_MAX_PEERS];
    const int nk = has_l2 ? 5 : 2;
    for (int r = 0; r < n_peers; ++r) {
        const float *rec = parts + ((parity * n_peers + r) * cap + i) * 8 + slot0;
        for (int k = 0; k < nk; ++k) v[k][r] = __ldcg(rec + k);
    }
    BprDots d;
    d.ps = tree_sum(v[0], n_peers);
    d.ns = tree_sum(v[1], n_peers);
    d.qa = d.qb = d.qc = 0.f;
    if (has_l2) { d.qa = tree_sum(v[2], n_peers); d.qb = tree_sum(v[3], n_peers); d.qc = tree_sum(v[4], n_peers); }
    bpr_finish(d, has_l2 != 0, i, sp, sig, l2);
}

// ------------------------------------------------------------------ loss finalisation (one CTA, fixed order)
__device__ float block_sum_ordered(const float *__restrict__ v, int64_t n, float *sm) {
    // thread t sums a contiguous slice sequentially, then thread 0 adds the 256 partials in order
    const int64_t per = (n + kThreads - 1) / kThreads;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = min(n, lo + per);
    float s = 0.f;
    for (int64_t k = lo; k < hi; ++k) s += v[k];
    sm[threadIdx.x] = s;
    __syncthreads();
    float tot = 0.f;
    if (threadIdx.x == 0)
        for (int k = 0; k < kThreads; ++k) tot += sm[k];
    __syncthreads();
    return tot;
}

__global__ void __launch_bounds__(kThreads) loss_finalize_kernel(const float *sp, const float *l2, const float *aux, int64_t B,
                                                                 int64_t B_aux, float l2_reg, float aux_reg, float *loss, double *acc) {
    __shared__ float sm[kThreads];
    const float s0 = block_sum_ordered(sp, B, sm);
    const float s1 = l2 ? block_sum_ordered(l2, B, sm) : 0.f;
    const float s2 = aux ? block_sum_ordered(aux, B_aux, sm) : 0.f;
    if (threadIdx.x == 0) {
        float v = s0 / (float)B;
        float reg = 0.f;
        if (l2) reg = l2_reg * (s1 / (float)B);
        if (aux) reg += aux_reg * (s2 / (float)B_aux);
        v += reg;
        loss[0] = v;
        if (acc) {   // AverageMeter.update(loss.item(), B): the reference weights by the LAST inputs' batch size
            const double n = (double)(aux ? B_aux : B);
            acc[0] += (double)v * n;
            acc[1] += n;
        }
    }
}

// ------------------------------------------------------------------ scatter plan: bitonic sort of (row id, slot)
__global__ void __launch_bounds__(1024) plan_kernel(const int64_t *__restrict__ tri, int64_t B, int64_t off, int n_pad,
                                                    int32_t *order, int32_t *seg_start, int64_t *seg_row, int32_t *n_seg,
                                                    uint32_t *touched, int64_t touched_words) {
    extern __shared__ uint64_t keys[];
    __shared__ int warp_tot[32];
    const int n = (int)(3 * B);
    if (touched)
        for (int64_t i = threadIdx.x; i < touched_words; i += blockDim.x) touched[i] = 0u;
    for (int s = threadIdx.x; s < n_pad; s += blockDim.x) {
        uint64_t k = ~0ULL;
        if (s < n) {
            const int kind = s / (int)B, i = s % (int)B;
            const int64_t id = tri[(int64_t)i * 3 + kind] + (kind ? off : 0);
            k = ((uint64_t)id << 32) | (uint32_t)s;
        }
        keys[s] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n_pad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < n_pad / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    // segment heads + exclusive scan of head flags (thread t owns a contiguous slice)
    const int per = (n_pad + blockDim.x - 1) / blockDim.x;
    const int lo = threadIdx.x * per, hi = min(n, lo + per);
    int cnt = 0;
    for (int s = lo; s < hi; ++s) cnt += (s == 0 || (keys[s] >> 32) != (keys[s - 1] >> 32));
    int incl = cnt;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += x;
        }
        warp_tot[lane] = v;
    }
    __syncthreads();
    int seg = incl - cnt + (wid ? warp_tot[wid - 1] : 0);
    for (int s = lo; s < hi; ++s) {
        const uint64_t k = keys[s];
        order[s] = (int32_t)(k & 0xffffffffu);
        if (s == 0 || (k >> 32) != (keys[s - 1] >> 32)) {
            seg_start[seg] = s;
            seg_row[seg] = (int64_t)(k >> 32);
            if (touched) atomicOr(touched + (k >> 37), 1u << ((k >> 32) & 31));   // OR is order independent
            ++seg;
        }
    }
    if (threadIdx.x == blockDim.x - 1) {
        const int total = warp_tot[31];
        n_seg[0] = total;
        seg_start[total] = n;
    }
}

// Fast path (3B <= 8192, ids < 2^19): 32-bit keys (id << 13 | slot), ITEMS keys per thread in registers.
// Bitonic network: strides below ITEMS are register exchanges, strides inside a warp are shuffles, only
// the strides across warps go through shared memory (15 of the 91 steps at 8192 keys).
constexpr int kPlanSlotBits = 13;

__device__ __forceinline__ void cmpx(uint32_t &a, uint32_t &b, bool up) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = up ? lo : hi;
    b = up ? hi : lo;
}

template <int ITEMS>
__global__ void __launch_bounds__(1024) plan_fast_kernel(const int64_t *__restrict__ tri, int64_t B, int64_t off, int32_t *order,
                                                         int32_t *seg_start, int64_t *seg_row, int32_t *n_seg, uint32_t *touched,
                                                         int64_t touched_words) {
    __shared__ uint32_t sm[ITEMS * 1024];
    __shared__ uint32_t last_key[1024];
    __shared__ int warp_tot[32];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int n = (int)(3 * B);
    constexpr int n_pad = ITEMS * 1024;
    if (touched)
        for (int64_t i = t; i < touched_words; i += 1024) touched[i] = 0u;
    uint32_t key[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int s = t * ITEMS + r;
        uint32_t k = 0xffffffffu;
        if (s < n) {
            const int kind = s / (int)B, i = s % (int)B;
            const int64_t id = tri[(int64_t)i * 3 + kind] + (kind ? off : 0);
            k = ((uint32_t)id << kPlanSlotBits) | (uint32_t)s;
        }
        key[r] = k;
    }
    for (int size = 2; size <= n_pad; size <<= 1) {
        int stride = size >> 1;
        for (; stride >= ITEMS; stride >>= 1) {
            const int m = stride / ITEMS;                     // partner thread = t ^ m, same register
            const bool lower = (t & m) == 0;
            if (m < 32) {
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) {
                    const uint32_t other = __shfl_xor_sync(0xffffffffu, key[r], m);
                    const bool up = (((t * ITEMS + r) & size) == 0);
                    key[r] = (lower == up) ? min(key[r], other) : max(key[r], other);
                }
            } else {
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) sm[r * 1024 + t] = key[r];
                __syncthreads();
#pragma unroll
                for (int r = 0; r < ITEMS; ++r) {
                    const uint32_t other = sm[r * 1024 + (t ^ m)];
                    const bool up = (((t * ITEMS + r) & size) == 0);
                    key[r] = (lower == up) ? min(key[r], other) : max(key[r], other);
                }
                __syncthreads();
            }
        }
#pragma unroll
        for (int s = ITEMS / 2; s > 0; s >>= 1) {
            if (s < size) {
#pragma unroll
                for (int r = 0; r < ITEMS; ++r)
                    if ((r & s) == 0) cmpx(key[r], key[r | s], (((t * ITEMS + r) & size) == 0));
            }
        }
    }
    // thread t now owns sorted positions [t * ITEMS, +ITEMS)
    last_key[t] = key[ITEMS - 1];
    __syncthreads();
    uint32_t prev_id = t ? (last_key[t - 1] >> kPlanSlotBits) : 0xffffffffu;
    int cnt = 0;
    bool head[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int s = t * ITEMS + r;
        const uint32_t id = key[r] >> kPlanSlotBits;
        head[r] = s < n && (s == 0 || id != prev_id);
        cnt += head[r] ? 1 : 0;
        prev_id = id;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += x;
        }
        warp_tot[lane] = v;
    }
    __syncthreads();
    int seg = incl - cnt + (wid ? warp_tot[wid - 1] : 0);
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const int s = t * ITEMS + r;
        if (s < n) {
            order[s] = (int32_t)(key[r] & ((1u << kPlanSlotBits) - 1u));
            if (head[r]) {
                const uint32_t id = key[r] >> kPlanSlotBits;
                seg_start[seg] = s;
                seg_row[seg] = (int64_t)id;
                if (touched) atomicOr(touched + (id >> 5), 1u << (id & 31));
                ++seg;
            }
        }
    }
    if (t == 1023) {
        const int total = warp_tot[31];
        n_seg[0] = total;
        seg_start[total] = n;
    }
}

// ------------------------------------------------------------------ backward: one group per touched row
template <int LANES>
__global__ void __launch_bounds__(kThreads) bpr_bwd_kernel(const float *__restrict__ T, const float *__restrict__ w,
                                                           const int64_t *__restrict__ tri, int64_t B, int64_t off, int D,
                                                           const float *__restrict__ sig, float cscale, float lam, int l2_on,
                                                           const int32_t *__restrict__ order, const int32_t *__restrict__ seg_start,
                                                           const int64_t *__restrict__ seg_row, const int32_t *__restrict__ n_seg,
                                                           float *G, int accumulate) {
    const int lane = threadIdx.x % LANES;
    const int64_t seg = (int64_t)blockIdx.x * (kThreads / LANES) + threadIdx.x / LANES;
    if (seg >= *n_seg) return;
    if (lane * 4 >= D) return;
    const int64_t row = seg_row[seg];
    const int s0 = seg_start[seg], s1 = seg_start[seg + 1];
    float4 ww = make_float4(1.f, 1.f, 1.f, 1.f);
    if (w) ww = ld4(w + lane * 4);
    const float4 self = ld4(T + row * D + lane * 4);
    float4 acc = f4zero();
    for (int s = s0; s < s1; ++s) {
        const int slot = order[s];
        const int kind = slot / (int)B, i = slot % (int)B;
        const float c = cscale * sig[i];
        float4 g;
        if (kind == 0) {
            const int64_t p = tri[(int64_t)i * 3 + 1] + off, n = tri[(int64_t)i * 3 + 2] + off;
            const float4 xp = ld4(T + p * D + lane * 4), xn = ld4(T + n * D + lane * 4);
            g = make_float4(c * ww.x * (xn.x - xp.x), c * ww.y * (xn.y - xp.y), c * ww.z * (xn.z - xp.z), c * ww.w * (xn.w - xp.w));
        } else {
            const int64_t u = tri[(int64_t)i * 3];
            const float4 xu = ld4(T + u * D + lane * 4);
            const float cs = kind == 1 ? -c : c;
            g = make_float4(cs * ww.x * xu.x, cs * ww.y * xu.y, cs * ww.z * xu.z, cs * ww.w * xu.w);
        }
        if (l2_on) fma4(g, lam, self);
        add4(acc, g);
    }
    float *dst = G + row * D + lane * 4;
    if (accumulate) add4(acc, ld4(dst));
    st4(dst, acc);
}

// dw: stage 1 sums 64 triples per CTA in order, stage 2 adds the CTA partials in order.  The gathers of 8 triples are
// issued together (they do not depend on the running sum), the fused multiply-adds stay in triple order.
__global__ void __launch_bounds__(128) dw_stage1(const float *__restrict__ T, const int64_t *__restrict__ tri, int64_t B, int64_t off,
                                                 int D, const float *__restrict__ sig, float cscale, float *scratch) {
    const int d = threadIdx.x;
    if (d >= D) return;
    const int64_t lo = (int64_t)blockIdx.x * 64, hi = min(B, lo + 64);
    float t = 0.f;
    for (int64_t i0 = lo; i0 < hi; i0 += 8) {
        float a[8], b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t i = i0 + q;
            a[q] = b[q] = 0.f;
            if (i < hi) {
                const int64_t u = tri[i * 3], p = tri[i * 3 + 1] + off, n = tri[i * 3 + 2] + off;
                a[q] = cscale * sig[i] * T[u * D + d];
                b[q] = T[n * D + d] - T[p * D + d];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (i0 + q < hi) t = fmaf(a[q], b[q], t);
    }
    scratch[(int64_t)blockIdx.x * D + d] = t;
}

__global__ void dw_stage2(const float *__restrict__ scratch, int64_t n_blocks, int D, float *dw) {
    const int d = threadIdx.x;
    if (d >= D) return;
    float t = 0.f;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (b0 + q < n_blocks) ? scratch[(b0 + q) * D + d] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (b0 + q < n_blocks) t += v[q];
    }
    dw[d] += t;
}

template <int LANES>
__global__ void __launch_bounds__(kThreads) l2_rows_kernel(const float *__restrict__ E, float *dE, int D, float coef,
                                                           const int32_t *__restrict__ seg_start, const int64_t *__restrict__ seg_row,
                                                           const int32_t *__restrict__ n_seg) {
    const int lane = threadIdx.x % LANES;
    const int64_t seg = (int64_t)blockIdx.x * (kThreads / LANES) + threadIdx.x / LANES;
    if (seg >= *n_seg || lane * 4 >= D) return;
    const int64_t row = seg_row[seg];
    const float m = coef * (float)(seg_start[seg + 1] - seg_start[seg]);
    float4 g = ld4(dE + row * D + lane * 4);
    fma4(g, m, ld4(E + row * D + lane * 4));
    st4(dE + row * D + lane * 4, g);
}

// ------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(kThreads) adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                        float *__restrict__ v, int64_t n4, int64_t n, float b1, float b2,
                                                        float step_size, float inv_sqrt_bc2, float eps,
                                                        const igcn_step_state *__restrict__ state) {
    if (state) { step_size = state->adam_step_size; inv_sqrt_bc2 = state->adam_inv_sqrt_bc2; }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = ld4(p + i * 4), gg = ld4(g + i * 4), mm = ld4(m + i * 4), vv = ld4(v + i * 4);
#define IGCN_ADAM1(c)                                                      \
        mm.c = mm.c + (gg.c - mm.c) * (1.f - b1);                          \
        vv.c = vv.c * b2 + (1.f - b2) * gg.c * gg.c;                       \
        pp.c = pp.c - step_size * (mm.c / (sqrtf(vv.c) * inv_sqrt_bc2 + eps));
        IGCN_ADAM1(x) IGCN_ADAM1(y) IGCN_ADAM1(z) IGCN_ADAM1(w)
        st4(p + i * 4, pp); st4(m + i * 4, mm); st4(v + i * 4, vv);
    }
    // tail (n not a multiple of 4)
    const int64_t t = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && t < n) {
        float mm = m[t], vv = v[t];
        const float gg = g[t];
        mm = mm + (gg - mm) * (1.f - b1);
        vv = vv * b2 + (1.f - b2) * gg * gg;
        p[t] -= step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
        m[t] = mm; v[t] = vv;
    }
}

__global__ void step_tick_kernel(igcn_step_state *s, float lr, float b1, float b2) {
    const uint64_t t = s->step + 1;
    s->step = t;
    s->adam_step_size = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
    s->adam_inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
}

template <typename F8, typename F16, typename F32>
static void by_lanes(int D, F8 f8, F16 f16, F32 f32) {
    if (D <= 32) f8(); else if (D <= 64) f16(); else f32();
}

}  // namespace igcn

using namespace igcn;

#define IGCN_CHECK_D(D) IGCN_CHECK_ARG((D) > 0 && (D) <= 128 && !((D)&3), "embedding size unsupported (need D % 4 == 0, D <= 128)")

extern "C" int igcn_sample_triples(const int64_t *rowptr, const int32_t *col, int64_t col_offset, int64_t n_users,
                                   int64_t n_items, int64_t B, uint64_t seed, uint64_t step, const uint64_t *step_dev, int64_t *out,
                                   void *stream) {
    IGCN_CHECK_ARG(rowptr && col && out, "null pointer");
    IGCN_CHECK_ARG(n_users > 0 && n_items > 0 && B >= 0, "bad sizes");
    if (B == 0) return 0;
    const uint64_t key = mix64(seed * 0x9e3779b97f4a7c15ULL + mix64(step + 0x1234567ULL));
    sample_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(rowptr, col, col_offset, n_users, n_items, B, key, step_dev, out);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_bpr_fwd(const float *table, const float *l2_table, const float *w, const int64_t *triples, int64_t B,
                            int64_t item_offset, int32_t D, float *sp, float *sig, float *l2, void *stream) {
    IGCN_CHECK_ARG(table && triples && sp && sig, "null pointer");
    IGCN_CHECK_D(D);
    IGCN_CHECK_ARG(!l2_table || l2, "l2_table given without l2 output");
    if (B <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    by_lanes(D,
             [&] { bpr_fwd_kernel<8><<<(unsigned)((B + 31) / 32), kThreads, 0, st>>>(table, l2_table, w, triples, B, item_offset, D, sp, sig, l2); },
             [&] { bpr_fwd_kernel<16><<<(unsigned)((B + 15) / 16), kThreads, 0, st>>>(table, l2_table, w, triples, B, item_offset, D, sp, sig, l2); },
             [&] { bpr_fwd_kernel<32><<<(unsigned)((B + 7) / 8), kThreads, 0, st>>>(table, l2_table, w, triples, B, item_offset, D, sp, sig, l2); });
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_bpr_partial(const float *table, const float *l2_table, const float *w, const int64_t *triples, int64_t B,
                                int64_t item_offset, int32_t D, float *const *parts_peer_host, int32_t n_peers, int32_t rank,
                                int32_t slot0, int64_t cap, const igcn_step_state *state_dev, void *stream) {
    IGCN_CHECK_ARG(table && triples && parts_peer_host && state_dev, "null pointer");
    IGCN_CHECK_D(D);
    IGCN_CHECK_ARG(n_peers == 2 || n_peers == 4 || n_peers == 8, "column sharding supports 2, 4 or 8 ranks");
    IGCN_CHECK_ARG(rank >= 0 && rank < n_peers && B <= cap && slot0 >= 0 && slot0 + (l2_table ? 5 : 2) <= 8, "bad rank / capacity / slot");
    if (B <= 0) return 0;
    PartArgs pa{};
    for (int p = 0; p < n_peers; ++p) pa.peer[p] = parts_peer_host[p];
    pa.n_peers = n_peers; pa.rank = rank; pa.slot0 = slot0; pa.state = state_dev; pa.cap = cap;
    cudaStream_t st = as_stream(stream);
    by_lanes(D,
             [&] { bpr_partial_kernel<8><<<(unsigned)((B + 31) / 32), kThreads, 0, st>>>(table, l2_table, w, triples, B, item_offset, D, pa); },
             [&] { bpr_partial_kernel<16><<<(unsigned)((B + 15) / 16), kThreads, 0, st>>>(table, l2_table, w, triples, B, item_offset, D, pa); },
             [&] { bpr_partial_kernel<32><<<(unsigned)((B + 7) / 8), kThreads, 0, st>>>(table, l2_table, w, triples, B, item_offset, D, pa); });
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_bpr_combine(const float *parts, int64_t B, int64_t cap, int32_t n_peers, int32_t slot0, int32_t has_l2,
                                const igcn_step_state *state_dev, float *sp, float *sig, float *l2, void *stream) {
    IGCN_CHECK_ARG(parts && state_dev && sp && sig, "null pointer");
    IGCN_CHECK_ARG(n_peers == 2 || n_peers == 4 || n_peers == 8, "column sharding supports 2, 4 or 8 ranks");
    IGCN_CHECK_ARG(B <= cap && slot0 >= 0 && slot0 + (has_l2 ? 5 : 2) <= 8, "bad capacity / slot");
    if (B <= 0) return 0;
    bpr_combine_kernel<<<(unsigned)((B + kThreads - 1) / kThreads), kThreads, 0, as_stream(stream)>>>(parts, B, cap, n_peers, slot0,
                                                                                                       has_l2, state_dev, sp, sig, l2);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_loss_finalize(const float *sp, const float *l2, const float *aux_sp, int64_t B, int64_t B_aux,
                                  float l2_reg, float aux_reg, float *loss, double *acc, void *stream) {
    IGCN_CHECK_ARG(sp && loss && B > 0, "null pointer or empty batch");
    IGCN_CHECK_ARG(!aux_sp || B_aux > 0, "empty aux batch");
    loss_finalize_kernel<<<1, kThreads, 0, as_stream(stream)>>>(sp, l2, aux_sp, B, B_aux, l2_reg, aux_reg, loss, acc);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_bpr_plan(const int64_t *triples, int64_t B, int64_t item_offset, int64_t n_rows, int32_t *order,
                             int32_t *seg_start, int64_t *seg_row, int32_t *n_seg, uint32_t *touched_bits, void *stream) {
    IGCN_CHECK_ARG(triples && order && seg_start && seg_row && n_seg, "null pointer");
    IGCN_CHECK_ARG(B > 0 && 3 * B <= 16384, "batch size must satisfy 0 < 3*B <= 16384");
    IGCN_CHECK_ARG(n_rows > 0, "n_rows (upper bound of the row ids) must be positive");
    const int64_t words = (n_rows + 31) / 32;
    if (3 * B <= 8192 && n_rows <= (1 << (32 - kPlanSlotBits))) {
        cudaStream_t st = as_stream(stream);
        const int64_t n = 3 * B;
        if (n <= 1024) plan_fast_kernel<1><<<1, 1024, 0, st>>>(triples, B, item_offset, order, seg_start, seg_row, n_seg, touched_bits, words);
        else if (n <= 2048) plan_fast_kernel<2><<<1, 1024, 0, st>>>(triples, B, item_offset, order, seg_start, seg_row, n_seg, touched_bits, words);
        else if (n <= 4096) plan_fast_kernel<4><<<1, 1024, 0, st>>>(triples, B, item_offset, order, seg_start, seg_row, n_seg, touched_bits, words);
        else plan_fast_kernel<8><<<1, 1024, 0, st>>>(triples, B, item_offset, order, seg_start, seg_row, n_seg, touched_bits, words);
        IGCN_CHECK_LAUNCH();
        return 0;
    }
    int n_pad = 2;
    while (n_pad < 3 * B) n_pad <<= 1;
    const size_t smem = (size_t)n_pad * sizeof(uint64_t);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
        if (e != cudaSuccess) { set_error("igcn_bpr_plan: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    plan_kernel<<<1, 1024, smem, as_stream(stream)>>>(triples, B, item_offset, n_pad, order, seg_start, seg_row, n_seg,
                                                      touched_bits, words);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_bpr_bwd(const float *table, const float *w, const int64_t *triples, int64_t B, int64_t item_offset,
                            int32_t D, const float *sig, float scale, float l2_coef, int32_t l2_on_table,
                            const int32_t *order, const int32_t *seg_start, const int64_t *seg_row, const int32_t *n_seg,
                            float *G, int32_t accumulate, float *dw, float *dw_scratch, void *stream) {
    IGCN_CHECK_ARG(table && triples && sig && order && seg_start && seg_row && n_seg && G, "null pointer");
    IGCN_CHECK_D(D);
    IGCN_CHECK_ARG(!dw || (w && dw_scratch), "dw needs w and dw_scratch");
    if (B <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    const float cscale = scale / (float)B;
    const float lam = scale * 2.f * l2_coef / (float)B;
    const int64_t max_seg = 3 * B;
#define IGCN_BWD_LAUNCH(L)                                                                                       \
    bpr_bwd_kernel<L><<<(unsigned)((max_seg + kThreads / L - 1) / (kThreads / L)), kThreads, 0, st>>>(           \
        table, w, triples, B, item_offset, D, sig, cscale, lam, l2_on_table, order, seg_start, seg_row, n_seg, G, accumulate)
    by_lanes(D, [&] { IGCN_BWD_LAUNCH(8); }, [&] { IGCN_BWD_LAUNCH(16); }, [&] { IGCN_BWD_LAUNCH(32); });
    if (dw) {
        const int64_t nb = (B + 63) / 64;
        dw_stage1<<<(unsigned)nb, 128, 0, st>>>(table, triples, B, item_offset, D, sig, cscale, dw_scratch);
        dw_stage2<<<1, 128, 0, st>>>(dw_scratch, nb, D, dw);
    }
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_bpr_dw(const float *table, const int64_t *triples, int64_t B, int64_t item_offset, int32_t D, const float *sig,
                           float scale, float *dw, float *dw_scratch, void *stream) {
    IGCN_CHECK_ARG(table && triples && sig && dw && dw_scratch, "null pointer");
    IGCN_CHECK_D(D);
    if (B <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    const int64_t nb = (B + 63) / 64;
    dw_stage1<<<(unsigned)nb, 128, 0, st>>>(table, triples, B, item_offset, D, sig, scale / (float)B, dw_scratch);
    dw_stage2<<<1, 128, 0, st>>>(dw_scratch, nb, D, dw);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_l2_rows_bwd(const float *E, float *dE, int32_t D, float coef, const int32_t *seg_start,
                                const int64_t *seg_row, const int32_t *n_seg, int64_t max_seg, void *stream) {
    IGCN_CHECK_ARG(E && dE && seg_start && seg_row && n_seg, "null pointer");
    IGCN_CHECK_D(D);
    if (max_seg <= 0) return 0;
    cudaStream_t st = as_stream(stream);
#define IGCN_L2_LAUNCH(L) \
    l2_rows_kernel<L><<<(unsigned)((max_seg + kThreads / L - 1) / (kThreads / L)), kThreads, 0, st>>>(E, dE, D, coef, seg_start, seg_row, n_seg)
    by_lanes(D, [&] { IGCN_L2_LAUNCH(8); }, [&] { IGCN_L2_LAUNCH(16); }, [&] { IGCN_L2_LAUNCH(32); });
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_adam(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                         float eps, int64_t t, const igcn_step_state *state_dev, void *stream) {
    IGCN_CHECK_ARG(p && g && m && v, "null pointer");
    IGCN_CHECK_ARG(state_dev || t >= 1, "step count starts at 1");
    if (t < 1) t = 1;
    if (n <= 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    const float step_size = (float)((double)lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + kThreads - 1) / kThreads;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adam_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(p, g, m, v, n4, n, beta1, beta2, step_size, inv_sqrt_bc2, eps, state_dev);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_step_tick(igcn_step_state *state_dev, float lr, float beta1, float beta2, void *stream) {
    IGCN_CHECK_ARG(state_dev, "null pointer");
    step_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(state_dev, lr, beta1, beta2);
    IGCN_CHECK_LAUNCH();
    return 0;
}
