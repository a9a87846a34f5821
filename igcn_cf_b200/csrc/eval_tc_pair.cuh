// Candidate kernel of the tensor-core scoring path with TWO user tiles (256 users) per CTA; included by eval_tc.cu.
//
// The time decomposition of score_tc_kernel (eval_tc.cu, IGCN_TC_EXPERIMENT) showed that 46 % of its time is a
// floor set by streaming the item image: every 128-user CTA pulls all item tiles through L2->SM.  Here an item
// tile (N = 128) is fetched once for 256 users: two A tiles per CTA, two UMMA issues per item tile into a ring
// of four TMEM accumulators (two per user tile, double-buffered), and EIGHT epilogue warps -- four per user
// tile, each thread still owns one user row and one candidate list.  To fit 256 lists in shared memory the
// lists are 56 slots (KEEP = 28, 16-column filter steps, so 16 free slots suffice before a step); k <= 20.
// Method, operand images, error bound and the exact finalize pass are those of eval_tc.cu.
// Warp roles (12 warps): 0 bulk-TMA producer, 1 MMA issuer, 2 / 3 mask builders (one per user tile),
// 4..11 epilogue (user tile h = (warp - 4) / 4, TMEM lane quarter = warp % 4).
#pragma once
#include "tc_common.cuh"

namespace igcn {

constexpr int T3_BN = 128;        // items per tile (UMMA N)
constexpr int T3_CAP = 56;        // candidate slots per user row
constexpr int T3_KEEP = 28;       // the threshold never rises above the T3_KEEP-th best upper bound
constexpr int T3_CHUNK = 16;      // columns filtered per step: free slots needed before a step
constexpr int T3_STAGES = 2;      // item-tile smem stages
constexpr int T3_ACC = 4;         // TMEM accumulators: index = 2 * user-tile half + (item tile & 1)
constexpr int T3_THREADS = 12 * 32;
constexpr int T3_STAGE_W = 20;    // words per row of the chunk staging area (16 B aligned, conflict-free STS.128)

struct Tc3Args {
    const uint8_t *a_img;
    const uint8_t *b_img;             // packed with 128-row item tiles
    int n_utiles, n_splits, n_head, kcores;      // the first n_head user-tile pairs are not split
    int64_t n_eval, n_items, item_lo, item_hi;
    const uint32_t *banned;
    const int32_t *mask_tile_ptr;     // [n_utiles, n_buckets + 1], buckets of 256 items
    int n_buckets;
    const uint16_t *mask_entries;     // (row << 8) | col within the 256-item bucket
    int32_t *cand_items;              // [n_eval, n_splits, T3_CAP]
    int32_t *cand_cnt;                // [n_eval, n_splits]
    float *cand_thr;                  // [n_eval, n_splits]
};

struct Tc3Smem {
    uint64_t full[T3_STAGES], empty[T3_STAGES], a_full, tmem_full[T3_ACC], tmem_empty[T3_ACC], mask_full[T3_ACC];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// Lane-parallel compaction (see compact_lanes in eval_tc.cu), for the shorter lists of this kernel.
__device__ __forceinline__ void compact_lanes3(uint64_t *mybuf, int &cnt, float &thr) {
    const bool act = cnt > T3_KEEP + 4;
    const int n = act ? cnt : 0;
    const int nmax = (__reduce_max_sync(0xffffffffu, n) + 3) & ~3;       // loops run in groups of 4 (T3_CAP % 4 == 0)
    const float *sc = reinterpret_cast<const float *>(mybuf);          // score of entry j at sc[2 * j]
    float vmax = -INFINITY, vmin = INFINITY;
    for (int j = 0; j < nmax; j += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = sc[2 * (j + u)];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j + u < n) { vmax = fmaxf(vmax, v[u]); vmin = fminf(vmin, v[u]); }
    }
    float lo = (thr == -INFINITY) ? vmin : thr;      // invariant: count(score >= lo) >= T3_KEEP
    float hi = vmax;
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {
        const float q = 0.25f * (hi - lo);
        const float m1 = lo + q, m2 = lo + 2.f * q, m3 = lo + 3.f * q;
        int c1 = 0, c2 = 0, c3 = 0;
        for (int j = 0; j < nmax; j += 4) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (j + u < n) ? sc[2 * (j + u)] : -INFINITY;
#pragma unroll
            for (int u = 0; u < 4; ++u) { c1 += v[u] >= m1 ? 1 : 0; c2 += v[u] >= m2 ? 1 : 0; c3 += v[u] >= m3 ? 1 : 0; }
        }
        if (c3 >= T3_KEEP) lo = m3;
        else if (c2 >= T3_KEEP) { lo = m2; hi = m3; }
        else if (c1 >= T3_KEEP) { lo = m1; hi = m2; }
        else hi = m1;
    }
    int w = 0;
    for (int j = 0; j < nmax; j += 4) {
        uint64_t e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = mybuf[j + u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j + u < n && entry_score(e[u]) >= lo) mybuf[w++] = e[u];
    }
    if (act) {
        if (w > T3_CAP - T3_CHUNK) { cnt = 0; thr = INFINITY; }   // flat scores: hand the user to the exact kernel
        else { cnt = w; thr = lo; }
    }
}

__global__ void __launch_bounds__(T3_THREADS, 1) score_tc3_kernel(const __grid_constant__ Tc3Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t a_bytes = (uint32_t)(TC_BM / 8) * a.kcores * 128;        // one user tile
    const uint32_t b_bytes = (uint32_t)(T3_BN / 8) * a.kcores * 128;
    uint8_t *sA = smem_raw;                                                   // [2 user tiles]
    uint8_t *sB = sA + 2 * a_bytes;
    uint64_t *cand = reinterpret_cast<uint64_t *>(sB + (size_t)T3_STAGES * b_bytes);          // [256 rows][CAP + 1]
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(cand + (size_t)2 * TC_BM * (T3_CAP + 1));   // [4 acc][4 words][128 rows]
    uint32_t *stage = bitmap + T3_ACC * TC_BM * 4;                                              // [8 warps x 32 rows][T3_STAGE_W]
    Tc3Smem *sm = reinterpret_cast<Tc3Smem *>(stage + 8 * 32 * T3_STAGE_W);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA -> (user-tile pair, item split): head pairs scan every item tile in one CTA, the others are split n_splits ways
    const bool head = (int)blockIdx.x < a.n_head;
    const int rest = (int)blockIdx.x - a.n_head;
    const int pair = head ? (int)blockIdx.x : a.n_head + rest / a.n_splits;
    const int sp = head ? 0 : rest % a.n_splits;
    const int ns = head ? 1 : a.n_splits;
    // item tiles intersecting [item_lo, item_hi), divided evenly over the splits
    const int64_t hi_eff = min(a.item_hi, a.n_items);
    const int t_first = (int)(max((int64_t)0, a.item_lo) / T3_BN);
    const int t_last = (int)((hi_eff + T3_BN - 1) / T3_BN);                 // exclusive
    const int n_t = max(0, t_last - t_first);
    const int per = (n_t + ns - 1) / ns;
    const int t0 = t_first + sp * per, t1 = min(t_last, t0 + per);
    const int n_it = max(0, t1 - t0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < T3_STAGES; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
        mbar_init(&sm->a_full, 1);
        for (int i = 0; i < T3_ACC; ++i) { mbar_init(&sm->tmem_full[i], 1); mbar_init(&sm->tmem_empty[i], 4); mbar_init(&sm->mask_full[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < T3_ACC * TC_BM * 4; i += T3_THREADS) bitmap[i] = 0u;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        // ===== bulk-TMA producer
        if (lane == 0 && n_it > 0) {
            mbar_arrive_expect_tx(&sm->a_full, 2 * a_bytes);
            for (int h = 0; h < 2; ++h) {
                const int ut = min(2 * pair + h, a.n_utiles - 1);            // an odd tile count repeats the last tile (unused)
                bulk_g2s(sA + (size_t)h * a_bytes, a.a_img + (size_t)ut * a_bytes, a_bytes, &sm->a_full);
            }
            for (int it = 0; it < n_it; ++it) {
                const int s = it % T3_STAGES;
                const uint32_t ph = (uint32_t)(it / T3_STAGES) & 1u;
                mbar_wait_parked(&sm->empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&sm->full[s], b_bytes);
                bulk_g2s(sB + (size_t)s * b_bytes, a.b_img + (size_t)(t0 + it) * b_bytes, b_bytes, &sm->full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): per item tile one UMMA group per user tile
        if (lane == 0 && n_it > 0) {
            const uint32_t idesc = umma_idesc_f16_m128(T3_BN);
            const uint32_t sbo = (uint32_t)a.kcores * 128, lbo = 128;
            const int ksteps = a.kcores / 2;
            mbar_wait_parked(&sm->a_full, 0);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % T3_STAGES;
                mbar_wait_parked(&sm->full[s], (uint32_t)(it / T3_STAGES) & 1u);
                const uint32_t b0 = smem_u32(sB + (size_t)s * b_bytes);
                for (int h = 0; h < 2; ++h) {
                    const int acc = 2 * h + (it & 1);
                    mbar_wait_parked(&sm->tmem_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + (size_t)h * a_bytes);
                    for (int ks = 0; ks < ksteps; ++ks)
                        tc_mma_f16(tmem_base + (uint32_t)acc * T3_BN, umma_desc(a0 + ks * 256, lbo, sbo), umma_desc(b0 + ks * 256, lbo, sbo),
                                   idesc, ks > 0 ? 1u : 0u);
                    tc_commit(&sm->tmem_full[acc]);    // accumulator ready for user tile h's epilogue warps
                }
                tc_commit(&sm->empty[s]);              // smem stage reusable once both groups of MMAs retire
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ===== mask builders (warp 2 + h serves user tile h): seen-item bucket + banned bitmap + item range
        const int h = warp - 2;
        const int ut = 2 * pair + h;
        for (int it = 0; it < n_it; ++it) {
            const int acc = 2 * h + (it & 1), t = t0 + it;
            mbar_wait_parked(&sm->tmem_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            uint32_t *bm = bitmap + (size_t)acc * TC_BM * 4;
            uint32_t common = 0;
            if (lane < 4) {
                const int64_t c0 = (int64_t)t * T3_BN + lane * 32;
                if (c0 < a.item_lo) common |= (a.item_lo - c0 >= 32) ? 0xffffffffu : ((1u << (a.item_lo - c0)) - 1u);
                if (c0 + 32 > hi_eff) common |= (c0 >= hi_eff) ? 0xffffffffu : ~((1u << (hi_eff - c0)) - 1u);
                if (a.banned && c0 < a.n_items) common |= __ldg(a.banned + (c0 >> 5));
            }
            if (__any_sync(0xffffffffu, common != 0u)) {
                for (int w = 0; w < 4; ++w) {
                    const uint32_t cw = __shfl_sync(0xffffffffu, common, w);
                    if (cw)
                        for (int r = lane; r < TC_BM; r += 32) bm[w * TC_BM + r] |= cw;
                }
                __syncwarp();
            }
            if (a.mask_tile_ptr && ut < a.n_utiles) {
                // entries are bucketed by 256 items: this 128-item tile is one half of bucket t / 2
                const int32_t *p = a.mask_tile_ptr + (size_t)ut * (a.n_buckets + 1) + (t >> 1);
                const int e0 = __ldg(p), e1 = __ldg(p + 1);
                const uint32_t half = (uint32_t)(t & 1) << 7;
                for (int e = e0 + lane; e < e1; e += 32) {
                    const uint32_t ent = a.mask_entries[e];
                    const uint32_t col = ent & 255u;
                    if ((col & 128u) == half) atomicOr(bm + ((col & 127u) >> 5) * TC_BM + (ent >> 8), 1u << (col & 31u));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->mask_full[acc]);
        }
    } else {
        // ===== epilogue: thread = user row (TMEM lane) of user tile h
        const int h = (warp - 4) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint64_t *mybuf = cand + ((size_t)h * TC_BM + row) * (T3_CAP + 1);
        uint32_t *mystage = stage + ((size_t)(warp - 4) * 32 + lane) * T3_STAGE_W;
        float thr = -INFINITY;
        int cnt = 0;
        for (int it = 0; it < n_it; ++it) {
            const int acc = 2 * h + (it & 1), t = t0 + it;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&sm->tmem_full[acc], ph);
            mbar_wait(&sm->mask_full[acc], ph);
            tc_fence_after();
            uint32_t *bm = bitmap + (size_t)acc * TC_BM * 4 + row;          // word w of this row at bm[w * 128]
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * T3_BN;
            uint32_t va[T3_CHUNK], vb[T3_CHUNK];
            tc_ld16(taddr, va);
            uint32_t mword = 0;
            auto filter = [&](uint32_t (&v)[T3_CHUNK], int ch) {
                if (__any_sync(0xffffffffu, cnt > T3_CAP - T3_CHUNK)) compact_lanes3(mybuf, cnt, thr);
                if ((ch & 1) == 0) {                     // one bitmap word covers two 16-column steps
                    mword = bm[(ch >> 1) * TC_BM];
                    bm[(ch >> 1) * TC_BM] = 0u;
                }
                const uint32_t m = (mword >> ((ch & 1) * 16)) & 0xffffu;
                const uint32_t item0 = (uint32_t)(t * T3_BN + ch * T3_CHUNK);
#pragma unroll
                for (int c = 0; c < T3_CHUNK; c += 4)
                    *reinterpret_cast<uint4 *>(mystage + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
                uint32_t h0 = 0, h1 = 0, h2 = 0, h3 = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    hit_if_gt(v[c], thr, h0, 1u << c);
                    hit_if_gt(v[c + 4], thr, h1, 1u << (c + 4));
                    hit_if_gt(v[c + 8], thr, h2, 1u << (c + 8));
                    hit_if_gt(v[c + 12], thr, h3, 1u << (c + 12));
                }
                uint32_t hits = (h0 | h1 | h2 | h3) & ~m;        // seen / banned / out-of-range columns never pass
                while (hits) {                                   // rare; two per trip (both staged scores in flight)
                    const int c0 = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const bool two = hits != 0u;
                    const int c1 = two ? __ffs(hits) - 1 : c0;
                    hits &= hits - 1;
                    const uint32_t s0 = mystage[c0], s1 = mystage[c1];
                    mybuf[cnt] = ((uint64_t)(item0 + c0) << 32) | s0;
                    if (two) mybuf[cnt + 1] = ((uint64_t)(item0 + c1) << 32) | s1;
                    cnt += two ? 2 : 1;
                }
            };
#pragma unroll 1
            for (int ch = 0; ch < T3_BN / T3_CHUNK; ch += 2) {
                tc_wait_ld();
                tc_ld16(taddr + (ch + 1) * T3_CHUNK, vb);
                filter(va, ch);
                tc_wait_ld();
                if (ch + 2 < T3_BN / T3_CHUNK) tc_ld16(taddr + (ch + 2) * T3_CHUNK, va);
                filter(vb, ch + 1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->tmem_empty[acc]);
        }
        // write this quarter's lists
        __syncwarp();
        const int ut = 2 * pair + h;
        for (int r = 0; r < 32; ++r) {
            const int64_t b = (int64_t)ut * TC_BM + q * 32 + r;
            if (b >= a.n_eval) break;
            const int n = __shfl_sync(0xffffffffu, cnt, r);
            const float th = __shfl_sync(0xffffffffu, thr, r);
            const uint64_t *src = cand + ((size_t)h * TC_BM + q * 32 + r) * (T3_CAP + 1);
            const size_t list = (size_t)b * a.n_splits + sp;
            int32_t *dst = a.cand_items + list * T3_CAP;
            for (int e = lane; e < n; e += 32) dst[e] = (int32_t)(uint32_t)(src[e] >> 32);
            if (lane == 0) {
                a.cand_cnt[list] = n;
                a.cand_thr[list] = th;
            }
            if (head && lane > 0 && lane < a.n_splits) {          // the list slots an unsplit pair does not use
                a.cand_cnt[list + lane] = 0;
                a.cand_thr[list + lane] = -INFINITY;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

}  // namespace igcn
