// Full-ranking evaluation on the 5th-generation tensor cores (tcgen05 + TMEM + bulk-TMA).
//
// Replaces torch.mm (reference model.py:122) + the -inf index_put (trainer.py:149-161) +
// torch.topk (trainer.py:163) for whole-catalogue ranking.  The B x I score matrix lives only in
// tensor memory; what leaves the SM is <= 96 candidate item ids per user.
//
// Exactness without an fp32 MMA kind: operands are rounded to fp16 (kind::f16, fp32 accumulate in
// TMEM) and ONE EXTRA K-BLOCK carries an error bound: user row gets c*|u| (+eps), item row gets
// |i| (+eps), both rounded UP, so the tensor core itself produces s_hat = s_fp16 + c|u||i| >= s
// (Cauchy-Schwarz on the fp16 rounding errors, c = 1e-3 > 2^-10 + accumulation slack).  The
// epilogue keeps, per user, every item whose upper bound s_hat beats a running threshold that is
// provably <= the 32nd largest s_hat; candidates are re-scored in exact fp32 (same FMA order as
// eval_exact.cu) by tc_finalize_kernel, and a user is accepted only if its k-th exact score is
// strictly above every dropped item's upper bound -- otherwise the user goes to the exact kernel.
//
// Items are packed RELATIVE TO THE MEAN ITEM ROW m: <u, i - m> = <u, i> - <u, m> ranks a user's items exactly
// like <u, i> does, but the bound c|u||i - m| no longer pays for a component all items share (propagated
// embeddings have a large one: every node aggregates the same hubs).  The verification adds <u, m> back.
//
// Kernel structure (one CTA per 128-user tile x item-range split, 7 warps):
//   warp 0  bulk-TMA producer: cp.async.bulk of pre-arranged operand images (no-swizzle K-major
//           core-matrix layout written by tc_pack_kernel, so a tile is one contiguous copy)
//   warp 1  TMEM allocator + single-thread tcgen05.mma issuer (M128 x N256 x K16, 5 per tile)
//   warp 2  mask helper: turns the (user tile, item tile) bucket of seen items, the banned bitmap
//           and the item range into a 128 x 256 bitmap in shared memory (and clears what it set two
//           tiles earlier: the epilogue never writes the bitmap)
//   warp 3-6 epilogue: tcgen05.ld 32x32b (thread = user row).  The filter is built around the ALU pipe
//           (16 lanes/clk per scheduler: one warp instruction every 2 cycles): a compare + predicated OR
//           per score costs 4 cycles per column and warp, 1,024 cycles per 256-column tile against 640
//           cycles of MMA, whatever the number of epilogue warps.  So the common case does no compare
//           at all: the 32 scores of a chunk are max-reduced with 3-input FMNMX (0.5 instruction per
//           score), ONE vote asks whether any row of the warp beats its threshold, and only then the
//           8-column groups that contain a hit are compared, masked and appended.  Items are scanned in
//           a caller-supplied order (most popular first): thresholds tighten within the first tiles and
//           most later chunks take the compare-free path.
// Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty double buffer (MMA <-> epilogue),
// bitmap full (helper -> epilogue).

#include <stdlib.h>

#include "tc_common.cuh"

namespace igcn {

constexpr int TC_BN = 256;        // items per tile     (UMMA N)
constexpr int TC_CAP = 96;        // candidate slots per user row
constexpr int TC_KEEP = 32;       // the threshold never rises above the TC_KEEP-th best upper bound
constexpr int TC_STAGES = 2;      // item-tile smem stages
constexpr int TC_MASKS = 4;       // ring of 128 x 256 mask bitmaps: the helper warp runs up to 4 tiles ahead of the epilogue
constexpr int TC_THREADS = 7 * 32;
constexpr int TC_STAGE_W = 12;    // words per row of the 8-score staging area (16 B aligned; 128-bit stores of 8 lanes hit 8 bank groups)

// ------------------------------------------------------------------ operand packing
__global__ void maxabs_kernel(const float *__restrict__ x, int64_t n, uint32_t *out) {
    uint32_t m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(fabsf(x[i])));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);   // max is order independent: deterministic
}

// One group of 16 lanes converts one row (D <= 64) into its tile image: kcores core matrices of
// 8 fp16 (16 B) per row; the last K block holds the bound entry in its first element.
// Row r of the image is rep row row_ids[r] (users), row0 + perm[r] (items in scan order) or row0 + r.
__global__ void __launch_bounds__(256) tc_pack_kernel(const float *__restrict__ rep, const int64_t *__restrict__ row_ids,
                                                      const int32_t *__restrict__ perm, int64_t row0, int64_t n_rows, int D,
                                                      int tile_rows, int kcores, int is_user,
                                                      const uint32_t *__restrict__ maxabs_bits, const float *__restrict__ center_sum,
                                                      float inv_n, uint8_t *__restrict__ img) {
    const int lane = threadIdx.x & 15;
    const int64_t r = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
    if (r >= n_rows) return;
    const uint32_t gmask = 0xffffu << ((threadIdx.x & 31) & 16);
    const float scale = tc_scale(maxabs_bits);
    const int64_t src = row_ids ? row_ids[r] : row0 + (perm ? (int64_t)perm[r] : r);
    float4 v = f4zero();
    if (lane * 4 < D) {
        v = ld4(rep + src * D + lane * 4);
        if (!is_user) {          // items are packed relative to the mean item row (see the header comment)
            const float4 m = ld4(center_sum + lane * 4);
            v.x -= m.x * inv_n; v.y -= m.y * inv_n; v.z -= m.z * inv_n; v.w -= m.w * inv_n;
        }
    }
    v = scale4(v, scale);
    float sq = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sq += __shfl_xor_sync(gmask, sq, o, 16);
    const float norm = sqrtf(sq) * 1.000001f;
    const int64_t tile = r / tile_rows;
    const int rr = (int)(r % tile_rows);
    const size_t group_bytes = (size_t)kcores * 128;
    uint8_t *base = img + (size_t)tile * (tile_rows / 8) * group_bytes + (size_t)(rr >> 3) * group_bytes + (size_t)(rr & 7) * 16;
    const int dcores = (D + 15) / 16 * 2;      // K cores holding embedding dims (D padded to 16)
    if ((lane >> 1) < dcores) {
        __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&h0);
        pk.y = *reinterpret_cast<uint32_t *>(&h1);
        *reinterpret_cast<uint2 *>(base + (size_t)(lane >> 1) * 128 + (lane & 1) * 8) = pk;
    }
    if (lane == 0) {
        const float b = is_user ? (TC_C * norm + TC_EPS_U) : (norm + TC_EPS_I);
        uint4 z = make_uint4(0u, 0u, 0u, 0u);
        z.x = (uint32_t)__half_as_ushort(__float2half_ru(b));
        *reinterpret_cast<uint4 *>(base + (size_t)dcores * 128) = z;
        *reinterpret_cast<uint4 *>(base + (size_t)(dcores + 1) * 128) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ------------------------------------------------------------------ candidate kernel
struct TcArgs {
    const uint8_t *a_img;
    const uint8_t *b_img;
    int n_utiles, n_itiles, n_splits, n_head, kcores;      // the first n_head user tiles are not split
    int64_t n_eval, n_items, item_lo, item_hi;
    const uint32_t *banned;
    const int32_t *mask_tile_ptr;     // [n_utiles, n_itiles + 1] or NULL
    const uint16_t *mask_entries;     // (row << 8) | col
    int32_t *cand_items;              // [n_eval, n_splits, TC_CAP]
    int32_t *cand_cnt;                // [n_eval, n_splits]
    float *cand_thr;                  // [n_eval, n_splits]   (scaled units; -inf = nothing was dropped)
    float *dump;                      // optional [n_utiles*128, n_itiles*256] of s_hat (tests)
    unsigned long long *stats;        // optional [5] filter statistics (igcn_tc_candidates_stats)
    int dbg;                          // timing experiments (IGCN_TC_DEBUG): 1 producer/MMA threads spin instead of parking,
                                      // 2 mask builder spins, 4 mask builder builds nothing (invalid results)
};

__device__ __forceinline__ void mbar_wait_helper(uint64_t *bar, uint32_t parity, bool spin) {
    if (spin) mbar_wait(bar, parity);
    else mbar_wait_parked(bar, parity);
}

struct TcSmem {
    uint64_t full[TC_STAGES], empty[TC_STAGES], a_full, tmem_full[2], tmem_empty[2], mask_full[TC_MASKS], mask_empty[TC_MASKS];
    uint32_t tmem_base;
    uint32_t pad;
};

// Candidate buffer entry: low word = raw fp32 score bits, high word = item id (what the predicated
// epilogue store writes).

// Lane-parallel compaction: every lane shrinks ITS OWN row's buffer, no cross-lane traffic except
// the common loop bound.  A float bisection between the current threshold and the row maximum finds
// lo with count(score >= lo) >= TC_KEEP (8 halvings: within a few entries of TC_KEEP); entries below
// lo are dropped and lo becomes the row's threshold.  A row that cannot be shrunk (ties) gives up:
// thr = +inf makes the finalize kernel route that user to the exact kernel.
__device__ __forceinline__ void compact_lanes(uint64_t *mybuf, int &cnt, float &thr) {
    const bool act = cnt > TC_KEEP + 8;
    const int n = act ? cnt : 0;
    const int nmax = (__reduce_max_sync(0xffffffffu, n) + 3) & ~3;       // loops run in groups of 4 (TC_CAP % 4 == 0)
    const float *sc = reinterpret_cast<const float *>(mybuf);          // score of entry j at sc[2 * j]
    float vmax = -INFINITY, vmin = INFINITY;
    for (int j = 0; j < nmax; j += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = sc[2 * (j + u)];             // 4 independent loads in flight
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j + u < n) { vmax = fmaxf(vmax, v[u]); vmin = fminf(vmin, v[u]); }
    }
    float lo = (thr == -INFINITY) ? vmin : thr;      // invariant: count(score >= lo) >= TC_KEEP
    float hi = vmax;
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {          // 4-ary search: 4 rounds = 1/256 of the range
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
        if (c3 >= TC_KEEP) lo = m3;
        else if (c2 >= TC_KEEP) { lo = m2; hi = m3; }
        else if (c1 >= TC_KEEP) { lo = m1; hi = m2; }
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
        if (w > TC_CAP - 32) { cnt = 0; thr = INFINITY; }   // flat scores: hand the user to the exact kernel
        else { cnt = w; thr = lo; }
    }
}

// VARIANT: 0 production, 1 = also dump every s_hat (tests), 4 = production + filter statistics (TcArgs.stats:
// chunks seen / chunks that left the compare-free path / 8-column groups compared / candidates appended /
// compactions); 2, 3, 5 = timing experiments selected with the IGCN_TC_EXPERIMENT environment variable (results
// are NOT valid): 2 reads the accumulators but does not filter (TMA + MMA + TMEM-read floor), 3 filters against
// thr = +inf (the compare-free path on every chunk, no hits, no compaction), 5 does not read TMEM at all (MMA
// issue + the mbarrier hand-offs only).
// (CTA pairs -- thread-block clusters of two that fetch half of every item tile each and multicast it into both CTAs'
// shared memory -- were built and measured: bit-identical, 0.512 vs 0.507 ms on the Yelp shape, hand-off floor 0.366
// vs 0.351 ms.  Halving the L2 reads changes nothing because every SM still has to take in 40 KB of item image per
// 128 x 256 tile: the floor is the per-SM operand ingest (~42 B/cycle), not L2 bandwidth.  Removed again.)
template <int VARIANT>
__global__ void __launch_bounds__(TC_THREADS, 1) score_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t a_bytes = (uint32_t)(TC_BM / 8) * a.kcores * 128;
    const uint32_t b_bytes = (uint32_t)(TC_BN / 8) * a.kcores * 128;
    uint8_t *sA = smem_raw;
    uint8_t *sB = sA + a_bytes;
    uint64_t *cand = reinterpret_cast<uint64_t *>(sB + (size_t)TC_STAGES * b_bytes);
    uint32_t *bitmap = reinterpret_cast<uint32_t *>(cand + (size_t)TC_BM * (TC_CAP + 1));
    uint32_t *stage = bitmap + TC_MASKS * TC_BM * 8;                // [128 rows][TC_STAGE_W]: one 8-score group per row
    TcSmem *sm = reinterpret_cast<TcSmem *>(stage + TC_BM * TC_STAGE_W);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA -> (user tile, item split): head tiles scan every item tile in one CTA, the others are split n_splits ways
    const bool head = (int)blockIdx.x < a.n_head;
    const int rest = (int)blockIdx.x - a.n_head;
    const int ut = head ? (int)blockIdx.x : a.n_head + rest / a.n_splits;
    const int sp = head ? 0 : rest % a.n_splits;
    const int ns = head ? 1 : a.n_splits;
    // item tiles intersecting [item_lo, item_hi), divided evenly over the splits
    const int64_t hi_eff = min(a.item_hi, a.n_items);
    const int t_first = (int)(max((int64_t)0, a.item_lo) / TC_BN);
    const int t_last = (int)((hi_eff + TC_BN - 1) / TC_BN);                 // exclusive
    const int n_t = max(0, t_last - t_first);
    // split sp takes tiles t_first + sp, + ns, + 2 ns, ...: every split sees the head of the scan order (the
    // popular items) first, so every list's threshold tightens early
    const int t0 = t_first + sp;
    const int n_it = sp < n_t ? (n_t - sp + ns - 1) / ns : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&sm->full[s], 1); mbar_init(&sm->empty[s], 1); }
        mbar_init(&sm->a_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&sm->tmem_full[i], 1); mbar_init(&sm->tmem_empty[i], 4); }
        for (int i = 0; i < TC_MASKS; ++i) { mbar_init(&sm->mask_full[i], 1); mbar_init(&sm->mask_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
            mbar_arrive_expect_tx(&sm->a_full, a_bytes);
            bulk_g2s(sA, a.a_img + (size_t)ut * a_bytes, a_bytes, &sm->a_full);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait_helper(&sm->empty[s], ph ^ 1u, a.dbg & 1);
                mbar_arrive_expect_tx(&sm->full[s], b_bytes);
                bulk_g2s(sB + (size_t)s * b_bytes, a.b_img + (size_t)(t0 + it * ns) * b_bytes, b_bytes, &sm->full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread)
        if (lane == 0 && n_it > 0) {
            const uint32_t idesc = umma_idesc_f16_m128(TC_BN);
            const uint32_t sbo = (uint32_t)a.kcores * 128, lbo = 128;
            const int ksteps = a.kcores / 2;
            mbar_wait_helper(&sm->a_full, 0, a.dbg & 1);
            for (int it = 0; it < n_it; ++it) {
                const int s = it % TC_STAGES, acc = it & 1;
                mbar_wait_helper(&sm->tmem_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u, a.dbg & 1);
                mbar_wait_helper(&sm->full[s], (uint32_t)(it / TC_STAGES) & 1u, a.dbg & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + (size_t)s * b_bytes);
                for (int ks = 0; ks < ksteps; ++ks)
                    tc_mma_f16(tmem_base + (uint32_t)acc * TC_BN, umma_desc(a0 + ks * 256, lbo, sbo), umma_desc(b0 + ks * 256, lbo, sbo),
                               idesc, ks > 0 ? 1u : 0u);
                tc_commit(&sm->empty[s]);          // smem stage reusable once these MMAs retire
                tc_commit(&sm->tmem_full[acc]);    // accumulator ready for the epilogue
            }
        }
    } else if (warp == 2) {
        // ===== mask helper: seen-item bucket + banned bitmap + item range -> bitmap[ring slot][8 words][128 rows].
        // The epilogue only reads the bitmaps.  The helper owns a ring of TC_MASKS of them and is NOT part of the
        // accumulator hand-off chain: it zeroes a slot as soon as the epilogue has released it (mask_empty), ORs in
        // the tile's entries -- whose global loads were issued one tile earlier, so the dependent chain tile_ptr ->
        // entries (two L2 round trips, ~1,900 cycles per tile when it sat between tmem_empty and mask_full) is off
        // the critical path -- and publishes it (mask_full) up to TC_MASKS tiles ahead of the epilogue.
        const int32_t *tp = a.mask_tile_ptr ? a.mask_tile_ptr + (size_t)ut * (a.n_itiles + 1) : nullptr;
        // Two-deep software pipeline of the bucket loads, so that no iteration waits for a load it has just issued:
        // the bucket bounds tile_ptr[t], tile_ptr[t + 1] are fetched TWO tiles ahead, the first 32 entries ONE tile
        // ahead (their address needs the bounds); with the bounds fetched one tile ahead the warp stalled an L2 round
        // trip per tile on them.
        auto bounds_of = [&](int it, int &b0, int &b1) {
            b0 = b1 = 0;
            if (tp && it < n_it) { const int t = t0 + it * ns; b0 = __ldg(tp + t); b1 = __ldg(tp + t + 1); }
        };
        auto first_of = [&](int b0, int b1) -> uint32_t { return (b0 + lane < b1) ? (uint32_t)a.mask_entries[b0 + lane] : 0u; };
        int e0, e1, n0, n1;
        bounds_of(0, e0, e1);
        bounds_of(1, n0, n1);
        uint32_t first = first_of(e0, e1);
        for (int it = 0; it < n_it; ++it) {
            const int slot = it % TC_MASKS, t = t0 + it * ns;
            int q0, q1;
            bounds_of(it + 2, q0, q1);                               // in flight for two tiles
            const uint32_t nfirst = first_of(n0, n1);                // in flight for one tile
            mbar_wait_helper(&sm->mask_empty[slot], ((uint32_t)(it / TC_MASKS) & 1u) ^ 1u, a.dbg & 2);
            uint32_t *bm = bitmap + (size_t)slot * TC_BM * 8;
            if (!(a.dbg & 4)) {
                uint32_t common = 0;
                if (lane < 8) {
                    const int64_t c0 = (int64_t)t * TC_BN + lane * 32;
                    if (c0 < a.item_lo) common |= (a.item_lo - c0 >= 32) ? 0xffffffffu : ((1u << (a.item_lo - c0)) - 1u);
                    if (c0 + 32 > hi_eff) common |= (c0 >= hi_eff) ? 0xffffffffu : ~((1u << (hi_eff - c0)) - 1u);
                    if (a.banned && c0 < a.n_items) common |= __ldg(a.banned + (c0 >> 5));
                }
                // word w of every row starts as the tile-wide word (item range, banned items), usually zero
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    const uint32_t cw = __shfl_sync(0xffffffffu, common, w);
                    const uint4 v4 = make_uint4(cw, cw, cw, cw);
                    *reinterpret_cast<uint4 *>(bm + w * TC_BM + lane * 4) = v4;
                }
                __syncwarp();
                if (e0 + lane < e1) atomicOr(bm + ((first & 255u) >> 5) * TC_BM + (first >> 8), 1u << (first & 31u));
                for (int e = e0 + 32 + lane; e < e1; e += 32) {      // buckets beyond 32 entries (rare)
                    const uint32_t ent = a.mask_entries[e];
                    atomicOr(bm + ((ent & 255u) >> 5) * TC_BM + (ent >> 8), 1u << (ent & 31u));
                }
                __syncwarp();
                __threadfence_block();
            }
            if (lane == 0) mbar_arrive(&sm->mask_full[slot]);
            e0 = n0; e1 = n1; first = nfirst;
            n0 = q0; n1 = q1;
        }
    } else {
        // ===== epilogue: thread = user row (TMEM lane)
        const int q = warp & 3;                         // warps 3,4,5,6 -> TMEM lane quarters 3,0,1,2
        const int row = q * 32 + lane;
        uint64_t *mybuf = cand + (size_t)row * (TC_CAP + 1);
        uint32_t *mystage = stage + (size_t)row * TC_STAGE_W;
        float thr = VARIANT == 3 ? INFINITY : -INFINITY;
        int cnt = 0;
        uint32_t st_chunks = 0, st_slow = 0, st_groups = 0, st_hits = 0, st_compact = 0;     // VARIANT 4 only
        for (int it = 0; it < n_it; ++it) {
            const int acc = it & 1, t = t0 + it * ns, slot = it % TC_MASKS;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&sm->tmem_full[acc], ph);
            mbar_wait(&sm->mask_full[slot], (uint32_t)(it / TC_MASKS) & 1u);
            tc_fence_after();
            if (VARIANT == 5) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&sm->tmem_empty[acc]); mbar_arrive(&sm->mask_empty[slot]); }
                continue;
            }
            const uint32_t *bm = bitmap + (size_t)slot * TC_BM * 8 + row;   // word w of this row at bm[w * 128]
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * TC_BN;
            uint32_t va[64], vb[64];
            tc_ld64(taddr, va);
            // 64 columns per TMEM load, the next load always in flight behind the filter of the previous 64.
            // The compare-free test covers all 64 columns with ONE vote: the two 32-column maxima are independent
            // FMNMX trees the compiler interleaves, so the dependent chain maxima -> compare -> vote -> branch (a
            // single epilogue warp per scheduler cannot hide it) is paid once per 64 columns.
            auto group_max = [&](const uint32_t *v, float (&gm)[4]) -> float {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t *f = &v[g * 8];
                    gm[g] = fmaxf(fmax3(f[0], f[1], f[2]), fmax3(f[3], f[4], fmax3(f[5], f[6], f[7])));
                }
                return fmaxf(fmax3(gm[0], gm[1], gm[2]), gm[3]);
            };
            // the compare path of one 32-column chunk some row of which may hold a hit
            auto compare_chunk = [&](const uint32_t *v, const float (&gm)[4], int ch) {
                const uint32_t item0 = (uint32_t)(t * TC_BN + ch * 32);
                if (VARIANT == 4) ++st_slow;
                if (__any_sync(0xffffffffu, cnt > TC_CAP - 32)) {
                    compact_lanes(mybuf, cnt, thr);
                    if (VARIANT == 4) ++st_compact;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (!__any_sync(0xffffffffu, gm[g] > thr)) continue;
                    if (VARIANT == 4) ++st_groups;
                    uint32_t h = 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) hit_if_gt(v[g * 8 + c], thr, h, 1u << c);
                    if (h) {
                        // rare per row (about one row in twenty of a compared group): stage the 8 scores so that the
                        // hit columns can be addressed, drop seen / banned / out-of-range columns, append
                        *reinterpret_cast<uint4 *>(mystage) = make_uint4(v[g * 8], v[g * 8 + 1], v[g * 8 + 2], v[g * 8 + 3]);
                        *reinterpret_cast<uint4 *>(mystage + 4) = make_uint4(v[g * 8 + 4], v[g * 8 + 5], v[g * 8 + 6], v[g * 8 + 7]);
                        h &= ~(bm[ch * TC_BM] >> (g * 8));
                        while (h) {
                            const int c0 = __ffs(h) - 1;
                            h &= h - 1;
                            mybuf[cnt++] = ((uint64_t)(item0 + g * 8 + c0) << 32) | mystage[c0];
                            if (VARIANT == 4) ++st_hits;
                        }
                    }
                }
            };
            auto filter64 = [&](const uint32_t *v, int ch) {          // columns [ch * 32, ch * 32 + 64)
                if (VARIANT == 2) {
                    uint32_t x = 0;
#pragma unroll
                    for (int c = 0; c < 64; ++c) x ^= v[c];
                    if (x == 0x7fc12345u) cnt = 1;           // keeps the loads alive, never true in practice
                    return;
                }
                if (VARIANT == 1) {
                    float *d = a.dump + ((size_t)ut * TC_BM + row) * ((size_t)a.n_itiles * TC_BN) + (size_t)t * TC_BN + ch * 32;
#pragma unroll
                    for (int c = 0; c < 64; ++c) d[c] = __uint_as_float(v[c]);
                }
                float gm0[4], gm1[4];
                const float m0 = group_max(v, gm0), m1 = group_max(v + 32, gm1);
                if (VARIANT == 4) st_chunks += 2;
                if (!__any_sync(0xffffffffu, fmaxf(m0, m1) > thr)) return;
                // ---- some row of the warp has a score above its threshold in these 64 columns
                if (__any_sync(0xffffffffu, m0 > thr)) compare_chunk(v, gm0, ch);
                if (__any_sync(0xffffffffu, m1 > thr)) compare_chunk(v + 32, gm1, ch + 1);
            };
#pragma unroll 1
            for (int ch = 0; ch < 8; ch += 4) {
                tc_wait_ld();
                tc_ld64(taddr + (ch + 2) * 32, vb);
                filter64(va, ch);
                tc_wait_ld();
                if (ch + 4 < 8) tc_ld64(taddr + (ch + 4) * 32, va);
                filter64(vb, ch + 2);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&sm->tmem_empty[acc]); mbar_arrive(&sm->mask_empty[slot]); }
        }
        // one last compaction: the lists leave the SM at ~KEEP entries instead of up to 96, which is what
        // igcn_tc_finalize has to re-score exactly and rank (its time more than halves); the threshold becomes the
        // KEEP-th best upper bound, still a valid bound on everything that was dropped
        if (VARIANT != 2 && VARIANT != 5 && n_it > 0) {
            __syncwarp();
            compact_lanes(mybuf, cnt, thr);
        }
        // dump this row's candidates
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int64_t b = (int64_t)ut * TC_BM + q * 32 + r;
            if (b >= a.n_eval) break;
            const int n = __shfl_sync(0xffffffffu, cnt, r);
            const float th = __shfl_sync(0xffffffffu, thr, r);
            const uint64_t *src = cand + (size_t)(q * 32 + r) * (TC_CAP + 1);
            int32_t *dst = a.cand_items + ((size_t)b * a.n_splits + sp) * TC_CAP;
            for (int e = lane; e < n; e += 32) dst[e] = (int32_t)(uint32_t)(src[e] >> 32);
            if (lane == 0) {
                a.cand_cnt[b * a.n_splits + sp] = n;
                a.cand_thr[b * a.n_splits + sp] = th;
            }
            if (head && lane > 0 && lane < a.n_splits) {          // the list slots an unsplit tile does not use
                a.cand_cnt[b * a.n_splits + lane] = 0;
                a.cand_thr[b * a.n_splits + lane] = -INFINITY;
            }
        }
        if (VARIANT == 4 && a.stats) {
            // chunk / group counters are per warp (lane 0 speaks), appended candidates per row (summed)
            const uint32_t hits = __reduce_add_sync(0xffffffffu, st_hits);
            if (lane == 0) {
                atomicAdd(a.stats + 0, (unsigned long long)st_chunks);
                atomicAdd(a.stats + 1, (unsigned long long)st_slow);
                atomicAdd(a.stats + 2, (unsigned long long)st_groups);
                atomicAdd(a.stats + 3, (unsigned long long)hits);
                atomicAdd(a.stats + 4, (unsigned long long)st_compact);
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

// ------------------------------------------------------------------ exact re-scoring + verification
// One warp per user: exact fp32 scores of all candidates (FMA chain, ascending d -- identical to
// eval_exact.cu), rank by (score desc, item asc), emit top-k, and verify that the k-th exact score
// is strictly above every dropped item's upper bound.  Unverified users are appended to a list.
//
// The candidates' item rows are staged through shared memory, 32 at a time: the warp copies them with coalesced 16-byte
// cp.async (D/4 lanes per row, 4 wavefronts per instruction), then every lane runs the chain over ITS candidate's
// row out of shared memory (row stride D + 4 floats: the 128-bit reads of a quarter warp hit 32 different banks).
// Every lane gathering its own 256-byte row straight from global memory cost 32 L1 wavefronts per load instruction
// (ncu round 2, Yelp shape: l1tex 85 % busy, 0.19 ms -- 40 % of the candidate kernel's time).  One bulk copy
// (cp.async.bulk) per lane and row was tried too: the compiler serialises it over the lanes (9 instructions each).
constexpr int FIN_ROWS = 32;

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(256) tc_finalize_kernel(const float *__restrict__ rep, const int64_t *__restrict__ user_ids,
                                                          int64_t n_eval, int64_t item_row0, int D, int n_splits,
                                                          const int32_t *__restrict__ cand_items, const int32_t *__restrict__ cand_cnt,
                                                          const float *__restrict__ cand_thr, const uint32_t *__restrict__ maxabs_bits,
                                                          const float *__restrict__ center_sum, float inv_n,
                                                          const int32_t *__restrict__ item_perm,
                                                          int k, int32_t *out_items, float *out_scores, int32_t *fb_count,
                                                          int64_t *fb_users, int32_t *fb_rows) {
    extern __shared__ __align__(16) uint64_t fin_keys[];  // [8 warps][n_splits * TC_CAP] keys | [8 warps][FIN_ROWS][D + 4] floats
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * 8 + wid;
    if (b >= n_eval) return;
    const int cap = n_splits * TC_CAP;
    const int DP = D + 4;
    uint64_t *keys = fin_keys + (size_t)wid * cap;
    float *rows = reinterpret_cast<float *>(fin_keys + (size_t)8 * cap) + (size_t)wid * FIN_ROWS * DP;
    const int64_t u = user_ids[b];
    const float *urow = rep + u * D;
    const int lpr = D >> 2;                              // lanes per row: one 16-byte piece each
    const int rows_per = 32 / lpr;                       // rows one copy instruction covers
    const int sub = lane / lpr, piece = lane - sub * lpr;
    const float *ibase = rep + item_row0 * D + piece * 4;                       // + item * D: this lane's piece of an item row
    const uint32_t sbase = smem_u32(rows + sub * DP + piece * 4);
    const uint32_t sstep = (uint32_t)(rows_per * DP * 4);
    const int r_end = sub < rows_per ? FIN_ROWS : 0;     // lanes beyond the last whole row of an instruction copy nothing
    float thr_max = -INFINITY;
    int n = 0;
    for (int sp = 0; sp < n_splits; ++sp) {
        const int c = cand_cnt[b * n_splits + sp];
        thr_max = fmaxf(thr_max, cand_thr[b * n_splits + sp]);
        const int32_t *src = cand_items + ((size_t)b * n_splits + sp) * TC_CAP;
        for (int e0 = 0; e0 < c; e0 += FIN_ROWS) {
            const int e = e0 + lane;
            int32_t item = 0;                            // lanes without a candidate name row 0: copied, never read
            if (e < c) item = item_perm ? __ldg(item_perm + src[e]) : src[e];        // scan position -> item id
            const int n_here = min(FIN_ROWS, c - e0);
            uint32_t sdst = sbase;
#pragma unroll 4
            for (int r0 = 0; r0 < n_here; r0 += rows_per, sdst += sstep) {
                const uint32_t it = (uint32_t)__shfl_sync(0xffffffffu, item, r0 + sub);
                if (r0 + sub < r_end) cp_async16(sdst, ibase + (size_t)(it * (uint32_t)D));
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (e < c) {
                const float *irow = rows + lane * DP;
                float s = 0.f;
                for (int d = 0; d < D; d += 4) {
                    const float4 x = ld4(urow + d), y = *reinterpret_cast<const float4 *>(irow + d);
                    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
                }
                keys[n + e] = ((uint64_t)f_order(s) << 32) | (uint32_t)(0x7fffffff - item);
            }
            __syncwarp();                                // the rows are overwritten by the next pass
        }
        n += c;
    }
    __syncwarp();
    float kth = -INFINITY;
    for (int e = lane; e < n; e += 32) {
        const uint64_t mine = keys[e];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += keys[j] > mine;
        if (rank < k) {
            out_items[b * k + rank] = 0x7fffffff - (int32_t)(mine & 0xffffffffu);
            out_scores[b * k + rank] = order_f((uint32_t)(mine >> 32));
        }
        if (rank == k - 1) kth = order_f((uint32_t)(mine >> 32));
    }
    for (int q = n + lane; q < k; q += 32) { out_items[b * k + q] = -1; out_scores[b * k + q] = -INFINITY; }
    kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, 16));
    kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, 8));
    kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, 4));
    kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, 2));
    kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, 1));
    // <u, m> and sum |u_d m_d| (m = mean item row): the tensor core scored u . (i - m)
    float um = 0.f, uam = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float p = urow[d] * (center_sum[d] * inv_n);
        um += p; uam += fabsf(p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { um += __shfl_xor_sync(0xffffffffu, um, o); uam += __shfl_xor_sync(0xffffffffu, uam, o); }
    if (lane == 0) {
        const float scale = tc_scale(maxabs_bits);
        // a dropped item j satisfies scale^2 * <u, i_j - m> <= s_hat_j <= thr_max, i.e. s_j <= thr_max / scale^2 + <u, m>;
        // scale^2 is a power of two (exact).  err covers the fp32 rounding of um, of the mean and of kth itself.
        const float err = 1e-5f * (uam + fabsf(kth));
        const bool ok = (thr_max == -INFINITY) || (n >= k && (kth - um - err) * scale * scale > thr_max);
        if (!ok) {
            const int slot = atomicAdd(fb_count, 1);
            fb_users[slot] = u;
            fb_rows[slot] = (int32_t)b;
        }
    }
}

}  // namespace igcn

using namespace igcn;

static int tc_kcores(int D) { return ((D + 15) / 16) * 2 + 2; }

extern "C" int igcn_tc_workspace(int64_t n_eval, int64_t n_items, int32_t D, int32_t n_splits, int64_t *a_img_bytes,
                                 int64_t *b_img_bytes, int64_t *cand_slots) {
    IGCN_CHECK_ARG(D > 0 && D <= 64 && !(D & 3), "tensor-core scoring supports D % 4 == 0, D <= 64");
    IGCN_CHECK_ARG(n_splits >= 1 && n_splits <= 8, "n_splits must be in [1, 8]");
    const int64_t kc = tc_kcores(D);
    *a_img_bytes = (n_eval + TC_BM - 1) / TC_BM * (TC_BM / 8) * kc * 128;
    *b_img_bytes = (n_items + TC_BN - 1) / TC_BN * (TC_BN / 8) * kc * 128;
    *cand_slots = n_eval * n_splits * TC_CAP;
    return 0;
}

extern "C" int igcn_tc_pack(const float *rep, int64_t n_rep_elems, const int64_t *user_ids, int64_t n_eval, int64_t item_row0,
                            int64_t n_items, int32_t D, const int32_t *item_perm, uint32_t *maxabs_bits, uint8_t *a_img,
                            uint8_t *b_img, float *center_sum, float *center_scratch, void *stream) {
    IGCN_CHECK_ARG(rep && user_ids && maxabs_bits && a_img && b_img && center_sum && center_scratch, "null pointer");
    IGCN_CHECK_ARG(D > 0 && D <= 64 && !(D & 3), "tensor-core scoring supports D % 4 == 0, D <= 64");
    cudaStream_t st = as_stream(stream);
    const int kc = tc_kcores(D);
    cudaMemsetAsync(maxabs_bits, 0, sizeof(uint32_t), st);
    maxabs_kernel<<<148 * 4, 256, 0, st>>>(rep, n_rep_elems, maxabs_bits);
    // column sums of the item rows in a fixed order (deterministic); the mean is sum * (1 / n_items)
    if (int rc = igcn_colsum_masked(rep, item_row0, item_row0 + n_items, D, nullptr, center_scratch, center_sum, stream)) return rc;
    const float inv_n = n_items > 0 ? 1.f / (float)n_items : 0.f;
    if (n_eval > 0)
        tc_pack_kernel<<<(unsigned)((n_eval + 15) / 16), 256, 0, st>>>(rep, user_ids, nullptr, 0, n_eval, D, TC_BM, kc, 1, maxabs_bits,
                                                                        center_sum, inv_n, a_img);
    if (n_items > 0)
        tc_pack_kernel<<<(unsigned)((n_items + 15) / 16), 256, 0, st>>>(rep, nullptr, item_perm, item_row0, n_items, D, TC_BN, kc, 0,
                                                                         maxabs_bits, center_sum, inv_n, b_img);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_tc_candidates(const uint8_t *a_img, const uint8_t *b_img, int64_t n_eval, int64_t n_items, int32_t D,
                                  int32_t n_splits, int32_t n_head, int64_t item_lo, int64_t item_hi, const uint32_t *banned_bits,
                                  const int32_t *mask_tile_ptr, const uint16_t *mask_entries, int32_t *cand_items,
                                  int32_t *cand_cnt, float *cand_thr, float *dump, unsigned long long *stats, void *stream) {
    IGCN_CHECK_ARG(a_img && b_img && cand_items && cand_cnt && cand_thr, "null pointer");
    IGCN_CHECK_ARG(D > 0 && D <= 64 && !(D & 3), "tensor-core scoring supports D % 4 == 0, D <= 64");
    IGCN_CHECK_ARG(n_splits >= 1 && n_splits <= 8, "n_splits must be in [1, 8]");
    IGCN_CHECK_ARG(!mask_tile_ptr || mask_entries, "mask_tile_ptr without mask_entries");
    if (n_eval <= 0) return 0;
    const int n_utiles = (int)((n_eval + TC_BM - 1) / TC_BM);
    IGCN_CHECK_ARG(n_head >= 0 && n_head <= n_utiles, "n_head must be in [0, number of user tiles]");
    const unsigned n_ctas = (unsigned)(n_head + (n_utiles - n_head) * n_splits);
    TcArgs a{};
    a.a_img = a_img; a.b_img = b_img;
    a.n_utiles = n_utiles;
    a.n_itiles = (int)((n_items + TC_BN - 1) / TC_BN);
    a.n_splits = n_splits; a.n_head = n_head; a.kcores = tc_kcores(D);
    a.n_eval = n_eval; a.n_items = n_items; a.item_lo = item_lo; a.item_hi = item_hi;
    a.banned = banned_bits; a.mask_tile_ptr = mask_tile_ptr; a.mask_entries = mask_entries;
    a.cand_items = cand_items; a.cand_cnt = cand_cnt; a.cand_thr = cand_thr; a.dump = dump; a.stats = stats;
    const size_t smem = (size_t)(TC_BM / 8 + TC_STAGES * (TC_BN / 8)) * a.kcores * 128 + (size_t)TC_BM * (TC_CAP + 1) * 8 +
                        (size_t)TC_MASKS * TC_BM * 8 * 4 + (size_t)TC_BM * TC_STAGE_W * 4 + sizeof(TcSmem) + 64;
    const char *dbg_env = getenv("IGCN_TC_DEBUG");
    a.dbg = dbg_env ? atoi(dbg_env) : 0;
    const char *exp_env = getenv("IGCN_TC_EXPERIMENT");      // read per call: tools/tc_floor.py sweeps it in one process
    const int experiment = exp_env ? atoi(exp_env) : 0;
    const int variant = dump ? 1 : stats ? 4 : (experiment == 2 || experiment == 3 || experiment == 5) ? experiment : 0;
    void (*kern)(TcArgs) = nullptr;
#define IGCN_TC_PICK(V) case V: kern = score_tc_kernel<V>; break
    switch (variant) { IGCN_TC_PICK(0); IGCN_TC_PICK(1); IGCN_TC_PICK(2); IGCN_TC_PICK(3); IGCN_TC_PICK(4); IGCN_TC_PICK(5); }
#undef IGCN_TC_PICK
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("igcn_tc_candidates: %s", cudaGetErrorString(e)); return (int)e; }
    kern<<<n_ctas, TC_THREADS, smem, as_stream(stream)>>>(a);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_tc_finalize(const float *rep, const int64_t *user_ids, int64_t n_eval, int64_t item_row0, int32_t D,
                                int32_t n_splits, const int32_t *cand_items, const int32_t *cand_cnt, const float *cand_thr,
                                const uint32_t *maxabs_bits, const float *center_sum, int64_t n_items, const int32_t *item_perm,
                                int32_t k, int32_t *out_items, float *out_scores, int32_t *fb_count, int64_t *fb_users,
                                int32_t *fb_rows, void *stream) {
    IGCN_CHECK_ARG(rep && user_ids && cand_items && cand_cnt && cand_thr && maxabs_bits && center_sum && out_items && out_scores,
                   "null pointer");
    IGCN_CHECK_ARG(fb_count && fb_users && fb_rows, "null fallback buffers");
    IGCN_CHECK_ARG(k > 0 && k <= TC_KEEP - 8, "tensor-core path supports k <= 24");
    IGCN_CHECK_ARG(D > 0 && D <= 64 && !(D & 3) && n_splits >= 1 && n_splits <= 8, "tensor-core scoring supports D % 4 == 0, D <= 64, 1..8 splits");
    if (n_eval <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(fb_count, 0, sizeof(int32_t), st);
    const size_t smem = (size_t)8 * n_splits * TC_CAP * sizeof(uint64_t) + (size_t)8 * FIN_ROWS * (D + 4) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(tc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("igcn_tc_finalize: %s", cudaGetErrorString(e)); return (int)e; }
    const float inv_n = n_items > 0 ? 1.f / (float)n_items : 0.f;
    tc_finalize_kernel<<<(unsigned)((n_eval + 7) / 8), 256, smem, st>>>(rep, user_ids, n_eval, item_row0, D, n_splits, cand_items,
                                                                         cand_cnt, cand_thr, maxabs_bits, center_sum, inv_n, item_perm, k,
                                                                         out_items, out_scores, fb_count, fb_users, fb_rows);
    IGCN_CHECK_LAUNCH();
    return 0;
}
