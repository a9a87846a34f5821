// Embedding propagation on B200: CSR row-split gather kernels.
//
//   igcn_spmm       Y = alpha * rowscale .* (A X + sum_j add_j)     K1/K4/K5 of SURVEY.md 2.2
//   igcn_inmo_fwd   X0 = s .* (F~ E)    template aggregation + dropout fused    K2/K3
//   igcn_inmo_bwd   dE = F~^T G         transposed, mask regenerated            K4
//   igcn_colsum_masked   gradient of the two global template rows
//
// Work decomposition (all three share it): a group of 8 lanes walks a run of non-zeros in CSR order;
// for D = 64 every lane carries two float4 accumulators, so one neighbour is two 128-bit loads per
// lane and each load instruction of a group reads one full 128-byte line.  Column ids and values
// are loaded 8 at a time (coalesced, one batch ahead) and broadcast with shuffles; 4 neighbours
// (8 loads per lane) are in flight per group.  Rows are visited in degree-descending order
// (igcn_csr.row_order).  A short row (<= IGCN_MEDIUM_NNZ non-zeros) is one group's work, a medium row
// is shared by the 4 groups of a warp (quarter each, ordered shuffle combine), a long row (>
// long_threshold) is pre-cut into chunks on the host (igcn_csr.chunk_*): chunk warps come first in
// the grid, write partial sums, and the last one to arrive (self-resetting counter) adds the
// partials in chunk order.  The dependent-load chain of any unit is thus <= 8 batches, and how a row
// is summed depends on its length only -- deterministic, independent of how rows are sharded over GPUs.
//
// Narrow tables (D = 16 or 8: the column-sharded training step on 4 / 8 GPUs) use 4 or 2 lanes per row instead of 8, so
// that a warp instruction still moves useful bytes in every lane.  How a row is summed must not depend on that: the
// pieces of a shared row are always cut for a TEAM of 4 subgroups on 8-non-zero boundaries (what the 8-lane kernel does),
// a warp simply holds 2 or 4 teams; every output element then sees the same additions in the same order as in the
// D = 64 kernel -- bit-identical column slices.
//
// Everything here is HBM/L2-bound gather work; there is no tensor-core shape to it.
#include <stdlib.h>

#include "common.cuh"

namespace igcn {

enum { MODE_SPMM = 0, MODE_INMO_FWD = 1, MODE_INMO_BWD = 2 };

struct PropArgs {
    igcn_csr g;
    const float *X;   // gathered table (X, E or G)
    float *Y;         // output rows
    const float *add[IGCN_MAX_ADD];
    int n_add;
    const float *rowscale;
    float alpha;
    int D;
    float *peer[IGCN_MAX_PEERS];   // row-sharded multi-GPU: the same output buffer on every rank (self included)
    int n_peers;                   // 0 = single GPU (write a.Y only)
    // SPMM variants (template parameter DROP doubles as the variant: 1 = row list, 2 = column filter)
    const int64_t *row_list;       // ascending GLOBAL row ids to compute
    const int32_t *n_list;         // device-side length of row_list
    const uint32_t *col_bits;      // bitmap over columns: non-zeros with a clear bit are skipped
    int64_t max_list;
    // INMO only
    const int32_t *tmpl;
    igcn_dropout drop;
    uint32_t thresh;
    float inv_keep;
    int64_t row0, n_users, glob_user, glob_item;
};

constexpr int kThreads = 256;
constexpr uint32_t kSelfCol = 0xffffffffu;

template <int LPR>
__device__ __forceinline__ uint32_t group_mask() {
    if (LPR == 32) return 0xffffffffu;
    const uint32_t base = (LPR == 16) ? 0xffffu : (LPR == 8) ? 0xffu : (LPR == 4) ? 0xfu : 0x3u;
    return base << ((threadIdx.x & 31) & ~(LPR - 1));
}

// Subgroups that share one medium row / one chunk, and the alignment of their pieces: 4 subgroups on 8-non-zero
// boundaries for every width up to 64 floats (LPR <= 8), the whole warp for the wide kernels.
template <int LPR>
struct Team {
    static constexpr int SUB = 32 / LPR;                    // subgroups per warp
    static constexpr int SIZE = SUB < 4 ? SUB : 4;          // subgroups per team
    static constexpr int PER_WARP = SUB / SIZE;             // teams per warp
    static constexpr int ALIGN = LPR < 8 ? 8 : LPR;         // piece boundaries
};

// Thread mapping: LPR lanes own one row, each lane carries V float4 accumulators; vector v of lane
// l covers floats [v*LPR*4 + l*4, +4), so every load instruction of a group reads one contiguous
// LPR*16-byte piece (a full 128-byte line for LPR = 8).  EXACT: D == LPR*V*4 (no per-vector guard).
template <int LPR, int V, bool EXACT>
struct RowVec {
    float4 v[V];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = f4zero();
    }
    static __device__ __forceinline__ bool on(int i, int lane, int D) { return EXACT || (i * LPR + lane) * 4 < D; }
};

// Sum over non-zeros [beg, end) of one row of w_e * T[col_e] (this lane's slices).
// Column ids / values are fetched LPR at a time, one batch AHEAD of the batch being gathered, so a row
// costs one dependent memory round trip per batch instead of two.
template <int LPR, int V, bool EXACT, int MODE, int DROP>
__device__ __forceinline__ void gather_range(const PropArgs &a, RowVec<LPR, V, EXACT> &acc, int64_t beg, int64_t end,
                                             int64_t grow, int lane, uint32_t gmask, uint64_t seed) {
    // 32-bit offsets relative to the first non-zero of the range keep the loop state in few registers
    const int32_t *__restrict__ colp = a.g.col + beg;
    const float *__restrict__ valp = (MODE == MODE_SPMM && a.g.val) ? a.g.val + beg : nullptr;
    const int len = (int)(end - beg);
    const int D = a.D;
    const float *__restrict__ Tl = a.X + lane * 4;
    constexpr bool COLS = (MODE == MODE_SPMM && DROP == 2);
    constexpr int Q0 = 8 / V;                           // neighbour rows in flight per group (8 x 16 B per lane) ...
    constexpr int Q = Q0 < LPR ? Q0 : LPR;              // ... but never more than one batch of column ids
    const int gshift = (threadIdx.x & 31) & ~(LPR - 1);

    int c_next = 0;
    float v_next = 1.f;
    if (lane < len) {
        c_next = __ldg(colp + lane);
        if (valp) v_next = __ldg(valp + lane);
    }
    for (int o = 0; o < len; o += LPR) {
        const int n = min(LPR, len - o);
        int c = c_next;
        float w = v_next;
        {   // prefetch the next batch
            const int e = o + LPR + lane;
            if (e < len) {
                c_next = __ldg(colp + e);
                if (valp) v_next = __ldg(valp + e);
            }
        }
        if (lane < n) {
            if (MODE == MODE_SPMM) {
                if (COLS && !((__ldg(a.col_bits + (c >> 5)) >> (c & 31)) & 1u)) w = 0.f;   // X[c] is an exact zero row
            } else {
                bool keep = true;
                if (DROP == 1) {
                    const uint32_t h = (MODE == MODE_INMO_FWD) ? edge_hash(seed, (uint32_t)grow, (uint32_t)c)
                                                               : edge_hash(seed, (uint32_t)c, (uint32_t)grow);
                    keep = h >= a.thresh;
                } else if (DROP == 2) {
                    const int64_t e = beg + o + lane;
                    const int64_t b = (MODE == MODE_INMO_FWD) ? e : __ldg(a.drop.tperm + e);
                    keep = (__ldg(a.drop.edge_keep + (b >> 5)) >> (b & 31)) & 1u;
                }
                if (MODE == MODE_INMO_FWD && a.tmpl) {
                    c = __ldg(a.tmpl + c);
                    keep = keep && (c >= 0);
                }
                w = keep ? 1.f : 0.f;
                if (!keep) c = 0;
            }
        } else {
            w = 0.f;
            c = 0;
        }
        if (COLS) {
            // few columns survive the filter (about a quarter, often none): visit only those, still in ascending
            // order, QC at a time -- padding every trip to 4 slots cost more instructions than it hid latency
            constexpr int QC = 2;
            uint32_t km = (__ballot_sync(gmask, w != 0.f) >> gshift) & ((1u << LPR) - 1u);
            while (km) {
                float4 x[QC][V];
                float ww[QC];
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    const bool ok = km != 0u;
                    const int j = ok ? __ffs(km) - 1 : 0;
                    km &= km - 1u;
                    const int cj = __shfl_sync(gmask, c, j, LPR);
                    ww[q] = ok ? __shfl_sync(gmask, w, j, LPR) : 0.f;
                    const float *p = Tl + (int64_t)cj * D;
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        x[q][i] = (ok && RowVec<LPR, V, EXACT>::on(i, lane, D)) ? ld4(p + i * LPR * 4) : f4zero();
                }
#pragma unroll
                for (int q = 0; q < QC; ++q)
#pragma unroll
                    for (int i = 0; i < V; ++i) fma4(acc.v[i], ww[q], x[q][i]);
            }
            continue;
        }
        for (int j0 = 0; j0 < n; j0 += Q) {
            float4 x[Q][V];
            float ww[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int cj = __shfl_sync(gmask, c, j0 + q, LPR);
                ww[q] = __shfl_sync(gmask, w, j0 + q, LPR);
                const bool ok = (j0 + q < n) && (MODE == MODE_SPMM || ww[q] != 0.f);
                const float *p = Tl + (int64_t)cj * D;
#pragma unroll
                for (int i = 0; i < V; ++i)
                    x[q][i] = (ok && RowVec<LPR, V, EXACT>::on(i, lane, D)) ? ld4(p + i * LPR * 4) : f4zero();
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
                for (int i = 0; i < V; ++i) fma4(acc.v[i], ww[q], x[q][i]);
        }
    }
}

template <int LPR, int V, bool EXACT, int MODE, int DROP>
__device__ __forceinline__ void finish_row(const PropArgs &a, int64_t r, RowVec<LPR, V, EXACT> &acc, int lane, uint64_t seed) {
    const int D = a.D;
    const int64_t grow = a.row0 + r;
    float s = 1.f;
    int64_t out_row = r;
    bool self = false;
    if (MODE == MODE_SPMM) {
        s = a.alpha;
        if (a.rowscale) s *= __ldg(a.rowscale + r);
    } else if (MODE == MODE_INMO_FWD) {
        self = true;
        if (DROP == 1) self = edge_hash(seed, (uint32_t)grow, kSelfCol) >= a.thresh;
        if (DROP == 2) self = (__ldg(a.drop.self_keep + (grow >> 5)) >> (grow & 31)) & 1u;
        s = __ldg(a.rowscale + r) * a.inv_keep;
    } else {
        out_row = a.tmpl ? (int64_t)__ldg(a.tmpl + grow) : grow;
        if (out_row < 0) return;
    }
    const int64_t gt = grow < a.n_users ? a.glob_user : a.glob_item;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        if (!RowVec<LPR, V, EXACT>::on(i, lane, D)) continue;
        const int off = (i * LPR + lane) * 4;
        float4 t = acc.v[i];
        if (MODE == MODE_SPMM)
            for (int j = 0; j < a.n_add; ++j) add4(t, ld4(a.add[j] + r * D + off));
        if (MODE == MODE_INMO_FWD && self) add4(t, ld4(a.X + gt * D + off));
        const float4 o = MODE == MODE_INMO_BWD ? t : scale4(t, s);
        if (a.n_peers == 0) {
            st4(a.Y + out_row * D + off, o);
        } else {
            // fused all-gather: the row goes straight into every rank's copy over NVLink (peer stores)
            for (int p = 0; p < a.n_peers; ++p) st4(a.peer[p] + out_row * D + off, o);
        }
    }
}

// Ordered in-team combine: subgroup 0 of the team ends up with ((p0 + p1) + p2) + ... of its subgroups' accumulators.
template <int LPR, int V, bool EXACT>
__device__ __forceinline__ void combine_subgroups(RowVec<LPR, V, EXACT> &acc, int lane, int tsub, int team_base) {
    constexpr int SIZE = Team<LPR>::SIZE;
    // only the team's own lanes take part: the other teams of the warp may be on a different path or gone
    const uint32_t tmask = (SIZE * LPR == 32) ? 0xffffffffu : (((1u << (SIZE * LPR)) - 1u) << team_base);
#pragma unroll
    for (int s = 1; s < SIZE; ++s) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float4 o;
            const int src = team_base + lane + s * LPR;
            o.x = __shfl_sync(tmask, acc.v[i].x, src);
            o.y = __shfl_sync(tmask, acc.v[i].y, src);
            o.z = __shfl_sync(tmask, acc.v[i].z, src);
            o.w = __shfl_sync(tmask, acc.v[i].w, src);
            if (tsub == 0) add4(acc.v[i], o);
        }
    }
}

// This subgroup's slice of the non-zero range [beg, end) when a team shares it (ALIGN-aligned pieces).
template <int LPR>
__device__ __forceinline__ void split_range(int64_t &beg, int64_t &end, int tsub) {
    constexpr int SIZE = Team<LPR>::SIZE, ALIGN = Team<LPR>::ALIGN;
    const int64_t q = (((end - beg) + SIZE - 1) / SIZE + ALIGN - 1) & ~(int64_t)(ALIGN - 1);
    const int64_t b = beg + tsub * q;
    end = min(end, b + q);
    beg = min(b, end);
}

// Work units, in this order (rows are visited by non-zero count, descending -- igcn_csr.row_order):
//   [0, n_chunks)                 one chunk of a long row (> long_threshold non-zeros) per TEAM (4 subgroups = one warp
//                                 at 8 lanes per row): the subgroups share the chunk, the combined partial goes to
//                                 g.partial and the last chunk of the row to finish adds the partials in chunk order
//   [n_chunks, +n_medium_rows)    one medium row (> IGCN_MEDIUM_NNZ non-zeros) per team, subgroups share it
//   the rest                      one short row per LPR-lane subgroup
// How a row is summed depends on its length only, never on the grid, the row block or the GPU count.
template <int LPR, int V, bool EXACT, int MODE, int DROP>
__global__ void __launch_bounds__(kThreads, 4) prop_kernel(const __grid_constant__ PropArgs a) {
    constexpr int SUB = 32 / LPR;
    constexpr int TSIZE = Team<LPR>::SIZE, TPW = Team<LPR>::PER_WARP;
    constexpr bool ROWS = (MODE == MODE_SPMM && DROP == 1);
    using Vec = RowVec<LPR, V, EXACT>;
    const int lane = threadIdx.x % LPR;
    const int sub = (threadIdx.x & 31) / LPR;
    const int team = sub / TSIZE, tsub = sub % TSIZE;          // team of subgroups inside the warp, subgroup inside the team
    const int team_base = team * TSIZE * LPR;
    const uint32_t gmask = group_mask<LPR>();
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t n_chunks = a.g.n_chunks;
    const int64_t n_long = a.g.row_order ? a.g.n_long_rows : 0, n_med = a.g.row_order ? a.g.n_medium_rows : 0;
    uint64_t seed = a.drop.seed;
    if (MODE != MODE_SPMM && DROP == 1 && a.drop.seed_dev) seed = mix64(seed ^ mix64(*a.drop.seed_dev + 0x2545f491ULL));
    const int D = a.D;
    Vec acc;
    acc.zero();

    // team units: [0, n_chunks) chunks of long rows, then (full kernel) the medium rows / (row-list kernel) the listed
    // rows; they fill the first team_warps warps, TPW teams each.  Short rows of the full kernel follow, one per subgroup.
    const int64_t n_shared = ROWS ? a.max_list : n_med;
    const int64_t team_warps = (n_chunks + n_shared + TPW - 1) / TPW;
    const int64_t unit = warp * TPW + team;

    if (warp < team_warps && unit < n_chunks) {
        // ---- one chunk of a long row
        const int ch = (int)unit;
        const int64_t r = a.g.chunk_row[ch];
        if (ROWS) {
            // is this long row on the list?  (ascending ids: binary search, same answer on every lane)
            const int64_t want = a.row0 + r;
            int lo = 0, hi = *a.n_list;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(a.row_list + mid) < want) lo = mid + 1; else hi = mid;
            }
            if (lo >= *a.n_list || __ldg(a.row_list + lo) != want) return;
        }
        int64_t beg = a.g.chunk_begin[ch], end = beg + a.g.chunk_len[ch];
        split_range<LPR>(beg, end, tsub);
        gather_range<LPR, V, EXACT, MODE, DROP>(a, acc, beg, end, a.row0 + r, lane, gmask, seed);
        combine_subgroups<LPR, V, EXACT>(acc, lane, tsub, team_base);
        if (tsub != 0) return;
        const int first = a.g.chunk_first[ch];
        const int count = a.g.chunk_count[ch];
#pragma unroll
        for (int i = 0; i < V; ++i)
            if (Vec::on(i, lane, D)) st4(a.g.partial + (int64_t)ch * D + (i * LPR + lane) * 4, acc.v[i]);
        // Two-level combine in a fixed order.  Chunks are grouped by kSuper = 32: the last chunk of a group to finish
        // adds the group's partials in chunk order; a row of more than 32 chunks then adds its group sums in group
        // order (again by whoever finishes last).  A single sequential pass over all partials -- 380 of them for the
        // most popular item of the Amazon shape, four loads in flight -- was a 30 us serial tail, invisible behind a
        // 100 us full-width layer but THE floor of the narrow layers of the column-sharded step.
        constexpr int kSuper = 32;
        auto last_to_arrive = [&](int32_t *counter, int expected) -> bool {
            __threadfence();
            __syncwarp(gmask);
            int old = 0;
            if (lane == 0) old = atomicAdd(counter, 1);
            old = __shfl_sync(gmask, old, 0, LPR);
            if (old != expected - 1) return false;
            __threadfence();
            if (lane == 0) *counter = 0;              // self-reset for the next launch
            return true;
        };
        auto ordered_sum = [&](int64_t slot0, int n, int stride) {      // acc = p[slot0] + p[slot0 + stride] + ... in order
            acc.zero();
            for (int k0 = 0; k0 < n; k0 += 4) {           // 4 partials in flight, added in order
                float4 pp[4][V];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < V; ++i)
                        pp[q][i] = (k0 + q < n && Vec::on(i, lane, D))
                                       ? __ldcg(reinterpret_cast<const float4 *>(a.g.partial + (slot0 + (int64_t)(k0 + q) * stride) * D + (i * LPR + lane) * 4))
                                       : f4zero();
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k0 + q < n)
#pragma unroll
                        for (int i = 0; i < V; ++i) add4(acc.v[i], pp[q][i]);
            }
        };
        const int g0 = first + (ch - first) / kSuper * kSuper;          // first chunk of this chunk's group
        const int g_count = min(kSuper, first + count - g0);
        if (!last_to_arrive(a.g.counters + g0, g_count)) return;
        ordered_sum(g0, g_count, 1);
        if (count > kSuper) {
#pragma unroll
            for (int i = 0; i < V; ++i)                    // the group's sum replaces its first partial (all of them are consumed)
                if (Vec::on(i, lane, D)) st4(a.g.partial + (int64_t)g0 * D + (i * LPR + lane) * 4, acc.v[i]);
            const int n_groups = (count + kSuper - 1) / kSuper;
            if (!last_to_arrive(a.g.counters + a.g.n_chunks + first, n_groups)) return;
            ordered_sum(first, n_groups, kSuper);
        }
        finish_row<LPR, V, EXACT, MODE, DROP>(a, r, acc, lane, seed);
        return;
    }

    int64_t r;
    bool shared_row;                                    // does the whole team work on row r?
    if (warp < team_warps) {
        const int64_t idx = unit - n_chunks;
        if (idx >= n_shared) return;                    // idle team of the last team warp
        if (ROWS) {
            // one listed row per team; its length decides how it is summed (same rule as the full kernel)
            if (idx >= *a.n_list) return;
            r = __ldg(a.row_list + idx) - a.row0;
            if (r < 0 || r >= a.g.n_rows) return;       // another block's row
        } else {
            r = __ldg(a.g.row_order + n_long + idx);
        }
        shared_row = true;
    } else {
        if (ROWS) return;
        const int64_t idx = n_long + n_med + (warp - team_warps) * SUB + sub;
        if (idx >= a.g.n_rows) return;
        r = a.g.row_order ? (int64_t)__ldg(a.g.row_order + idx) : idx;
        shared_row = false;
    }
    int64_t beg = __ldg(a.g.rowptr + r), end = __ldg(a.g.rowptr + r + 1);
    if (ROWS) {
        const int64_t nnz = end - beg;
        if (n_chunks > 0 && nnz > a.g.long_threshold) return;          // the chunk teams own it
        if (!(a.g.row_order && nnz > IGCN_MEDIUM_NNZ)) {                // short row: subgroup 0 of the team alone, like the full kernel
            if (tsub != 0) return;
            shared_row = false;
        }
    }
    if (shared_row) {
        split_range<LPR>(beg, end, tsub);
        gather_range<LPR, V, EXACT, MODE, DROP>(a, acc, beg, end, a.row0 + r, lane, gmask, seed);
        combine_subgroups<LPR, V, EXACT>(acc, lane, tsub, team_base);
        if (tsub != 0) return;
    } else {
        gather_range<LPR, V, EXACT, MODE, DROP>(a, acc, beg, end, a.row0 + r, lane, gmask, seed);
    }
    finish_row<LPR, V, EXACT, MODE, DROP>(a, r, acc, lane, seed);
}

template <int LPR>
static int64_t warp_units(const PropArgs &a, bool rows_variant) {
    constexpr int SUB = 32 / LPR, TPW = Team<LPR>::PER_WARP;
    const int64_t n_long = a.g.row_order ? a.g.n_long_rows : 0, n_med = a.g.row_order ? a.g.n_medium_rows : 0;
    const int64_t n_shared = rows_variant ? a.max_list : n_med;
    const int64_t team_warps = (a.g.n_chunks + n_shared + TPW - 1) / TPW;
    if (rows_variant) return team_warps;
    return team_warps + (a.g.n_rows - n_long - n_med + SUB - 1) / SUB;
}

template <int LPR, int V, bool EXACT, int MODE, int DROP>
static void launch_one(const PropArgs &a, cudaStream_t st) {
    const int64_t warps = warp_units<LPR>(a, MODE == MODE_SPMM && DROP == 1);
    constexpr int WPB = kThreads / 32;
    if (warps <= 0) return;
    prop_kernel<LPR, V, EXACT, MODE, DROP><<<(unsigned)((warps + WPB - 1) / WPB), kThreads, 0, st>>>(a);
}

template <int MODE, int DROP>
static int launch_lanes(const PropArgs &a, cudaStream_t st) {
    if (a.g.n_chunks + ((MODE == MODE_SPMM && DROP == 1) ? a.max_list : a.g.n_rows) == 0) return 0;
    const int D = a.D;
    // D = 64: 8 lanes x 2 vectors per row.  (4 lanes x 4 vectors -- 8 rows per warp like the D = 32 shape below -- was
    // measured: 16 accumulator + 32 in-flight registers per lane spill at the 64-register budget, Yelp-shaped step
    // 0.406 -> 0.624 ms.)
    if (D == 64) launch_one<8, 2, true, MODE, DROP>(a, st);
    else if (D == 32) launch_one<4, 2, true, MODE, DROP>(a, st);      // 8 rows per warp (2-GPU column shards: Amazon-shaped step 0.736 -> 0.626 ms against 8 lanes x 1 vector)
    else if (D == 16) launch_one<4, 1, true, MODE, DROP>(a, st);
    else if (D == 8) launch_one<2, 1, true, MODE, DROP>(a, st);
    else if (D == 128) launch_one<16, 2, true, MODE, DROP>(a, st);
    else if (D < 32) launch_one<8, 1, false, MODE, DROP>(a, st);
    else if (D < 64) launch_one<8, 2, false, MODE, DROP>(a, st);
    else launch_one<16, 2, false, MODE, DROP>(a, st);
    return 0;
}

template <int MODE>
static int launch_drop(const PropArgs &a, cudaStream_t st) {
    switch (a.drop.mode) {
        case 0: return launch_lanes<MODE, 0>(a, st);
        case 1: return launch_lanes<MODE, 1>(a, st);
        default: return launch_lanes<MODE, 2>(a, st);
    }
}

static int check_common(const igcn_csr *g, int32_t D) {
    if (!g) { set_error("null csr"); return -1; }
    if (D <= 0 || D > 128 || (D & 3)) { set_error("embedding size %d unsupported (need D %% 4 == 0, D <= 128)", D); return -1; }
    if (g->n_chunks > 0 && (!g->partial || !g->counters || !g->chunk_row)) { set_error("chunk plan incomplete"); return -1; }
    if (g->n_long_rows < 0 || g->n_medium_rows < 0 || (int64_t)g->n_long_rows + g->n_medium_rows > g->n_rows) {
        set_error("row class counts out of range"); return -1;
    }
    if (!g->row_order && g->n_chunks > 0) { set_error("a chunk plan needs row_order"); return -1; }
    return 0;
}

static int fill_drop(PropArgs &a, const igcn_dropout *drop) {
    if (drop) a.drop = *drop; else { a.drop = igcn_dropout{}; }
    if (a.drop.mode < 0 || a.drop.mode > 2) { set_error("dropout mode %d", a.drop.mode); return -1; }
    if (a.drop.mode != 0 && !(a.drop.p >= 0.f && a.drop.p < 1.f)) { set_error("dropout p out of range"); return -1; }
    if (a.drop.mode == 2 && (!a.drop.edge_keep || !a.drop.self_keep)) { set_error("mode 2 needs keep bits"); return -1; }
    a.thresh = a.drop.mode == 1 ? drop_threshold(a.drop.p) : 0u;
    a.inv_keep = a.drop.mode == 0 ? 1.f : 1.f / (1.f - a.drop.p);
    return 0;
}

// ------------------------------------------------------------------ masked column sums
// A CTA of GROUPS x LANES threads sums 256 rows: group g adds rows g, g + GROUPS, ... in order, then the GROUPS partials
// are added in order.  GROUPS is the same for every width (the CTA size changes instead), so each column's sum is
// formed in the same order whatever D is -- the column-sharded training step (D / ranks columns per GPU) has to
// reproduce the single-GPU bits.
constexpr int kColsumGroups = 16;
template <int LANES>
__global__ void __launch_bounds__(kColsumGroups * LANES) colsum_stage1(const float *__restrict__ G, int64_t row_begin, int64_t row_end,
                                                                      int D, igcn_dropout drop, uint32_t thresh, float *scratch) {
    constexpr int GROUPS = kColsumGroups;
    __shared__ float4 sm[GROUPS][LANES];
    const int lane = threadIdx.x % LANES, grp = threadIdx.x / LANES;
    const bool active = lane * 4 < D;
    uint64_t seed = drop.seed;
    if (drop.mode == 1 && drop.seed_dev) seed = mix64(seed ^ mix64(*drop.seed_dev + 0x2545f491ULL));
    const int64_t base = row_begin + (int64_t)blockIdx.x * 256;
    const int64_t stop = min(row_end, base + 256);
    float4 acc = f4zero();
    // the group's 16 rows are fetched together (independent loads), then added in row order
    float4 v[256 / GROUPS];
#pragma unroll
    for (int q = 0; q < 256 / GROUPS; ++q) {
        const int64_t r = base + grp + (int64_t)q * GROUPS;
        bool keep = r < stop;
        if (keep && drop.mode == 1) keep = edge_hash(seed, (uint32_t)r, kSelfCol) >= thresh;
        if (keep && drop.mode == 2) keep = (__ldg(drop.self_keep + (r >> 5)) >> (r & 31)) & 1u;
        v[q] = (keep && active) ? ld4(G + r * D + lane * 4) : f4zero();
    }
#pragma unroll
    for (int q = 0; q < 256 / GROUPS; ++q) add4(acc, v[q]);
    sm[grp][lane] = acc;
    __syncthreads();
    if (grp == 0 && active) {
        float4 t = sm[0][lane];
        for (int k = 1; k < GROUPS; ++k) add4(t, sm[k][lane]);
        st4(scratch + (int64_t)blockIdx.x * D + lane * 4, t);
    }
}

__global__ void colsum_stage2(const float *__restrict__ scratch, int64_t n_blocks, int D, float *out) {
    const int d = threadIdx.x;
    if (d >= D) return;
    float t = 0.f;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += 8) {           // 8 independent loads, added in block order
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (b0 + q < n_blocks) ? scratch[(b0 + q) * D + d] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (b0 + q < n_blocks) t += v[q];
    }
    out[d] = t;
}

}  // namespace igcn

using namespace igcn;

static int fill_peers(PropArgs &a, float *const *peer_host, int32_t n_peers) {
    if (n_peers < 0 || n_peers > IGCN_MAX_PEERS || (n_peers > 0 && !peer_host)) { set_error("bad peer list"); return -1; }
    a.n_peers = n_peers;
    for (int p = 0; p < n_peers; ++p) a.peer[p] = peer_host[p];
    return 0;
}

extern "C" int igcn_spmm(const igcn_csr *g, const float *X, float *Y, int32_t D, const float *const *add_host,
                         int32_t n_add, const float *rowscale, float alpha, float *const *peer_y_host, int32_t n_peers,
                         void *stream) {
    if (check_common(g, D)) return -1;
    IGCN_CHECK_ARG(X && Y, "null X/Y");
    IGCN_CHECK_ARG(n_add >= 0 && n_add <= IGCN_MAX_ADD, "n_add out of range");
    PropArgs a{};
    if (fill_peers(a, peer_y_host, n_peers)) return -1;
    a.g = *g; a.X = X; a.Y = Y; a.D = D; a.n_add = n_add; a.rowscale = rowscale; a.alpha = alpha;
    for (int j = 0; j < n_add; ++j) a.add[j] = add_host[j];
    launch_lanes<MODE_SPMM, 0>(a, as_stream(stream));
    IGCN_CHECK_LAUNCH();
    return 0;
}

static int spmm_common(PropArgs &a, const igcn_csr *g, const float *X, float *Y, int32_t D, const float *const *add_host,
                       int32_t n_add, const float *rowscale, float alpha, float *const *peer_y_host, int32_t n_peers) {
    if (check_common(g, D)) return -1;
    IGCN_CHECK_ARG(X && Y, "null X/Y");
    IGCN_CHECK_ARG(n_add >= 0 && n_add <= IGCN_MAX_ADD, "n_add out of range");
    if (fill_peers(a, peer_y_host, n_peers)) return -1;
    a.g = *g; a.X = X; a.Y = Y; a.D = D; a.n_add = n_add; a.rowscale = rowscale; a.alpha = alpha;
    for (int j = 0; j < n_add; ++j) a.add[j] = add_host[j];
    return 0;
}

extern "C" int igcn_spmm_rows(const igcn_csr *g, const float *X, float *Y, int32_t D, const float *const *add_host,
                              int32_t n_add, const float *rowscale, float alpha, const int64_t *row_list,
                              const int32_t *n_list, int64_t max_list, int64_t row0, float *const *peer_y_host,
                              int32_t n_peers, void *stream) {
    PropArgs a{};
    if (spmm_common(a, g, X, Y, D, add_host, n_add, rowscale, alpha, peer_y_host, n_peers)) return -1;
    IGCN_CHECK_ARG(row_list && n_list && max_list >= 0, "row list missing");
    a.row_list = row_list; a.n_list = n_list; a.max_list = max_list; a.row0 = row0;
    launch_lanes<MODE_SPMM, 1>(a, as_stream(stream));
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_spmm_cols(const igcn_csr *g, const float *X, float *Y, int32_t D, const float *const *add_host,
                              int32_t n_add, const float *rowscale, float alpha, const uint32_t *col_bits,
                              float *const *peer_y_host, int32_t n_peers, void *stream) {
    PropArgs a{};
    if (spmm_common(a, g, X, Y, D, add_host, n_add, rowscale, alpha, peer_y_host, n_peers)) return -1;
    IGCN_CHECK_ARG(col_bits, "column bitmap missing");
    a.col_bits = col_bits;
    launch_lanes<MODE_SPMM, 2>(a, as_stream(stream));
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_inmo_fwd(const igcn_csr *g, const int32_t *tmpl, const float *rowscale, const igcn_dropout *drop,
                             const float *E, float *X0, int32_t D, int64_t row0, int64_t n_users, int64_t glob_user,
                             int64_t glob_item, float *const *peer_x0_host, int32_t n_peers, void *stream) {
    if (check_common(g, D)) return -1;
    IGCN_CHECK_ARG(E && X0 && rowscale, "null E/X0/rowscale");
    PropArgs a{};
    if (fill_peers(a, peer_x0_host, n_peers)) return -1;
    a.g = *g; a.X = E; a.Y = X0; a.D = D; a.rowscale = rowscale; a.alpha = 1.f; a.tmpl = tmpl;
    a.row0 = row0; a.n_users = n_users; a.glob_user = glob_user; a.glob_item = glob_item;
    if (fill_drop(a, drop)) return -1;
    launch_drop<MODE_INMO_FWD>(a, as_stream(stream));
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_inmo_bwd(const igcn_csr *g, const int32_t *tmpl, const igcn_dropout *drop, const float *G,
                             float *dE, int32_t D, int64_t row0, float *const *peer_de_host, int32_t n_peers, void *stream) {
    if (check_common(g, D)) return -1;
    IGCN_CHECK_ARG(G && dE, "null G/dE");
    PropArgs a{};
    if (fill_peers(a, peer_de_host, n_peers)) return -1;
    a.g = *g; a.X = G; a.Y = dE; a.D = D; a.alpha = 1.f; a.tmpl = tmpl; a.row0 = row0;
    if (fill_drop(a, drop)) return -1;
    IGCN_CHECK_ARG(a.drop.mode != 2 || a.drop.tperm, "mode 2 backward needs tperm");
    launch_drop<MODE_INMO_BWD>(a, as_stream(stream));
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_colsum_masked(const float *G, int64_t row_begin, int64_t row_end, int32_t D,
                                  const igcn_dropout *drop, float *scratch, float *out, void *stream) {
    IGCN_CHECK_ARG(G && scratch && out, "null pointer");
    IGCN_CHECK_ARG(D > 0 && D <= 128 && !(D & 3), "embedding size unsupported");
    IGCN_CHECK_ARG(row_end >= row_begin, "bad row range");
    PropArgs a{};
    if (fill_drop(a, drop)) return -1;
    cudaStream_t st = as_stream(stream);
    const int64_t n_blocks = (row_end - row_begin + 255) / 256;
    if (n_blocks > 0) {
        if (D <= 32) colsum_stage1<8><<<(unsigned)n_blocks, kColsumGroups * 8, 0, st>>>(G, row_begin, row_end, D, a.drop, a.thresh, scratch);
        else if (D <= 64) colsum_stage1<16><<<(unsigned)n_blocks, kColsumGroups * 16, 0, st>>>(G, row_begin, row_end, D, a.drop, a.thresh, scratch);
        else colsum_stage1<32><<<(unsigned)n_blocks, kColsumGroups * 32, 0, st>>>(G, row_begin, row_end, D, a.drop, a.thresh, scratch);
    }
    colsum_stage2<<<1, 128, 0, st>>>(scratch, n_blocks, D, out);
    IGCN_CHECK_LAUNCH();
    return 0;
}
