// EXPERIMENTAL (opt-in, IGCN_SPMM_HOT=<rows>): full propagation layer with the hottest gathered rows staged in
// shared memory -- the "shared-memory staging of hot item rows" of BASELINE.json's north_star.
//
// prop_kernel (propagate.cu) runs at ~0.8 of the L2->SM fabric rate on the paper-sized graphs (the whole table
// is L2-resident); the only way below that is to serve part of the gathers from inside the SM.
// Here ONE persistent CTA of 1,024 threads per SM (same 32 warps / 64 registers as 4 x 256) first copies the
// n_hot highest-degree rows of X into shared memory (768 rows = 192 KB, once per launch: 29 MB chip-wide
// against the 717 MB a Yelp-shaped layer gathers), then its warps walk the same warp units as prop_kernel
// (chunks of long rows, medium rows, four short rows per warp; units dealt round-robin, longest first).  The
// column array is pre-encoded on the device (graph.CsrDevice.hot_plan): a hot column c is stored as
// -(slot + 1), so the inner loop picks the shared-memory or the global address with one select and issues the
// same generic 128-bit loads.  Every row is summed in exactly prop_kernel's order -- results are bit-identical
// (tests/test_gpu_step_kernels.py, run with IGCN_EXPERIMENTAL=1).  D = 64, single GPU, no dropout only.
// Measured coverage on the synthetic graphs: 768 rows serve 18-20 % of the gathers; whether that pays for the
// extra select per load and the static schedule is what the first B200 run of this file has to show.
#include "common.cuh"

namespace igcn {

struct HotArgs {
    igcn_csr g;
    const int32_t *col_enc;   // g.col with hot columns replaced by -(slot + 1)
    const int32_t *hot_ids;   // [n_hot] rows of X staged in shared memory, slot order
    int n_hot;
    const float *X;
    float *Y;
    const float *add[IGCN_MAX_ADD];
    int n_add;
    const float *rowscale;
    float alpha;
    int64_t total_units;
};

constexpr int kHotThreads = 1024;
constexpr int kD = 64, kLPR = 8, kSUB = 4;

__device__ __forceinline__ void hot_gather(const HotArgs &a, const float *hot, float4 (&acc)[2], int64_t beg, int64_t end,
                                           int lane, uint32_t gmask) {
    const int32_t *__restrict__ colp = a.col_enc + beg;
    const float *__restrict__ valp = a.g.val ? a.g.val + beg : nullptr;
    const int len = (int)(end - beg);
    const float *Tl = a.X + lane * 4;
    const float *Hl = hot + lane * 4;
    int c_next = 0;
    float v_next = 1.f;
    if (lane < len) {
        c_next = __ldg(colp + lane);
        if (valp) v_next = __ldg(valp + lane);
    }
    for (int o = 0; o < len; o += kLPR) {
        const int n = min(kLPR, len - o);
        int c = c_next;
        float w = v_next;
        {
            const int e = o + kLPR + lane;
            if (e < len) {
                c_next = __ldg(colp + e);
                if (valp) v_next = __ldg(valp + e);
            }
        }
        if (lane >= n) { w = 0.f; c = 0; }
        for (int j0 = 0; j0 < n; j0 += 4) {
            float4 x[4][2];
            float ww[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int cj = __shfl_sync(gmask, c, j0 + q, kLPR);
                ww[q] = __shfl_sync(gmask, w, j0 + q, kLPR);
                const bool ok = j0 + q < n;
                const float *p = cj < 0 ? Hl + (int64_t)(-cj - 1) * kD : Tl + (int64_t)cj * kD;   // generic address
                x[q][0] = ok ? ld4(p) : f4zero();
                x[q][1] = ok ? ld4(p + kLPR * 4) : f4zero();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                fma4(acc[0], ww[q], x[q][0]);
                fma4(acc[1], ww[q], x[q][1]);
            }
        }
    }
}

__device__ __forceinline__ void hot_combine(float4 (&acc)[2], int lane, int sub) {
#pragma unroll
    for (int s = 1; s < kSUB; ++s) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float4 o;
            o.x = __shfl_sync(0xffffffffu, acc[i].x, lane + s * kLPR);
            o.y = __shfl_sync(0xffffffffu, acc[i].y, lane + s * kLPR);
            o.z = __shfl_sync(0xffffffffu, acc[i].z, lane + s * kLPR);
            o.w = __shfl_sync(0xffffffffu, acc[i].w, lane + s * kLPR);
            if (sub == 0) add4(acc[i], o);
        }
    }
}

__device__ __forceinline__ void hot_split(int64_t &beg, int64_t &end, int sub) {
    const int64_t q = (((end - beg) + kSUB - 1) / kSUB + kLPR - 1) & ~(int64_t)(kLPR - 1);
    const int64_t b = beg + sub * q;
    end = min(end, b + q);
    beg = min(b, end);
}

__device__ __forceinline__ void hot_finish(const HotArgs &a, int64_t r, float4 (&acc)[2], int lane) {
    float s = a.alpha;
    if (a.rowscale) s *= __ldg(a.rowscale + r);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int off = (i * kLPR + lane) * 4;
        float4 t = acc[i];
        for (int j = 0; j < a.n_add; ++j) add4(t, ld4(a.add[j] + r * kD + off));
        st4(a.Y + r * kD + off, scale4(t, s));
    }
}

__global__ void __launch_bounds__(kHotThreads, 1) prop_hot_kernel(const __grid_constant__ HotArgs a) {
    extern __shared__ __align__(16) float hot[];              // [n_hot][64]
    for (int idx = threadIdx.x; idx < a.n_hot * (kD / 4); idx += kHotThreads) {
        const int64_t row = __ldg(a.hot_ids + (idx >> 4));
        reinterpret_cast<float4 *>(hot)[idx] = ld4(a.X + row * kD + (idx & 15) * 4);
    }
    __syncthreads();

    const int lane = threadIdx.x % kLPR;
    const int sub = (threadIdx.x & 31) / kLPR;
    const uint32_t gmask = 0xffu << ((threadIdx.x & 31) & ~(kLPR - 1));
    const int64_t n_chunks = a.g.n_chunks, n_long = a.g.n_long_rows, n_med = a.g.n_medium_rows;
    constexpr int WPB = kHotThreads / 32;
    for (int64_t warp = (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5); warp < a.total_units; warp += (int64_t)gridDim.x * WPB) {
        // ---- which unit: a chunk of a long row, a medium row (both shared by the warp's four groups) or four short rows
        const bool is_chunk = warp < n_chunks;
        const bool shared_row = warp < n_chunks + n_med;
        int64_t r, beg, end;
        if (is_chunk) {
            r = a.g.chunk_row[warp];
            beg = a.g.chunk_begin[warp];
            end = beg + a.g.chunk_len[warp];
        } else {
            const int64_t idx = shared_row ? n_long + (warp - n_chunks) : n_long + n_med + (warp - n_chunks - n_med) * kSUB + sub;
            if (idx >= a.g.n_rows) continue;                 // ragged last unit of short rows: no later shuffle needs this group
            r = __ldg(a.g.row_order + idx);
            beg = __ldg(a.g.rowptr + r);
            end = __ldg(a.g.rowptr + r + 1);
        }
        if (shared_row) hot_split(beg, end, sub);
        float4 acc[2] = {f4zero(), f4zero()};
        hot_gather(a, hot, acc, beg, end, lane, gmask);
        if (shared_row) {
            hot_combine(acc, lane, sub);
            if (sub != 0) continue;
        }
        if (is_chunk) {
            // same protocol as prop_kernel: partial sums to g.partial, the last chunk of the row to arrive adds them in order
            const int ch = (int)warp;
            const int first = a.g.chunk_first[ch];
            const int count = a.g.chunk_count[ch];
#pragma unroll
            for (int i = 0; i < 2; ++i) st4(a.g.partial + (int64_t)ch * kD + (i * kLPR + lane) * 4, acc[i]);
            __threadfence();
            __syncwarp(gmask);
            int old = 0;
            if (lane == 0) old = atomicAdd(a.g.counters + first, 1);
            old = __shfl_sync(gmask, old, 0, kLPR);
            if (old != count - 1) continue;
            __threadfence();
            if (lane == 0) a.g.counters[first] = 0;
            acc[0] = acc[1] = f4zero();
            for (int k0 = 0; k0 < count; k0 += 4) {
                float4 pp[4][2];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        pp[q][i] = (k0 + q < count)
                                       ? __ldcg(reinterpret_cast<const float4 *>(a.g.partial + (int64_t)(first + k0 + q) * kD + (i * kLPR + lane) * 4))
                                       : f4zero();
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (k0 + q < count) {
                        add4(acc[0], pp[q][0]);
                        add4(acc[1], pp[q][1]);
                    }
            }
        }
        hot_finish(a, r, acc, lane);
    }
}

}  // namespace igcn

using namespace igcn;

extern "C" int igcn_spmm_hot(const igcn_csr *g, const int32_t *col_enc, const int32_t *hot_ids, int32_t n_hot, const float *X,
                             float *Y, int32_t D, const float *const *add_host, int32_t n_add, const float *rowscale,
                             float alpha, void *stream) {
    IGCN_CHECK_ARG(g && col_enc && hot_ids && X && Y, "null pointer");
    IGCN_CHECK_ARG(D == kD, "hot-row staging is built for D == 64");
    IGCN_CHECK_ARG(g->row_order, "needs the degree-sorted row order");
    IGCN_CHECK_ARG(n_hot >= 0 && (size_t)n_hot * kD * 4 <= 200 * 1024, "n_hot must fit 200 KB of shared memory (<= 800 rows)");
    IGCN_CHECK_ARG(n_add >= 0 && n_add <= IGCN_MAX_ADD, "n_add out of range");
    IGCN_CHECK_ARG(g->n_chunks == 0 || (g->partial && g->counters && g->chunk_row), "chunk plan incomplete");
    if (g->n_rows == 0) return 0;
    HotArgs a{};
    a.g = *g; a.col_enc = col_enc; a.hot_ids = hot_ids; a.n_hot = n_hot; a.X = X; a.Y = Y;
    a.n_add = n_add; a.rowscale = rowscale; a.alpha = alpha;
    for (int j = 0; j < n_add; ++j) a.add[j] = add_host[j];
    const int64_t n_long = g->n_long_rows, n_med = g->n_medium_rows;
    a.total_units = g->n_chunks + n_med + (g->n_rows - n_long - n_med + kSUB - 1) / kSUB;
    // one process drives one GPU: the SM count and the shared-memory opt-in are set up on the first call (which the
    // callers make eagerly, before any CUDA-graph capture) and reused
    static int sms = 0;
    static size_t smem_ok = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const size_t smem = (size_t)n_hot * kD * 4;
    if (smem > smem_ok) {
        cudaError_t e = cudaFuncSetAttribute(prop_hot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("igcn_spmm_hot: %s", cudaGetErrorString(e)); return (int)e; }
        smem_ok = smem;
    }
    const int64_t ctas = min((int64_t)sms, (a.total_units + kHotThreads / 32 - 1) / (kHotThreads / 32));
    prop_hot_kernel<<<(unsigned)ctas, kHotThreads, smem, as_stream(stream)>>>(a);
    IGCN_CHECK_LAUNCH();
    return 0;
}
