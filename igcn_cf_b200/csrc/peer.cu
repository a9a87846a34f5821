// Peer memory over NVLink 5 / NVSwitch for the row-sharded propagation.
//
// One process per GPU.  Every rank allocates the same set of buffers with igcn_peer_alloc, ships the
// 64-byte IPC handles to the other ranks (the host code uses torch.distributed for that) and maps
// theirs with igcn_peer_open.  The propagation kernels then store every finished row into all ranks'
// copies (igcn_spmm's peer list): the all-gather that would follow a layer is part of the SpMM.
// igcn_peer_barrier is the only synchronisation between layers: a one-CTA kernel that publishes this
// rank's epoch in every rank's flag array (st.release.sys) and waits until every rank has published
// the same epoch (ld.acquire.sys).  It lives on the launch stream, so it can be captured in the
// step's CUDA graph; a bounded spin turns a lost peer into a trap (the context dies, every
// later CUDA call of the process fails loudly) instead of a hung GPU or a silently wrong layer.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace igcn {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct BarrierArgs {
    uint32_t *flags[IGCN_MAX_PEERS];   // flags[p] = rank p's flag array [IGCN_MAX_PEERS]
    int n_peers, rank;
    uint32_t *epoch;                   // local counter
    uint32_t *status;                  // local: set to 1 when the spin timed out
    long long timeout_cycles;
};

__global__ void peer_barrier_kernel(const __grid_constant__ BarrierArgs a) {
    __shared__ uint32_t e_sh;
    if (threadIdx.x == 0) e_sh = *a.epoch + 1;
    __syncthreads();
    const uint32_t e = e_sh;
    const int p = threadIdx.x;
    if (p < a.n_peers) {
        // everything this stream did before the barrier happens-before the release store
        __threadfence_system();
        st_release_sys(a.flags[p] + a.rank, e);
        const uint32_t *mine = a.flags[a.rank] + p;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            if (clock64() - t0 > a.timeout_cycles) {
                // a lost peer is fatal: record it where the host can read it (mapped status word) and kill the
                // context -- the layers behind this barrier would otherwise read rows that never arrived
                *a.status = 1u;
                __threadfence_system();
                __trap();
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.epoch = e;
}

struct PushArgs {
    float *peer[IGCN_MAX_PEERS];
    int n_peers, rank;
    int64_t off, n4;                   // element offset of the block, number of float4 in it
};

// Bulk form of the all-gather: this rank's freshly written row block streams out to every other rank's copy
// in long coalesced bursts (one 512-byte store per warp instruction and peer).
__global__ void __launch_bounds__(256) peer_push_kernel(const __grid_constant__ PushArgs a) {
    const float4 *__restrict__ src = reinterpret_cast<const float4 *>(a.peer[a.rank] + a.off);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n4; i += stride) {
        const float4 v = __ldcg(src + i);
        for (int p = 0; p < a.n_peers; ++p)
            if (p != a.rank) __stcg(reinterpret_cast<float4 *>(a.peer[p] + a.off) + i, v);
    }
}

struct PushColsArgs {
    float *peer[IGCN_MAX_PEERS];
    int n_peers;
    const float *src;                  // [rows, ds] contiguous
    int64_t rows;
    int ds, d, col0;                   // slice width, full row width, first column of the slice
};

// Column-sharded training: every rank owns columns [col0, col0 + ds) of a [rows, d] table.  This writes the rank's
// slice into that column range of EVERY rank's full-width copy (its own included): the all-gather of the parameters
// that an evaluation or a checkpoint needs, once per epoch.
__global__ void __launch_bounds__(256) peer_push_cols_kernel(const __grid_constant__ PushColsArgs a) {
    const int per_row = a.ds / 4;
    const int64_t n4 = a.rows * per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const int64_t r = i / per_row;
        const int c = (int)(i % per_row) * 4;
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(a.src + r * a.ds + c));
        for (int p = 0; p < a.n_peers; ++p) __stcg(reinterpret_cast<float4 *>(a.peer[p] + r * a.d + a.col0 + c), v);
    }
}

}  // namespace igcn

using namespace igcn;

#define IGCN_CUDA(call)                                                            \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            igcn::set_error("%s: %s", __func__, cudaGetErrorString(e__));          \
            return (int)e__;                                                       \
        }                                                                          \
    } while (0)

extern "C" int igcn_peer_alloc(int64_t bytes, void **ptr_out, uint8_t *handle_out) {
    IGCN_CHECK_ARG(bytes > 0 && ptr_out && handle_out, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == IGCN_PEER_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    IGCN_CUDA(cudaMalloc(&p, (size_t)bytes));
    IGCN_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("igcn_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        return (int)e;
    }
    IGCN_CUDA(cudaDeviceSynchronize());
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return 0;
}

extern "C" int igcn_peer_open(const uint8_t *handle, void **ptr_out) {
    IGCN_CHECK_ARG(handle && ptr_out, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    IGCN_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return 0;
}

extern "C" int igcn_peer_close(void *ptr) {
    IGCN_CHECK_ARG(ptr, "null pointer");
    IGCN_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int igcn_peer_free(void *ptr) {
    IGCN_CHECK_ARG(ptr, "null pointer");
    IGCN_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int igcn_peer_barrier(uint32_t *const *flags_host, int32_t n_peers, int32_t rank, uint32_t *epoch_dev,
                                 uint32_t *status_dev, void *stream) {
    IGCN_CHECK_ARG(flags_host && epoch_dev && status_dev, "null pointer");
    IGCN_CHECK_ARG(n_peers >= 1 && n_peers <= IGCN_MAX_PEERS && rank >= 0 && rank < n_peers, "bad rank / peer count");
    BarrierArgs a{};
    for (int p = 0; p < n_peers; ++p) a.flags[p] = flags_host[p];
    a.n_peers = n_peers; a.rank = rank; a.epoch = epoch_dev; a.status = status_dev;
    // host-side skew of seconds between ranks is normal (checkpoint I/O, a CUDA-graph capture on one rank), so the
    // default is generous; IGCN_PEER_TIMEOUT_S overrides it
    static long long timeout_cycles = 0;
    if (!timeout_cycles) {
        const char *env = getenv("IGCN_PEER_TIMEOUT_S");
        double sec = env ? atof(env) : 120.0;
        if (!(sec > 0.0)) sec = 120.0;
        timeout_cycles = (long long)(sec * 1.9e9);
    }
    a.timeout_cycles = timeout_cycles;
    peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(a);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_peer_push(float *const *peer_host, int32_t n_peers, int32_t rank, int64_t elem_offset, int64_t n_elems,
                              void *stream) {
    IGCN_CHECK_ARG(peer_host, "null pointer");
    IGCN_CHECK_ARG(n_peers >= 1 && n_peers <= IGCN_MAX_PEERS && rank >= 0 && rank < n_peers, "bad rank / peer count");
    IGCN_CHECK_ARG(elem_offset >= 0 && n_elems >= 0 && !(elem_offset & 3) && !(n_elems & 3), "block must be float4 aligned");
    if (n_elems == 0 || n_peers == 1) return 0;
    PushArgs a{};
    for (int p = 0; p < n_peers; ++p) a.peer[p] = peer_host[p];
    a.n_peers = n_peers; a.rank = rank; a.off = elem_offset; a.n4 = n_elems / 4;
    int64_t blocks = (a.n4 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    peer_push_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
    IGCN_CHECK_LAUNCH();
    return 0;
}

extern "C" int igcn_peer_push_cols(float *const *peer_host, int32_t n_peers, const float *src, int64_t rows, int32_t ds, int32_t d,
                                   int32_t col0, void *stream) {
    IGCN_CHECK_ARG(peer_host && src, "null pointer");
    IGCN_CHECK_ARG(n_peers >= 1 && n_peers <= IGCN_MAX_PEERS, "bad peer count");
    IGCN_CHECK_ARG(ds > 0 && !(ds & 3) && !(d & 3) && !(col0 & 3) && col0 >= 0 && col0 + ds <= d, "slice must be float4 aligned and inside the row");
    if (rows <= 0) return 0;
    PushColsArgs a{};
    for (int p = 0; p < n_peers; ++p) a.peer[p] = peer_host[p];
    a.n_peers = n_peers; a.src = src; a.rows = rows; a.ds = ds; a.d = d; a.col0 = col0;
    int64_t blocks = (rows * (ds / 4) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    peer_push_cols_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
    IGCN_CHECK_LAUNCH();
    return 0;
}
