// Microbenchmark: TMEM -> register read bandwidth (tcgen05.ld 32x32b.x32) per SM as a function of the
// number of reading warps.  Evidence for the epilogue bound of the scoring kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(256, 1) bw_kernel(int iters, int n_warps, long long *cycles, uint32_t *sink) {
    __shared__ uint32_t tmem_base_sh;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_sh;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < n_warps) {
        // warp w may only touch TMEM lanes [32*(w%4), +32)
        const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t va[32], vb[32];
        t0 = clock64();
        tc_ld32(taddr, va);
        for (int i = 0; i < iters; ++i) {
            tc_wait_ld();
            tc_ld32(taddr + ((2 * i + 1) & 15) * 32, vb);
#pragma unroll
            for (int c = 0; c < 32; ++c) acc ^= va[c];
            tc_wait_ld();
            tc_ld32(taddr + ((2 * i + 2) & 15) * 32, va);
#pragma unroll
            for (int c = 0; c < 32; ++c) acc ^= vb[c];
        }
        tc_wait_ld();
        t1 = clock64();
    }
    if (warp < n_warps && (threadIdx.x & 31) == 0) cycles[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
}

int main() {
    long long *cyc;
    uint32_t *sink;
    cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const int iters = 20000;
    for (int n_warps : {1, 2, 4, 8}) {
        cudaMemset(cyc, 0, 148 * 8 * sizeof(long long));
        bw_kernel<<<148, 256>>>(iters, n_warps, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[8];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < n_warps; ++w) mx = h[w] > mx ? h[w] : mx;
        const double bytes = (double)n_warps * (2.0 * iters + 1) * 32 * 32 * 4;
        printf("warps %d: %lld cycles, %.1f B/cycle/SM, %.2f fp32 values/cycle/SM\n", n_warps, mx, bytes / mx, bytes / mx / 4);
    }
    return 0;
}
