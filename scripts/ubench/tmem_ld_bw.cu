// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a, for 4 and 8 warps per CTA.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu ; run: ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: dependent (ld; wait; ld; wait ...) -> latency.  mode 1: 4 loads in flight then wait -> throughput
__global__ void k(int iters, int mode, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; i++) {
            uint32_t r[32];
            ld32(base + 32 * (i & 3), r);
            wait_ld();
            acc += r[0] ^ r[13] ^ r[31];
        }
    } else {
        for (int i = 0; i < iters; i += 2) {
            uint32_t r0[32], r1[32];
            ld32(base, r0);
            ld32(base + 32, r1);
            wait_ld();
            acc += r0[0] ^ r0[31] ^ r1[5] ^ r1[31];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
    const int iters = 4096;
    for (int warps = 1; warps <= 8; warps *= 2)
        for (int mode = 0; mode < 2; mode++) {
            k<<<148, 32 * warps>>>(iters, mode, out, sink);
            k<<<148, 32 * warps>>>(iters, mode, out, sink);
            long long h[148];
            if (cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            const double clk = (double)h[0];
            const double bytes = (double)iters * warps * 4096.0;
            printf("warps %d mode %s: %.1f clk per x32 load per warp, %.1f B/clk/SM\n", warps, mode ? "2-in-flight" : "dependent", clk / iters, bytes / clk);
        }
    return 0;
}
