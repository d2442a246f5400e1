// L2 gather ceiling microbenchmark ("report achieved gather GB/s against ... the measured L2 ceiling").
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/l2_gather_bw scripts/ubench/l2_gather_bw.cu
//   ./l2_gather_bw            -> one JSON line
//
// The renderer's hash-grid tables (7.4 + 4.4 + 4.4 MB) are L2-resident, so the field kernels' gathers are bounded by what L2/L1 can
// serve, not by HBM.  This measures that ceiling for the access shapes that matter, 8-byte (float2) loads through the read-only path,
// 32 independent loads in flight per thread, a grid of 148 x 8 CTAs x 256 threads:
//   random      : every load hits a uniformly random entry of a 16 MB table (1 table, and 2 tables alternating)  -> each 8-B load costs
//                 one 32-B L2 sector; no L1 reuse.  Worst case: the rate of UNCOALESCED sectors L2 can return.
//   coherent_K  : the 32 lanes of a warp read entries inside a window of K consecutive entries around a random base (K = 64, 256): the
//                 situation of neighbouring samples on a fine grid level (lanes share sectors/lines -> L1 hits + sector reuse).
//   dense_small : random entries of a 39 KB table (3-D level 0: 4,920 entries) -> L1-resident ceiling.
// Reported: GB/s of REQUESTED bytes (8 B per load), the same accounting as roofline.achieved in bench.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t xorshift(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// mode 0: random; mode 1: coherent window (window entries, power of two)
template <int NT>
__global__ void __launch_bounds__(256) k_gather(const float2* __restrict__ t0, const float2* __restrict__ t1, uint32_t mask, uint32_t window,
                                                int iters, float2* __restrict__ out) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t sw = (blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 40503u + 977u;   // per-warp stream for the window base
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; it++) {
        float2 v[32];
        #pragma unroll
        for (int k = 0; k < 32; k++) {
            uint32_t idx;
            if (window) {
                const uint32_t base = xorshift(sw) & mask;
                idx = (base + (xorshift(s) & (window - 1))) & mask;
            } else {
                idx = xorshift(s) & mask;
            }
            const float2* t = (NT == 2 && (k & 1)) ? t1 : t0;
            v[k] = __ldg(t + idx);
        }
        #pragma unroll
        for (int k = 0; k < 32; k++) { acc.x += v[k].x; acc.y += v[k].y; }
    }
    if (acc.x == 123.456f) out[0] = acc;     // keep the loads alive
}

static float run(int nt, const float2* t0, const float2* t1, uint32_t entries, uint32_t window, int iters, int grid, float2* out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        if (nt == 1) k_gather<1><<<grid, 256>>>(t0, t1, entries - 1, window, iters, out);
        else k_gather<2><<<grid, 256>>>(t0, t1, entries - 1, window, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;     // rep 0 = warm-up (table first touch from HBM)
    }
    const double bytes = (double)grid * 256 * iters * 32 * 8;
    return (float)(bytes / (best * 1e-3) / 1e9);
}

int main() {
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint32_t E = 2u << 20;                 // 2 Mi entries x 8 B = 16 MB
    float2 *t0, *t1, *out;
    CK(cudaMalloc(&t0, (size_t)E * 8)); CK(cudaMalloc(&t1, (size_t)E * 8)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(t0, 0, (size_t)E * 8)); CK(cudaMemset(t1, 0, (size_t)E * 8));
    const int grid = sms * 8, iters = 64;
    const float r1 = run(1, t0, t1, E, 0, iters, grid, out);
    const float r2 = run(2, t0, t1, E, 0, iters, grid, out);
    const float c64 = run(1, t0, t1, E, 64, iters, grid, out);
    const float c256 = run(1, t0, t1, E, 256, iters, grid, out);
    const float c64_2 = run(2, t0, t1, E, 64, iters, grid, out);
    const float small = run(1, t0, t1, 4096, 0, iters, grid, out);
    printf("{\"what\": \"8-byte gathers, requested GB/s (32 loads in flight/thread, %d CTAs x 256 thr)\", \"sms\": %d, "
           "\"random_16MB_1table\": %.1f, \"random_16MB_2tables\": %.1f, \"coherent64_1table\": %.1f, \"coherent256_1table\": %.1f, "
           "\"coherent64_2tables\": %.1f, \"l1_resident_32KB\": %.1f}\n", grid, sms, r1, r2, c64, c256, c64_2, small);
    return 0;
}
