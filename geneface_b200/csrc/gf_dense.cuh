// fp32 SIMT dense layers on a tile of 128 samples held k-major in shared memory
// (act[k][s], s = 0..127).  Used by the reference-arithmetic ("precision 0") field and by the
// torso field.  Block = 256 threads; each thread owns an 8 (samples) x 8 (outputs) register
// tile; weights stream from global (pre-transposed [K][ldw]) through a cp.async double buffer.
#pragma once
#include "gf_common.cuh"

namespace gf {

constexpr int TILE_S = 128;      // samples per tile
constexpr int DENSE_KC = 16;     // k-chunk
constexpr int DENSE_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// out[n][s] = act( sum_k in[k][s] * Wt[k][n] + bias[n] ),  n < N <= 128 (N multiple of 8), K multiple of 16 not required.
// `in` and `out` are shared-memory k-major tiles; wstage = shared [2][DENSE_KC][128] floats.
// Wt rows must be 16-byte aligned (ldw % 4 == 0).  All 256 threads must call.
__device__ __forceinline__ void dense_tile(const float* __restrict__ in, int K, const float* __restrict__ Wt, int ldw, int N,
                                           float* __restrict__ out, const float* __restrict__ bias, bool relu,
                                           float* __restrict__ wstage) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;   // ty: sample group (8 samples), tx: output group (8 outputs)
    const int s0 = ty * 8, n0 = tx * 8;
    float acc[8][8];
    #pragma unroll
    for (int i = 0; i < 8; i++)
        #pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    const int nchunks = (K + DENSE_KC - 1) / DENSE_KC;
    const int vec_per_row = N >> 2;                        // float4 per weight row actually used
    auto stage = [&](int c, int buf) {
        // copy rows k = c*16 .. +15 (clipped to K), N floats each, into wstage[buf][kk][0..N)
        const int k0 = c * DENSE_KC;
        const int total = DENSE_KC * vec_per_row;
        for (int i = tid; i < total; i += DENSE_THREADS) {
            const int kk = i / vec_per_row, v = i - kk * vec_per_row;
            float* dst = wstage + (buf * DENSE_KC + kk) * 128 + v * 4;
            if (k0 + kk < K) cp_async16(dst, Wt + (size_t)(k0 + kk) * ldw + v * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };
    stage(0, 0);
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) { stage(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* w = wstage + ((c & 1) * DENSE_KC) * 128;
        const int k0 = c * DENSE_KC;
        const int kmax = min(DENSE_KC, K - k0);
        if (n0 < N) {
            #pragma unroll 4
            for (int kk = 0; kk < kmax; kk++) {
                const float4 a0 = *reinterpret_cast<const float4*>(in + (size_t)(k0 + kk) * TILE_S + s0);
                const float4 a1 = *reinterpret_cast<const float4*>(in + (size_t)(k0 + kk) * TILE_S + s0 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(w + kk * 128 + n0);
                const float4 b1 = *reinterpret_cast<const float4*>(w + kk * 128 + n0 + 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                #pragma unroll
                for (int i = 0; i < 8; i++)
                    #pragma unroll
                    for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    if (n0 < N) {
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const float bj = bias ? __ldg(bias + n0 + j) : 0.f;
            float v[8];
            #pragma unroll
            for (int i = 0; i < 8; i++) {
                v[i] = acc[i][j] + bj;
                if (relu) v[i] = fmaxf(v[i], 0.f);
            }
            float* o = out + (size_t)(n0 + j) * TILE_S + s0;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
}

// Narrow layers (N = 64 or 32: the torso nets).  dense_tile's 8 x 8 register tile leaves (128 - N) / 128 of the block idle; here the 256
// threads are re-tiled as (128 / SPT sample groups) x (N / 8 output groups) with SPT = N / 16 samples per thread (4 for N = 64, 2 for
// N = 32), so every thread works.  Each (sample, output) pair still accumulates over k in order: results are bit-identical to dense_tile.
template <int SPT>
__device__ __forceinline__ void dense_tile_narrow(const float* __restrict__ in, int K, const float* __restrict__ Wt, int ldw,
                                                  float* __restrict__ out, const float* __restrict__ bias, bool relu,
                                                  float* __restrict__ wstage) {
    constexpr int N = SPT * 16;                         // output width served by 256 threads
    constexpr int OG = N / 8;                           // output groups of 8
    const int tid = threadIdx.x;
    const int ty = tid / OG, tx = tid % OG;
    const int s0 = ty * SPT, n0 = tx * 8;
    float acc[SPT][8];
    #pragma unroll
    for (int i = 0; i < SPT; i++)
        #pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    const int nchunks = (K + DENSE_KC - 1) / DENSE_KC;
    constexpr int vec_per_row = N >> 2;
    auto stage = [&](int c, int buf) {
        const int k0 = c * DENSE_KC;
        constexpr int total = DENSE_KC * vec_per_row;
        for (int i = tid; i < total; i += DENSE_THREADS) {
            const int kk = i / vec_per_row, v = i - kk * vec_per_row;
            float* dst = wstage + (buf * DENSE_KC + kk) * 128 + v * 4;
            if (k0 + kk < K) cp_async16(dst, Wt + (size_t)(k0 + kk) * ldw + v * 4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };
    stage(0, 0);
    for (int c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) { stage(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* w = wstage + ((c & 1) * DENSE_KC) * 128;
        const int k0 = c * DENSE_KC;
        const int kmax = min(DENSE_KC, K - k0);
        #pragma unroll 4
        for (int kk = 0; kk < kmax; kk++) {
            float a[SPT];
            #pragma unroll
            for (int i = 0; i < SPT; i++) a[i] = in[(size_t)(k0 + kk) * TILE_S + s0 + i];
            const float4 b0 = *reinterpret_cast<const float4*>(w + kk * 128 + n0);
            const float4 b1 = *reinterpret_cast<const float4*>(w + kk * 128 + n0 + 4);
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            #pragma unroll
            for (int i = 0; i < SPT; i++)
                #pragma unroll
                for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    #pragma unroll
    for (int j = 0; j < 8; j++) {
        const float bj = bias ? __ldg(bias + n0 + j) : 0.f;
        #pragma unroll
        for (int i = 0; i < SPT; i++) {
            float v = acc[i][j] + bj;
            if (relu) v = fmaxf(v, 0.f);
            out[(size_t)(n0 + j) * TILE_S + s0 + i] = v;
        }
    }
    __syncthreads();
}

// Tiny output layers (N <= 4): out[j][s] = sum_k in[k][s] * W[j][k], W row-major [N][K] in global.
// Threads 0..127 take the even k, 128..255 the odd k; partial sums meet in `scratch` (shared [4][128]).
__device__ __forceinline__ void dense_small(const float* __restrict__ in, int K, const float* __restrict__ W, int N,
                                            float* __restrict__ out, float* __restrict__ scratch) {
    const int tid = threadIdx.x, s = tid & 127, half = tid >> 7;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = half; k < K; k += 2) {
        const float a = in[(size_t)k * TILE_S + s];
        #pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < N) acc[j] = fmaf(a, __ldg(W + (size_t)j * K + k), acc[j]);
    }
    if (half == 1) {
        #pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < N) scratch[j * TILE_S + s] = acc[j];
    }
    __syncthreads();
    if (half == 0) {
        #pragma unroll
        for (int j = 0; j < 4; j++)
            if (j < N) out[j * TILE_S + s] = acc[j] + scratch[j * TILE_S + s];
    }
    __syncthreads();
}

}  // namespace gf
