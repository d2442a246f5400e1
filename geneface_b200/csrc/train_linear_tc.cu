// libgfrender: tensor-core linear layers for the TRAINING step: forward, data gradient and weight gradient of the
// reference's bias-free MLPs (modules/radnerfs/cond_encoder.py:92-111: Linear(bias=False) -> ReLU -> ... -> Linear) on tcgen05, replacing the
// library GEMMs autograd ran for them (fp32 SIMT sgemm was 23-34 % of a step).  Arithmetic = the reference's own
// training arithmetic under `amp: true` (fp16 operands, fp32 accumulation, fp32 master weights and gradients).
//
// Everything works on 128-sample tiles in the tile-major fp16 layout of adnerf_mlp_tc.cu: [tile][64-column chunk][128 rows x 128 B, 16-byte units
// XOR-swizzled by row & 7] = the shared-memory image of a SWIZZLE_128B UMMA operand, staged by one linear cp.async.bulk per chunk.  The SAME bytes
// serve all three products -- only the descriptors change:
//
//   forward   Y  = X  W^T    A = X tile   (K-major: rows = samples, 128 B along features), B = weight image [n rows][k] (K-major)
//   dgrad     dX = dY W      A = dY tile  (K-major),                                         B = the SAME weight image read MN-major (contraction along its rows)
//   wgrad     dW = dY^T X    A = dY tiles read MN-major (M = 128 features, K = samples),     B = X tiles read MN-major; accumulated over all of a CTA's
//                                                                                            tiles in tensor memory, then one fp32 reduction per entry
//
// MN-major SWIZZLE_128B operand (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): one K index = one 128-byte row of 64
// consecutive MN elements, 8 rows = one 1024-byte swizzle atom, SBO = stride between 8-row groups along K (1024 B), LBO = stride between 64-element
// atoms along MN (our chunk stride).
//
//   k_tl_pack    fp32 / fp16 rows [M][ld] (x optional device scale) -> tiles          k_tl_wimg   fp32 W[N][K] -> fp16 image
//   k_tl_gemm    forward / dgrad over all tiles (persistent, TMA producer / MMA issuer / 4 epilogue warps, two accumulator buffers):
//                epilogue = [x ReLU mask of a saved activation] -> [ReLU] -> fp16 tiles and / or fp32 rows (x optional device scale)
//   k_tl_wgrad   weight gradient
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "gf_tc.cuh"

namespace gf {

constexpr int TL_THREADS = 192;
constexpr uint32_t TL_CHUNK = 128 * 128;
constexpr uint32_t TL_SMEM_LIMIT = 232448;

// MN-major SWIZZLE_128B descriptor: start address, LBO (bytes between 64-element atoms along MN), SBO = 1024 (8-row groups along K)
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, int swap = 0) {
    const uint64_t lbo = (lbo_bytes >> 4) & 0x3FFF, sbo = 64;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((swap ? sbo : lbo) << 16) | ((swap ? lbo : sbo) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor with operand majorness bits (15: A is MN-major, 16: B is MN-major)
__host__ __device__ constexpr uint32_t idesc_f16_t(uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------------------------- pack
// rows -> tiles.  One thread per (row, 16-byte unit) of the column range [col0, col1) of the tiles (col0, col1 multiples of 8): columns
// col0 .. col0 + K - 1 come from src[r * ld + (col - col0)] (ld = 0: one row broadcast to every sample), the rest of the range is zero.  Several calls
// with adjacent ranges assemble a concatenated input without materialising it.  src_f16: source is __half; scale: optional device scalar.
__global__ void k_tl_pack(const void* __restrict__ src, int src_f16, uint32_t ld, uint32_t K, uint32_t M, uint32_t chunks, uint32_t col0, uint32_t col1,
                          const float* __restrict__ scale, uint8_t* __restrict__ tiles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ntiles = (M + 127) / 128, units = (col1 - col0) >> 3;
    if (t >= ntiles * 128 * units) return;
    const uint32_t r = t / units, u = (col0 >> 3) + t % units, tile = r >> 7, row = r & 127, c = u >> 3, uu = u & 7;
    const float s = scale ? *scale : 1.0f;
    __align__(16) __half h[8];
    #pragma unroll
    for (int e = 0; e < 8; e++) {
        const uint32_t col = u * 8 + e - col0;
        float v = 0.f;
        if (r < M && col < K) v = src_f16 ? __half2float(reinterpret_cast<const __half*>(src)[(size_t)r * ld + col]) : reinterpret_cast<const float*>(src)[(size_t)r * ld + col];
        h[e] = __float2half_rn(v * s);
    }
    *reinterpret_cast<uint4*>(tiles + ((size_t)tile * chunks + c) * TL_CHUNK + sw128(row, uu)) = *reinterpret_cast<const uint4*>(h);
}

// W[N][K] fp32 -> image: `chunks` blocks of [rows_pad x 128 B]; rows >= N and columns >= K are zero
__global__ void k_tl_wimg(const float* __restrict__ W, uint32_t N, uint32_t K, uint32_t rows_pad, uint32_t chunks, uint8_t* __restrict__ img) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows_pad * chunks * 64) return;
    const uint32_t n = t / (chunks * 64), k = t % (chunks * 64);
    const float v = (n < N && k < K) ? W[(size_t)n * K + k] : 0.f;
    *reinterpret_cast<__half*>(img + (size_t)(k >> 6) * rows_pad * 128 + sw128(n, (k & 63) >> 3) + (k & 7) * 2) = __float2half_rn(v);
}

// ---------------------------------------------------------------------------------------------------------------------------------- gemm
struct TlGemmArgs {
    const uint8_t* w_img;
    uint32_t w_rows, w_chunks;      // image: w_chunks blocks of [w_rows x 128 B]
    int dgrad;                      // 0: D = A W^T (N = w_rows, contraction over the image's columns); 1: D = A W (N = 64 w_chunks, contraction over its rows)
    const uint8_t* a;
    uint32_t a_chunks;              // chunks per A tile (forward: == w_chunks; dgrad: ceil(w_rows / 64))
    uint8_t* out;                   // fp16 tiles or null
    uint32_t out_chunks;
    int relu;
    const uint8_t* mask;            // saved activation tiles (same shape as out / the fp32 rows): result zeroed where the activation is <= 0
    uint32_t mask_chunks;
    float* out_f32;                 // fp32 rows [M][ld_f32], columns [0, n_f32) or null
    uint32_t ld_f32, n_f32;
    const float* out_scale;         // device scalar multiplied into the fp32 rows or null
    uint32_t M, nslot;
    int mn_swap;                    // diagnostics (GF_TL_MN_SWAP=1): exchange the LBO / SBO fields of the MN-major descriptors
};

__global__ void __launch_bounds__(TL_THREADS, 1) k_tl_gemm(const TlGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t wbytes = a.w_chunks * a.w_rows * 128;
    const uint32_t W_OFF = 0, A_OFF = (wbytes + 1023) & ~1023u, BAR_OFF = A_OFF + a.nslot * TL_CHUNK;
    const uint32_t bar_w = sbase + BAR_OFF, bar_afull = bar_w + 8, bar_aempty = bar_afull + 8 * a.nslot, bar_dfull = bar_aempty + 8 * a.nslot,
                   bar_dempty = bar_dfull + 16, tmem_slot = bar_dempty + 16;
    const uint32_t N = a.dgrad ? 64 * a.w_chunks : a.w_rows;
    const uint32_t ksteps = a.dgrad ? a.w_rows / 16 : 4 * a.w_chunks;
    const uint32_t num_tiles = (a.M + 127) / 128;
    const uint32_t my_tiles = num_tiles > blockIdx.x ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (uint32_t s = 0; s < a.nslot; s++) { mbar_init(bar_afull + 8 * s, 1); mbar_init(bar_aempty + 8 * s, 1); }
        mbar_init(bar_dfull, 1); mbar_init(bar_dfull + 8, 1);
        mbar_init(bar_dempty, 128); mbar_init(bar_dempty + 8, 128);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(smem + (tmem_slot - sbase)), 0);
    const uint32_t warp_u = __shfl_sync(0xffffffffu, warp, 0);

    if (warp_u == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one_sync()) {
            mbar_expect_tx(bar_w, wbytes);
            for (uint32_t c = 0; c < a.w_chunks; c++) bulk_g2s(sbase + W_OFF + c * a.w_rows * 128, a.w_img + (size_t)c * a.w_rows * 128, a.w_rows * 128, bar_w);
            uint32_t it = 0;
            for (uint32_t j = 0; j < my_tiles; j++) {
                const size_t tile = blockIdx.x + (size_t)j * gridDim.x;
                for (uint32_t c = 0; c < a.a_chunks; c++, it++) {
                    const uint32_t slot = it % a.nslot, n = it / a.nslot;
                    mbar_wait(bar_aempty + 8 * slot, (n & 1) ^ 1);
                    mbar_expect_tx(bar_afull + 8 * slot, TL_CHUNK);
                    bulk_g2s(sbase + A_OFF + slot * TL_CHUNK, a.a + (tile * a.a_chunks + c) * TL_CHUNK, TL_CHUNK, bar_afull + 8 * slot);
                }
            }
        }
    } else if (warp_u == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (elect_one_sync()) {
            mbar_wait(bar_w, 0);
            const uint32_t idesc = idesc_f16_t(N, 0, a.dgrad ? 1 : 0);
            const uint32_t w_addr = sbase + W_OFF, wchunk = a.w_rows * 128;
            uint32_t it = 0;
            for (uint32_t j = 0; j < my_tiles; j++) {
                const uint32_t buf = j & 1;
                mbar_wait(bar_dempty + 8 * buf, ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * 256;
                for (uint32_t c = 0; c < a.a_chunks; c++, it++) {
                    const uint32_t slot = it % a.nslot, n = it / a.nslot;
                    mbar_wait(bar_afull + 8 * slot, n & 1);
                    tc_fence_after();
                    const uint32_t a_addr = sbase + A_OFF + slot * TL_CHUNK;
                    #pragma unroll 1
                    for (uint32_t k = 0; k < 4; k++) {
                        const uint32_t ks = 4 * c + k;
                        if (ks >= ksteps) break;
                        // forward: B = rows n of image chunk c, columns 16k.. (K-major).  dgrad: B = image rows 16 ks .. 16 ks + 15 (the contraction
                        // index) of ALL column chunks: MN-major, 64-element atoms `wchunk` bytes apart
                        const uint64_t bdesc = a.dgrad ? smem_desc_mn(w_addr + ks * 2048, wchunk, a.mn_swap) : smem_desc(w_addr + c * wchunk + 32 * k);
                        mma_ss(d, smem_desc(a_addr + 32 * k), bdesc, idesc, ks ? 1 : 0);
                    }
                    mma_commit(bar_aempty + 8 * slot);
                }
                mma_commit(bar_dfull + 8 * buf);
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (4 warps = 128 TMEM lanes = 128 tile rows)
        const uint32_t q = warp & 3, row = q * 32 + lane;
        const float oscale = a.out_scale ? *a.out_scale : 1.0f;
        for (uint32_t j = 0; j < my_tiles; j++) {
            const uint32_t buf = j & 1;
            const size_t tile = blockIdx.x + (size_t)j * gridDim.x;
            const size_t i = tile * 128 + row;
            mbar_wait(bar_dfull + 8 * buf, (j >> 1) & 1);
            tc_fence_after();
            const uint32_t t_d = tmem_base + ((q * 32) << 16) + buf * 256;
            const uint32_t groups = (N + 31) / 32;
            for (uint32_t g = 0; g < groups; g++) {
                float v[32];
                if (32 * g + 32 <= N) tmem_ld32(t_d + 32 * g, v);
                else {                                   // N is a multiple of 16: the last group may hold 16 columns
                    float v4[4];
                    #pragma unroll
                    for (int e = 0; e < 32; e++) v[e] = 0.f;
                    #pragma unroll
                    for (int e = 0; e < 4; e++) { tmem_ld4(t_d + 32 * g + 4 * e, v4); v[4 * e] = v4[0]; v[4 * e + 1] = v4[1]; v[4 * e + 2] = v4[2]; v[4 * e + 3] = v4[3]; }
                }
                if (a.mask && (g >> 1) < a.mask_chunks) {
                    const uint8_t* mt = a.mask + (tile * a.mask_chunks + (g >> 1)) * TL_CHUNK;
                    #pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint4 m = *reinterpret_cast<const uint4*>(mt + sw128(row, (g & 1) * 4 + u));
                        const __half2* mh = reinterpret_cast<const __half2*>(&m);
                        #pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float2 f = __half22float2(mh[e]);
                            if (!(f.x > 0.f)) v[8 * u + 2 * e] = 0.f;
                            if (!(f.y > 0.f)) v[8 * u + 2 * e + 1] = 0.f;
                        }
                    }
                }
                if (a.relu) {
                    #pragma unroll
                    for (int e = 0; e < 32; e++) v[e] = fmaxf(v[e], 0.f);
                }
                if (a.out && (g >> 1) < a.out_chunks) {
                    uint32_t p[16];
                    #pragma unroll
                    for (int e = 0; e < 16; e++) p[e] = pack_h2(v[2 * e], v[2 * e + 1]);
                    uint8_t* dst = a.out + (tile * a.out_chunks + (g >> 1)) * TL_CHUNK;
                    #pragma unroll
                    for (int u = 0; u < 4; u++)
                        *reinterpret_cast<uint4*>(dst + sw128(row, (g & 1) * 4 + u)) = make_uint4(p[4 * u], p[4 * u + 1], p[4 * u + 2], p[4 * u + 3]);
                }
                if (a.out_f32 && i < a.M) {
                    // a thread owns one row: 16-byte stores when the row pitch allows it (n_f32 is then a multiple of 4 -- the host pads the pitch of
                    // odd-width outputs).  Scalar stores made the row-major outputs (32 lanes x 4 B, each in another row) the slowest launches of a step.
                    float* dst = a.out_f32 + i * a.ld_f32;
                    if ((a.ld_f32 & 3) == 0 && (a.n_f32 & 3) == 0) {
                        #pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            if (32 * g + e < a.n_f32)
                                *reinterpret_cast<float4*>(dst + 32 * g + e) = make_float4(v[e] * oscale, v[e + 1] * oscale, v[e + 2] * oscale, v[e + 3] * oscale);
                    } else {
                        #pragma unroll
                        for (int e = 0; e < 32; e++)
                            if (32 * g + e < a.n_f32) dst[32 * g + e] = v[e] * oscale;
                    }
                }
            }
            // out tiles wider than the accumulator (N = 16 / 144 rounded up to whole chunks): zero the rest so that later products read defined values
            if (a.out) {
                for (uint32_t g = groups; g < 2 * a.out_chunks; g++) {
                    uint8_t* dst = a.out + (tile * a.out_chunks + (g >> 1)) * TL_CHUNK;
                    #pragma unroll
                    for (int u = 0; u < 4; u++) *reinterpret_cast<uint4*>(dst + sw128(row, (g & 1) * 4 + u)) = make_uint4(0, 0, 0, 0);
                }
            }
            tc_fence_before();
            mbar_arrive(bar_dempty + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------------------- wgrad
struct TlWgradArgs {
    const uint8_t* p;               // M-side tiles: features [64 p_c0, 64 p_c0 + 128) are the product's 128 rows
    uint32_t p_chunks, p_c0;
    const uint8_t* q;               // N-side tiles: features [0, N)
    uint32_t q_chunks;
    uint32_t N;                     // multiple of 16, <= 256
    float* dw;                      // fp32, += (atomic): transposed == 0: dw[m * ld + n] (m < rows_m, n < cols_n); 1: dw[n * ld + m]
    uint32_t ld, rows_m, cols_n;
    int transposed;
    const float* scale;             // device scalar multiplied into the result or null
    uint32_t M;
    int mn_swap;
};

__global__ void __launch_bounds__(TL_THREADS, 1) k_tl_wgrad(const TlWgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t qn = (a.N + 63) / 64;                            // N-side chunks staged per tile
    const uint32_t slot_bytes = (2 + qn) * TL_CHUNK, nslot = 2;
    const uint32_t BAR_OFF = nslot * slot_bytes;
    const uint32_t bar_full = sbase + BAR_OFF, bar_empty = bar_full + 8 * nslot, bar_done = bar_empty + 8 * nslot, tmem_slot = bar_done + 8;
    const uint32_t num_tiles = (a.M + 127) / 128;
    const uint32_t my_tiles = num_tiles > blockIdx.x ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (uint32_t s = 0; s < nslot; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(smem + (tmem_slot - sbase)), 0);
    const uint32_t warp_u = __shfl_sync(0xffffffffu, warp, 0);
    if (my_tiles) {
        if (warp_u == 0) {
            if (elect_one_sync()) {
                for (uint32_t j = 0; j < my_tiles; j++) {
                    const size_t tile = blockIdx.x + (size_t)j * gridDim.x;
                    const uint32_t slot = j % nslot, n = j / nslot, dst = sbase + slot * slot_bytes;
                    mbar_wait(bar_empty + 8 * slot, (n & 1) ^ 1);
                    mbar_expect_tx(bar_full + 8 * slot, slot_bytes);
                    bulk_g2s(dst, a.p + (tile * a.p_chunks + a.p_c0) * TL_CHUNK, 2 * TL_CHUNK, bar_full + 8 * slot);
                    bulk_g2s(dst + 2 * TL_CHUNK, a.q + tile * a.q_chunks * TL_CHUNK, qn * TL_CHUNK, bar_full + 8 * slot);
                }
            }
        } else if (warp_u == 1) {
            if (elect_one_sync()) {
                const uint32_t idesc = idesc_f16_t(a.N, 1, 1);
                for (uint32_t j = 0; j < my_tiles; j++) {
                    const uint32_t slot = j % nslot, n = j / nslot, base = sbase + slot * slot_bytes;
                    mbar_wait(bar_full + 8 * slot, n & 1);
                    tc_fence_after();
                    #pragma unroll 1
                    for (uint32_t ks = 0; ks < 8; ks++)       // 16 samples per step: rows 16 ks .. of every chunk
                        mma_ss(tmem_base, smem_desc_mn(base + ks * 2048, TL_CHUNK, a.mn_swap), smem_desc_mn(base + 2 * TL_CHUNK + ks * 2048, TL_CHUNK, a.mn_swap), idesc, (j | ks) ? 1 : 0);
                    mma_commit(bar_empty + 8 * slot);
                }
                mma_commit(bar_done);
            }
        } else {
            const uint32_t q = warp & 3, m = q * 32 + lane;
            const float s = a.scale ? *a.scale : 1.0f;
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const uint32_t t_d = tmem_base + ((q * 32) << 16);
            for (uint32_t g = 0; g < a.N / 16; g++) {
                float v[16];
                #pragma unroll
                for (int e = 0; e < 4; e++) {
                    float v4[4];
                    tmem_ld4(t_d + 16 * g + 4 * e, v4);
                    v[4 * e] = v4[0]; v[4 * e + 1] = v4[1]; v[4 * e + 2] = v4[2]; v[4 * e + 3] = v4[3];
                }
                if (m < a.rows_m) {
                    #pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const uint32_t n = 16 * g + e;
                        if (n < a.cols_n) atomicAdd(a.transposed ? a.dw + (size_t)n * a.ld + m : a.dw + (size_t)m * a.ld + n, v[e] * s);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace gf

// ======================================================================================================================================
// C ABI
// ======================================================================================================================================
using namespace gf;

extern "C" {

// bytes of one tensor in tile layout: ceil(M / 128) tiles x chunks x 16 KB
GF_API size_t gf_tl_tiles_bytes(uint32_t M, uint32_t chunks) { return (size_t)((M + 127) / 128) * chunks * TL_CHUNK; }

// rows [M][ld] (fp32, or fp16 if src_f16; ld = 0 broadcasts one row) columns [0, K) -> columns [col0, col0 + K) of fp16 tiles with `chunks` 64-column
// chunks; the rest of [col0, col1) is zero filled (col1 = 0: up to the tile width).  col0, col1 multiples of 8.  Optional device scale.
GF_API int gf_tl_pack(const void* src, int src_f16, uint32_t ld, uint32_t K, uint32_t M, uint32_t chunks, uint32_t col0, uint32_t col1, const float* scale,
                      void* tiles, gf_stream_t stream) {
    GF_REQUIRE(src && tiles, "tl_pack: null pointer");
    if (!col1) col1 = 64 * chunks;
    GF_REQUIRE(chunks >= 1 && (col0 & 7) == 0 && (col1 & 7) == 0 && col0 + K <= col1 && col1 <= 64 * chunks, "tl_pack: bad column range");
    if (!M || col1 == col0) return GF_OK;
    const size_t total = (size_t)((M + 127) / 128) * 128 * ((col1 - col0) >> 3);
    k_tl_pack<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, src_f16, ld, K, M, chunks, col0, col1, scale, (uint8_t*)tiles);
    return check_launch("tl_pack");
}

// W [N][K] fp32 -> fp16 weight image: `chunks` blocks of [rows_pad x 128 B] (rows_pad: multiple of 16 >= N; 64 chunks >= K)
GF_API int gf_tl_weight_image(const float* W, uint32_t N, uint32_t K, uint32_t rows_pad, uint32_t chunks, void* img, gf_stream_t stream) {
    GF_REQUIRE(W && img, "tl_weight_image: null pointer");
    GF_REQUIRE(rows_pad % 16 == 0 && rows_pad >= N && rows_pad <= 256 && K <= 64 * chunks && chunks >= 1 && chunks <= 4, "tl_weight_image: bad shape");
    const uint32_t total = rows_pad * chunks * 64;
    k_tl_wimg<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, N, K, rows_pad, chunks, (uint8_t*)img);
    return check_launch("tl_weight_image");
}

static int g_tl_sms = 0;
static int tl_sms() {
    if (!g_tl_sms) {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        g_tl_sms = n > 0 ? n : 148;
    }
    return g_tl_sms;
}

// forward (dgrad = 0): D = A W^T with the image's rows as outputs; data gradient (dgrad = 1): D = A W with the image's columns as outputs.
// a: A tiles (a_chunks per tile).  Result -> fp16 tiles `out` (out_chunks per tile, optional ReLU) and / or fp32 rows out_f32 [M][ld_f32] columns [0, n_f32)
// (x *out_scale); `mask`: tiles (mask_chunks per tile) of the saved ReLU output this gradient flows back through, or NULL.
GF_API int gf_tl_gemm(const void* a, uint32_t a_chunks, const void* w_img, uint32_t w_rows, uint32_t w_chunks, int dgrad, uint32_t M, void* out,
                      uint32_t out_chunks, int relu, const void* mask, uint32_t mask_chunks, float* out_f32, uint32_t ld_f32, uint32_t n_f32,
                      const float* out_scale, gf_stream_t stream) {
    GF_REQUIRE(a && w_img, "tl_gemm: null pointer");
    GF_REQUIRE(w_rows % 16 == 0 && w_rows >= 16 && w_rows <= 256 && w_chunks >= 1 && w_chunks <= 4, "tl_gemm: bad weight image shape");
    GF_REQUIRE(dgrad ? a_chunks == (w_rows + 63) / 64 : a_chunks == w_chunks, "tl_gemm: A chunks do not match the contraction length");
    GF_REQUIRE(out || out_f32, "tl_gemm: no output");
    if (!M) return GF_OK;
    TlGemmArgs g;
    memset(&g, 0, sizeof(g));
    g.w_img = (const uint8_t*)w_img; g.w_rows = w_rows; g.w_chunks = w_chunks; g.dgrad = dgrad; g.a = (const uint8_t*)a; g.a_chunks = a_chunks;
    g.out = (uint8_t*)out; g.out_chunks = out_chunks; g.relu = relu; g.mask = (const uint8_t*)mask; g.mask_chunks = mask_chunks;
    g.out_f32 = out_f32; g.ld_f32 = ld_f32; g.n_f32 = n_f32; g.out_scale = out_scale; g.M = M;
    { const char* e = getenv("GF_TL_MN_SWAP"); g.mn_swap = e && e[0] == '1'; }
    const uint32_t wbytes = (w_chunks * w_rows * 128 + 1023) & ~1023u;
    uint32_t nslot = (TL_SMEM_LIMIT - 1024 - wbytes - 512) / TL_CHUNK;
    if (nslot > 8) nslot = 8;
    g.nslot = nslot;
    const uint32_t smem = 1024 + wbytes + nslot * TL_CHUNK + 512;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(k_tl_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_SMEM_LIMIT) != cudaSuccess) { cudaGetLastError(); set_error("tl_gemm: smem attribute"); return GF_ERR_CUDA; }
        attr = true;
    }
    const uint32_t tiles = (M + 127) / 128;
    const uint32_t grid = tiles < (uint32_t)tl_sms() ? tiles : (uint32_t)tl_sms();
    k_tl_gemm<<<grid, TL_THREADS, smem, (cudaStream_t)stream>>>(g);
    return check_launch("tl_gemm");
}

// weight gradient: dw += scale * P[:, 64 p_c0 : 64 p_c0 + 128]^T Q[:, 0:N]  (contraction over the M samples), P / Q in tile layout.
// transposed = 0: dw[m * ld + n]; 1: dw[n * ld + m]; only m < rows_m, n < cols_n are written.  dw must be zero-initialised by the caller.
GF_API int gf_tl_wgrad(const void* p, uint32_t p_chunks, uint32_t p_c0, const void* q, uint32_t q_chunks, uint32_t N, uint32_t M, float* dw, uint32_t ld,
                       uint32_t rows_m, uint32_t cols_n, int transposed, const float* scale, gf_stream_t stream) {
    GF_REQUIRE(p && q && dw, "tl_wgrad: null pointer");
    GF_REQUIRE(p_c0 + 2 <= p_chunks, "tl_wgrad: the M side needs 128 features");
    GF_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && (N + 63) / 64 <= q_chunks, "tl_wgrad: bad N");
    if (!M) return GF_OK;
    TlWgradArgs g;
    memset(&g, 0, sizeof(g));
    g.p = (const uint8_t*)p; g.p_chunks = p_chunks; g.p_c0 = p_c0; g.q = (const uint8_t*)q; g.q_chunks = q_chunks; g.N = N; g.dw = dw; g.ld = ld;
    g.rows_m = rows_m; g.cols_n = cols_n; g.transposed = transposed; g.scale = scale; g.M = M;
    { const char* e = getenv("GF_TL_MN_SWAP"); g.mn_swap = e && e[0] == '1'; }
    const uint32_t smem = 1024 + 2 * (2 + (N + 63) / 64) * TL_CHUNK + 256;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(k_tl_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_SMEM_LIMIT) != cudaSuccess) { cudaGetLastError(); set_error("tl_wgrad: smem attribute"); return GF_ERR_CUDA; }
        attr = true;
    }
    const uint32_t tiles = (M + 127) / 128;
    const uint32_t grid = tiles < (uint32_t)tl_sms() ? tiles : (uint32_t)tl_sms();
    k_tl_wgrad<<<grid, TL_THREADS, smem, (cudaStream_t)stream>>>(g);
    return check_launch("tl_wgrad");
}
}
