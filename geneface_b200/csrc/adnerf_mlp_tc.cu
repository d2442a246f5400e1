// libgfrender: the vanilla AD-NeRF backbone (modules/nerfs/adnerf/backbone.py:82-135: 8 x hid density trunk with the input re-injected
// after layer 4, 1 density output, 3 x hid/2 colour head on [trunk, view embedding], 3 colour outputs) on tcgen05 tensor cores.
//
// Round 1 ran these layers as fp32 library GEMMs (39 ms per 64x64x(64+192) frame).  Here every layer is one launch of ONE persistent,
// warp-specialised tcgen05 kernel (k_dense_tc) computing   out = act(A1 @ W1^T [+ A2 @ W2^T] + bias)   over 128-sample tiles:
//
//   warp 0      TMA producer : weights of the layer once per CTA (<= 160 KB, resident), then the tile's activation K-chunks
//                              (128 rows x 64 fp16 = 16 KB each, ONE contiguous cp.async.bulk per chunk) into a shared-memory ring
//   warp 1      MMA issuer   : 4 x tcgen05.mma (K = 16) per chunk into one of TWO accumulator buffers in tensor memory (2 x 256 columns);
//                              tcgen05.commit releases the ring slot / publishes the accumulator
//   warps 2..5  epilogue     : tcgen05.ld -> + bias -> ReLU -> fp16 -> next layer's activation tile in HBM (or fp32 raw sigma / rgb columns);
//                              runs on the tile that just finished while the MMA warp is already in the next one
//
// Activations travel between layers as fp16 in OUR tile-major layout: [tile][k-chunk][128 rows x 128 B, 16-byte units XOR-swizzled by
// row & 7] = exactly the shared-memory image a K-major SWIZZLE_128B UMMA operand needs, so a chunk is staged by one linear bulk copy (no
// tensor map) and written by the producing layer's epilogue.  The frequency embeddings of the sample positions (63 -> 64 columns) and of
// the view direction (27 -> 64) are written in the same layout by k_adnerf_embed_tiles and enter layers 0 / 5 and the first colour layer
// as an extra K-chunk; the per-frame condition vector enters layers 0 and 5 through their bias (b + W[:, cond] cond), the density output
// rides on the first colour layer as row hid/2 (N = hid/2 + 16).  Arithmetic: fp16 operands, fp32 accumulation, fp32 bias.
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "gf_tc.cuh"

namespace gf {

constexpr int DT_THREADS = 192;
constexpr uint32_t DT_CHUNK = 128 * 128;                 // one K-chunk of one tile: 128 rows x 64 fp16
constexpr int DT_MAX_SLOTS = 8;
constexpr uint32_t DT_SMEM_LIMIT = 232448;               // 227 KB

struct DenseArgs {
    const uint8_t* w_img;       // fp16 weight image: (a1_chunks + a2_chunks) chunks of [N rows x 128 B], SW128
    const float* bias;          // [N] fp32 or null
    const uint8_t* a1;          // activation tiles, a1_chunks x 16 KB per tile
    const uint8_t* a2;          // second operand source (embedding tiles, 1 chunk per tile) or null
    uint32_t a1_chunks, a2_chunks;
    uint8_t* out;               // fp16 ReLU output tiles (relu_cols / 64 chunks per tile) or null
    uint32_t relu_cols;         // accumulator columns [0, relu_cols) -> ReLU -> fp16 -> out
    float* raw;                 // [M, 4] fp32 raw network output or null
    uint32_t raw_src_col, raw_cols, raw_dst_col;   // accumulator columns [raw_src_col, +raw_cols) (+ bias, no activation) -> raw[i*4 + raw_dst_col + j]
    uint32_t M, N, nslot;
};

template <int DUMMY>
__global__ void __launch_bounds__(DT_THREADS, 1) k_dense_tc(const DenseArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nk = a.a1_chunks + a.a2_chunks;
    const uint32_t wchunk = a.N * 128;
    const uint32_t W_OFF = 0, A_OFF = nk * wchunk, BIAS_OFF = A_OFF + a.nslot * DT_CHUNK, BAR_OFF = BIAS_OFF + 1024;
    // barriers: [0] weights, [1 .. nslot] a_full, [1+nslot .. 2 nslot] a_empty, then d_full[2], d_empty[2]
    const uint32_t bar_w = sbase + BAR_OFF, bar_afull = bar_w + 8, bar_aempty = bar_afull + 8 * a.nslot, bar_dfull = bar_aempty + 8 * a.nslot,
                   bar_dempty = bar_dfull + 16, tmem_slot = bar_dempty + 16;
    float* bias = reinterpret_cast<float*>(smem + BIAS_OFF);
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
    for (uint32_t i = tid; i < 256; i += DT_THREADS) bias[i] = (a.bias && i < a.N) ? a.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // warp-uniform copies (see field_tc_split.cu): with per-thread values every tcgen05.mma went through an ELECT / R2UR broadcast loop
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(smem + (tmem_slot - sbase)), 0);
    const uint32_t warp_u = __shfl_sync(0xffffffffu, warp, 0);

    if (warp_u == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (elect_one_sync()) {
            mbar_expect_tx(bar_w, nk * wchunk);
            for (uint32_t c = 0; c < nk; c++) bulk_g2s(sbase + W_OFF + c * wchunk, a.w_img + (size_t)c * wchunk, wchunk, bar_w);
            uint32_t it = 0;
            for (uint32_t j = 0; j < my_tiles; j++) {
                const size_t tile = blockIdx.x + (size_t)j * gridDim.x;
                for (uint32_t c = 0; c < nk; c++, it++) {
                    const uint32_t slot = it % a.nslot, n = it / a.nslot;
                    mbar_wait(bar_aempty + 8 * slot, (n & 1) ^ 1);
                    const uint8_t* src = c < a.a1_chunks ? a.a1 + (tile * a.a1_chunks + c) * DT_CHUNK : a.a2 + (tile * a.a2_chunks + (c - a.a1_chunks)) * DT_CHUNK;
                    mbar_expect_tx(bar_afull + 8 * slot, DT_CHUNK);
                    bulk_g2s(sbase + A_OFF + slot * DT_CHUNK, src, DT_CHUNK, bar_afull + 8 * slot);
                }
            }
        }
    } else if (warp_u == 1) {
        // ---------------------------------------------------------------- MMA issuer (one elected lane; elect.sync directly after the uniform test)
        if (elect_one_sync()) {
            mbar_wait(bar_w, 0);
            const uint32_t idesc = idesc_f16(a.N);
            uint32_t it = 0;
            for (uint32_t j = 0; j < my_tiles; j++) {
                const uint32_t buf = j & 1;
                mbar_wait(bar_dempty + 8 * buf, ((j >> 1) & 1) ^ 1);        // the epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d = tmem_base + buf * 256;
                for (uint32_t c = 0; c < nk; c++, it++) {
                    const uint32_t slot = it % a.nslot, n = it / a.nslot;
                    mbar_wait(bar_afull + 8 * slot, n & 1);
                    tc_fence_after();
                    const uint32_t a_addr = sbase + A_OFF + slot * DT_CHUNK, w_addr = sbase + W_OFF + c * wchunk;
                    #pragma unroll
                    for (int k = 0; k < 4; k++) mma_ss(d, smem_desc(a_addr + 32 * k), smem_desc(w_addr + 32 * k), idesc, (c | k) ? 1 : 0);
                    mma_commit(bar_aempty + 8 * slot);                      // slot reusable once these MMAs have read it
                }
                mma_commit(bar_dfull + 8 * buf);                            // accumulator complete
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (4 warps = 128 TMEM lanes = 128 tile rows)
        const uint32_t q = warp & 3, row = q * 32 + lane;
        for (uint32_t j = 0; j < my_tiles; j++) {
            const uint32_t buf = j & 1;
            const size_t tile = blockIdx.x + (size_t)j * gridDim.x;
            mbar_wait(bar_dfull + 8 * buf, (j >> 1) & 1);
            tc_fence_after();
            const uint32_t t_d = tmem_base + ((q * 32) << 16) + buf * 256;
            const uint32_t out_chunks = a.relu_cols >> 6;
            for (uint32_t g = 0; g < (a.relu_cols >> 5); g++) {
                float v[32];
                tmem_ld32(t_d + 32 * g, v);
                uint32_t p[16];
                #pragma unroll
                for (int i = 0; i < 16; i++) p[i] = pack_relu_h2(v[2 * i] + bias[32 * g + 2 * i], v[2 * i + 1] + bias[32 * g + 2 * i + 1]);
                uint8_t* dst = a.out + (tile * out_chunks + (g >> 1)) * DT_CHUNK;
                #pragma unroll
                for (int u = 0; u < 4; u++)
                    *reinterpret_cast<uint4*>(dst + sw128(row, (g & 1) * 4 + u)) = make_uint4(p[4 * u], p[4 * u + 1], p[4 * u + 2], p[4 * u + 3]);
            }
            if (a.raw) {
                float r4[4];
                tmem_ld4(t_d + a.raw_src_col, r4);
                const size_t i = tile * 128 + row;
                if (i < a.M) {
                    for (uint32_t c = 0; c < a.raw_cols; c++) a.raw[4 * i + a.raw_dst_col + c] = r4[c] + bias[a.raw_src_col + c];
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

// --------------------------------------------------------------------------------------------------------------------------------------
// embeddings in tile layout: P tile = frequency embedding of the sample position (3 + 2*3*Lp = 63 of 64 columns), V tile = embedding of the
// ray's unit view direction (3 + 2*3*Lv = 27 of 64 columns).  Same [x, sin(2^k x), cos(2^k x)]_k order as commons/embedders.py:5-45.
__global__ void k_adnerf_embed_tiles(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z,
                                     const float* __restrict__ viewdirs, uint32_t R, uint32_t S, uint32_t Lp, uint32_t Lv,
                                     uint8_t* __restrict__ P, uint8_t* __restrict__ V) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t Mpad = ((R * S + 127) / 128) * 128;
    if (i >= Mpad) return;
    const uint32_t tile = i >> 7, row = i & 127;
    __align__(16) __half e[64];
    #pragma unroll
    for (int k = 0; k < 64; k++) e[k] = __float2half_rn(0.f);
    const bool valid = i < R * S;
    const uint32_t r = valid ? i / S : 0;
    if (valid) {
        const float zz = z[i];
        float p[3];
        #pragma unroll
        for (int c = 0; c < 3; c++) { p[c] = rays_o[3 * (size_t)r + c] + rays_d[3 * (size_t)r + c] * zz; e[c] = __float2half_rn(p[c]); }
        float f = 1.0f;
        for (uint32_t k = 0; k < Lp; k++, f *= 2.0f) {
            #pragma unroll
            for (int c = 0; c < 3; c++) {
                float s, co;
                sincosf(p[c] * f, &s, &co);
                e[3 + 6 * k + c] = __float2half_rn(s);
                e[3 + 6 * k + 3 + c] = __float2half_rn(co);
            }
        }
    }
    uint8_t* dst = P + (size_t)tile * DT_CHUNK;
    #pragma unroll
    for (int u = 0; u < 8; u++) *reinterpret_cast<uint4*>(dst + sw128(row, u)) = *reinterpret_cast<const uint4*>(&e[8 * u]);
    #pragma unroll
    for (int k = 0; k < 64; k++) e[k] = __float2half_rn(0.f);
    if (valid) {
        float d[3];
        #pragma unroll
        for (int c = 0; c < 3; c++) { d[c] = viewdirs[3 * (size_t)r + c]; e[c] = __float2half_rn(d[c]); }
        float f = 1.0f;
        for (uint32_t k = 0; k < Lv; k++, f *= 2.0f) {
            #pragma unroll
            for (int c = 0; c < 3; c++) {
                float s, co;
                sincosf(d[c] * f, &s, &co);
                e[3 + 6 * k + c] = __float2half_rn(s);
                e[3 + 6 * k + 3 + c] = __float2half_rn(co);
            }
        }
    }
    dst = V + (size_t)tile * DT_CHUNK;
    #pragma unroll
    for (int u = 0; u < 8; u++) *reinterpret_cast<uint4*>(dst + sw128(row, u)) = *reinterpret_cast<const uint4*>(&e[8 * u]);
}

// weight block -> fp16 SW128 image: rows n < Npad (zero beyond N), columns k < 64 * chunks (zero beyond K), source W[n * ldw + col0 + k]
__global__ void k_pack_dense(const float* __restrict__ W, uint32_t ldw, uint32_t col0, uint32_t K, uint32_t N, uint32_t row0, uint32_t Npad,
                             uint32_t chunks, uint8_t* __restrict__ img) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t kk = chunks * 64;
    if (t >= N * kk) return;
    const uint32_t n = t / kk, k = t % kk;
    const float v = k < K ? W[(size_t)n * ldw + col0 + k] : 0.f;
    const uint32_t rown = row0 + n;
    *reinterpret_cast<__half*>(img + (size_t)(k >> 6) * Npad * 128 + sw128(rown, (k & 63) >> 3) + (k & 7) * 2) = __float2half_rn(v);
}

// per-frame biases of the two layers that see the condition vector: out[l][n] = b_l[n] + sum_c Wc_l[n][c] cond[c]
__global__ void k_adnerf_bias_fold(const float* __restrict__ wc0, const float* __restrict__ b0, const float* __restrict__ wc5,
                                   const float* __restrict__ b5, const float* __restrict__ cond, uint32_t H, uint32_t C, float* __restrict__ out) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= 2 * H) return;
    const uint32_t l = n / H, r = n % H;
    const float* w = (l ? wc5 : wc0) + (size_t)r * C;
    float acc = (l ? b5 : b0)[r];
    for (uint32_t c = 0; c < C; c++) acc = fmaf(w[c], cond[c], acc);
    out[n] = acc;
}

__global__ void k_copy_cols(const float* __restrict__ W, uint32_t ldw, uint32_t col0, uint32_t rows, uint32_t cols, float* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    out[t] = W[(size_t)(t / cols) * ldw + col0 + t % cols];
}

}  // namespace gf

// ======================================================================================================================================
// C ABI
// ======================================================================================================================================
using namespace gf;

struct GfAdnerfLayer {
    size_t w_off;                   // into the image blob
    size_t b_off;                   // into the bias blob (floats); (size_t)-1: per-frame folded bias (slot 0 / 1 of the workspace)
    int fold_slot;
    uint32_t N, a1_chunks, a2_kind; // a2_kind: 0 none, 1 position tiles, 2 view tiles
    uint32_t a1_src;                // 0: position tiles, 1: activations
    uint32_t relu_cols, raw_src_col, raw_cols, raw_dst_col;
};

struct GfAdnerfMlp {
    uint32_t hid, cond_dim, Lp, Lv;
    uint8_t* img;
    float* fblob;                   // biases [12][256] | Wc0 [hid][cond] | b0 [hid] | Wc5 [hid][cond] | b5 [hid]
    size_t wc0, b0, wc5, b5;
    GfAdnerfLayer layer[12];
    int num_sms;
};

static uint32_t dense_smem_bytes(uint32_t N, uint32_t nk, uint32_t* nslot_out) {
    const uint32_t w = nk * N * 128;
    uint32_t fixed = 1024 /*alignment*/ + w + 1024 /*bias*/ + 256 /*barriers*/;
    uint32_t nslot = (DT_SMEM_LIMIT - fixed) / DT_CHUNK;
    if (nslot > DT_MAX_SLOTS) nslot = DT_MAX_SLOTS;
    *nslot_out = nslot;
    return fixed + nslot * DT_CHUNK;
}

extern "C" {

GF_API int gf_adnerf_mlp_create(const GfAdnerfDesc* d, GfAdnerfMlp** out, gf_stream_t stream) {
    GF_REQUIRE(d && out, "adnerf_mlp_create: null pointer");
    GF_REQUIRE(d->hid == 128 || d->hid == 256, "adnerf_mlp_create: hidden size must be 128 or 256");
    GF_REQUIRE(d->pos_multires >= 1 && d->pos_multires <= 10 && d->view_multires >= 1 && d->view_multires <= 10, "adnerf_mlp_create: multires must be in [1,10]");
    GF_REQUIRE(d->cond_dim >= 1 && d->cond_dim <= 1024, "adnerf_mlp_create: cond_dim out of range");
    for (int i = 0; i < 8; i++) GF_REQUIRE(d->dens_w[i] && d->dens_b[i], "adnerf_mlp_create: null density layer");
    for (int i = 0; i < 3; i++) GF_REQUIRE(d->col_w[i] && d->col_b[i], "adnerf_mlp_create: null colour layer");
    GF_REQUIRE(d->dens_out_w && d->dens_out_b && d->col_out_w && d->col_out_b, "adnerf_mlp_create: null output layer");
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t H = d->hid, Hc = H / 2, C = d->cond_dim;
    const uint32_t PD = 3 + 6 * d->pos_multires, VD = 3 + 6 * d->view_multires;      // 63, 27
    const uint32_t din = PD + C;
    GfAdnerfMlp* m = new GfAdnerfMlp();
    memset(m, 0, sizeof(*m));
    m->hid = H; m->cond_dim = C; m->Lp = d->pos_multires; m->Lv = d->view_multires;
    // ---- layer table ----
    const uint32_t hc = H / 64, cc = Hc / 64 ? Hc / 64 : 1;
    size_t woff = 0;
    auto add = [&](int l, uint32_t N, uint32_t a1_src, uint32_t a1_chunks, uint32_t a2_kind, uint32_t relu_cols, uint32_t rs, uint32_t rc, uint32_t rd, int fold) {
        GfAdnerfLayer& L = m->layer[l];
        L.N = N; L.a1_src = a1_src; L.a1_chunks = a1_chunks; L.a2_kind = a2_kind; L.relu_cols = relu_cols;
        L.raw_src_col = rs; L.raw_cols = rc; L.raw_dst_col = rd; L.fold_slot = fold;
        L.w_off = woff; L.b_off = (size_t)l * 256;
        woff += (size_t)(a1_chunks + (a2_kind ? 1 : 0)) * N * 128;
    };
    add(0, H, 0, 1, 0, H, 0, 0, 0, 0);
    for (int l = 1; l <= 4; l++) add(l, H, 1, hc, 0, H, 0, 0, 0, -1);
    add(5, H, 1, hc, 1, H, 0, 0, 0, 1);
    add(6, H, 1, hc, 0, H, 0, 0, 0, -1);
    add(7, H, 1, hc, 0, H, 0, 0, 0, -1);
    add(8, Hc + 16, 1, hc, 2, Hc, Hc, 1, 3, -1);                  // first colour layer + density output row; raw[..., 3] = sigma
    add(9, Hc, 1, cc, 0, Hc, 0, 0, 0, -1);
    add(10, Hc, 1, cc, 0, Hc, 0, 0, 0, -1);
    add(11, 16, 1, cc, 0, 0, 0, 3, 0, -1);                        // colour output: raw[..., 0:3]
    GF_REQUIRE(Hc % 64 == 0, "adnerf_mlp_create: hid/2 must be a multiple of 64");
    const size_t fl = (size_t)12 * 256 + 2 * ((size_t)H * C + H);
    if (cudaMalloc(&m->img, woff) != cudaSuccess || cudaMalloc(&m->fblob, fl * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        if (m->img) cudaFree(m->img);
        delete m;
        set_error("adnerf_mlp_create: cudaMalloc failed");
        return GF_ERR_CUDA;
    }
    cudaMemsetAsync(m->img, 0, woff, st);
    cudaMemsetAsync(m->fblob, 0, fl * sizeof(float), st);
    m->wc0 = (size_t)12 * 256; m->b0 = m->wc0 + (size_t)H * C; m->wc5 = m->b0 + H; m->b5 = m->wc5 + (size_t)H * C;
    auto pack = [&](const float* W, uint32_t ldw, uint32_t col0, uint32_t K, uint32_t N, uint32_t row0, uint32_t Npad, uint32_t chunks, size_t off) {
        const uint32_t total = N * chunks * 64;
        k_pack_dense<<<(total + 255) / 256, 256, 0, st>>>(W, ldw, col0, K, N, row0, Npad, chunks, m->img + off);
    };
    auto bias = [&](int l, const float* b, uint32_t n, uint32_t dst0) { cudaMemcpyAsync(m->fblob + (size_t)l * 256 + dst0, b, n * sizeof(float), cudaMemcpyDeviceToDevice, st); };
    // density trunk (backbone.py:99-117): layer 0 sees [pos, cond]; layer 5 sees [pos, cond, h]
    pack(d->dens_w[0], din, 0, PD, H, 0, H, 1, m->layer[0].w_off);
    for (int l = 1; l <= 7; l++) {
        if (l == 5) {
            pack(d->dens_w[5], din + H, din, H, H, 0, H, hc, m->layer[5].w_off);                            // h part  (K chunks 0 .. hc-1)
            pack(d->dens_w[5], din + H, 0, PD, H, 0, H, 1, m->layer[5].w_off + (size_t)hc * H * 128);       // pos part (last chunk)
        } else {
            pack(d->dens_w[l], H, 0, H, H, 0, H, hc, m->layer[l].w_off);
            bias(l, d->dens_b[l], H, 0);
        }
    }
    k_copy_cols<<<(H * C + 255) / 256, 256, 0, st>>>(d->dens_w[0], din, PD, H, C, m->fblob + m->wc0);
    k_copy_cols<<<(H * C + 255) / 256, 256, 0, st>>>(d->dens_w[5], din + H, PD, H, C, m->fblob + m->wc5);
    cudaMemcpyAsync(m->fblob + m->b0, d->dens_b[0], H * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(m->fblob + m->b5, d->dens_b[5], H * sizeof(float), cudaMemcpyDeviceToDevice, st);
    // first colour layer on [h, view] (backbone.py:121-126) + the density output (:119) as row Hc
    {
        const uint32_t N8 = Hc + 16;
        pack(d->col_w[0], H + VD, 0, H, Hc, 0, N8, hc, m->layer[8].w_off);
        pack(d->dens_out_w, H, 0, H, 1, Hc, N8, hc, m->layer[8].w_off);
        pack(d->col_w[0], H + VD, H, VD, Hc, 0, N8, 1, m->layer[8].w_off + (size_t)hc * N8 * 128);
        bias(8, d->col_b[0], Hc, 0);
        bias(8, d->dens_out_b, 1, Hc);
    }
    pack(d->col_w[1], Hc, 0, Hc, Hc, 0, Hc, cc, m->layer[9].w_off);  bias(9, d->col_b[1], Hc, 0);
    pack(d->col_w[2], Hc, 0, Hc, Hc, 0, Hc, cc, m->layer[10].w_off); bias(10, d->col_b[2], Hc, 0);
    pack(d->col_out_w, Hc, 0, Hc, 3, 0, 16, cc, m->layer[11].w_off); bias(11, d->col_out_b, 3, 0);
    if (cudaFuncSetAttribute(k_dense_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DT_SMEM_LIMIT) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(m->img); cudaFree(m->fblob); delete m;
        set_error("adnerf_mlp_create: cannot reserve dynamic shared memory");
        return GF_ERR_CUDA;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (m->num_sms <= 0) m->num_sms = 148;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
        cudaFree(m->img); cudaFree(m->fblob); delete m;
        set_error("adnerf_mlp_create: %s", cudaGetErrorString(e));
        return GF_ERR_CUDA;
    }
    *out = m;
    return GF_OK;
}

GF_API void gf_adnerf_mlp_destroy(GfAdnerfMlp* m) {
    if (!m) return;
    if (m->img) cudaFree(m->img);
    if (m->fblob) cudaFree(m->fblob);
    delete m;
}

// workspace: folded biases (2 x 256 floats) | position tiles | view tiles | activations ping | activations pong
GF_API uint64_t gf_adnerf_mlp_workspace_bytes(const GfAdnerfMlp* m, uint32_t n_samples) {
    if (!m) return 0;
    const uint64_t tiles = ((uint64_t)n_samples + 127) / 128;
    return 4096 + tiles * DT_CHUNK * (2 + 2 * (uint64_t)(m->hid / 64));
}

// raw[R, S, 4] = backbone(embed(rays_o + rays_d z), cond, embed(viewdirs))   (volume_rendering.py:153-155 + backbone.py:99-135)
GF_API int gf_adnerf_mlp_forward(const GfAdnerfMlp* m, const float* rays_o, const float* rays_d, const float* z_vals, const float* viewdirs,
                                 const float* cond, uint32_t R, uint32_t S, float* raw, void* workspace, uint64_t workspace_bytes,
                                 gf_stream_t stream) {
    GF_REQUIRE(m && rays_o && rays_d && z_vals && viewdirs && cond && raw && workspace, "adnerf_mlp_forward: null pointer");
    GF_REQUIRE((uint64_t)R * S < (1ull << 31), "adnerf_mlp_forward: too many samples");
    const uint32_t M = R * S;
    if (M == 0) return GF_OK;
    GF_REQUIRE(((uintptr_t)workspace & 1023) == 0, "adnerf_mlp_forward: workspace must be 1024-byte aligned");
    GF_REQUIRE(workspace_bytes >= gf_adnerf_mlp_workspace_bytes(m, M), "adnerf_mlp_forward: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t tiles = ((uint64_t)M + 127) / 128;
    const uint32_t H = m->hid, hc = H / 64;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* fold = reinterpret_cast<float*>(ws);
    uint8_t* P = ws + 4096;
    uint8_t* V = P + tiles * DT_CHUNK;
    uint8_t* act[2] = {V + tiles * DT_CHUNK, V + tiles * DT_CHUNK + tiles * DT_CHUNK * hc};
    k_adnerf_bias_fold<<<(2 * H + 127) / 128, 128, 0, st>>>(m->fblob + m->wc0, m->fblob + m->b0, m->fblob + m->wc5, m->fblob + m->b5, cond, H, m->cond_dim, fold);
    k_adnerf_embed_tiles<<<(uint32_t)((tiles * 128 + 127) / 128), 128, 0, st>>>(rays_o, rays_d, z_vals, viewdirs, R, S, m->Lp, m->Lv, P, V);
    int rc = check_launch("adnerf_mlp_forward(embed)");
    if (rc) return rc;
    int cur = 0;
    for (int l = 0; l < 12; l++) {
        const GfAdnerfLayer& L = m->layer[l];
        DenseArgs a;
        memset(&a, 0, sizeof(a));
        a.w_img = m->img + L.w_off;
        a.bias = L.fold_slot >= 0 ? fold + (size_t)L.fold_slot * H : m->fblob + L.b_off;
        a.a1 = L.a1_src == 0 ? P : act[cur];
        a.a1_chunks = L.a1_chunks;
        a.a2 = L.a2_kind == 1 ? P : (L.a2_kind == 2 ? V : nullptr);
        a.a2_chunks = L.a2_kind ? 1 : 0;
        a.out = L.relu_cols ? act[cur ^ 1] : nullptr;
        a.relu_cols = L.relu_cols;
        a.raw = L.raw_cols ? raw : nullptr;
        a.raw_src_col = L.raw_src_col; a.raw_cols = L.raw_cols; a.raw_dst_col = L.raw_dst_col;
        a.M = M; a.N = L.N;
        const uint32_t smem = dense_smem_bytes(L.N, a.a1_chunks + a.a2_chunks, &a.nslot);
        const uint32_t grid = tiles < (uint64_t)m->num_sms ? (uint32_t)tiles : (uint32_t)m->num_sms;
        k_dense_tc<0><<<grid, DT_THREADS, smem, st>>>(a);
        if (L.relu_cols) cur ^= 1;
    }
    return check_launch("adnerf_mlp_forward(layers)");
}

}  // extern "C"
