// libgfrender: tcgen05 / TMEM field kernel (precision = 1), sm_100a only.
//
// Evaluates the RAD-NeRF head field (radnerf.py:73-105) for a dense list of samples:
//   3D grid gather -> ambient MLP (32+cond -> 128 -> 128 -> 2, tanh) -> 2D grid gather
//   -> sigma MLP (64 -> 128 -> 128 -> 1+128) -> colour MLP (16 SH + 128 geo + 4 ind -> 128 -> 3)
// One CTA per SM, persistent, 512 threads = TWO independent tile streams of 256 threads.  A tile is 128 samples =
// the 128 TMEM lanes; every sample row is served by TWO threads (warps w and w+4 of a stream share TMEM lane
// quadrant w%4): they split the grid levels of the gathers and the accumulator columns of the epilogues, so 16
// warps keep the 4 schedulers busy while activations never leave TMEM/registers:
//   * gathered layer inputs (grid features, SH) are written as fp16 into a 128x64 K-major SWIZZLE_128B shared-memory
//     tile and consumed by tcgen05.mma in SS mode;
//   * hidden activations are read from the fp32 accumulator with tcgen05.ld (32x32b: lane = row), bias/ReLU'd,
//     packed to fp16 and written back to TMEM with tcgen05.st; the next layer's tcgen05.mma takes A from TMEM (TS);
//   * all weights (184 KB fp16, pre-swizzled at model-create time into the exact shared-memory image) are staged once
//     per CTA by TMA bulk copies (cp.async.bulk, mbarrier complete_tx);
//   * per-frame condition vector and individual code are folded into fp32 bias vectors; the linear pair
//     sigma_net.net[2] -> color_net.net[0] (no activation between them, radnerf.py:90-101) is pre-multiplied into one
//     128 -> 128(+sigma) layer.
// Precision: fp16 operands / fp32 accumulation everywhere EXCEPT the ambient branch, whose output is a coordinate
// into a 2048^2 grid (1e-4 of coordinate error = 0.2 cells): there both operands are split hi+lo (a = a_hi + a_lo,
// three MMAs a_hi*w_hi + a_lo*w_hi + a_hi*w_lo, ~22-bit operands) and the 128->2 output layer runs in fp32 on the
// CUDA cores straight from the accumulator.  Measured effect: sigma error drops ~30x (scripts/tc_error_model.py).
// MMA issue: one thread per stream, tcgen05.commit -> mbarrier; everyone else waits on the mbarrier.  While one
// stream is in a gather/epilogue phase the tensor core works on the other stream's layer.
#include <cuda_fp16.h>

#include <cstring>

#include <cstdlib>
#include <cstring>

#include "gf_model.cuh"
#include "gf_tc.cuh"

namespace gf {

// ------------------------------------------------------------------------------------------
// shared-memory image of the weights (bytes).  Every block is a [rows x 64 halfs] K-major tile
// with 128-byte rows, 16-byte units XOR-swizzled by (row & 7), 1024-byte aligned.
// ------------------------------------------------------------------------------------------
constexpr uint32_t WB_A0A = 0;                               // [128]: k 0..31 = Wa0_hi, k 32..63 = Wa0_hi (pairs with [F_hi | F_lo])
constexpr uint32_t WB_A0B = WB_A0A + 128 * 128;              // [128]: k 0..31 = Wa0_lo, k 32..47 = colour-L0 SH columns, rest 0
constexpr uint32_t WB_A1H = WB_A0B + 128 * 128;              // 2 chunks x [128]: Wa1_hi
constexpr uint32_t WB_A1L = WB_A1H + 2 * 128 * 128;          // 2 chunks x [128]: Wa1_lo
constexpr uint32_t WB_SIG0 = WB_A1L + 2 * 128 * 128;         // [128] k 0..63
constexpr uint32_t WB_SIG1 = WB_SIG0 + 128 * 128;            // 2 chunks x [128]
constexpr uint32_t WB_MRG = WB_SIG1 + 2 * 128 * 128;         // 2 chunks x [144]: rows 0..127 = W_c0[:,16:144] @ W_s2[1:], row 128 = W_s2[0]
constexpr uint32_t WB_COL1 = WB_MRG + 2 * 144 * 128;         // 2 chunks x [16]: rows 0..2 = colour layer 1
constexpr uint32_t WB_TOTAL = WB_COL1 + 2 * 16 * 128;
static_assert(WB_TOTAL == 188416, "weight image size");
static_assert(WB_SIG0 % 1024 == 0 && WB_MRG % 1024 == 0 && WB_COL1 % 1024 == 0, "1024-byte aligned blocks");

constexpr uint32_t SM_W = 0;
constexpr uint32_t SM_F = WB_TOTAL;                          // 2 streams x [128 rows x 128 B] feature tiles
constexpr uint32_t SM_BIAS = SM_F + 2 * 128 * 128;           // 2 x 128 floats: cond bias, individual-code bias
constexpr uint32_t SM_XCH = SM_BIAS + 2 * 128 * 4;           // 2 streams x 2 halves x 128 rows x float2: ambient-logit partial sums
constexpr uint32_t SM_BAR = SM_XCH + 2 * 2 * 128 * 8;        // mbarriers: [0] weights, [1], [2] stream MMA
constexpr uint32_t SM_TMEM = SM_BAR + 4 * 8;                 // tmem base address
constexpr uint32_t SM_TOTAL = SM_TMEM + 16;
constexpr uint32_t TC_SMEM_BYTES = SM_TOTAL + 1024;          // + slack to 1024-align the dynamic base
static_assert(TC_SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

// TMEM columns per stream (256): accumulator D at +0 (<=144 cols).  A operand (64 cols = 128 halfs):
// ambient phase (D is 128 wide): A_hi at +128, A_lo at +192;  later layers: A at +144.
constexpr uint32_t TM_STREAM = 256, TM_AHI = 128, TM_ALO = 192, TM_A = 144;
constexpr int TC_THREADS = 512;

// ------------------------------------------------------------------------------------------
// pack kernel: fp32 reference weights -> fp16 swizzled image (global), run once per model
// ------------------------------------------------------------------------------------------
struct TcPackSrc {
    const float *a0, *a1, *a2, *s0, *s1, *s2, *c0, *c1;
    int cond, ind, G;
};

__device__ __forceinline__ void put_half(uint8_t* img, uint32_t block, uint32_t rows, uint32_t n, uint32_t k, __half v) {
    const uint32_t chunk = k >> 6, kk = k & 63;
    const uint32_t off = block + chunk * rows * 128 + sw128(n, kk >> 3) + (kk & 7) * 2;
    *reinterpret_cast<__half*>(img + off) = v;
}
__device__ __forceinline__ __half hi_of(float v) { return __float2half_rn(v); }
__device__ __forceinline__ __half lo_of(float v) { return __float2half_rn(v - __half2float(__float2half_rn(v))); }

__global__ void k_tc_pack(TcPackSrc s, uint8_t* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (n, k) of a 144 x 128 index space
    const int n = i / 128, k = i % 128;
    if (n >= 144) return;
    const int a_in = 32 + s.cond, c_in = 16 + s.G + s.ind;
    if (n < 128) {
        if (k < 32) {
            const float w = s.a0[(size_t)n * a_in + k];
            put_half(img, WB_A0A, 128, n, k, hi_of(w));
            put_half(img, WB_A0A, 128, n, 32 + k, hi_of(w));
            put_half(img, WB_A0B, 128, n, k, lo_of(w));
        } else if (k < 48) put_half(img, WB_A0B, 128, n, k, hi_of(s.c0[(size_t)n * c_in + (k - 32)]));   // SH columns of colour layer 0
        else if (k < 64) put_half(img, WB_A0B, 128, n, k, hi_of(0.f));
        const float w1 = s.a1[(size_t)n * 128 + k];
        put_half(img, WB_A1H, 128, n, k, hi_of(w1));
        put_half(img, WB_A1L, 128, n, k, lo_of(w1));
        if (k < 64) put_half(img, WB_SIG0, 128, n, k, hi_of(s.s0[(size_t)n * 64 + k]));
        put_half(img, WB_SIG1, 128, n, k, hi_of(s.s1[(size_t)n * 128 + k]));
        // merged: sum_j W_c0[n][16 + j] * W_s2[1 + j][k]
        float acc = 0.f;
        for (int j = 0; j < s.G; j++) acc = fmaf(s.c0[(size_t)n * c_in + 16 + j], s.s2[(size_t)(1 + j) * 128 + k], acc);
        put_half(img, WB_MRG, 144, n, k, hi_of(acc));
    } else {
        put_half(img, WB_MRG, 144, n, k, hi_of(n == 128 ? s.s2[k] : 0.f));
    }
    if (n < 16) put_half(img, WB_COL1, 16, n, k, hi_of(n < 3 ? s.c1[(size_t)n * 128 + k] : 0.f));
}

// This thread's half (64 columns starting at 64*half) of the accumulator -> (+bias) -> ReLU -> fp16 -> A operand.
// SPLIT: also emit the fp16 residual into a second A region (hi + lo ~ 22-bit operand).
template <bool SPLIT>
__device__ __forceinline__ void epilogue_relu_to_A(uint32_t t_d, uint32_t t_a, uint32_t t_alo, uint32_t half, const float* __restrict__ bias_smem, float* dbg) {
    #pragma unroll 1
    for (int c = 0; c < 2; c++) {
        const int col = 64 * half + 32 * c;
        float v[32];
        tmem_ld32(t_d + col, v);
        if (dbg) {
            #pragma unroll
            for (int i = 0; i < 32; i++) dbg[col + i] = v[i];
        }
        #pragma unroll
        for (int i = 0; i < 32; i++) {
            if (bias_smem) v[i] += bias_smem[col + i];
            v[i] = fmaxf(v[i], 0.f);
        }
        uint32_t p[16];
        #pragma unroll
        for (int i = 0; i < 16; i++) p[i] = pack_h2(v[2 * i], v[2 * i + 1]);
        tmem_st16(t_a + (col >> 1), p);
        if (SPLIT) {
            #pragma unroll
            for (int i = 0; i < 16; i++) p[i] = pack_h2(h_resid(v[2 * i]), h_resid(v[2 * i + 1]));
            tmem_st16(t_alo + (col >> 1), p);
        }
    }
    tmem_wait_st();
}

// 3-D grid, levels 8*HALF .. 8*HALF+7 of one row: fp16 hi into F[row][k 16*HALF ..], fp16 residual into F[row][k 32+16*HALF ..]
template <int HALF>
__device__ __forceinline__ void gather3_half(const GridDesc& g, float ux, float uy, float uz, uint8_t* F, uint32_t row) {
    #pragma unroll
    for (int b = 0; b < 2; b++) {                       // 4 levels per batch: up to 32 gathers in flight
        float2 f[4];
        grid3_levels<4, true>(g, 8 * HALF + 4 * b, ux, uy, uz, f);
        *reinterpret_cast<uint4*>(F + sw128(row, 2 * HALF + b)) =
            make_uint4(pack_h2(f[0].x, f[0].y), pack_h2(f[1].x, f[1].y), pack_h2(f[2].x, f[2].y), pack_h2(f[3].x, f[3].y));
        *reinterpret_cast<uint4*>(F + sw128(row, 4 + 2 * HALF + b)) =
            make_uint4(pack_h2(h_resid(f[0].x), h_resid(f[0].y)), pack_h2(h_resid(f[1].x), h_resid(f[1].y)),
                       pack_h2(h_resid(f[2].x), h_resid(f[2].y)), pack_h2(h_resid(f[3].x), h_resid(f[3].y)));
    }
}

// 2-D ambient grid, levels 8*HALF .. 8*HALF+7 -> F[row][k 32+16*HALF ..]
template <int HALF>
__device__ __forceinline__ void gather2_half(const GridDesc& g, float vx, float vy, uint8_t* F, uint32_t row) {
    float2 f[8];
    grid2_levels<8>(g, 8 * HALF, vx, vy, f);                          // 32 gathers in flight
    #pragma unroll
    for (int u = 0; u < 2; u++)
        *reinterpret_cast<uint4*>(F + sw128(row, 4 + 2 * HALF + u)) =
            make_uint4(pack_h2(f[4 * u].x, f[4 * u].y), pack_h2(f[4 * u + 1].x, f[4 * u + 1].y),
                       pack_h2(f[4 * u + 2].x, f[4 * u + 2].y), pack_h2(f[4 * u + 3].x, f[4 * u + 3].y));
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
struct TcArgs {
    GridDesc pos, amb;
    float bound, inv2b;
    const uint8_t* wimg;        // WB_TOTAL bytes, global
    const float* bias_ind;      // [128] fp32 (packed fp32 blob, c_bind) or null
    float w_amb2[256];          // [2][128] fp32 ambient output layer, by value: FFMA reads it straight from the constant bank
    FieldTcIO io;
    float* dbg;                 // [9][128][144] floats or null: accumulators of tile 0 after each layer
};

// ambient output layer (128 -> 2) over this thread's 64 accumulator columns, fp32, weights from the constant bank
template <int HALF>
__device__ __forceinline__ void amb2_partial(const TcArgs& a, uint32_t t_d, float* dbg, float& s0, float& s1) {
    #pragma unroll
    for (int c = 0; c < 2; c++) {
        constexpr int base = 64 * HALF;
        const int col = base + 32 * c;
        float v[32];
        tmem_ld32(t_d + col, v);
        if (dbg) {
            #pragma unroll
            for (int j = 0; j < 32; j++) dbg[1 * 128 * 144 + col + j] = v[j];
        }
        #pragma unroll
        for (int j = 0; j < 32; j++) {
            const float r = fmaxf(v[j], 0.f);
            s0 = fmaf(r, a.w_amb2[base + 32 * c + j], s0);
            s1 = fmaf(r, a.w_amb2[128 + base + 32 * c + j], s1);
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_field_tc(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint32_t stream = tid >> 8, half = (tid >> 7) & 1, row = tid & 127;
    const uint32_t bar_w = sbase + SM_BAR, bar_s = sbase + SM_BAR + 8 * (1 + stream);
    float* bias_cond = reinterpret_cast<float*>(smem + SM_BIAS);
    float* bias_ind = bias_cond + 128;
    float2* xch = reinterpret_cast<float2*>(smem + SM_XCH) + stream * 256;     // [half][row]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
    const uint32_t M = a.io.M_dev ? *a.io.M_dev : a.io.M_host;

    // ---- one-time setup: barriers, TMEM, weights via TMA bulk copy, biases -------------------
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(sbase + SM_BAR + 8, 1);
        mbar_init(sbase + SM_BAR + 16, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(sbase + SM_TMEM, 512);
    if (tid < 128) bias_cond[tid] = a.io.bias_amb[tid];
    else if (tid < 256) bias_ind[tid - 128] = a.bias_ind ? a.bias_ind[tid - 128] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        mbar_expect_tx(bar_w, WB_TOTAL);
        const uint32_t cuts[7] = {0, WB_A1H, WB_A1L, WB_SIG0, WB_SIG1, WB_MRG, WB_TOTAL};   // six copies of <= 40 KB
        #pragma unroll
        for (int i = 0; i < 6; i++) bulk_g2s(sbase + SM_W + cuts[i], a.wimg + cuts[i], cuts[i + 1] - cuts[i], bar_w);
    }
    mbar_wait(bar_w, 0);

    const uint32_t t_lane = ((warp & 3) * 32) << 16;                       // this warp's TMEM lane quadrant
    const uint32_t t_d = tmem_base + t_lane + stream * TM_STREAM;            // accumulator (thread view)
    const uint32_t m_d = tmem_base + stream * TM_STREAM;                     // MMA view (lane 0)
    uint8_t* F = smem + SM_F + stream * (128 * 128);
    const uint32_t f_addr = sbase + SM_F + stream * (128 * 128);
    const uint32_t w_addr = sbase + SM_W;
    const bool leader = (tid & 255) == 0;
    const uint32_t bar_id = 1 + stream;
    uint32_t phase = 0;

    const uint32_t num_tiles = (M + 127) / 128;
    for (uint32_t tile = blockIdx.x * 2 + stream; tile < num_tiles; tile += gridDim.x * 2) {
        float* dbg = (a.dbg && tile == 0) ? a.dbg + (size_t)row * 144 : nullptr;
        const uint32_t i = tile * 128 + row;
        const bool valid = i < M;
        float x = 0.f, y = 0.f, z = 0.f, dx = 0.f, dy = 0.f, dz = 1.f;
        if (valid) {
            if (a.io.pos4) {
                const float4 p = a.io.pos4[i];
                x = p.x; y = p.y; z = p.z;
                if (half == 0) {
                    const int ray = __float_as_int(p.w);
                    dx = __ldg(a.io.rays_d + 3 * (size_t)ray); dy = __ldg(a.io.rays_d + 3 * (size_t)ray + 1); dz = __ldg(a.io.rays_d + 3 * (size_t)ray + 2);
                }
            } else {
                x = a.io.xyzs[3 * (size_t)i]; y = a.io.xyzs[3 * (size_t)i + 1]; z = a.io.xyzs[3 * (size_t)i + 2];
                dx = a.io.dirs[3 * (size_t)i]; dy = a.io.dirs[3 * (size_t)i + 1]; dz = a.io.dirs[3 * (size_t)i + 2];
            }
        }
        // ---- 3D grid: this thread takes levels 8*half .. 8*half+7 -> fp16 hi into F[row][k 16h..16h+15], residual into k 32+16h.. ------
        {
            // rows past the end of the list sample the centre (results are discarded)
            // (x + b) / (2b) as a multiply: exact for power-of-two bounds, <= 1 ulp otherwise (features are rounded to fp16 anyway)
            const float ux = valid ? (x + a.bound) * a.inv2b : 0.5f, uy = (y + a.bound) * a.inv2b, uz = (z + a.bound) * a.inv2b;
            // `half` is warp-uniform: two copies of the code so that every per-level constant is an immediate constant-bank operand
            if (half == 0) gather3_half<0>(a.pos, ux, uy, uz, F, row);
            else gather3_half<1>(a.pos, ux, uy, uz, F, row);
        }
        fence_async_smem();
        tc_fence_before();
        bar_stream(bar_id);
        // ---- ambient layer 0, split precision: [F_hi | F_lo] (K=64) x [W_hi | W_hi]  +  F_hi (K=32) x W_lo --------------------------------
        if (leader) {
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(m_d, smem_desc(f_addr + 32 * k), smem_desc(w_addr + WB_A0A + 32 * k), idesc_f16(128), k);
            #pragma unroll
            for (int k = 0; k < 2; k++) mma_ss(m_d, smem_desc(f_addr + 32 * k), smem_desc(w_addr + WB_A0B + 32 * k), idesc_f16(128), 1);
            mma_commit(bar_s);
        }
        mbar_wait(bar_s, phase); phase ^= 1;
        tc_fence_after();
        epilogue_relu_to_A<true>(t_d, t_d + TM_AHI, t_d + TM_ALO, half, bias_cond, dbg ? dbg + 0 * 128 * 144 : nullptr);
        tc_fence_before();
        bar_stream(bar_id);
        // ---- ambient layer 1, split precision (A from TMEM): A_hi W_hi + A_lo W_hi + A_hi W_lo ------------------------------------------------
        if (leader) {
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < 8; k++)
                mma_ts(m_d, m_d + TM_AHI + 8 * k, smem_desc(w_addr + WB_A1H + (k >> 2) * (128 * 128) + 32 * (k & 3)), idesc_f16(128), k);
            #pragma unroll
            for (int k = 0; k < 8; k++)
                mma_ts(m_d, m_d + TM_ALO + 8 * k, smem_desc(w_addr + WB_A1H + (k >> 2) * (128 * 128) + 32 * (k & 3)), idesc_f16(128), 1);
            #pragma unroll
            for (int k = 0; k < 8; k++)
                mma_ts(m_d, m_d + TM_AHI + 8 * k, smem_desc(w_addr + WB_A1L + (k >> 2) * (128 * 128) + 32 * (k & 3)), idesc_f16(128), 1);
            mma_commit(bar_s);
        }
        mbar_wait(bar_s, phase); phase ^= 1;
        tc_fence_after();
        // ---- ambient layer 2 (128 -> 2) in fp32 on the CUDA cores, straight from the accumulator; tanh -------------------------------------------
        float ax, ay;
        {
            float s0 = 0.f, s1 = 0.f;
            if (half == 0) amb2_partial<0>(a, t_d, dbg, s0, s1);
            else amb2_partial<1>(a, t_d, dbg, s0, s1);
            xch[half * 128 + row] = make_float2(s0, s1);
            tc_fence_before();
            bar_stream(bar_id);
            const float2 o = xch[(half ^ 1) * 128 + row];
            // fixed summation order (half 0 + half 1) so both threads of a row get identical coordinates
            const float l0 = half ? o.x + s0 : s0 + o.x, l1 = half ? o.y + s1 : s1 + o.y;
            if (dbg && half == 0) { dbg[2 * 128 * 144 + 0] = l0; dbg[2 * 128 * 144 + 1] = l1; }
            ax = tanhf(l0); ay = tanhf(l1);
        }
        // ---- 2D ambient grid: levels 8*half .. +7 -> F[row][k 32+16h .. 32+16h+15] -------------------------------------------------------------------
        {
            const float vx = (ax + 1.0f) * 0.5f, vy = (ay + 1.0f) * 0.5f;
            if (half == 0) gather2_half<0>(a.amb, vx, vy, F, row);
            else gather2_half<1>(a.amb, vx, vy, F, row);
        }
        fence_async_smem();
        tc_fence_before();
        bar_stream(bar_id);
        // ---- sigma layer 0: D = F[:, 0:64] @ Ws0^T -----------------------------------------------------------------------------------------------------
        if (leader) {
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < 4; k++) mma_ss(m_d, smem_desc(f_addr + 32 * k), smem_desc(w_addr + WB_SIG0 + 32 * k), idesc_f16(128), k);
            mma_commit(bar_s);
        }
        mbar_wait(bar_s, phase); phase ^= 1;
        tc_fence_after();
        epilogue_relu_to_A<false>(t_d, t_d + TM_A, 0, half, nullptr, dbg ? dbg + 3 * 128 * 144 : nullptr);
        tc_fence_before();
        bar_stream(bar_id);
        // ---- sigma layer 1 --------------------------------------------------------------------------------------------------------------------------------
        if (leader) {
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < 8; k++)
                mma_ts(m_d, m_d + TM_A + 8 * k, smem_desc(w_addr + WB_SIG1 + (k >> 2) * (128 * 128) + 32 * (k & 3)), idesc_f16(128), k);
            mma_commit(bar_s);
        }
        mbar_wait(bar_s, phase); phase ^= 1;
        tc_fence_after();
        epilogue_relu_to_A<false>(t_d, t_d + TM_A, 0, half, nullptr, dbg ? dbg + 4 * 128 * 144 : nullptr);
        // SH(dir) -> F[row][k 32..47] (the sigma-layer-0 MMA that read this tile has completed)
        if (half == 0) {
            float sh[16];
            sh4(dx, dy, dz, sh);
            uint32_t p[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) p[j] = pack_h2(sh[2 * j], sh[2 * j + 1]);
            *reinterpret_cast<uint4*>(F + sw128(row, 4)) = make_uint4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<uint4*>(F + sw128(row, 5)) = make_uint4(p[4], p[5], p[6], p[7]);
        }
        fence_async_smem();
        tc_fence_before();
        bar_stream(bar_id);
        // ---- merged sigma layer 2 x colour layer 0 (N = 144: cols 0..127 colour pre-activation, col 128 sigma logit),
        //      then += SH part (SS, K = 16, N = 128) ---------------------------------------------------------------------------------------------------------
        if (leader) {
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < 8; k++)
                mma_ts(m_d, m_d + TM_A + 8 * k, smem_desc(w_addr + WB_MRG + (k >> 2) * (144 * 128) + 32 * (k & 3)), idesc_f16(144), k);
            mma_ss(m_d, smem_desc(f_addr + 64), smem_desc(w_addr + WB_A0B + 64), idesc_f16(128), 1);
            mma_commit(bar_s);
        }
        mbar_wait(bar_s, phase); phase ^= 1;
        tc_fence_after();
        float sg[4] = {0.f, 0.f, 0.f, 0.f};
        if (half == 0) {
            tmem_ld4(t_d + 128, sg);
            if (dbg) dbg[5 * 128 * 144 + 128] = sg[0];
        }
        epilogue_relu_to_A<false>(t_d, t_d + TM_A, 0, half, bias_ind, dbg ? dbg + 5 * 128 * 144 : nullptr);
        tc_fence_before();
        bar_stream(bar_id);
        // ---- colour layer 1 (N = 16; 3 real outputs) -> sigmoid -------------------------------------------------------------------------------------------------
        if (leader) {
            tc_fence_after();
            #pragma unroll
            for (int k = 0; k < 8; k++)
                mma_ts(m_d, m_d + TM_A + 8 * k, smem_desc(w_addr + WB_COL1 + (k >> 2) * (16 * 128) + 32 * (k & 3)), idesc_f16(16), k);
            mma_commit(bar_s);
        }
        mbar_wait(bar_s, phase); phase ^= 1;
        tc_fence_after();
        if (half == 0) {
            float c[4];
            tmem_ld4(t_d, c);
            if (dbg) { dbg[6 * 128 * 144 + 0] = c[0]; dbg[6 * 128 * 144 + 1] = c[1]; dbg[6 * 128 * 144 + 2] = c[2]; }
            if (valid) {
                const float sigma = __expf(sg[0]);
                const float cr = __fdividef(1.0f, 1.0f + __expf(-c[0]));
                const float cg = __fdividef(1.0f, 1.0f + __expf(-c[1]));
                const float cb = __fdividef(1.0f, 1.0f + __expf(-c[2]));
                if (a.io.out4) a.io.out4[i] = make_float4(sigma, cr, cg, cb);
                if (a.io.sigmas) a.io.sigmas[i] = sigma;
                if (a.io.rgbs) { a.io.rgbs[3 * (size_t)i] = cr; a.io.rgbs[3 * (size_t)i + 1] = cg; a.io.rgbs[3 * (size_t)i + 2] = cb; }
                if (a.io.ambient) { a.io.ambient[2 * (size_t)i] = ax; a.io.ambient[2 * (size_t)i + 1] = ay; }
            }
        }
        tc_fence_before();   // order this tile's TMEM reads before the next tile's first MMA (issued after the next bar.sync)
    }
    // ---- teardown -----------------------------------------------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    if (a.io.stat_samples && blockIdx.x == 0 && tid == 0) atomicAdd(a.io.stat_samples, (unsigned long long)M);
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static int ensure_tc_pack(GfModel* m, cudaStream_t st) {
    if (m->tc_blob) return GF_OK;
    const GfModelDesc& d = m->desc;
    if (d.hidden_dim != 128 || d.geo_feat_dim != 128) {
        set_error("precision=1 (tcgen05) supports hidden_dim == 128 and geo_feat_dim == 128 only; use precision=0");
        return GF_ERR_UNSUPPORTED;
    }
    uint8_t* img = nullptr;
    if (cudaMalloc(&img, WB_TOTAL) != cudaSuccess) { cudaGetLastError(); set_error("tc pack: cudaMalloc failed"); return GF_ERR_CUDA; }
    cudaMemsetAsync(img, 0, WB_TOTAL, st);
    TcPackSrc s;
    s.a0 = d.ambient_w0; s.a1 = d.ambient_w1; s.a2 = d.ambient_w2; s.s0 = d.sigma_w0; s.s1 = d.sigma_w1; s.s2 = d.sigma_w2;
    s.c0 = d.color_w0; s.c1 = d.color_w1; s.cond = (int)d.cond_dim; s.ind = (int)d.ind_dim; s.G = (int)d.geo_feat_dim;
    k_tc_pack<<<(144 * 128 + 255) / 256, 256, 0, st>>>(s, img);
    int rc = check_launch("tc pack");
    if (rc) { cudaFree(img); return rc; }
    if (cudaFuncSetAttribute(k_field_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(img);
        set_error("tc pack: cannot reserve %u bytes of dynamic shared memory", TC_SMEM_BYTES);
        return GF_ERR_CUDA;
    }
    // host copy of the fp32 ambient output layer (travels by value in the kernel arguments)
    if (cudaMemcpyAsync(m->w_amb2_host, m->w + m->dev.a_w2, sizeof(float) * 256, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(img);
        set_error("tc pack: cannot read back the ambient output layer");
        return GF_ERR_CUDA;
    }
    m->tc_blob = img;
    m->tc_bytes = WB_TOTAL;
    return GF_OK;
}

int field_tc_split_launch(const GfModel* model, const FieldTcIO& io, cudaStream_t st);   // field_tc_split.cu

// GF_TC_MODE=fused selects the single fused kernel of this file; the default is the two-kernel warp-specialised pipeline.
static bool tc_mode_fused() {
    static const int mode = [] { const char* e = getenv("GF_TC_MODE"); return (e && !strcmp(e, "fused")) ? 1 : 0; }();
    return mode == 1;
}

// kernels one precision-1 field evaluation launches (gpu_launches bookkeeping of gf_render_frame)
int field_tc_kernel_count() { return tc_mode_fused() ? 1 : 2; }

int field_tc_launch(const GfModel* model, const FieldTcIO& io, cudaStream_t st) {
    if (!tc_mode_fused()) return field_tc_split_launch(model, io, st);
    GfModel* m = const_cast<GfModel*>(model);
    int rc = ensure_tc_pack(m, st);
    if (rc) return rc;
    TcArgs a;
    a.pos = model->dev.pos; a.amb = model->dev.amb; a.bound = model->dev.bound; a.inv2b = 0.5f / model->dev.bound;
    a.wimg = (const uint8_t*)m->tc_blob;
    a.bias_ind = model->dev.ind ? model->dev.w + model->dev.c_bind : nullptr;
    memcpy(a.w_amb2, m->w_amb2_host, sizeof(a.w_amb2));
    a.io = io;
    a.dbg = m->tc_dbg;
    uint32_t grid = (uint32_t)model->num_sms;
    if (!io.M_dev) {
        const uint32_t tiles = (io.M_host + 127) / 128;
        const uint32_t need = (tiles + 1) / 2;
        if (need < grid) grid = need ? need : 1;
    }
    k_field_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a);
    return check_launch("field_tc");
}

}  // namespace gf

extern "C" {
// Diagnostics: make the next precision-1 launches dump the fp32 accumulators of tile 0 after each MMA stage into dbg
// (device float[9*128*144]); pass NULL to switch it off.  Used by tests/test_parity_gpu.py.
GF_API int gf_tc_debug(GfModel* model, float* dbg) {
    if (!model) return GF_ERR_INVALID;
    model->tc_dbg = dbg;
    return GF_OK;
}
}
