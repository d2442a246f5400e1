// libgfrender: `_gridencoder`, `_shencoder`, `_freqencoder` operator sets (fine-grained boundary).
//
// Replaces
//   modules/radnerfs/encoders/gridencoder/src/gridencoder.cu  (gridencoder.h:11-14)
//   modules/radnerfs/encoders/shencoder/src/shencoder.cu      (shencoder.h:8-9)
//   modules/radnerfs/encoders/freqencoder/src/freqencoder.cu  (freqencoder.h:8-9)
//
// Grid encoder: same (sample, level) decomposition and the same fp32 rounding sequence as the
// reference kernel (gridencoder.cu:87-244) -- level geometry via exp2f, FFMA position, corner
// order idx = 0..2^D-1, FFMA accumulation -- so fp32 outputs are bit-identical; the C
// feature channels of a corner are fetched with ONE vector load (8 B for C=2, 16 B for C=4)
// instead of C scalar loads.  Backward scatters with vector (float2 / half2) reductions.
// SH: evaluated from the definition Y_l^m = (-1)^m K_l^m Q_l^m(z) {Re,Im}(x+iy)^m by recurrence
// (Q_l^m = d^m/dz^m P_l), which is what the reference's expanded polynomials
// (shencoder.cu:43-121) are; results agree to a few ulp.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "gf_common.cuh"

namespace gf {

// ======================================================================================
// grid encoder
// ======================================================================================
__device__ __forceinline__ uint32_t fast_hash(const uint32_t* pos_grid, int D) {
    // gridencoder.cu:50-63
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t result = 0;
    #pragma unroll
    for (int i = 0; i < 5; ++i)
        if (i < D) result ^= pos_grid[i] * primes[i];
    return result;
}

// gridencoder.cu:66-84.  The stride loop stops once stride > hashmap_size: in tiled mode the
// remaining dimensions are DROPPED from the index (SURVEY.md "tiled-grid index quirk").
template <int D>
__device__ __forceinline__ uint32_t grid_index(uint32_t gridtype, bool align_corners, uint32_t hashmap_size, uint32_t resolution,
                                               const uint32_t pos_grid[D]) {
    uint32_t stride = 1, index = 0;
    #pragma unroll
    for (int d = 0; d < D; d++) {
        if (stride <= hashmap_size) {
            index += pos_grid[d] * stride;
            stride *= align_corners ? resolution : (resolution + 1);
        }
    }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash(pos_grid, D);
    return index % hashmap_size;   // entry index (multiply by C for the element offset)
}

__device__ __forceinline__ void level_geometry(uint32_t level, float S, uint32_t H, float& scale, uint32_t& resolution) {
    // gridencoder.cu:138-139 : exp2f(level * S) * H - 1.0f  (nvcc contracts the tail to one FFMA)
    scale = __fmaf_rn(exp2f(__fmul_rn((float)level, S)), (float)H, -1.0f);
    resolution = (uint32_t)ceilf(scale) + 1;
}

// one vector load of the C channels of table entry `e`
template <typename T, int C>
__device__ __forceinline__ void load_entry(const T* __restrict__ grid, uint32_t e, float (&out)[C]) {
    if constexpr (std::is_same<T, float>::value) {
        if constexpr (C == 1) out[0] = __ldg(grid + e);
        else if constexpr (C == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(grid) + e); out[0] = v.x; out[1] = v.y; }
        else if constexpr (C == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(grid) + e); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
        else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)e);
            const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)e + 1);
            out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
        }
    } else {
        const __half* g = reinterpret_cast<const __half*>(grid) + (size_t)e * C;
        #pragma unroll
        for (int c = 0; c < C; c++) out[c] = __half2float(__ldg(g + c));
    }
}

template <typename T>
__device__ __forceinline__ T from_float(float x);
template <>
__device__ __forceinline__ float from_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ __half from_float<__half>(float x) { return __float2half_rn(x); }
__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(__half x) { return __half2float(x); }

// K13.  grid = (ceil(B/256), L); thread = one (sample, level).
// The 2^D corner entries are gathered ONCE into registers (one vector load each) and serve both the interpolated output and, when asked
// for, the input Jacobian dy_dx: the derivative along dimension g is the same multilinear sum over the other dimensions applied to the
// corner DIFFERENCES corner[idx | bit g] - corner[idx].  (The reference gathers the corners a second time for dy_dx, 2^(D-1) * 2 * D more
// loads per sample and level, gridencoder.cu:200-243.)  The fp32 operation order of both results is the reference's -- corner order
// idx = 0 .. 2^D - 1 with FFMA accumulation for the output; scale * prod(w_d) * diff * deriv summed over the sub-corners in ascending
// order for the Jacobian -- so the outputs stay bit-identical to it.
template <typename T, int D, int C>
__global__ void __launch_bounds__(256) k_grid_forward(const float* __restrict__ inputs, const T* __restrict__ grid_all,
                                                       const int* __restrict__ offsets, T* __restrict__ outputs, uint32_t B, uint32_t L,
                                                       float S, uint32_t H, T* __restrict__ dy_dx, uint32_t gridtype, bool align_corners,
                                                       uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    T* out = outputs + ((size_t)level * B + b) * C;
    T* jac = dy_dx ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;

    float frac[D], dfrac[D];           // interpolation weight of the upper corner per dimension, and its derivative w.r.t. the cell coordinate
    uint32_t cell[D];
    bool inside = true;
    float scale; uint32_t resolution;
    level_geometry(level, S, H, scale, resolution);
    #pragma unroll
    for (int d = 0; d < D; d++) {
        const float x = inputs[(size_t)b * D + d];
        inside = inside && x >= 0 && x <= 1;
        float p = __fmaf_rn(x, scale, align_corners ? 0.0f : 0.5f);
        cell[d] = (uint32_t)floorf(p);
        p = __fsub_rn(p, (float)cell[d]);
        if (interp == 1) { dfrac[d] = 6 * p * (1.0f - p); p = p * p * (3.0f - 2.0f * p); }     // smoothstep
        else dfrac[d] = 1.0f;
        frac[d] = p;
    }
    if (!inside) {                     // gridencoder.cu:110-135: samples outside [0,1]^D encode to zero
        #pragma unroll
        for (int c = 0; c < C; c++) out[c] = from_float<T>(0.f);
        if (jac) {
            #pragma unroll
            for (int i = 0; i < D * C; i++) jac[i] = from_float<T>(0.f);
        }
        return;
    }
    const T* table = grid_all + (size_t)(uint32_t)offsets[level] * C;
    const uint32_t entries = (uint32_t)(offsets[level + 1] - offsets[level]);

    float corner[1 << D][C];
    #pragma unroll
    for (int idx = 0; idx < (1 << D); idx++) {
        uint32_t pg[D];
        #pragma unroll
        for (int d = 0; d < D; d++) pg[d] = cell[d] + ((idx >> d) & 1);
        load_entry<T, C>(table, grid_index<D>(gridtype, align_corners, entries, resolution, pg), corner[idx]);
    }
    T acc[C];
    #pragma unroll
    for (int c = 0; c < C; c++) acc[c] = from_float<T>(0.f);
    #pragma unroll
    for (int idx = 0; idx < (1 << D); idx++) {
        float w = 1;
        #pragma unroll
        for (int d = 0; d < D; d++) w = __fmul_rn(w, ((idx >> d) & 1) ? frac[d] : __fsub_rn(1.0f, frac[d]));
        #pragma unroll
        for (int c = 0; c < C; c++) acc[c] = from_float<T>(__fmaf_rn(w, corner[idx][c], to_float(acc[c])));
    }
    if constexpr (std::is_same<T, float>::value && C == 2) {
        *reinterpret_cast<float2*>(out) = make_float2(acc[0], acc[1]);
    } else {
        #pragma unroll
        for (int c = 0; c < C; c++) out[c] = acc[c];
    }
    if (!jac) return;
    #pragma unroll
    for (int g = 0; g < D; g++) {
        T dg[C];
        #pragma unroll
        for (int c = 0; c < C; c++) dg[c] = from_float<T>(0.f);
        #pragma unroll
        for (int idx = 0; idx < (1 << D); idx++) {
            if (idx & (1 << g)) continue;                                   // lower corner along g; its partner is idx | (1 << g)
            float w = scale;
            #pragma unroll
            for (int d = 0; d < D; d++)
                if (d != g) w *= ((idx >> d) & 1) ? frac[d] : 1 - frac[d];
            #pragma unroll
            for (int c = 0; c < C; c++) {
                const float diff = corner[idx | (1 << g)][c] - corner[idx][c];
                if constexpr (std::is_same<T, float>::value) dg[c] += w * diff * dfrac[g];
                else dg[c] = from_float<T>(to_float(dg[c]) + w * to_float(from_float<T>(diff)) * dfrac[g]);
            }
        }
        #pragma unroll
        for (int c = 0; c < C; c++) jac[g * C + c] = dg[c];
    }
}

// K14, B200 form: level-major grid (blockIdx.y = level), one thread per sample carrying all C channels, grid-stride over the samples.
//   * SMALL DENSE LEVELS (table <= the CTA's shared-memory budget: 3-D levels 0-1 = 4,920 / 13,824 entries, 2-D levels 0-6 of the May
//     configuration) are where global atomics collide hardest: every sample of the batch lands in a few thousand entries (level 0: ~800
//     updates per entry per step at 0.5 M samples).  `priv_ctas` CTAs per such level each accumulate their share of the samples into a
//     PRIVATE shared-memory copy of the level (shared-memory reductions, no L2 round trip, no cross-SM contention) and flush only the
//     touched entries with vector reductions: <= priv_ctas * entries global reductions instead of 2^D * B.
//   * when EVERY lane of a warp targets the same entry -- samples of one ray inside one coarse cell, or ambient coordinates clustered in a
//     few cells of the 2-D grid: the dominant pattern behind the measured contention -- the warp tree-reduces the 32 contributions with
//     shuffles and commits once instead of serialising 32 same-address reductions;
//   * LARGER LEVELS OF THE 2-D GRIDS (ambient / torso: the coordinates are network outputs and cluster -- a whole batch can sit in a handful of
//     cells of every level, so the per-warp commits above still serialise on a few addresses; measured: 37 % of a 65,536-ray step): `cache_ctas`
//     CTAs per level accumulate through a direct-mapped shared-memory cache (8,192 slots: tag + C floats; a slot is claimed with one atomicCAS,
//     a conflicting entry goes straight to global memory) and flush the claimed slots once at the end;
//   * everything else is one 8-byte vector reduction per channel pair (RED.E.ADD.F32x2 / .F16x2) per corner.
// The sum order differs from the reference's (which is itself non-deterministic: atomics); parity is checked against an fp64
// re-accumulation.  Measured (profiles/r02_summary.md): 65,536-ray step 25.1 ms (plain reductions) -> 17.9 (privatised) -> 16.0 (+ warp
// aggregation); at 4,096 rays privatisation loses, so it is switched on from 131,072 samples.
constexpr uint32_t GRID_BWD_PRIV_BYTES = 13824 * 8;           // 3-D level 1 of the May configuration (C = 2): 110,592 B, 2 CTAs / SM
constexpr uint32_t GRID_BWD_CACHE_LOG2 = 13, GRID_BWD_CACHE_SLOTS = 1u << GRID_BWD_CACHE_LOG2;   // tags 32 KB + values C x 32 KB
template <int C, typename Commit>
__device__ __forceinline__ void grid_update(uint32_t e, float (&v)[C], Commit&& commit) {
    if (__activemask() == 0xffffffffu && __match_any_sync(0xffffffffu, e) == 0xffffffffu) {
        #pragma unroll
        for (int c = 0; c < C; c++) {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
        }
        if ((threadIdx.x & 31) == 0) commit(e, v);
    } else {
        commit(e, v);
    }
}

template <typename T, int C>
__device__ __forceinline__ void grid_reduce_global(T* __restrict__ grad_grid, uint32_t e, const float (&v)[C]) {
    T* dst = grad_grid + (size_t)e * C;
    if constexpr (C == 1) {
        if constexpr (std::is_same<T, float>::value) atomicAdd(dst, v[0]);
        else atomicAdd(reinterpret_cast<__half*>(dst), __float2half_rn(v[0]));
    } else {
        #pragma unroll
        for (int c = 0; c < C; c += 2) {
            if constexpr (std::is_same<T, float>::value) atomicAdd(reinterpret_cast<float2*>(dst + c), make_float2(v[c], v[c + 1]));
            else atomicAdd(reinterpret_cast<__half2*>(dst + c), __floats2half2_rn(v[c], v[c + 1]));
        }
    }
}

template <typename T, int D, int C>
__global__ void __launch_bounds__(256) k_grid_backward_b200(const T* __restrict__ grad, const float* __restrict__ inputs,
                                                             const int* __restrict__ offsets, T* __restrict__ grad_grid_all, uint32_t B,
                                                             uint32_t L, float S, uint32_t H, uint32_t gridtype, bool align_corners,
                                                             uint32_t interp, uint32_t priv_ctas, uint32_t priv_entries, uint32_t cache_ctas) {
    extern __shared__ float tab[];                        // private copy of a small level: [hashmap_size][C] fp32; or the cache: tags, then values
    const uint32_t level = blockIdx.y;
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    const bool priv = hashmap_size <= priv_entries;
    const bool cached = !priv && cache_ctas != 0;
    if (priv && blockIdx.x >= priv_ctas) return;
    if (cached && blockIdx.x >= cache_ctas) return;
    const uint32_t nctas = priv ? (priv_ctas < gridDim.x ? priv_ctas : gridDim.x) : cached ? (cache_ctas < gridDim.x ? cache_ctas : gridDim.x) : gridDim.x;
    uint32_t* ctag = reinterpret_cast<uint32_t*>(tab);
    float* cval = tab + GRID_BWD_CACHE_SLOTS;
    if (cached) {
        for (uint32_t e = threadIdx.x; e < GRID_BWD_CACHE_SLOTS; e += blockDim.x) {
            ctag[e] = 0xffffffffu;
            #pragma unroll
            for (int c = 0; c < C; c++) cval[(size_t)e * C + c] = 0.f;
        }
        __syncthreads();
    }
    T* grad_grid = grad_grid_all + (size_t)(uint32_t)offsets[level] * C;
    float scale; uint32_t resolution;
    level_geometry(level, S, H, scale, resolution);
    if (priv) {
        for (uint32_t e = threadIdx.x; e < hashmap_size * C; e += blockDim.x) tab[e] = 0.f;
        __syncthreads();
    }
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += nctas * blockDim.x) {
        const float* in = inputs + (size_t)b * D;
        float pos[D];
        uint32_t pos_grid[D];
        bool oob = false;
        #pragma unroll
        for (int d = 0; d < D; d++) {
            const float x = in[d];
            oob = oob || x < 0 || x > 1;
            pos[d] = __fmaf_rn(x, scale, align_corners ? 0.0f : 0.5f);
            pos_grid[d] = (uint32_t)floorf(pos[d]);
            pos[d] = __fsub_rn(pos[d], (float)pos_grid[d]);
            if (interp == 1) pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
        }
        if (oob) continue;                                  // gridencoder.cu:281-286: out-of-range inputs get no gradient
        const T* g = grad + (size_t)level * B * C + (size_t)b * C;
        float gc[C];
        #pragma unroll
        for (int c = 0; c < C; c++) gc[c] = to_float(g[c]);
        #pragma unroll
        for (int idx = 0; idx < (1 << D); idx++) {
            float w = 1;
            uint32_t pgl[D];
            #pragma unroll
            for (int d = 0; d < D; d++) {
                if ((idx & (1 << d)) == 0) { w *= 1 - pos[d]; pgl[d] = pos_grid[d]; }
                else { w *= pos[d]; pgl[d] = pos_grid[d] + 1; }
            }
            const uint32_t e = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pgl);
            float v[C];
            #pragma unroll
            for (int c = 0; c < C; c++) v[c] = w * gc[c];
            if (priv) {
                grid_update<C>(e, v, [&](uint32_t ee, const float (&vv)[C]) {
                    #pragma unroll
                    for (int c = 0; c < C; c++) atomicAdd(&tab[(size_t)ee * C + c], vv[c]);
                });
            } else if (cached) {
                grid_update<C>(e, v, [&](uint32_t ee, const float (&vv)[C]) {
                    const uint32_t slot = (ee * 2654435761u) >> (32 - GRID_BWD_CACHE_LOG2);
                    const uint32_t old = atomicCAS(&ctag[slot], 0xffffffffu, ee);
                    if (old == 0xffffffffu || old == ee) {
                        #pragma unroll
                        for (int c = 0; c < C; c++) atomicAdd(&cval[(size_t)slot * C + c], vv[c]);
                    } else {
                        grid_reduce_global<T, C>(grad_grid, ee, vv);
                    }
                });
            } else {
                grid_update<C>(e, v, [&](uint32_t ee, const float (&vv)[C]) { grid_reduce_global<T, C>(grad_grid, ee, vv); });
            }
        }
    }
    if (priv) {
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < hashmap_size; e += blockDim.x) {
            float v[C];
            bool any = false;
            #pragma unroll
            for (int c = 0; c < C; c++) { v[c] = tab[(size_t)e * C + c]; any = any || v[c] != 0.f; }
            if (any) grid_reduce_global<T, C>(grad_grid, e, v);
        }
    }
    if (cached) {
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < GRID_BWD_CACHE_SLOTS; e += blockDim.x) {
            const uint32_t ee = ctag[e];
            if (ee == 0xffffffffu) continue;
            float v[C];
            #pragma unroll
            for (int c = 0; c < C; c++) v[c] = cval[(size_t)e * C + c];
            grid_reduce_global<T, C>(grad_grid, ee, v);
        }
    }
}


// GF_GRID_BWD=plain forces plain vector reductions (no privatisation), GF_GRID_BWD=priv forces privatisation at any batch size: A/B runs
static int grid_bwd_mode() {
    static const int v = [] { const char* e = getenv("GF_GRID_BWD"); return !e ? 0 : (!strcmp(e, "plain") || !strcmp(e, "legacy")) ? 1 : !strcmp(e, "priv") ? 2 : !strcmp(e, "nocache") ? 3 : 0; }();
    return v;
}

// K15  gridencoder.cu:342-368
template <typename T, int D, int C>
__global__ void k_grid_input_backward(const T* __restrict__ grad, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs, uint32_t B,
                                      uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const T* dd = dy_dx + (size_t)b * L * D * C;
    T result = from_float<T>(0.f);
    for (uint32_t l = 0; l < L; l++) {
        #pragma unroll
        for (int c = 0; c < C; c++) {
            const float a = to_float(grad[(size_t)l * B * C + (size_t)b * C + c]), bb = to_float(dd[l * D * C + d * C + c]);
            if constexpr (std::is_same<T, float>::value) result = __fmaf_rn(a, bb, result);
            else result = from_float<T>(to_float(result) + to_float(from_float<T>(a * bb)));
        }
    }
    grad_inputs[t] = result;
}

// K16  gridencoder.cu:505-609 (float only)
template <int D, int C>
__global__ void k_grad_tv(const float* __restrict__ inputs, const float* __restrict__ grid_all, float* __restrict__ grad_all,
                          const int* __restrict__ offsets, float weight, uint32_t B, uint32_t L, float S, uint32_t H,
                          uint32_t gridtype, bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    const float* in = inputs + (size_t)b * D;
    const float* grid = grid_all + (size_t)(uint32_t)offsets[level] * C;
    float* grad = grad_all + (size_t)(uint32_t)offsets[level] * C;
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    float scale; uint32_t resolution;
    level_geometry(level, S, H, scale, resolution);
    uint32_t pos_grid[D];
    #pragma unroll
    for (int d = 0; d < D; d++) {
        const float x = in[d];
        if (x < 0 || x > 1) return;
        pos_grid[d] = (uint32_t)floorf(__fmaf_rn(x, scale, align_corners ? 0.0f : 0.5f));
    }
    float results[C], idelta[C], center[C];
    #pragma unroll
    for (int c = 0; c < C; c++) results[c] = idelta[c] = 0.f;
    const uint32_t e0 = grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pos_grid);
    load_entry<float, C>(grid, e0, center);
    const float w = weight / (2 * D);
    #pragma unroll
    for (int d = 0; d < D; d++) {
        const uint32_t cur = pos_grid[d];
        float nb[C];
        if (cur < resolution) {
            pos_grid[d] = cur + 1;
            load_entry<float, C>(grid, grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pos_grid), nb);
            #pragma unroll
            for (int c = 0; c < C; c++) { const float gv = center[c] - nb[c]; results[c] += gv; idelta[c] += gv * gv; }
        }
        if (cur > 0) {
            pos_grid[d] = cur - 1;
            load_entry<float, C>(grid, grid_index<D>(gridtype, align_corners, hashmap_size, resolution, pos_grid), nb);
            #pragma unroll
            for (int c = 0; c < C; c++) { const float gv = center[c] - nb[c]; results[c] += gv; idelta[c] += gv * gv; }
        }
        pos_grid[d] = cur;
    }
    #pragma unroll
    for (int c = 0; c < C; c++) atomicAdd(grad + (size_t)e0 * C + c, w * results[c] * rsqrtf(idelta[c] + 1e-9f));
}

// ---- dispatch ---------------------------------------------------------------------------
template <typename T, int D, int C>
static int launch_grid_forward(const float* inputs, const void* emb, const int* offsets, void* outputs, uint32_t B, uint32_t L, float S,
                               uint32_t H, void* dy_dx, uint32_t gridtype, bool ac, uint32_t interp, cudaStream_t st) {
    const dim3 grid(div_up(B, 256), L, 1);
    k_grid_forward<T, D, C><<<grid, 256, 0, st>>>(inputs, (const T*)emb, offsets, (T*)outputs, B, L, S, H, (T*)dy_dx, gridtype, ac, interp);
    return check_launch("grid_encode_forward");
}

template <typename T, int D, int C>
static int launch_grid_backward(const void* grad, const float* inputs, const int* offsets, void* grad_emb, uint32_t B, uint32_t L, float S,
                                uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype, bool ac, uint32_t interp,
                                cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_grid_backward_b200<T, D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRID_BWD_PRIV_BYTES);
        attr = true;
    }
    // Privatisation pays only when the batch is large enough that (a) contention on the small levels is real and (b) every privatising
    // CTA still sees thousands of samples (zero + flush cost one pass over the level's table each).
    const int mode = grid_bwd_mode();
    const bool use_priv = mode == 2 || ((mode == 0 || mode == 3) && B >= 131072);
    // the shared-memory cache for the larger levels: 2-D grids only (clustered network-output coordinates), batches large enough to fill the CTAs
    static_assert((1 + C) * GRID_BWD_CACHE_SLOTS * 4 <= GRID_BWD_PRIV_BYTES || C > 2, "cache does not fit");
    uint32_t cache_ctas = 0;
    if (D == 2 && C <= 2 && (mode == 2 || (mode == 0 && B >= 65536))) { cache_ctas = B / 2048; cache_ctas = cache_ctas < 8 ? 8 : (cache_ctas > 128 ? 128 : cache_ctas); }
    uint32_t priv_ctas = B / 4096;
    priv_ctas = priv_ctas < 8 ? 8 : (priv_ctas > 64 ? 64 : priv_ctas);
    uint32_t gx = div_up(B, 256 * 4);
    gx = gx < priv_ctas ? priv_ctas : (gx > 1024 ? 1024 : gx);
    if (gx < cache_ctas) gx = cache_ctas;
    k_grid_backward_b200<T, D, C><<<dim3(gx, L, 1), 256, GRID_BWD_PRIV_BYTES, st>>>(
        (const T*)grad, inputs, offsets, (T*)grad_emb, B, L, S, H, gridtype, ac, interp, priv_ctas,
        use_priv ? GRID_BWD_PRIV_BYTES / (uint32_t)(C * sizeof(float)) : 0u, cache_ctas);
    int rc = check_launch("grid_encode_backward");
    if (rc) return rc;
    if (dy_dx && grad_inputs) {
        k_grid_input_backward<T, D, C><<<div_up(B * D, 256), 256, 0, st>>>((const T*)grad, (const T*)dy_dx, (T*)grad_inputs, B, L);
        rc = check_launch("grid_encode_backward(inputs)");
    }
    return rc;
}

#define GF_DISPATCH_DC(FN, T, ...)                                                          \
    switch (D * 16 + C) {                                                                   \
        case 2 * 16 + 1: return FN<T, 2, 1>(__VA_ARGS__);                                   \
        case 2 * 16 + 2: return FN<T, 2, 2>(__VA_ARGS__);                                   \
        case 2 * 16 + 4: return FN<T, 2, 4>(__VA_ARGS__);                                   \
        case 2 * 16 + 8: return FN<T, 2, 8>(__VA_ARGS__);                                   \
        case 3 * 16 + 1: return FN<T, 3, 1>(__VA_ARGS__);                                   \
        case 3 * 16 + 2: return FN<T, 3, 2>(__VA_ARGS__);                                   \
        case 3 * 16 + 4: return FN<T, 3, 4>(__VA_ARGS__);                                   \
        case 3 * 16 + 8: return FN<T, 3, 8>(__VA_ARGS__);                                   \
        case 4 * 16 + 1: return FN<T, 4, 1>(__VA_ARGS__);                                   \
        case 4 * 16 + 2: return FN<T, 4, 2>(__VA_ARGS__);                                   \
        case 4 * 16 + 4: return FN<T, 4, 4>(__VA_ARGS__);                                   \
        case 4 * 16 + 8: return FN<T, 4, 8>(__VA_ARGS__);                                   \
        case 5 * 16 + 1: return FN<T, 5, 1>(__VA_ARGS__);                                   \
        case 5 * 16 + 2: return FN<T, 5, 2>(__VA_ARGS__);                                   \
        case 5 * 16 + 4: return FN<T, 5, 4>(__VA_ARGS__);                                   \
        case 5 * 16 + 8: return FN<T, 5, 8>(__VA_ARGS__);                                   \
        default: break;                                                                     \
    }

// ======================================================================================
// spherical harmonics
// ======================================================================================
struct SHConst {
    float K[64];   // (-1)^m * normalisation, index l*l + l + m (same for +-m)
};

static SHConst make_sh_const() {
    SHConst c;
    const double PI_ = 3.14159265358979323846;
    for (int l = 0; l < 8; l++)
        for (int m = -l; m <= l; m++) {
            const int am = m < 0 ? -m : m;
            double f1 = 1, f2 = 1;
            for (int i = 2; i <= l - am; i++) f1 *= i;
            for (int i = 2; i <= l + am; i++) f2 *= i;
            double K = std::sqrt((2 * l + 1) / (4 * PI_) * f1 / f2);
            if (am > 0) K *= std::sqrt(2.0);
            if (am & 1) K = -K;
            c.K[l * l + l + m] = (float)K;
        }
    return c;
}

// inputs [B,3] -> outputs [B,deg^2]; dy_dx [B,3,deg^2] (dx | dy | dz blocks) or null.
__global__ void __launch_bounds__(256) k_sh_forward(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B, uint32_t D,
                                                     uint32_t deg, float* __restrict__ dy_dx, const SHConst kc) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t C2 = deg * deg;
    const float x = inputs[(size_t)b * D], y = inputs[(size_t)b * D + 1], z = inputs[(size_t)b * D + 2];
    float* out = outputs + (size_t)b * C2;
    float* ddx = dy_dx ? dy_dx + (size_t)b * D * C2 : nullptr;
    float* ddy = ddx ? ddx + C2 : nullptr;
    float* ddz = ddx ? ddy + C2 : nullptr;
    // A_m = Re (x+iy)^m, S_m = Im (x+iy)^m
    float A[9], Sm[9];
    A[0] = 1.f; Sm[0] = 0.f;
    #pragma unroll
    for (int m = 1; m <= 8; m++) {
        A[m] = A[m - 1] * x - Sm[m - 1] * y;
        Sm[m] = A[m - 1] * y + Sm[m - 1] * x;
    }
    #pragma unroll
    for (int m = 0; m < 8; m++) {
        if (m >= (int)deg) break;
        // column m of Q_l^m(z), l = m..deg-1, and of Q_l^{m+1}(z) (= dQ_l^m/dz)
        float dfact = 1.f;
        #pragma unroll
        for (int k = 1; k <= 8; k++) if (k <= m) dfact *= (float)(2 * k - 1);
        float q_prev2 = 0.f, q_prev1 = 0.f;       // Q_{l-2}^m, Q_{l-1}^m
        float r_prev2 = 0.f, r_prev1 = 0.f;       // same for order m+1
        const float dfact1 = dfact * (float)(2 * m + 1);
        #pragma unroll
        for (int l = 0; l < 8; l++) {
            if (l < m || l >= (int)deg) continue;
            float q, r;
            if (l == m) q = dfact;
            else if (l == m + 1) q = (float)(2 * m + 1) * z * q_prev1;
            else q = ((float)(2 * l - 1) * z * q_prev1 - (float)(l + m - 1) * q_prev2) * (1.0f / (float)(l - m));
            if (l < m + 1) r = 0.f;
            else if (l == m + 1) r = dfact1;
            else if (l == m + 2) r = (float)(2 * m + 3) * z * r_prev1;
            else r = ((float)(2 * l - 1) * z * r_prev1 - (float)(l + m) * r_prev2) * (1.0f / (float)(l - m - 1));
            q_prev2 = q_prev1; q_prev1 = q;
            r_prev2 = r_prev1; r_prev1 = r;
            const float K = kc.K[l * l + l + m];
            const int ip = l * l + l + m, in_ = l * l + l - m;
            out[ip] = K * q * A[m];
            if (m > 0) out[in_] = K * q * Sm[m];
            if (ddx) {
                if (m == 0) {
                    ddx[ip] = 0.f; ddy[ip] = 0.f; ddz[ip] = K * r;
                } else {
                    const float km = K * q * (float)m;
                    ddx[ip] = km * A[m - 1];
                    ddy[ip] = -km * Sm[m - 1];
                    ddz[ip] = K * r * A[m];
                    ddx[in_] = km * Sm[m - 1];
                    ddy[in_] = km * A[m - 1];
                    ddz[in_] = K * r * Sm[m];
                }
            }
        }
    }
}

// shencoder.cu:359-382 (accumulates into grad_inputs, as the reference does)
__global__ void k_sh_backward(const float* __restrict__ grad, uint32_t B, uint32_t D, uint32_t deg, const float* __restrict__ dy_dx,
                              float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = t / D;
    if (b >= B) return;
    const uint32_t d = t - b * D, C2 = deg * deg;
    const float* g = grad + (size_t)b * C2;
    const float* dd = dy_dx + (size_t)b * D * C2 + (size_t)d * C2;
    float acc = grad_inputs[t];
    for (uint32_t c = 0; c < C2; c++) acc = fmaf(g[c], dd[c], acc);
    grad_inputs[t] = acc;
}

// ======================================================================================
// frequency encoder   freqencoder.cu:30-94
// ======================================================================================
__global__ void k_freq_forward(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                               float* __restrict__ outputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * C) return;
    const uint32_t b = t / C, c = t - b * C;
    const float* in = inputs + (size_t)b * D;
    if (c < D) outputs[t] = in[c];
    else {
        const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
        const float phase = (float)(col % 2) * (3.141592653589793f / 2);
        outputs[t] = __sinf(__fadd_rn(scalbnf(in[d], (int)freq), phase));
    }
}

__global__ void k_freq_backward(const float* __restrict__ grad, const float* __restrict__ outputs, uint32_t B, uint32_t D, uint32_t deg,
                                uint32_t C, float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float result = g[d];
    g += D; o += D;
    for (uint32_t f = 0; f < deg; f++) {
        result += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
        g += 2 * D; o += 2 * D;
    }
    grad_inputs[t] = result;
}

}  // namespace gf

namespace gf {
template <int D, int C>
static int launch_tv(const float* inputs, const float* emb, float* grad, const int* offsets, float weight, uint32_t B, uint32_t L, float S,
                     uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    const dim3 grid(div_up(B, 256), L, 1);
    k_grad_tv<D, C><<<grid, 256, 0, st>>>(inputs, emb, grad, offsets, weight, B, L, S, H, gridtype, ac);
    return check_launch("grad_total_variation");
}

}  // namespace gf

// ======================================================================================
// C ABI
// ======================================================================================
using namespace gf;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

GF_API int gf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs, uint32_t B,
                                  uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void* dy_dx, uint32_t gridtype,
                                  int align_corners, uint32_t interp, int dtype, gf_stream_t stream) {
    GF_REQUIRE(inputs && embeddings && offsets && outputs, "grid_encode_forward: null pointer");
    GF_REQUIRE(dtype == 0 || dtype == 1, "grid_encode_forward: dtype must be 0 (float32) or 1 (float16)");
    GF_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    GF_REQUIRE(D >= 2 && D <= 5, "GridEncoding: D must be 2, 3, 4, or 5.");
    GF_REQUIRE(gridtype <= 1 && interp <= 1, "grid_encode_forward: bad gridtype/interp");
    if (B == 0 || L == 0) return GF_OK;
    const bool ac = align_corners != 0;
    if (dtype == 0) {
        GF_DISPATCH_DC(launch_grid_forward, float, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac, interp, ST(stream));
    } else {
        GF_DISPATCH_DC(launch_grid_forward, __half, inputs, embeddings, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac, interp, ST(stream));
    }
    set_error("grid_encode_forward: unsupported D/C");
    return GF_ERR_UNSUPPORTED;
}

GF_API int gf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                                   void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                   const void* dy_dx, void* grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                                   gf_stream_t stream) {
    (void)embeddings;
    GF_REQUIRE(grad && inputs && offsets && grad_embeddings, "grid_encode_backward: null pointer");
    GF_REQUIRE(dtype == 0 || dtype == 1, "grid_encode_backward: dtype must be 0 (float32) or 1 (float16)");
    GF_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    GF_REQUIRE(D >= 2 && D <= 5, "GridEncoding: D must be 2, 3, 4, or 5.");
    if (B == 0 || L == 0) return GF_OK;
    const bool ac = align_corners != 0;
    if (dtype == 0) {
        GF_DISPATCH_DC(launch_grid_backward, float, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs, gridtype, ac, interp, ST(stream));
    } else {
        GF_DISPATCH_DC(launch_grid_backward, __half, grad, inputs, offsets, grad_embeddings, B, L, S, H, dy_dx, grad_inputs, gridtype, ac, interp, ST(stream));
    }
    set_error("grid_encode_backward: unsupported D/C");
    return GF_ERR_UNSUPPORTED;
}

GF_API int gf_grad_total_variation(const float* inputs, const float* embeddings, float* grad, const int32_t* offsets, float weight,
                                   uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                   int align_corners, gf_stream_t stream) {
    GF_REQUIRE(inputs && embeddings && grad && offsets, "grad_total_variation: null pointer");
    GF_REQUIRE(C == 1 || C == 2 || C == 4 || C == 8, "GridEncoding: C must be 1, 2, 4, or 8.");
    if (B == 0 || L == 0) return GF_OK;
    const bool ac = align_corners != 0;
    switch (D * 16 + C) {
        case 2 * 16 + 1: return launch_tv<2, 1>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 2 * 16 + 2: return launch_tv<2, 2>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 2 * 16 + 4: return launch_tv<2, 4>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 2 * 16 + 8: return launch_tv<2, 8>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 3 * 16 + 1: return launch_tv<3, 1>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 3 * 16 + 2: return launch_tv<3, 2>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 3 * 16 + 4: return launch_tv<3, 4>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        case 3 * 16 + 8: return launch_tv<3, 8>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype, ac, ST(stream));
        default: break;
    }
    set_error("grad_total_variation: D must be 2 or 3, C in {1,2,4,8}");
    return GF_ERR_UNSUPPORTED;
}

GF_API int gf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx,
                                gf_stream_t stream) {
    GF_REQUIRE(inputs && outputs, "sh_encode_forward: null pointer");
    GF_REQUIRE(D == 3, "SH encoder only support input dim == 3");
    GF_REQUIRE(degree >= 1 && degree <= 8, "SH encoder only supports degree in [1, 8]");
    if (B == 0) return GF_OK;
    static const SHConst kc = make_sh_const();
    k_sh_forward<<<div_up(B, 256), 256, 0, ST(stream)>>>(inputs, outputs, B, D, degree, dy_dx, kc);
    return check_launch("sh_encode_forward");
}

GF_API int gf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                                 float* grad_inputs, gf_stream_t stream) {
    (void)inputs;
    GF_REQUIRE(grad && dy_dx && grad_inputs, "sh_encode_backward: null pointer");
    GF_REQUIRE(D == 3 && degree >= 1 && degree <= 8, "sh_encode_backward: bad D/degree");
    if (B == 0) return GF_OK;
    k_sh_backward<<<div_up(B * D, 256), 256, 0, ST(stream)>>>(grad, B, D, degree, dy_dx, grad_inputs);
    return check_launch("sh_encode_backward");
}

GF_API int gf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t degree, uint32_t C, float* outputs,
                                  gf_stream_t stream) {
    GF_REQUIRE(inputs && outputs, "freq_encode_forward: null pointer");
    GF_REQUIRE(C == D + D * 2 * degree, "freq_encode_forward: C must equal D + 2*D*degree");
    if (B == 0) return GF_OK;
    k_freq_forward<<<div_up(B * C, 128), 128, 0, ST(stream)>>>(inputs, B, D, degree, C, outputs);
    return check_launch("freq_encode_forward");
}

GF_API int gf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t degree, uint32_t C,
                                   float* grad_inputs, gf_stream_t stream) {
    GF_REQUIRE(grad && outputs && grad_inputs, "freq_encode_backward: null pointer");
    GF_REQUIRE(C == D + D * 2 * degree, "freq_encode_backward: C must equal D + 2*D*degree");
    if (B == 0) return GF_OK;
    k_freq_backward<<<div_up(B * D, 128), 128, 0, ST(stream)>>>(grad, outputs, B, D, degree, C, grad_inputs);
    return check_launch("freq_encode_backward");
}

}  // extern "C"
