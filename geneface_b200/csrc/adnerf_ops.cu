// libgfrender: the non-GEMM operators of the vanilla AD-NeRF path (reference: modules/nerfs).
//
//   gf_adnerf_get_rays        commons/ray_samplers.py:11-44     OpenGL-convention pinhole rays + unit view directions
//   gf_adnerf_embed           commons/embedders.py:5-45         [x, sin(2^k x), cos(2^k x)]_k frequency embedding
//   gf_adnerf_embed_points    volume_rendering.py:153 + embed   pts = o + d z, embedded in the same pass (no [R,S,3] round trip)
//   gf_adnerf_raw2outputs     volume_rendering.py:9-59          sigma/rgb -> weights, rgb/depth/disp/acc maps (warp scan per ray)
//   gf_adnerf_sample_pdf      volume_rendering.py:62-96,177-182 inverse-CDF importance samples merged + sorted with the coarse z
//
// The 8x256 / 3x128 MLPs of the backbone stay plain library GEMMs on the host side (geneface_b200/adnerf.py).  Inference only.
#include <cuda_runtime.h>

#include <cstdint>

#include "gf_common.cuh"

namespace gf {

// ---------------------------------------------------------------------------------------------------------- rays
__global__ void k_adnerf_rays(uint32_t H, uint32_t W, float focal, float cx, float cy, const float* __restrict__ c2w /*3x4 row-major*/,
                              float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ viewdirs) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= H * W) return;
    const float i = (float)(n % W), j = (float)(n / W);
    // camera-space direction: x right, y up, looking along -z
    const float dx = (i - cx) / focal, dy = -(j - cy) / focal, dz = -1.0f;
    float d[3];
    #pragma unroll
    for (int r = 0; r < 3; r++) d[r] = dx * c2w[4 * r] + dy * c2w[4 * r + 1] + dz * c2w[4 * r + 2];
    const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    #pragma unroll
    for (int r = 0; r < 3; r++) {
        rays_o[3 * (size_t)n + r] = c2w[4 * r + 3];
        rays_d[3 * (size_t)n + r] = d[r];
        if (viewdirs) viewdirs[3 * (size_t)n + r] = d[r] * inv;
    }
}

// ---------------------------------------------------------------------------------------------------------- embedding
// out row = [x (D), sin(f0 x) (D), cos(f0 x) (D), sin(f1 x), cos(f1 x), ...], f_k = 2^k (log bands, include_input)
__device__ __forceinline__ void embed_row(const float* x, uint32_t D, uint32_t L, float* out) {
    for (uint32_t c = 0; c < D; c++) out[c] = x[c];
    float f = 1.0f;
    for (uint32_t k = 0; k < L; k++, f *= 2.0f) {
        for (uint32_t c = 0; c < D; c++) {
            float s, co;
            sincosf(x[c] * f, &s, &co);
            out[D + (2 * k) * D + c] = s;
            out[D + (2 * k + 1) * D + c] = co;
        }
    }
}

__global__ void k_adnerf_embed(const float* __restrict__ x, uint32_t n, uint32_t D, uint32_t L, float* __restrict__ out, uint32_t ld) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[8];
    for (uint32_t c = 0; c < D; c++) v[c] = x[(size_t)i * D + c];
    embed_row(v, D, L, out + (size_t)i * ld);
}

// one thread per (ray, sample)
__global__ void k_adnerf_embed_points(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z,
                                      uint32_t R, uint32_t S, uint32_t L, float* __restrict__ out, uint32_t ld) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * S) return;
    const uint32_t r = i / S;
    const float zz = z[i];
    float p[3];
    #pragma unroll
    for (int c = 0; c < 3; c++) p[c] = rays_o[3 * (size_t)r + c] + rays_d[3 * (size_t)r + c] * zz;
    embed_row(p, 3, L, out + (size_t)i * ld);
}

// ---------------------------------------------------------------------------------------------------------- raw2outputs
// One warp per ray; lanes stride the samples; transmittance by a warp-level multiplicative scan carried across 32-sample chunks.
//   dist_s  = (z_{s+1} - z_s) |d|   (last: 1e10 |d|)
//   alpha_s = 1 - exp(-(relu(sigma_s) + 1e-6) dist_s)
//   T_s     = prod_{k<s} (1 - alpha_k + 1e-10),  w_s = alpha_s T_s
//   rgb_s   = sigmoid(raw_rgb_s), except the LAST sample, whose colour is the background colour of the ray
__global__ void k_adnerf_raw2outputs(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                                     const float* __restrict__ bc_rgb, uint32_t R, uint32_t S, int white_bkgd, float* __restrict__ rgb_map,
                                     float* __restrict__ disp_map, float* __restrict__ acc_map, float* __restrict__ weights,
                                     float* __restrict__ depth_map, float* __restrict__ rgb_map_fg) {
    const uint32_t ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= R) return;
    const float dn = sqrtf(rays_d[3 * (size_t)ray] * rays_d[3 * (size_t)ray] + rays_d[3 * (size_t)ray + 1] * rays_d[3 * (size_t)ray + 1] +
                           rays_d[3 * (size_t)ray + 2] * rays_d[3 * (size_t)ray + 2]);
    const float* zr = z + (size_t)ray * S;
    const float4* rr = reinterpret_cast<const float4*>(raw) + (size_t)ray * S;
    float carry = 1.0f;                                   // transmittance in front of the current chunk
    float ar = 0.f, ag = 0.f, ab = 0.f, fr = 0.f, fg = 0.f, fb = 0.f, ad = 0.f, aw = 0.f;
    for (uint32_t s0 = 0; s0 < S; s0 += 32) {
        const uint32_t s = s0 + lane;
        const bool in = s < S;
        float alpha = 0.f, zz = 0.f;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) {
            v = rr[s];
            zz = zr[s];
            const float dist = (s + 1 < S ? zr[s + 1] - zz : 1e10f) * dn;
            alpha = 1.0f - expf(-(fmaxf(v.w, 0.f) + 1e-6f) * dist);
        }
        const float t = in ? 1.0f - alpha + 1e-10f : 1.0f;
        float incl = t;                                    // inclusive product over the lanes of the chunk
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl *= u;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float w = alpha * carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        if (in) {
            if (weights) weights[(size_t)ray * S + s] = w;
            const bool last = s + 1 == S;
            const float cr = last ? bc_rgb[3 * (size_t)ray] : 1.0f / (1.0f + expf(-v.x));
            const float cg = last ? bc_rgb[3 * (size_t)ray + 1] : 1.0f / (1.0f + expf(-v.y));
            const float cb = last ? bc_rgb[3 * (size_t)ray + 2] : 1.0f / (1.0f + expf(-v.z));
            ar += w * cr; ag += w * cg; ab += w * cb;
            if (!last) { fr += w * cr; fg += w * cg; fb += w * cb; }
            ad += w * zz; aw += w;
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ar += __shfl_xor_sync(0xffffffffu, ar, o); ag += __shfl_xor_sync(0xffffffffu, ag, o); ab += __shfl_xor_sync(0xffffffffu, ab, o);
        fr += __shfl_xor_sync(0xffffffffu, fr, o); fg += __shfl_xor_sync(0xffffffffu, fg, o); fb += __shfl_xor_sync(0xffffffffu, fb, o);
        ad += __shfl_xor_sync(0xffffffffu, ad, o); aw += __shfl_xor_sync(0xffffffffu, aw, o);
    }
    if (lane == 0) {
        const float wb = white_bkgd ? 1.0f - aw : 0.0f;
        rgb_map[3 * (size_t)ray] = ar + wb; rgb_map[3 * (size_t)ray + 1] = ag + wb; rgb_map[3 * (size_t)ray + 2] = ab + wb;
        if (rgb_map_fg) { rgb_map_fg[3 * (size_t)ray] = fr; rgb_map_fg[3 * (size_t)ray + 1] = fg; rgb_map_fg[3 * (size_t)ray + 2] = fb; }
        if (depth_map) depth_map[ray] = ad;
        if (acc_map) acc_map[ray] = aw;
        if (disp_map) disp_map[ray] = 1.0f / fmaxf(1e-10f, ad / aw);
    }
}

// ---------------------------------------------------------------------------------------------------------- sample_pdf + merge
// One block per ray.  bins = mid-points of z (S-1 values), pdf weights = w[1 : S-1] (S-2 values) + 1e-5, cdf = [0, cumsum(pdf)].
// Sample i: u_i (det: i/(N-1), else caller's uniform numbers), ind = #(cdf <= u) (searchsorted right), below = max(ind-1, 0),
// above = min(ind, S-2), t = (u - cdf[below]) / (cdf[above]-cdf[below] or 1 when < 1e-5), sample = bins[below] + t (bins[above]-bins[below]).
// Output row: the S coarse depths and the N new ones, sorted ascending (bitonic sort in shared memory).
constexpr int PDF_MAX = 512;        // S + N padded to a power of two
// MERGE = false: the plain sample_pdf(bins [R,S], weights [R,S-1]) of the reference: z holds the bins themselves, w the pdf
// weights, and z_out receives the N samples unsorted.
template <bool MERGE>
__global__ void __launch_bounds__(128) k_adnerf_sample_pdf(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ u_in,
                                                           uint32_t R, uint32_t S, uint32_t N, float* __restrict__ z_out,
                                                           float* __restrict__ samples_out) {
    __shared__ float cdf[PDF_MAX];
    __shared__ float val[PDF_MAX];
    const uint32_t ray = blockIdx.x, tid = threadIdx.x;
    if (ray >= R) return;
    const float* zr = z + (size_t)ray * S;
    const uint32_t nb = MERGE ? S - 1 : S;          // bins, = cdf entries
    if (tid == 0) {
        const float* wr = MERGE ? w + (size_t)ray * S + 1 : w + (size_t)ray * (S - 1);      // nb - 1 pdf weights
        float tot = 0.f;
        for (uint32_t k = 0; k + 1 < nb; k++) tot += wr[k] + 1e-5f;
        float c = 0.f;
        cdf[0] = 0.f;
        for (uint32_t k = 0; k + 1 < nb; k++) { c += (wr[k] + 1e-5f) / tot; cdf[k + 1] = c; }
    }
    if (MERGE)
        for (uint32_t s = tid; s < S; s += blockDim.x) val[s] = zr[s];
    __syncthreads();
    for (uint32_t i = tid; i < N; i += blockDim.x) {
        const float u = u_in ? u_in[(size_t)ray * N + i] : (N > 1 ? (float)i / (float)(N - 1) : 0.f);
        uint32_t lo = 0, hi = nb;                   // first index with cdf > u
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
        }
        const uint32_t below = lo > 0 ? lo - 1 : 0, above = lo < nb - 1 ? lo : nb - 1;
        const float c0 = cdf[below], c1 = cdf[above];
        float den = c1 - c0;
        if (den < 1e-5f) den = 1.0f;
        const float t = (u - c0) / den;
        const float b0 = MERGE ? 0.5f * (zr[below + 1] + zr[below]) : zr[below], b1 = MERGE ? 0.5f * (zr[above + 1] + zr[above]) : zr[above];
        const float smp = b0 + t * (b1 - b0);
        if (MERGE) val[S + i] = smp; else z_out[(size_t)ray * N + i] = smp;
        if (samples_out) samples_out[(size_t)ray * N + i] = smp;
    }
    if (!MERGE) return;
    uint32_t P = 1;
    while (P < S + N) P <<= 1;
    for (uint32_t s = S + N + tid; s < P; s += blockDim.x) val[s] = 3.0e38f;
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < P; i += blockDim.x) {
                const uint32_t l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const float a = val[i], b = val[l];
                    if ((a > b) == up) { val[i] = b; val[l] = a; }
                }
            }
            __syncthreads();
        }
    for (uint32_t s = tid; s < S + N; s += blockDim.x) z_out[(size_t)ray * (S + N) + s] = val[s];
}

}  // namespace gf

using namespace gf;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

GF_API int gf_adnerf_get_rays(uint32_t H, uint32_t W, float focal, float cx, float cy, const float* c2w, float* rays_o, float* rays_d,
                              float* viewdirs, gf_stream_t stream) {
    GF_REQUIRE(c2w && rays_o && rays_d, "adnerf_get_rays: null pointer");
    if (H * W == 0) return GF_OK;
    k_adnerf_rays<<<div_up(H * W, 256), 256, 0, ST(stream)>>>(H, W, focal, cx, cy, c2w, rays_o, rays_d, viewdirs);
    return check_launch("adnerf_get_rays");
}

GF_API int gf_adnerf_embed(const float* x, uint32_t n, uint32_t D, uint32_t multi_res, float* out, uint32_t ld, gf_stream_t stream) {
    GF_REQUIRE(x && out, "adnerf_embed: null pointer");
    GF_REQUIRE(D >= 1 && D <= 8, "adnerf_embed: input dim %u not in 1..8", D);
    GF_REQUIRE(ld >= D * (1 + 2 * multi_res), "adnerf_embed: row stride %u < embedding width %u", ld, D * (1 + 2 * multi_res));
    if (n == 0) return GF_OK;
    k_adnerf_embed<<<div_up(n, 256), 256, 0, ST(stream)>>>(x, n, D, multi_res, out, ld);
    return check_launch("adnerf_embed");
}

GF_API int gf_adnerf_embed_points(const float* rays_o, const float* rays_d, const float* z_vals, uint32_t R, uint32_t S, uint32_t multi_res,
                                  float* out, uint32_t ld, gf_stream_t stream) {
    GF_REQUIRE(rays_o && rays_d && z_vals && out, "adnerf_embed_points: null pointer");
    GF_REQUIRE(ld >= 3 * (1 + 2 * multi_res), "adnerf_embed_points: row stride %u < embedding width %u", ld, 3 * (1 + 2 * multi_res));
    if ((uint64_t)R * S == 0) return GF_OK;
    GF_REQUIRE((uint64_t)R * S < 0xffffffffull, "adnerf_embed_points: too many samples");
    k_adnerf_embed_points<<<div_up(R * S, 256), 256, 0, ST(stream)>>>(rays_o, rays_d, z_vals, R, S, multi_res, out, ld);
    return check_launch("adnerf_embed_points");
}

GF_API int gf_adnerf_raw2outputs(const float* raw, const float* z_vals, const float* rays_d, const float* bc_rgb, uint32_t R, uint32_t S,
                                 int white_bkgd, float* rgb_map, float* disp_map, float* acc_map, float* weights, float* depth_map,
                                 float* rgb_map_fg, gf_stream_t stream) {
    GF_REQUIRE(raw && z_vals && rays_d && bc_rgb && rgb_map, "adnerf_raw2outputs: null pointer");
    GF_REQUIRE(S >= 1, "adnerf_raw2outputs: no samples");
    GF_REQUIRE((reinterpret_cast<uintptr_t>(raw) & 15) == 0, "adnerf_raw2outputs: raw must be 16-byte aligned");
    if (R == 0) return GF_OK;
    k_adnerf_raw2outputs<<<div_up(R, 4), 128, 0, ST(stream)>>>(raw, z_vals, rays_d, bc_rgb, R, S, white_bkgd, rgb_map, disp_map, acc_map, weights,
                                                              depth_map, rgb_map_fg);
    return check_launch("adnerf_raw2outputs");
}

GF_API int gf_adnerf_sample_pdf(const float* z_vals, const float* weights, const float* u, uint32_t R, uint32_t S, uint32_t N_importance,
                                int merge, float* z_out, float* samples_out, gf_stream_t stream) {
    GF_REQUIRE(z_vals && weights && z_out, "adnerf_sample_pdf: null pointer");
    GF_REQUIRE(S >= 3, "adnerf_sample_pdf: needs at least 3 bins / coarse samples, got %u", S);
    GF_REQUIRE(S + N_importance <= (uint32_t)PDF_MAX, "adnerf_sample_pdf: S + N_importance = %u exceeds %d", S + N_importance, PDF_MAX);
    if (R == 0) return GF_OK;
    if (merge) k_adnerf_sample_pdf<true><<<R, 128, 0, ST(stream)>>>(z_vals, weights, u, R, S, N_importance, z_out, samples_out);
    else k_adnerf_sample_pdf<false><<<R, 128, 0, ST(stream)>>>(z_vals, weights, u, R, S, N_importance, z_out, samples_out);
    return check_launch("adnerf_sample_pdf");
}

}  // extern "C"
