// libgfrender: the `_raymarching_face` operator set (fine-grained drop-in boundary).
//
// Replaces modules/radnerfs/raymarching/src/raymarching.cu (12 host entry points,
// raymarching.h:7-20).  Same per-element arithmetic as the reference -- the rounding sequence
// of the occupancy march is pinned in gf_common.cuh -- but:
//   * every launch goes on the caller's stream and is error-checked,
//   * march_rays_train lays rays out DETERMINISTICALLY (count -> single-block scan -> write)
//     instead of racing two global atomics per ray (raymarching.cu:446-447),
//   * grids are sized from the element count, block = 128/256 threads.
#include "gf_common.cuh"

namespace gf {

static constexpr uint32_t NT = 128;

// ------------------------------------------------------------------------------------ K1
__global__ void k_near_far_from_aabb(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                     const float* __restrict__ aabb, uint32_t N, float min_near, float* __restrict__ nears,
                                     float* __restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* o = rays_o + 3 * (size_t)n;
    const float* d = rays_d + 3 * (size_t)n;
    const Ray r = make_ray(o[0], o[1], o[2], d[0], d[1], d[2]);
    float near, far;
    near_far_aabb(r, aabb, min_near, near, far);
    nears[n] = near;
    fars[n] = far;
}

// ------------------------------------------------------------------------------------ K2
// raymarching.cu:162-198
__global__ void k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius, uint32_t N,
                               float* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[3 * (size_t)n], oy = rays_o[3 * (size_t)n + 1], oz = rays_o[3 * (size_t)n + 2];
    const float dx = rays_d[3 * (size_t)n], dy = rays_d[3 * (size_t)n + 1], dz = rays_d[3 * (size_t)n + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float C = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * C)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    const float RPI = 0.3183098861837907f;
    coords[2 * (size_t)n] = 2 * theta * RPI - 1;
    coords[2 * (size_t)n + 1] = phi * RPI;
}

// ------------------------------------------------------------------------------------ K3/K4
__global__ void k_morton3D(const int* __restrict__ coords, uint32_t N, int* __restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int)morton3D(coords[3 * (size_t)n], coords[3 * (size_t)n + 1], coords[3 * (size_t)n + 2]);
}
__global__ void k_morton3D_invert(const int* __restrict__ indices, uint32_t N, int* __restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ind = indices[n];
    coords[3 * (size_t)n + 0] = (int)morton3D_invert((uint32_t)(ind >> 0));
    coords[3 * (size_t)n + 1] = (int)morton3D_invert((uint32_t)(ind >> 1));
    coords[3 * (size_t)n + 2] = (int)morton3D_invert((uint32_t)(ind >> 2));
}

// ------------------------------------------------------------------------------------ K5
// raymarching.cu:267-289.  One thread per output byte; the 8 floats are read as two float4.
__global__ void k_packbits(const float* __restrict__ grid, uint32_t N, float thresh, uint8_t* __restrict__ bitfield) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n);
    const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n + 1);
    uint32_t bits = 0;
    bits |= (a.x > thresh) ? 1u : 0u;
    bits |= (a.y > thresh) ? 2u : 0u;
    bits |= (a.z > thresh) ? 4u : 0u;
    bits |= (a.w > thresh) ? 8u : 0u;
    bits |= (b.x > thresh) ? 16u : 0u;
    bits |= (b.y > thresh) ? 32u : 0u;
    bits |= (b.z > thresh) ? 64u : 0u;
    bits |= (b.w > thresh) ? 128u : 0u;
    bitfield[n] = (uint8_t)bits;
}

// ------------------------------------------------------------------------------------ K6
// raymarching.cu:304-335
__global__ void k_morton3D_dilation(const float* __restrict__ grid, uint32_t C, uint32_t H, float* __restrict__ out) {
    const uint32_t H3 = H * H * H;
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= C * H3) return;
    const uint32_t c = n / H3, ind = n - c * H3;
    const uint32_t x = morton3D_invert(ind >> 0), y = morton3D_invert(ind >> 1), z = morton3D_invert(ind >> 2);
    const float* g = grid + (size_t)c * H3;
    float res = g[ind];
    if (x + 1 < H) res = fmaxf(res, __ldg(g + morton3D(x + 1, y, z)));
    if (x > 0) res = fmaxf(res, __ldg(g + morton3D(x - 1, y, z)));
    if (y + 1 < H) res = fmaxf(res, __ldg(g + morton3D(x, y + 1, z)));
    if (y > 0) res = fmaxf(res, __ldg(g + morton3D(x, y - 1, z)));
    if (z + 1 < H) res = fmaxf(res, __ldg(g + morton3D(x, y, z + 1)));
    if (z > 0) res = fmaxf(res, __ldg(g + morton3D(x, y, z - 1)));
    out[n] = res;
}

// ------------------------------------------------------------------------------------ K7
// raymarching.cu:352-518 as three launches: count, scan, write.
__device__ __forceinline__ float perturbed_t0(const MarchConst& m, float near, float noise) {
    // raymarching.cu:392 / :873 : t += clamp(t*dt_gamma, dt_min, dt_max) * noise  (one FFMA)
    return __fmaf_rn(noise, clampf(__fmul_rn(near, m.dt_gamma), m.dt_min, m.dt_max), near);
}

__global__ void k_march_train_count(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                    const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps,
                                    uint32_t N, uint32_t C, uint32_t H, const float* __restrict__ nears,
                                    const float* __restrict__ fars, const float* __restrict__ noises, int* __restrict__ rays) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const MarchConst m = make_march_const(bound, dt_gamma, max_steps, C, H, grid);
    const float* o = rays_o + 3 * (size_t)n;
    const float* d = rays_d + 3 * (size_t)n;
    const Ray r = make_ray(o[0], o[1], o[2], d[0], d[1], d[2]);
    const float far = fars[n];
    float t = perturbed_t0(m, nears[n], noises[n]);
    uint32_t num_steps = 0;
    Probe p;
    while (num_steps < max_steps && march_next(m, r, far, t, p)) {
        num_steps++;
        t = __fadd_rn(t, p.dt);
    }
    rays[3 * (size_t)n + 2] = (int)num_steps;      // the count's final place in row n = (ray id, offset, count)
}

// Single block: exclusive scan of the per-ray counts (already in column 2 of `rays`) into sample offsets (column 1).
// Offsets are handed out in ROTATED ray order (first ray = rot): when the sample total exceeds M (the running mean of the previous
// steps, raymarching.py:225-228) the rays that lose their samples are the last ones in allocation order -- in the reference whichever
// lose its atomic race, here a contiguous run starting at a per-call pseudo-random ray instead of always the highest indices (which
// would systematically starve the bottom rows of an ordered ray set such as the lips rectangle).  rot is derived from the first
// perturbation noise (0 when perturb is off: plain index order).  Row n always describes ray n; counter[1] only reports N.
__global__ void k_march_train_scan(uint32_t N, int* rays, int* counter, const float* __restrict__ noises) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int base_point = counter[0];
    const uint32_t rot = N > 1 ? (__float_as_uint(noises[0]) >> 3) % N : 0;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t start = 0; start < N; start += blockDim.x) {
        const uint32_t p = start + tid;
        uint32_t n = p + rot;
        if (n >= N) n -= N;
        const int c = p < N ? rays[3 * (size_t)n + 2] : 0;
        int v = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (uint32_t)o) v += u;
        }
        if (lane == 31) warp_sums[wid] = v;
        __syncthreads();
        if (wid == 0) {
            int w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += u;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int incl = v + (wid ? warp_sums[wid - 1] : 0) + carry;
        if (p < N) {
            rays[3 * (size_t)n] = (int)n;
            rays[3 * (size_t)n + 1] = base_point + incl - c;
        }
        __syncthreads();
        if (tid == blockDim.x - 1) carry = incl;
        __syncthreads();
    }
    if (tid == 0) {
        counter[0] = base_point + carry;
        counter[1] += (int)N;
    }
}

__global__ void k_march_train_write(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                    const uint8_t* __restrict__ grid, float bound, float dt_gamma, uint32_t max_steps,
                                    uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* __restrict__ nears,
                                    const float* __restrict__ fars, const float* __restrict__ noises,
                                    const int* __restrict__ rays, const int* __restrict__ counter, float* __restrict__ xyzs,
                                    float* __restrict__ dirs, float* __restrict__ deltas) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    (void)counter;
    const int* row = rays + 3 * (size_t)n;
    const uint32_t point_index = (uint32_t)row[1], num_steps = (uint32_t)row[2];
    if (num_steps == 0 || point_index + num_steps > M) return;
    const MarchConst m = make_march_const(bound, dt_gamma, max_steps, C, H, grid);
    const float* o = rays_o + 3 * (size_t)n;
    const float* d = rays_d + 3 * (size_t)n;
    const Ray r = make_ray(o[0], o[1], o[2], d[0], d[1], d[2]);
    const float far = fars[n];
    float t = perturbed_t0(m, nears[n], noises[n]);
    float* px = xyzs + 3 * (size_t)point_index;
    float* pd = dirs + 3 * (size_t)point_index;
    float* pl = deltas + 2 * (size_t)point_index;
    uint32_t step = 0;
    Probe p;
    while (step < num_steps && march_next(m, r, far, t, p)) {
        px[0] = p.x; px[1] = p.y; px[2] = p.z;
        pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
        t = __fadd_rn(t, p.dt);
        pl[0] = p.dt; pl[1] = t;
        px += 3; pd += 3; pl += 2; step++;
    }
}

// ------------------------------------------------------------------------------------ K8
// raymarching.cu:535-583 (accumulates into the caller's grad buffers; row = slot n)
__global__ void k_march_train_backward(const float* __restrict__ grad_xyzs, const float* __restrict__ grad_dirs,
                                       const int* __restrict__ rays, const float* __restrict__ deltas, uint32_t N, uint32_t M,
                                       float* __restrict__ grad_rays_o, float* __restrict__ grad_rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    float go0 = grad_rays_o[3 * (size_t)n], go1 = grad_rays_o[3 * (size_t)n + 1], go2 = grad_rays_o[3 * (size_t)n + 2];
    float gd0 = grad_rays_d[3 * (size_t)n], gd1 = grad_rays_d[3 * (size_t)n + 1], gd2 = grad_rays_d[3 * (size_t)n + 2];
    for (uint32_t s = 0; s < num_steps; s++) {
        const size_t i = (size_t)offset + s;
        const float gx = grad_xyzs[3 * i], gy = grad_xyzs[3 * i + 1], gz = grad_xyzs[3 * i + 2];
        const float tt = deltas[2 * i + 1];
        go0 += gx; go1 += gy; go2 += gz;
        gd0 += fmaf(gx, tt, grad_dirs[3 * i]);
        gd1 += fmaf(gy, tt, grad_dirs[3 * i + 1]);
        gd2 += fmaf(gz, tt, grad_dirs[3 * i + 2]);
    }
    grad_rays_o[3 * (size_t)n] = go0; grad_rays_o[3 * (size_t)n + 1] = go1; grad_rays_o[3 * (size_t)n + 2] = go2;
    grad_rays_d[3 * (size_t)n] = gd0; grad_rays_d[3 * (size_t)n + 1] = gd1; grad_rays_d[3 * (size_t)n + 2] = gd2;
}

// ------------------------------------------------------------------------------------ K9
// raymarching.cu:603-687
__global__ void k_composite_train_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                      const float* __restrict__ ambient, const float* __restrict__ deltas,
                                      const int* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                                      float* __restrict__ weights_sum, float* __restrict__ ambient_sum, float* __restrict__ depth,
                                      float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) {
        weights_sum[index] = 0; ambient_sum[index] = 0; depth[index] = 0;
        image[index * 3] = 0; image[index * 3 + 1] = 0; image[index * 3 + 2] = 0;
        return;
    }
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0, amb = 0;
    for (uint32_t s = 0; s < num_steps; s++) {
        const size_t i = (size_t)offset + s;
        const float2 dl = __ldg(reinterpret_cast<const float2*>(deltas) + i);
        const float alpha = 1.0f - __expf(-__ldg(sigmas + i) * dl.x);
        const float weight = alpha * T;
        r = fmaf(weight, __ldg(rgbs + 3 * i), r);
        g = fmaf(weight, __ldg(rgbs + 3 * i + 1), g);
        b = fmaf(weight, __ldg(rgbs + 3 * i + 2), b);
        d = fmaf(weight, dl.y, d);
        ws += weight;
        amb += __ldg(ambient + i);
        T *= 1.0f - alpha;
        if (T < T_thresh) break;
    }
    weights_sum[index] = ws; ambient_sum[index] = amb; depth[index] = d;
    image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
}

// ------------------------------------------------------------------------------------ K10
// raymarching.cu:711-809
__global__ void k_composite_train_bwd(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_ambient_sum,
                                      const float* __restrict__ grad_image, const float* __restrict__ sigmas,
                                      const float* __restrict__ rgbs, const float* __restrict__ deltas,
                                      const int* __restrict__ rays, const float* __restrict__ weights_sum,
                                      const float* __restrict__ image, uint32_t M, uint32_t N, float T_thresh,
                                      float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, float* __restrict__ grad_ambient) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    const float gws = grad_weights_sum[index], gas = grad_ambient_sum[index];
    const float gi0 = grad_image[3 * (size_t)index], gi1 = grad_image[3 * (size_t)index + 1], gi2 = grad_image[3 * (size_t)index + 2];
    const float r_final = image[3 * (size_t)index], g_final = image[3 * (size_t)index + 1], b_final = image[3 * (size_t)index + 2];
    const float ws_final = weights_sum[index];
    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
    for (uint32_t s = 0; s < num_steps; s++) {
        const size_t i = (size_t)offset + s;
        const float2 dl = __ldg(reinterpret_cast<const float2*>(deltas) + i);
        const float c0 = __ldg(rgbs + 3 * i), c1 = __ldg(rgbs + 3 * i + 1), c2 = __ldg(rgbs + 3 * i + 2);
        const float alpha = 1.0f - __expf(-__ldg(sigmas + i) * dl.x);
        const float weight = alpha * T;
        r = fmaf(weight, c0, r);
        g = fmaf(weight, c1, g);
        b = fmaf(weight, c2, b);
        ws += weight;
        T *= 1.0f - alpha;
        grad_rgbs[3 * i] = gi0 * weight;
        grad_rgbs[3 * i + 1] = gi1 * weight;
        grad_rgbs[3 * i + 2] = gi2 * weight;
        grad_ambient[i] = gas;
        grad_sigmas[i] = dl.x * (gi0 * (T * c0 - (r_final - r)) + gi1 * (T * c1 - (g_final - g)) +
                                 gi2 * (T * c2 - (b_final - b)) + gws * (1 - ws_final));
        if (T < T_thresh) break;
    }
}

// ------------------------------------------------------------------------------------ K11
// raymarching.cu:827-929
__global__ void k_march_rays(uint32_t n_alive, uint32_t n_step, const int* __restrict__ rays_alive,
                             const float* __restrict__ rays_t, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                             float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                             const uint8_t* __restrict__ grid, const float* __restrict__ fars, float* __restrict__ xyzs,
                             float* __restrict__ dirs, float* __restrict__ deltas, const float* __restrict__ noises) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const MarchConst m = make_march_const(bound, dt_gamma, max_steps, C, H, grid);
    const float* o = rays_o + 3 * (size_t)index;
    const float* d = rays_d + 3 * (size_t)index;
    const Ray r = make_ray(o[0], o[1], o[2], d[0], d[1], d[2]);
    const float far = fars[index];
    float t = perturbed_t0(m, rays_t[index], noises[n]);
    float* px = xyzs + 3 * (size_t)n * n_step;
    float* pd = dirs + 3 * (size_t)n * n_step;
    float* pl = deltas + 2 * (size_t)n * n_step;
    uint32_t step = 0;
    Probe p;
    while (step < n_step && march_next(m, r, far, t, p)) {
        px[0] = p.x; px[1] = p.y; px[2] = p.z;
        pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
        t = __fadd_rn(t, p.dt);
        pl[0] = p.dt; pl[1] = t;
        px += 3; pd += 3; pl += 2; step++;
    }
}

// ------------------------------------------------------------------------------------ K12
// raymarching.cu:942-1029
__global__ void k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int* __restrict__ rays_alive,
                                 float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                 const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ depth,
                                 float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float* sg = sigmas + (size_t)n * n_step;
    const float* rg = rgbs + 3 * (size_t)n * n_step;
    const float* dl = deltas + 2 * (size_t)n * n_step;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index];
    float r = image[3 * (size_t)index], g = image[3 * (size_t)index + 1], b = image[3 * (size_t)index + 2];
    uint32_t step = 0;
    while (step < n_step) {
        if (dl[0] == 0) break;
        const float alpha = 1.0f - __expf(-sg[0] * dl[0]);
        const float T = 1 - weight_sum;
        const float weight = alpha * T;
        weight_sum += weight;
        t = dl[1];
        d = fmaf(weight, t, d);
        r = fmaf(weight, rg[0], r);
        g = fmaf(weight, rg[1], g);
        b = fmaf(weight, rg[2], b);
        if (T < T_thresh) break;
        sg++; rg += 3; dl += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
    weights_sum[index] = weight_sum; depth[index] = d;
    image[3 * (size_t)index] = r; image[3 * (size_t)index + 1] = g; image[3 * (size_t)index + 2] = b;
}

}  // namespace gf

// ======================================================================================
// C ABI
// ======================================================================================
using namespace gf;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

GF_API int gf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                                 float* nears, float* fars, gf_stream_t stream) {
    GF_REQUIRE(rays_o && rays_d && aabb && nears && fars, "near_far_from_aabb: null pointer");
    if (N == 0) return GF_OK;
    k_near_far_from_aabb<<<div_up(N, NT), NT, 0, ST(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return check_launch("near_far_from_aabb");
}

GF_API int gf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords, gf_stream_t stream) {
    GF_REQUIRE(rays_o && rays_d && coords, "sph_from_ray: null pointer");
    if (N == 0) return GF_OK;
    k_sph_from_ray<<<div_up(N, NT), NT, 0, ST(stream)>>>(rays_o, rays_d, radius, N, coords);
    return check_launch("sph_from_ray");
}

GF_API int gf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, gf_stream_t stream) {
    GF_REQUIRE(coords && indices, "morton3D: null pointer");
    if (N == 0) return GF_OK;
    k_morton3D<<<div_up(N, NT), NT, 0, ST(stream)>>>(coords, N, indices);
    return check_launch("morton3D");
}

GF_API int gf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, gf_stream_t stream) {
    GF_REQUIRE(coords && indices, "morton3D_invert: null pointer");
    if (N == 0) return GF_OK;
    k_morton3D_invert<<<div_up(N, NT), NT, 0, ST(stream)>>>(indices, N, coords);
    return check_launch("morton3D_invert");
}

GF_API int gf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, gf_stream_t stream) {
    GF_REQUIRE(grid && bitfield, "packbits: null pointer");
    GF_REQUIRE(((uintptr_t)grid & 15) == 0, "packbits: grid must be 16-byte aligned");
    if (N == 0) return GF_OK;
    k_packbits<<<div_up(N, 256), 256, 0, ST(stream)>>>(grid, N, density_thresh, bitfield);
    return check_launch("packbits");
}

GF_API int gf_morton3D_dilation(const float* grid, uint32_t C, uint32_t H, float* grid_dilation, gf_stream_t stream) {
    GF_REQUIRE(grid && grid_dilation, "morton3D_dilation: null pointer");
    GF_REQUIRE(H > 0 && H <= 1024, "morton3D_dilation: H out of range (10-bit morton)");
    const uint32_t total = C * H * H * H;
    if (total == 0) return GF_OK;
    k_morton3D_dilation<<<div_up(total, 256), 256, 0, ST(stream)>>>(grid, C, H, grid_dilation);
    return check_launch("morton3D_dilation");
}

GF_API int gf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                               uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                               const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                               const float* noises, gf_stream_t stream) {
    GF_REQUIRE(rays_o && rays_d && grid && nears && fars && xyzs && dirs && deltas && rays && counter && noises,
               "march_rays_train: null pointer");
    GF_REQUIRE(C >= 1 && C <= 8 && H >= 1 && max_steps >= 1, "march_rays_train: bad C/H/max_steps");
    if (N == 0) return GF_OK;
    // pass 1: per-ray sample counts straight into column 2 of `rays`; pass 2: offsets (column 1) by a single-block scan; pass 3: samples.
    // Row n of `rays` always describes ray n (the reference fills rows in atomic-arrival order, raymarching.cu:452-457).
    k_march_train_count<<<div_up(N, NT), NT, 0, ST(stream)>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H,
                                                               nears, fars, noises, rays);
    int rc = check_launch("march_rays_train(count)");
    if (rc) return rc;
    k_march_train_scan<<<1, 1024, 0, ST(stream)>>>(N, rays, counter, noises);
    rc = check_launch("march_rays_train(scan)");
    if (rc) return rc;
    k_march_train_write<<<div_up(N, NT), NT, 0, ST(stream)>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M,
                                                               nears, fars, noises, rays, counter, xyzs, dirs, deltas);
    return check_launch("march_rays_train(write)");
}

GF_API int gf_march_rays_train_backward(const float* grad_xyzs, const float* grad_dirs, const int32_t* rays, const float* deltas,
                                        uint32_t N, uint32_t M, float* grad_rays_o, float* grad_rays_d, gf_stream_t stream) {
    GF_REQUIRE(grad_xyzs && grad_dirs && rays && deltas && grad_rays_o && grad_rays_d, "march_rays_train_backward: null pointer");
    if (N == 0) return GF_OK;
    k_march_train_backward<<<div_up(N, NT), NT, 0, ST(stream)>>>(grad_xyzs, grad_dirs, rays, deltas, N, M, grad_rays_o, grad_rays_d);
    return check_launch("march_rays_train_backward");
}

GF_API int gf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ambient, const float* deltas,
                                           const int32_t* rays, uint32_t M, uint32_t N, float T_thresh, float* weights_sum,
                                           float* ambient_sum, float* depth, float* image, gf_stream_t stream) {
    GF_REQUIRE(sigmas && rgbs && ambient && deltas && rays && weights_sum && ambient_sum && depth && image,
               "composite_rays_train_forward: null pointer");
    if (N == 0) return GF_OK;
    k_composite_train_fwd<<<div_up(N, NT), NT, 0, ST(stream)>>>(sigmas, rgbs, ambient, deltas, rays, M, N, T_thresh, weights_sum,
                                                                 ambient_sum, depth, image);
    return check_launch("composite_rays_train_forward");
}

GF_API int gf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_ambient_sum, const float* grad_image,
                                            const float* sigmas, const float* rgbs, const float* ambient, const float* deltas,
                                            const int32_t* rays, const float* weights_sum, const float* ambient_sum,
                                            const float* image, uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                            float* grad_rgbs, float* grad_ambient, gf_stream_t stream) {
    (void)ambient; (void)ambient_sum;
    GF_REQUIRE(grad_weights_sum && grad_ambient_sum && grad_image && sigmas && rgbs && deltas && rays && weights_sum && image &&
                   grad_sigmas && grad_rgbs && grad_ambient,
               "composite_rays_train_backward: null pointer");
    if (N == 0) return GF_OK;
    k_composite_train_bwd<<<div_up(N, NT), NT, 0, ST(stream)>>>(grad_weights_sum, grad_ambient_sum, grad_image, sigmas, rgbs, deltas,
                                                                 rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs,
                                                                 grad_ambient);
    return check_launch("composite_rays_train_backward");
}

GF_API int gf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t, const float* rays_o,
                         const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                         const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                         const float* noises, gf_stream_t stream) {
    (void)nears;
    GF_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises, "march_rays: null pointer");
    GF_REQUIRE(C >= 1 && C <= 8 && H >= 1 && max_steps >= 1 && n_step >= 1, "march_rays: bad C/H/max_steps/n_step");
    if (n_alive == 0) return GF_OK;
    k_march_rays<<<div_up(n_alive, NT), NT, 0, ST(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma,
                                                              max_steps, C, H, grid, fars, xyzs, dirs, deltas, noises);
    return check_launch("march_rays");
}

GF_API int gf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                             const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* depth,
                             float* image, gf_stream_t stream) {
    GF_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, "composite_rays: null pointer");
    if (n_alive == 0) return GF_OK;
    k_composite_rays<<<div_up(n_alive, NT), NT, 0, ST(stream)>>>(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas,
                                                                  weights_sum, depth, image);
    return check_launch("composite_rays");
}

}  // extern "C"
