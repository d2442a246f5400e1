// Shared device helpers for libgfrender (sm_100a).
//
// Arithmetic in the occupancy march is written with explicit __f*_rn intrinsics so that the
// rounding sequence is fixed in source (the compiler may neither contract nor un-contract it).
// The sequence reproduces what nvcc emits for the reference kernels
// (modules/radnerfs/raymarching/src/raymarching.cu:42-81, 875-928), which is what makes the
// occupancy-grid indices and per-ray sample counts bit-exact against the reference.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <float.h>

#include "../../include/gfrender.h"

namespace gf {

// ---- error plumbing (host) -------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> GF_ERR_CUDA

#define GF_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            gf::set_error(__VA_ARGS__);       \
            return GF_ERR_INVALID;            \
        }                                     \
    } while (0)

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ---- small math ------------------------------------------------------------------------
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
__device__ __forceinline__ float signf(float x) { return copysignf(1.0f, x); }

__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// exponent of frexpf for finite x >= 0 (0 -> 0), same value the reference's frexpf call yields
__device__ __forceinline__ int frexp_exponent(float x) {
    int e;
    frexpf(x, &e);
    return e;
}

// ---- occupancy march state (one ray) -----------------------------------------------------
struct MarchConst {
    float bound, dt_gamma, dt_min, dt_max;
    float Hf, rH, H3, Cm1;     // (float)H, 1/H, (float)(H^3), (float)(C-1)
    float Hm1;                 // (float)(H-1)
    const uint8_t* grid;
};

__device__ __forceinline__ MarchConst make_march_const(float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                                       uint32_t H, const uint8_t* grid) {
    MarchConst m;
    m.bound = bound;
    m.dt_gamma = dt_gamma;
    // raymarching.cu:866-867
    m.dt_max = __fdiv_rn(__fmul_rn((float)(1 << (C - 1)), 3.4641015529632568359f), (float)H);
    m.dt_min = fminf(m.dt_max, __fdiv_rn(3.4641015529632568359f, (float)max_steps));
    m.Hf = (float)H;
    m.rH = __fdiv_rn(1.0f, (float)H);
    m.H3 = (float)(H * H * H);
    m.Cm1 = (float)C - 1.0f;
    m.Hm1 = (float)(H - 1);
    m.grid = grid;
    return m;
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
};

__device__ __forceinline__ Ray make_ray(float ox, float oy, float oz, float dx, float dy, float dz) {
    Ray r;
    r.ox = ox; r.oy = oy; r.oz = oz; r.dx = dx; r.dy = dy; r.dz = dz;
    r.rdx = __fdiv_rn(1.0f, dx); r.rdy = __fdiv_rn(1.0f, dy); r.rdz = __fdiv_rn(1.0f, dz);
    return r;
}

struct Probe {
    float x, y, z, dt;
    uint32_t index;   // bit index into the occupancy bitfield
};

// Advance `t` until the next OCCUPIED sample (or t >= far).  Returns true and fills `p` when an
// occupied sample was found at the returned t (t is NOT yet advanced past it); false when the ray
// left the volume.  Mirrors one trip of the while-loop body at raymarching.cu:875-928.
__device__ __forceinline__ bool march_next(const MarchConst& m, const Ray& r, float far, float& t, Probe& p) {
    while (t < far) {
        const float x = clampf(__fmaf_rn(r.dx, t, r.ox), -m.bound, m.bound);
        const float y = clampf(__fmaf_rn(r.dy, t, r.oy), -m.bound, m.bound);
        const float z = clampf(__fmaf_rn(r.dz, t, r.oz), -m.bound, m.bound);
        const float dt = clampf(__fmul_rn(t, m.dt_gamma), m.dt_min, m.dt_max);
        // mip level: max(mip_from_pos, mip_from_dt), each clamped to [0, C-1]
        const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        const int l1 = (int)fminf(m.Cm1, fmaxf(0.0f, (float)frexp_exponent(mx)));
        const int l2 = (int)fminf(m.Cm1, fmaxf(0.0f, (float)frexp_exponent(__fmul_rn(__fmul_rn(dt, m.Hf), 0.5f))));
        const int level = max(l1, l2);
        const float mip_bound = fminf(scalbnf(1.0f, level), m.bound);
        const float mip_rbound = __fdiv_rn(1.0f, mip_bound);
        // (int)clamp(0.5*(x*rb+1)*H, 0, H-1): the reference's double detour is exact in float
        // because 0.5*v is exact and v*H is a single correctly rounded product.
        const float fx = __fmul_rn(__fmul_rn(0.5f, __fmaf_rn(x, mip_rbound, 1.0f)), m.Hf);
        const float fy = __fmul_rn(__fmul_rn(0.5f, __fmaf_rn(y, mip_rbound, 1.0f)), m.Hf);
        const float fz = __fmul_rn(__fmul_rn(0.5f, __fmaf_rn(z, mip_rbound, 1.0f)), m.Hf);
        const int nx = (int)clampf(fx, 0.0f, m.Hm1);
        const int ny = (int)clampf(fy, 0.0f, m.Hm1);
        const int nz = (int)clampf(fz, 0.0f, m.Hm1);
        const uint32_t index = (uint32_t)__fmaf_rn((float)level, m.H3, (float)morton3D(nx, ny, nz));
        const bool occ = (__ldg(m.grid + (index >> 3)) >> (index & 7)) & 1;
        if (occ) {
            p.x = x; p.y = y; p.z = z; p.dt = dt; p.index = index;
            return true;
        }
        // distance to the next voxel boundary (raymarching.cu:919-926)
        const float ax = __fmul_rn(__fadd_rn(__fadd_rn((float)nx, 0.5f), __fmul_rn(0.5f, signf(r.dx))), m.rH);
        const float ay = __fmul_rn(__fadd_rn(__fadd_rn((float)ny, 0.5f), __fmul_rn(0.5f, signf(r.dy))), m.rH);
        const float az = __fmul_rn(__fadd_rn(__fadd_rn((float)nz, 0.5f), __fmul_rn(0.5f, signf(r.dz))), m.rH);
        const float tx = __fmul_rn(__fmaf_rn(__fmaf_rn(ax, 2.0f, -1.0f), mip_bound, -x), r.rdx);
        const float ty = __fmul_rn(__fmaf_rn(__fmaf_rn(ay, 2.0f, -1.0f), mip_bound, -y), r.rdy);
        const float tz = __fmul_rn(__fmaf_rn(__fmaf_rn(az, 2.0f, -1.0f), mip_bound, -z), r.rdz);
        const float tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
        do {
            t = __fadd_rn(t, clampf(__fmul_rn(t, m.dt_gamma), m.dt_min, m.dt_max));
        } while (t < tt);
    }
    return false;
}

// slab test, raymarching.cu:91-145.  Miss => near = far = FLT_MAX.
__device__ __forceinline__ void near_far_aabb(const Ray& r, const float* __restrict__ aabb, float min_near, float& near_o,
                                              float& far_o) {
    float near = __fmul_rn(__fsub_rn(aabb[0], r.ox), r.rdx);
    float far = __fmul_rn(__fsub_rn(aabb[3], r.ox), r.rdx);
    if (near > far) { float c = near; near = far; far = c; }
    float near_y = __fmul_rn(__fsub_rn(aabb[1], r.oy), r.rdy);
    float far_y = __fmul_rn(__fsub_rn(aabb[4], r.oy), r.rdy);
    if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
    if (near > far_y || near_y > far) { near_o = far_o = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(__fsub_rn(aabb[2], r.oz), r.rdz);
    float far_z = __fmul_rn(__fsub_rn(aabb[5], r.oz), r.rdz);
    if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
    if (near > far_z || near_z > far) { near_o = far_o = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_o = near;
    far_o = far;
}

}  // namespace gf
