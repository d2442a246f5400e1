// Device-side building blocks of the RAD-NeRF field shared by the fused kernels:
// per-level grid geometry, C=2 fp32 tiled/hash grid sampling (same rounding sequence as
// k_grid_forward / reference gridencoder.cu:87-196), degree-4 SH, and the packed model layout.
#pragma once
#include "gf_common.cuh"

namespace gf {

// ---- per-level geometry, computed ON DEVICE once per model (exp2f must be the GPU's) ----------
struct GridLevels {
    float scale[16];
    uint32_t res[16];      // resolution = ceil(scale)+1
    uint32_t hsize[16];    // entries in the level
    uint32_t offset[16];   // first entry of the level
};

struct GridDesc {
    const float2* table;   // [sum hsize] entries of 2 floats
    GridLevels lv;
    uint32_t gridtype;     // 0 hash, 1 tiled
    uint32_t interp;       // 0 linear, 1 smoothstep
};

__device__ __forceinline__ uint32_t grid_index3(uint32_t gridtype, uint32_t hs, uint32_t res, uint32_t x, uint32_t y, uint32_t z) {
    // gridencoder.cu:66-84 with D=3, align_corners=false
    uint32_t stride = 1, index = 0;
    const uint32_t r1 = res + 1;
    if (stride <= hs) { index += x * stride; stride *= r1; }
    if (stride <= hs) { index += y * stride; stride *= r1; }
    if (stride <= hs) { index += z * stride; stride *= r1; }
    if (gridtype == 0 && stride > hs) index = (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);
    return index % hs;
}

__device__ __forceinline__ uint32_t grid_index2(uint32_t gridtype, uint32_t hs, uint32_t res, uint32_t x, uint32_t y) {
    uint32_t stride = 1, index = 0;
    const uint32_t r1 = res + 1;
    if (stride <= hs) { index += x * stride; stride *= r1; }
    if (stride <= hs) { index += y * stride; stride *= r1; }
    if (gridtype == 0 && stride > hs) index = (x * 1u) ^ (y * 2654435761u);
    return index % hs;
}

__device__ __forceinline__ float smooth_(float v) { return v * v * (3.0f - 2.0f * v); }

// x,y,z already mapped to [0,1] (grid.py:149).  Returns the 2 interpolated channels of `level`.
__device__ __forceinline__ float2 grid3_sample(const GridDesc& g, int level, float x, float y, float z) {
    if (x < 0 || x > 1 || y < 0 || y > 1 || z < 0 || z > 1) return make_float2(0.f, 0.f);
    const float scale = g.lv.scale[level];
    const uint32_t res = g.lv.res[level], hs = g.lv.hsize[level];
    const float2* __restrict__ tab = g.table + g.lv.offset[level];
    float px = __fmaf_rn(x, scale, 0.5f), py = __fmaf_rn(y, scale, 0.5f), pz = __fmaf_rn(z, scale, 0.5f);
    const uint32_t gx = (uint32_t)floorf(px), gy = (uint32_t)floorf(py), gz = (uint32_t)floorf(pz);
    px = __fsub_rn(px, (float)gx); py = __fsub_rn(py, (float)gy); pz = __fsub_rn(pz, (float)gz);
    if (g.interp == 1) { px = smooth_(px); py = smooth_(py); pz = smooth_(pz); }
    const float qx = __fsub_rn(1.0f, px), qy = __fsub_rn(1.0f, py), qz = __fsub_rn(1.0f, pz);
    float2 v[8];
    #pragma unroll
    for (int i = 0; i < 8; i++)
        v[i] = __ldg(tab + grid_index3(g.gridtype, hs, res, gx + (i & 1), gy + ((i >> 1) & 1), gz + (i >> 2)));
    float r0 = 0.f, r1 = 0.f;
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        // w = ((1 * wx) * wy) * wz in the reference's order (d = 0,1,2)
        const float w = __fmul_rn(__fmul_rn((i & 1) ? px : qx, ((i >> 1) & 1) ? py : qy), (i >> 2) ? pz : qz);
        r0 = __fmaf_rn(w, v[i].x, r0);
        r1 = __fmaf_rn(w, v[i].y, r1);
    }
    return make_float2(r0, r1);
}

__device__ __forceinline__ float2 grid2_sample(const GridDesc& g, int level, float x, float y) {
    if (x < 0 || x > 1 || y < 0 || y > 1) return make_float2(0.f, 0.f);
    const float scale = g.lv.scale[level];
    const uint32_t res = g.lv.res[level], hs = g.lv.hsize[level];
    const float2* __restrict__ tab = g.table + g.lv.offset[level];
    float px = __fmaf_rn(x, scale, 0.5f), py = __fmaf_rn(y, scale, 0.5f);
    const uint32_t gx = (uint32_t)floorf(px), gy = (uint32_t)floorf(py);
    px = __fsub_rn(px, (float)gx); py = __fsub_rn(py, (float)gy);
    if (g.interp == 1) { px = smooth_(px); py = smooth_(py); }
    const float qx = __fsub_rn(1.0f, px), qy = __fsub_rn(1.0f, py);
    float2 v[4];
    #pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __ldg(tab + grid_index2(g.gridtype, hs, res, gx + (i & 1), gy + (i >> 1)));
    float r0 = 0.f, r1 = 0.f;
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const float w = __fmul_rn((i & 1) ? px : qx, (i >> 1) ? py : qy);
        r0 = __fmaf_rn(w, v[i].x, r0);
        r1 = __fmaf_rn(w, v[i].y, r1);
    }
    return make_float2(r0, r1);
}

// map a world coordinate in [-bound, bound] to [0,1] exactly as grid.py:149 does in fp32
__device__ __forceinline__ float to_unit(float x, float bound) { return __fdiv_rn(__fadd_rn(x, bound), 2.0f * bound); }

// Real SH, degree 4 (16 values), Condon-Shortley phase, unit-norm input assumed
// (values of shencoder.cu:43-68; written from the definition with shared sub-terms).
__device__ __forceinline__ void sh4(float x, float y, float z, float* o) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * zz - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * (xx - yy);
    const float a = yy - 3.0f * xx;       // -3x^2 + y^2
    const float b = 3.0f * yy - xx;       // -x^2 + 3y^2
    const float c = 1.0f - 5.0f * zz;
    o[9] = 0.59004358992664352f * y * a;
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * c;
    o[12] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
    o[13] = 0.45704579946446572f * x * c;
    o[14] = 1.4453057213202769f * z * (xx - yy);
    o[15] = 0.59004358992664352f * x * b;
}

}  // namespace gf
