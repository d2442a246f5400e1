"""ctypes/numpy front-end of oracle/gf_oracle.c  (TEST INFRASTRUCTURE ONLY).

Mirrors the argument order of the reference's pybind functions
(modules/radnerfs/raymarching/src/raymarching.h:7-20,
 encoders/gridencoder/src/gridencoder.h:11-14, shencoder.h, freqencoder.h)
but on numpy arrays.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgforacle.so")
_SRC = os.path.join(_HERE, "gf_oracle.c")


def build(force=False):
    """gcc the C restatement into oracle/libgforacle.so (a few seconds)."""
    if (not force) and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


u32 = ctypes.c_uint32
f32 = ctypes.c_float
cint = ctypes.c_int


# ----------------------------------------------------------------- raymarching
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3), _f(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().gfo_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().gfo_sph_from_ray(_p(rays_o), _p(rays_d), f32(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i(coords)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().gfo_morton3D(_p(coords), u32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i(indices)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().gfo_morton3D_invert(_p(indices), u32(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().gfo_packbits(_p(grid), u32(N), f32(thresh), _p(out))
    return out


def morton3D_dilation(grid):
    grid = _f(grid)
    C, H3 = grid.shape
    H = int(round(H3 ** (1 / 3)))
    out = np.empty_like(grid)
    lib().gfo_morton3D_dilation(_p(grid), u32(C), u32(H), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, dt_gamma, max_steps, M=None):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    if M is None:
        M = N * max_steps
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    lib().gfo_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32(bound), f32(dt_gamma), u32(max_steps),
                               u32(N), u32(C), u32(H), u32(M), _p(_f(nears)), _p(_f(fars)),
                               _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter), _p(_f(noises)))
    return xyzs, dirs, deltas, rays, counter


def march_rays_train_backward(grad_xyzs, grad_dirs, rays, deltas):
    grad_xyzs, grad_dirs, rays, deltas = _f(grad_xyzs), _f(grad_dirs), _i(rays), _f(deltas)
    N, M = rays.shape[0], grad_xyzs.shape[0]
    go, gd = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    lib().gfo_march_rays_train_backward(_p(grad_xyzs), _p(grad_dirs), _p(rays), _p(deltas), u32(N), u32(M), _p(go), _p(gd))
    return go, gd


def composite_rays_train_forward(sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, ambient, deltas, rays = _f(sigmas), _f(rgbs), _f(ambient), _f(deltas), _i(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, amb, depth = np.empty(N, np.float32), np.empty(N, np.float32), np.empty(N, np.float32)
    image = np.empty((N, 3), np.float32)
    lib().gfo_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(ambient), _p(deltas), _p(rays), u32(M), u32(N),
                                           f32(T_thresh), _p(ws), _p(amb), _p(depth), _p(image))
    return ws, amb, depth, image


def composite_rays_train_backward(g_ws, g_amb, g_img, sigmas, rgbs, ambient, deltas, rays, ws, amb, image, T_thresh=1e-4):
    sigmas, rgbs, ambient, deltas, rays = _f(sigmas), _f(rgbs), _f(ambient), _f(deltas), _i(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr, ga = np.zeros(M, np.float32), np.zeros((M, 3), np.float32), np.zeros(M, np.float32)
    lib().gfo_composite_rays_train_backward(_p(_f(g_ws)), _p(_f(g_amb)), _p(_f(g_img)), _p(sigmas), _p(rgbs), _p(ambient),
                                            _p(deltas), _p(rays), _p(_f(ws)), _p(_f(amb)), _p(_f(image)),
                                            u32(M), u32(N), f32(T_thresh), _p(gs), _p(gr), _p(ga))
    return gs, gr, ga


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars,
               align=-1, noises=None, dt_gamma=0.0, max_steps=1024, with_indices=False):
    rays_o, rays_d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    if noises is None:
        noises = np.zeros(n_alive, np.float32)
    rays_alive, rays_t = _i(rays_alive), _f(rays_t)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    lib().gfo_march_rays(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                         f32(bound), f32(dt_gamma), u32(max_steps), u32(C), u32(H), _p(bitfield),
                         _p(_f(nears)), _p(_f(fars)), _p(xyzs), _p(dirs), _p(deltas), _p(_f(noises)))
    if with_indices:
        idx = np.empty((n_alive, n_step), np.int32)
        lib().gfo_march_rays_indices(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                                     f32(bound), f32(dt_gamma), u32(max_steps), u32(C), u32(H), _p(bitfield),
                                     _p(_f(fars)), _p(_f(noises)), _p(idx))
        return xyzs, dirs, deltas, idx
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In place on rays_alive, rays_t, weights_sum, depth, image (all must be contiguous numpy of the right dtype).
    Returns the number of samples each alive ray composited in this call."""
    assert rays_alive.dtype == np.int32 and rays_t.dtype == np.float32
    composited = np.zeros(n_alive, np.int32)
    lib().gfo_composite_rays_counted(u32(n_alive), u32(n_step), f32(T_thresh), _p(rays_alive), _p(rays_t),
                                     _p(_f(sigmas)), _p(_f(rgbs)), _p(_f(deltas)), _p(weights_sum), _p(depth), _p(image),
                                     _p(composited))
    return composited


# -------------------------------------------------------------------- encoders
def grid_encode_forward(inputs, embeddings, offsets, S, H, calc_grad_inputs=False, gridtype=0, align_corners=False, interp=0):
    """Returns outputs [L,B,C] (reference layout before the permute, grid.py:47) and dy_dx [B, L*D*C] or None."""
    inputs, embeddings, offsets = _f(inputs), _f(embeddings), _i(offsets)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    out = np.empty((L, B, C), np.float32)
    dy_dx = np.empty((B, L * D * C), np.float32) if calc_grad_inputs else None
    lib().gfo_grid_encode_forward(_p(inputs), _p(embeddings), _p(offsets), _p(out), u32(B), u32(D), u32(C), u32(L),
                                  f32(S), u32(H), _p(dy_dx), u32(gridtype), cint(int(align_corners)), u32(interp))
    return out, dy_dx


def grid_encode_backward(grad, inputs, embeddings, offsets, S, H, dy_dx=None, gridtype=0, align_corners=False, interp=0):
    grad, inputs, embeddings, offsets = _f(grad), _f(inputs), _f(embeddings), _i(offsets)
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    gg = np.zeros_like(embeddings)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    lib().gfo_grid_encode_backward(_p(grad), _p(inputs), _p(embeddings), _p(offsets), _p(gg), u32(B), u32(D), u32(C), u32(L),
                                   f32(S), u32(H), _p(_f(dy_dx)) if dy_dx is not None else None, _p(gi),
                                   u32(gridtype), cint(int(align_corners)), u32(interp))
    return gg, gi


def grad_total_variation(inputs, embeddings, grad, offsets, weight, S, H, gridtype=0, align_corners=False):
    inputs, embeddings, offsets = _f(inputs), _f(embeddings), _i(offsets)
    assert grad.dtype == np.float32 and grad.flags.c_contiguous
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    lib().gfo_grad_total_variation(_p(inputs), _p(embeddings), _p(grad), _p(offsets), f32(weight), u32(B), u32(D), u32(C),
                                   u32(L), f32(S), u32(H), u32(gridtype), cint(int(align_corners)))
    return grad


def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    inputs = _f(inputs)
    B, D = inputs.shape
    out = np.empty((B, degree * degree), np.float32)
    dy_dx = np.empty((B, D * degree * degree), np.float32) if calc_grad_inputs else None
    lib().gfo_sh_encode_forward(_p(inputs), _p(out), u32(B), u32(D), u32(degree), _p(dy_dx))
    return out, dy_dx


def sh_encode_backward(grad, inputs, degree, dy_dx):
    grad, inputs, dy_dx = _f(grad), _f(inputs), _f(dy_dx)
    B, D = inputs.shape
    gi = np.zeros((B, D), np.float32)
    lib().gfo_sh_encode_backward(_p(grad), _p(inputs), u32(B), u32(D), u32(degree), _p(dy_dx), _p(gi))
    return gi


def freq_encode_forward(inputs, degree, output_dim):
    inputs = _f(inputs)
    B, D = inputs.shape
    out = np.empty((B, output_dim), np.float32)
    lib().gfo_freq_encode_forward(_p(inputs), u32(B), u32(D), u32(degree), u32(output_dim), _p(out))
    return out


def freq_encode_backward(grad, outputs, degree, input_dim):
    grad, outputs = _f(grad), _f(outputs)
    B, C = outputs.shape
    gi = np.zeros((B, input_dim), np.float32)
    lib().gfo_freq_encode_backward(_p(grad), _p(outputs), u32(B), u32(input_dim), u32(degree), u32(C), _p(gi))
    return gi
