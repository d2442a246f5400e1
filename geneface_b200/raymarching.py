"""Drop-in for modules/radnerfs/raymarching/raymarching.py (reference @ 15ff4e5c).

Same ten public callables, same positional arguments, same returned tensors; each calls the
sm_100a kernels in libgfrender.so through the C ABI (include/gfrender.h) on torch's current
stream.  Outputs are allocated here in torch and passed in, exactly as the reference wrappers do
(raymarching.py:41-44, 230-233, 307-310, 384-386).
"""
import numpy as np
import torch
from torch.autograd import Function

from . import _lib
from ._lib import c_f32, c_u32, check, ptr, stream_ptr


def _cuda_f32(t):
    if not t.is_cuda:
        t = t.cuda()
    return t.float().contiguous()


class _near_far_from_aabb(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """raymarching.py:18-46 -> nears [N], fars [N]"""
        rays_o = _cuda_f32(rays_o).view(-1, 3)
        rays_d = _cuda_f32(rays_d).view(-1, 3)
        aabb = _cuda_f32(aabb)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        check(_lib.lib().gf_near_far_from_aabb(ptr(rays_o), ptr(rays_d), ptr(aabb), N, c_f32(min_near), ptr(nears), ptr(fars),
                                               stream_ptr()), "near_far_from_aabb")
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, radius):
        """raymarching.py:51-78 -> coords [N,2]"""
        rays_o = _cuda_f32(rays_o).view(-1, 3)
        rays_d = _cuda_f32(rays_d).view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
        check(_lib.lib().gf_sph_from_ray(ptr(rays_o), ptr(rays_d), c_f32(radius), N, ptr(coords), stream_ptr()), "sph_from_ray")
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """raymarching.py:83-101"""
        if not coords.is_cuda:
            coords = coords.cuda()
        coords = coords.int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        check(_lib.lib().gf_morton3D(ptr(coords), N, ptr(indices), stream_ptr()), "morton3D")
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """raymarching.py:105-124"""
        if not indices.is_cuda:
            indices = indices.cuda()
        indices = indices.int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        check(_lib.lib().gf_morton3D_invert(ptr(indices), N, ptr(coords), stream_ptr()), "morton3D_invert")
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    def forward(ctx, grid, thresh, bitfield=None):
        """raymarching.py:129-154"""
        grid = _cuda_f32(grid)
        C, H3 = grid.shape
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        check(_lib.lib().gf_packbits(ptr(grid), N, c_f32(thresh), ptr(bitfield), stream_ptr()), "packbits")
        return bitfield


packbits = _packbits.apply


class _morton3D_dilation(Function):
    @staticmethod
    def forward(ctx, grid):
        """raymarching.py:159-180"""
        grid = _cuda_f32(grid)
        C, H3 = grid.shape
        H = int(np.cbrt(H3) + 0.5)
        out = torch.empty_like(grid)
        check(_lib.lib().gf_morton3D_dilation(ptr(grid), C, H, ptr(out), stream_ptr()), "morton3D_dilation")
        return out


morton3D_dilation = _morton3D_dilation.apply


class _march_rays_train(Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        """raymarching.py:186-258.  Layout differs from the reference only in being deterministic: row n of `rays` is ray n (the
        reference's row order is whatever its atomics produce) and sample offsets are an exclusive scan of the per-ray counts
        starting at ray rot = bits(noises[0]) % N.  When the sample total exceeds M = mean_count (the normal training regime,
        raymarching.py:225-228) the rays that lose their samples are therefore a pseudo-random contiguous run that changes every
        step -- like the reference's race losers -- not systematically the highest ray indices."""
        rays_o = _cuda_f32(rays_o).view(-1, 3)
        rays_d = _cuda_f32(rays_d).view(-1, 3)
        if not density_bitfield.is_cuda:
            density_bitfield = density_bitfield.cuda()
        density_bitfield = density_bitfield.contiguous()
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else torch.zeros(N, dtype=torch.float32, device=dev)
        nears, fars = nears.float().contiguous(), fars.float().contiguous()   # named: must outlive the launch
        check(_lib.lib().gf_march_rays_train(ptr(rays_o), ptr(rays_d), ptr(density_bitfield), c_f32(bound), c_f32(dt_gamma),
                                             max_steps, N, C, H, M, ptr(nears), ptr(fars),
                                             ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(step_counter), ptr(noises),
                                             stream_ptr()), "march_rays_train")
        if force_all_rays or mean_count <= 0:
            m = step_counter[0].item()
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        ctx.save_for_backward(rays, deltas)
        return xyzs, dirs, deltas, rays

    @staticmethod
    def backward(ctx, grad_xyzs, grad_dirs, grad_deltas, grad_rays):
        rays, deltas = ctx.saved_tensors
        N, M = rays.shape[0], grad_xyzs.shape[0]
        grad_rays_o = torch.zeros(N, 3, device=rays.device)
        grad_rays_d = torch.zeros(N, 3, device=rays.device)
        grad_xyzs, grad_dirs, deltas = grad_xyzs.float().contiguous(), grad_dirs.float().contiguous(), deltas.contiguous()
        check(_lib.lib().gf_march_rays_train_backward(ptr(grad_xyzs), ptr(grad_dirs),
                                                      ptr(rays), ptr(deltas), N, M, ptr(grad_rays_o), ptr(grad_rays_d),
                                                      stream_ptr()), "march_rays_train_backward")
        return (grad_rays_o, grad_rays_d) + (None,) * 13


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    def forward(ctx, sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
        """raymarching.py:283-315"""
        sigmas, rgbs, ambient = sigmas.float().contiguous(), rgbs.float().contiguous(), ambient.float().contiguous()
        deltas = deltas.float().contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        ambient_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        check(_lib.lib().gf_composite_rays_train_forward(ptr(sigmas), ptr(rgbs), ptr(ambient), ptr(deltas), ptr(rays), M, N,
                                                         c_f32(T_thresh), ptr(weights_sum), ptr(ambient_sum), ptr(depth), ptr(image),
                                                         stream_ptr()), "composite_rays_train_forward")
        ctx.save_for_backward(sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, ambient_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_ambient_sum, grad_depth, grad_image):
        sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        grad_ambient = torch.zeros_like(ambient)
        gws, gas, gim = grad_weights_sum.float().contiguous(), grad_ambient_sum.float().contiguous(), grad_image.float().contiguous()
        check(_lib.lib().gf_composite_rays_train_backward(
            ptr(gws), ptr(gas), ptr(gim), ptr(sigmas), ptr(rgbs), ptr(ambient), ptr(deltas), ptr(rays), ptr(weights_sum),
            ptr(ambient_sum), ptr(image), M, N, c_f32(T_thresh), ptr(grad_sigmas), ptr(grad_rgbs), ptr(grad_ambient),
            stream_ptr()), "composite_rays_train_backward")
        return grad_sigmas, grad_rgbs, grad_ambient, None, None, None


composite_rays_train = _composite_rays_train.apply


class _march_rays(Function):
    @staticmethod
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024):
        """raymarching.py:347-396"""
        rays_o = _cuda_f32(rays_o).view(-1, 3)
        rays_d = _cuda_f32(rays_d).view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        dev = rays_o.device
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb else torch.zeros(n_alive, dtype=torch.float32, device=dev)
        check(_lib.lib().gf_march_rays(n_alive, n_step, ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), c_f32(bound),
                                       c_f32(dt_gamma), max_steps, C, H, ptr(density_bitfield), ptr(near), ptr(far), ptr(xyzs),
                                       ptr(dirs), ptr(deltas), ptr(noises), stream_ptr()), "march_rays")
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        """raymarching.py:401-420 (in place on rays_alive, rays_t, weights_sum, depth, image)"""
        sigmas, rgbs = sigmas.float().contiguous(), rgbs.float().contiguous()   # named: must outlive the launch
        check(_lib.lib().gf_composite_rays(n_alive, n_step, c_f32(T_thresh), ptr(rays_alive), ptr(rays_t),
                                           ptr(sigmas), ptr(rgbs), ptr(deltas),
                                           ptr(weights_sum), ptr(depth), ptr(image), stream_ptr()), "composite_rays")
        return tuple()


composite_rays = _composite_rays.apply
