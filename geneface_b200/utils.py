"""Host-side helpers of the render path (drop-in for the functions the path imports from
modules/radnerfs/utils.py; none of that file's unused heavy imports -- trimesh, mcubes, lpips,
tensorboardX, imageio, matplotlib -- are needed here)."""
import math

import numpy as np
import torch
from torch.autograd import Function


class _trunc_exp(Function):
    """utils.py:36-49: exp forward, gradient through exp(clamp(x, -15, 15))."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply


def nerf_matrix_to_ngp(pose, scale=4, offset=(0, 0, 0)):
    """utils.py:53-60: axis permutation + scale of the camera-to-world matrix."""
    p = np.asarray(pose, dtype=np.float32)
    out = np.eye(4, dtype=np.float32)
    for r, src in enumerate((1, 2, 0)):
        out[r, 0] = p[src, 0]
        out[r, 1] = -p[src, 1]
        out[r, 2] = -p[src, 2]
        out[r, 3] = p[src, 3] * scale + offset[r]
    return out


def custom_meshgrid(*args):
    return torch.meshgrid(*args, indexing='ij')


def get_audio_features(features, att_mode, index, smo_win_size=5):
    """Window of per-frame condition features around frame `index` (utils.py:71-104; the reference reads smo_win_size from the
    global hparams).  att_mode 0: the frame itself; 1: the `smo_win_size` frames before it; 2: a centred window.  Frames outside
    the sequence are zero rows.  Built by index arithmetic (one gather + one mask), no concatenation."""
    if att_mode == 0:
        return features[[index]]
    if att_mode == 1:
        lo, hi = index - smo_win_size, index
    elif att_mode == 2:
        lo, hi = index - smo_win_size // 2, index + smo_win_size - smo_win_size // 2
    else:
        raise NotImplementedError(f'wrong att_mode: {att_mode}')
    n = features.shape[0]
    idx = torch.arange(lo, hi, device=features.device)
    inside = (idx >= 0) & (idx < n)
    win = features[idx.clamp(0, n - 1)]
    return win * inside.view(-1, *([1] * (features.dim() - 1))).to(win.dtype)


def matrix_to_euler_angles_xyz(matrix):
    """utils.py:160-199 specialised to convention 'XYZ' (the only one the path uses)."""
    central = torch.asin(matrix[..., 0, 2])
    a0 = torch.atan2(-matrix[..., 1, 2], matrix[..., 2, 2])
    a2 = torch.atan2(-matrix[..., 0, 1], matrix[..., 0, 0])
    return torch.stack((a0, central, a2), -1)


def convert_poses(poses):
    """utils.py:263-269: [B,4,4] c2w -> [B,6] = (euler XYZ, translation)."""
    out = torch.empty(poses.shape[0], 6, dtype=torch.float32, device=poses.device)
    out[:, :3] = matrix_to_euler_angles_xyz(poses[:, :3, :3].float())
    out[:, 3:] = poses[:, :3, 3]
    return out


def get_bg_coords(H, W, device):
    """utils.py:273-278"""
    X = torch.arange(H, device=device) / (H - 1) * 2 - 1
    Y = torch.arange(W, device=device) / (W - 1) * 2 - 1
    xs, ys = custom_meshgrid(X, Y)
    return torch.cat([xs.reshape(-1, 1), ys.reshape(-1, 1)], dim=-1).unsqueeze(0)


def pixel_indices(H, W, N=-1, patch_size=1, rect=None, device='cpu'):
    """Flat (row-major, y * W + x) pixel indices of one ray batch, or None for the whole image -- the three sampling modes of the
    reference's get_rays (utils.py:303-341), by index arithmetic: random pixels (with repetition), random patch_size^2 patches
    (top-left corners drawn rows first, then columns, so a seeded torch RNG selects the same pixels as the reference), or every
    pixel of the rows x columns box rect = (r0, r1, c0, c1) in ascending order."""
    if rect is not None:
        r0, r1, c0, c1 = rect
        return (torch.arange(r0, r1, device=device).view(-1, 1) * W + torch.arange(c0, c1, device=device).view(1, -1)).reshape(-1)
    if N <= 0:
        return None
    N = min(N, H * W)
    if patch_size > 1:
        k = N // (patch_size ** 2)
        top = torch.randint(0, H - patch_size, size=[k], device=device)
        left = torch.randint(0, W - patch_size, size=[k], device=device)
        ar = torch.arange(patch_size, device=device)
        return ((top.view(-1, 1, 1) + ar.view(1, -1, 1)) * W + (left.view(-1, 1, 1) + ar.view(1, 1, -1))).reshape(-1)
    return torch.randint(0, H * W, size=[N], device=device)


def get_rays(poses, intrinsics, H, W, N=-1, patch_size=1, rect=None):
    """Drop-in for modules/radnerfs/utils.py:282-363: poses [B,4,4] c2w (CUDA), intrinsics (fx, fy, cx, cy) ->
    dict(rays_o [B,n,3], rays_d [B,n,3], inds [B,n], i [B,n], j [B,n]).  The rays come from the `gf_get_rays` operator of
    libgfrender (the same pixel -> ray arithmetic as the fused renderer's in-kernel ray generation)."""
    from . import _lib
    _lib.require_cuda()
    device = poses.device
    B = poses.shape[0]
    inds = pixel_indices(H, W, N, patch_size, rect, device)
    n = H * W if inds is None else inds.numel()
    P = poses.detach().float().contiguous()
    rays_o = torch.empty(B, n, 3, dtype=torch.float32, device=device)
    rays_d = torch.empty(B, n, 3, dtype=torch.float32, device=device)
    pi = torch.empty(n, dtype=torch.float32, device=device)
    pj = torch.empty(n, dtype=torch.float32, device=device)
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    idx = None if inds is None else inds.to(torch.int64).contiguous()
    _lib.check(_lib.lib().gf_get_rays(_lib.ptr(P), B, fx, fy, cx, cy, H, W, _lib.ptr(idx), n, _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(pi),
                                      _lib.ptr(pj), _lib.stream_ptr()), "gf_get_rays")
    if inds is None:
        inds = torch.arange(H * W, device=device)
    return {'rays_o': rays_o, 'rays_d': rays_d, 'inds': inds.expand(B, n), 'i': pi.expand(B, n), 'j': pj.expand(B, n)}


def orbit_pose(radius=3.35, yaw_deg=0.0):
    """The config-2 camera of SURVEY.md section 8d: OrbitCamera(r, fovy) default pose (radnerf_gui.py:21-64)
    = [[0,-1,0,0],[0,0,-1,r],[1,0,0,0],[0,0,0,1]], optionally yawed about the world up axis (y)."""
    base = np.array([[0, -1, 0, 0], [0, 0, -1, radius], [1, 0, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
    a = math.radians(yaw_deg)
    Ry = np.array([[math.cos(a), 0, math.sin(a), 0], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0], [0, 0, 0, 1]])
    return (Ry @ base).astype(np.float32)


def intrinsics_from_fovy(H, W, fovy_deg=21.24):
    f = H / (2 * math.tan(math.radians(fovy_deg) / 2))
    return (f, f, W / 2, H / 2)
