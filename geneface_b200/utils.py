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
    """utils.py:71-104 (the reference reads smo_win_size from the global hparams)."""
    if att_mode == 0:
        return features[[index]]
    if att_mode == 1:
        left = index - smo_win_size
        pad_left = max(0, -left)
        left = max(left, 0)
        auds = features[left:index]
        if pad_left > 0:
            auds = torch.cat([torch.zeros(pad_left, *auds.shape[1:], device=auds.device, dtype=auds.dtype), auds], dim=0)
        return auds
    if att_mode == 2:
        left = index - smo_win_size // 2
        right = index + (smo_win_size - smo_win_size // 2)
        pad_left = max(0, -left)
        left = max(left, 0)
        pad_right = max(0, right - features.shape[0])
        right = min(right, features.shape[0])
        auds = features[left:right]
        if pad_left > 0:
            auds = torch.cat([torch.zeros_like(auds[:pad_left]), auds], dim=0)
        if pad_right > 0:
            auds = torch.cat([auds, torch.zeros_like(auds[:pad_right])], dim=0)
        return auds
    raise NotImplementedError(f'wrong att_mode: {att_mode}')


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


def get_rays(poses, intrinsics, H, W, N=-1, patch_size=1, rect=None):
    """utils.py:282-363.  poses [B,4,4] c2w, intrinsics (fx,fy,cx,cy) -> dict(rays_o, rays_d, inds, i, j)."""
    device = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = intrinsics
    if rect is not None:
        xmin, xmax, ymin, ymax = rect
        N = (xmax - xmin) * (ymax - ymin)
    i, j = custom_meshgrid(torch.linspace(0, W - 1, W, device=device), torch.linspace(0, H - 1, H, device=device))
    i = i.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    j = j.t().reshape([1, H * W]).expand([B, H * W]) + 0.5
    results = {}
    if N > 0:
        N = min(N, H * W)
        if patch_size > 1:
            num_patch = N // (patch_size ** 2)
            inds_x = torch.randint(0, H - patch_size, size=[num_patch], device=device)
            inds_y = torch.randint(0, W - patch_size, size=[num_patch], device=device)
            inds = torch.stack([inds_x, inds_y], dim=-1)
            pi, pj = custom_meshgrid(torch.arange(patch_size, device=device), torch.arange(patch_size, device=device))
            offsets = torch.stack([pi.reshape(-1), pj.reshape(-1)], dim=-1)
            inds = (inds.unsqueeze(1) + offsets.unsqueeze(0)).view(-1, 2)
            inds = (inds[:, 0] * W + inds[:, 1]).expand([B, N])
        elif rect is not None:
            mask = torch.zeros(H, W, dtype=torch.bool, device=device)
            mask[xmin:xmax, ymin:ymax] = 1
            inds = torch.where(mask.view(-1))[0].unsqueeze(0)
        else:
            inds = torch.randint(0, H * W, size=[N], device=device).expand([B, N])
        i = torch.gather(i, -1, inds)
        j = torch.gather(j, -1, inds)
    else:
        inds = torch.arange(H * W, device=device).expand([B, H * W])
    results['i'], results['j'], results['inds'] = i, j, inds
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ poses[:, :3, :3].transpose(-1, -2)
    rays_o = poses[..., :3, 3][..., None, :].expand_as(rays_d)
    results['rays_o'], results['rays_d'] = rays_o, rays_d
    return results


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
