"""Deterministic synthetic models and scenes: there are no checkpoints or datasets in
this environment, so parity tests and bench.py render seeded random-weight models of the exact
architecture of the May configs (egs/datasets/videos/May/lm3d_radnerf{,_torso}.yaml resolved through
egs/egs_bases/radnerf/{lm3d_radnerf,base}.yaml).
"""
import math

import numpy as np
import torch

from .utils import convert_poses, intrinsics_from_fovy, orbit_pose


def may_hparams(**over):
    """Resolved hparams of the May lm3d_radnerf(_torso) configs (values: SURVEY.md section 8 header)."""
    hp = dict(
        bound=1, grid_size=128, min_near=0.05, density_thresh=10, density_thresh_torso=0.01, cuda_ray=True,
        individual_embedding_num=13000, individual_embedding_dim=4, torso_individual_embedding_dim=8,
        cond_type='idexp_lm3d_normalized', cond_out_dim=64, cond_win_size=1, smo_win_size=5, with_att=True,
        grid_type='tiledgrid', grid_interpolation_type='linear', log2_hashmap_size=16, desired_resolution=2048,
        num_layers_ambient=3, hidden_dim_ambient=128, ambient_out_dim=2,
        num_layers_sigma=3, hidden_dim_sigma=128, geo_feat_dim=128,
        num_layers_color=2, hidden_dim_color=128,
        torso_shrink=0.8, torso_head_aware=False,
        max_steps=16, dt_gamma=1 / 256, camera_scale=4,
    )
    hp.update(over)
    return hp


def sphere_bitfield(cascade, grid_size, radius=0.45):
    """Bitfield 'S': voxels (all cascades, morton order) whose centre lies inside ||x|| < radius."""
    H = grid_size
    idx = np.arange(H ** 3, dtype=np.uint32)

    def compact(v):
        v = v & 0x49249249
        v = (v | (v >> 2)) & 0xc30c30c3
        v = (v | (v >> 4)) & 0x0f00f00f
        v = (v | (v >> 8)) & 0xff0000ff
        v = (v | (v >> 16)) & 0x0000ffff
        return v

    x, y, z = compact(idx), compact(idx >> 1), compact(idx >> 2)
    bits = []
    for cas in range(cascade):
        b = min(2 ** cas, 2 ** (cascade - 1))
        c = [((v.astype(np.float64) + 0.5) / H * 2 - 1) * b for v in (x, y, z)]
        bits.append((c[0] ** 2 + c[1] ** 2 + c[2] ** 2) < radius ** 2)
    occ = np.concatenate(bits)
    return torch.from_numpy(np.packbits(occ, bitorder='little'))


def make_bitfield(kind, cascade, grid_size, seed=1):
    n = cascade * grid_size ** 3 // 8
    if kind == 'S':
        return sphere_bitfield(cascade, grid_size)
    if kind == 'F':
        return torch.full((n,), 255, dtype=torch.uint8)
    if kind == 'R':
        g = torch.Generator().manual_seed(seed)
        occ = (torch.rand(cascade * grid_size ** 3, generator=g) < 0.3).numpy()
        return torch.from_numpy(np.packbits(occ, bitorder='little'))
    raise ValueError(kind)


def randomize_(model, seed=0, emb_scale=0.5, sigma_scale=4.0):
    """Config-2 initialisation: embeddings ~ U(-1,1)*0.5 (the default U(+-1e-4) gives a featureless field),
    nn.Linear-style U(+-1/sqrt(fan_in)) weights, sigma logit row scaled by `sigma_scale` (4 => sigma spans
    ~[e^-4, e^4]; 0.25 => sigma stays near 1 so that no ray terminates early in the 128-sample workload)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('embeddings') and p.dim() == 2 and p.shape[1] == 2:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * emb_scale)
            elif 'individual' in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif p.dim() >= 2:
                bound = 1 / math.sqrt(p[0].numel())
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.1)
        model.sigma_net.net[-1].weight[0].mul_(sigma_scale)
    return model


def build_model(torso=False, bitfield='S', seed=0, device='cuda', sigma_scale=4.0, **hp_over):
    from .renderer import RADNeRF, RADNeRFTorso
    hp = may_hparams(**hp_over)
    torch.manual_seed(seed)
    model = (RADNeRFTorso if torso else RADNeRF)(hp)
    randomize_(model, seed=seed, sigma_scale=sigma_scale)
    model.density_bitfield.copy_(make_bitfield(bitfield, model.cascade, model.grid_size))
    if torso:
        # occupancy 1 on the lower image half (y_pix >= H/2).  bg_coords[n] = (X[y_pix], Y[x_pix]) (utils.py:273-278)
        # and grid_sample reads grid[row <- coord 1, col <- coord 0], so the lower half is columns >= H/2.
        g = torch.zeros(model.grid_size, model.grid_size)
        g[:, model.grid_size // 2:] = 1.0
        model.density_grid_torso.copy_(g.reshape(-1))
        model.mean_density_torso = 0             # as after a checkpoint load
    model.eval()
    return model.to(device), hp


def frame_inputs(H=512, W=512, yaw_deg=0.0, cond_seed=0, bg_seed=2, device='cuda', smo_win=5, cond_win=1, cond_dim=204):
    """Config-2 camera / cond / background for one frame."""
    pose = torch.from_numpy(orbit_pose(3.35, yaw_deg)).unsqueeze(0)
    intr = intrinsics_from_fovy(H, W, 21.24)
    g = torch.Generator().manual_seed(cond_seed)
    cond = torch.randn(smo_win, cond_win, cond_dim, generator=g)
    gb = torch.Generator().manual_seed(bg_seed)
    bg = torch.rand(1, H * W, 3, generator=gb)
    return dict(pose=pose.to(device), intrinsics=intr, cond=cond.to(device), bg_color=bg.to(device),
                poses6=convert_poses(pose).to(device), H=H, W=W)


def state_to_numpy(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
