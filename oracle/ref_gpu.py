"""The reference renderer on a GPU, assembled from the UNMODIFIED compiled reference kernels (oracle/_ref/*.so)
plus plain torch fp32 for the dense layers -- i.e. what modules/radnerfs/{renderer,radnerf,radnerf_torso}.py do,
restated minimally so it can run where /root/reference does not exist (the GPU box).

TEST INFRASTRUCTURE ONLY (parity oracle + "kernel to beat" timing).  Restated call sites:
  raymarching.py:18-46 (near_far), :347-396 (march_rays), :401-420 (composite_rays)
  grid.py:24-63,145-161 (grid_encode + bound mapping), sphere_harmonics.py:14-37, freq.py:15-35
  radnerf.py:73-105 (field), renderer.py:314-362 (eval loop + finish), radnerf_torso.py:51-84,155-196 (torso)
Weights come from a state_dict with the reference's key names.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return all(os.path.exists(os.path.join(_REF, n + ".so")) for n in ("_raymarching_face", "_gridencoder", "_shencoder", "_freqencoder"))


def _mods():
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    import _raymarching_face as RM
    import _gridencoder as GE
    import _shencoder as SH
    import _freqencoder as FQ
    return RM, GE, SH, FQ


class RefRenderer:
    def __init__(self, sd, hp, torso=False):
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        self.RM, self.GE, self.SH, self.FQ = _mods()
        self.sd = {k: v.detach().cuda().contiguous() for k, v in sd.items()}
        self.hp = hp
        self.bound = float(hp['bound'])
        self.cascade = 1 + int(np.ceil(np.log2(hp['bound'])))
        self.H = hp['grid_size']
        self.torso = torso
        self.pls_pos = np.exp2(np.log2(hp['desired_resolution'] * hp['bound'] / 16) / 15)
        self.pls_2d = np.exp2(np.log2(hp['desired_resolution'] / 16) / 15)
        self.pls_torso = np.exp2(np.log2(2048 / 16) / 15)
        self.gridtype = {'hashgrid': 0, 'tiledgrid': 1}[hp['grid_type']]
        self.interp = {'linear': 0, 'smoothstep': 1}[hp['grid_interpolation_type']]

    # -- encoders ---------------------------------------------------------------------------------
    def grid(self, x, bound, prefix, pls, gridtype, interp):
        emb, offsets = self.sd[prefix + '.embeddings'], self.sd[prefix + '.offsets']
        inputs = ((x + bound) / (2 * bound)).contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, emb.shape[1]
        out = torch.empty(L, B, C, device=x.device)
        self.GE.grid_encode_forward(inputs, emb, offsets, out, B, D, C, L, float(np.log2(pls)), 16, None, gridtype, False, interp)
        return out.permute(1, 0, 2).reshape(B, L * C)

    def sh(self, d):
        d = d.contiguous()
        out = torch.empty(d.shape[0], 16, device=d.device)
        self.SH.sh_encode_forward(d, out, d.shape[0], 3, 4, None)
        return out

    def freq(self, x, deg):
        x = x.contiguous()
        B, D = x.shape
        C = D + D * 2 * deg
        out = torch.empty(B, C, device=x.device)
        self.FQ.freq_encode_forward(x, B, D, deg, C, out)
        return out

    def mlp(self, x, prefix, n):
        for l in range(n):
            x = F.linear(x, self.sd[f'{prefix}.net.{l}.weight'])
            if l != n - 1:
                x = F.relu(x)
        return x

    # -- radnerf.py:73-105 --------------------------------------------------------------------------
    def field(self, position, direction, cond_feat, ind_code):
        M = position.shape[0]
        cond = cond_feat.view(1, -1).repeat(M, 1)
        pos_feat = self.grid(position, self.bound, 'position_embedder', self.pls_pos, self.gridtype, self.interp)
        ambient_pos = torch.tanh(self.mlp(torch.cat([pos_feat, cond], 1), 'ambient_net', 3).float())
        amb_feat = self.grid(ambient_pos, 1, 'ambient_embedder', self.pls_2d, self.gridtype, self.interp)
        h = self.mlp(torch.cat([pos_feat, amb_feat], -1), 'sigma_net', 3)
        sigma = torch.exp(h[..., 0])
        parts = [self.sh(direction), h[..., 1:]]
        if ind_code is not None:
            parts.append(ind_code.view(1, -1).repeat(M, 1))
        color = torch.sigmoid(self.mlp(torch.cat(parts, -1), 'color_net', 2))
        return sigma, color, ambient_pos

    # -- renderer.py:263-367 (eval).  Also returns the (n_alive, n_step) trace of the host loop and the per-ray
    #    count of MARCHED (non-padding) samples ---------------------------------------------------------
    def render_head(self, rays_o, rays_d, cond_feat, dt_gamma, max_steps, T_thresh=1e-4, trace=None):
        RM = self.RM
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        dev = rays_o.device
        nears = torch.empty(N, device=dev); fars = torch.empty(N, device=dev)
        RM.near_far_from_aabb(rays_o, rays_d, self.sd['aabb_infer'], N, float(self.hp['min_near']), nears, fars)
        ind_code = self.sd['individual_embeddings'][0] if 'individual_embeddings' in self.sd else None
        weights_sum = torch.zeros(N, device=dev); depth = torch.zeros(N, device=dev); image = torch.zeros(N, 3, device=dev)
        rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        n_marched = torch.zeros(N, dtype=torch.int32, device=dev)
        bitfield = self.sd['density_bitfield']
        step = 0
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            M = n_alive * n_step
            M += 128 - (M % 128)
            xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev); deltas = torch.zeros(M, 2, device=dev)
            noises = torch.zeros(n_alive, device=dev)
            RM.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, float(dt_gamma), int(max_steps), self.cascade, self.H,
                          bitfield, nears, fars, xyzs, dirs, deltas, noises)
            sigmas, rgbs, _ = self.field(xyzs, dirs, cond_feat, ind_code)
            ids = rays_alive.long().clone()
            n_marched[ids] += (deltas[:n_alive * n_step, 0].view(n_alive, n_step) != 0).sum(1).int()
            RM.composite_rays(n_alive, n_step, float(T_thresh), rays_alive, rays_t, sigmas.contiguous(), rgbs.contiguous(), deltas, weights_sum, depth, image)
            if trace is not None:
                trace.append((n_alive, n_step))
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
        return weights_sum, depth, image, nears, fars, n_marched



    def finish(self, image, weights_sum, depth, nears, fars, bg_color):
        image = (image + (1 - weights_sum).unsqueeze(-1) * bg_color).clamp(0, 1)
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        return image, depth

    # -- radnerf_torso.py:51-84 / 155-188 ---------------------------------------------------------------
    def forward_torso(self, x, poses, c):
        x = x * float(self.hp.get('torso_shrink', 0.8))
        enc_pose = self.freq(poses.view(1, 6), 4)
        enc_x = self.freq(x, 10)
        parts = [enc_x, enc_pose.repeat(x.shape[0], 1)]
        if c is not None:
            parts.append(c.view(1, -1).repeat(x.shape[0], 1))
        h = torch.cat(parts, -1)
        dx = self.mlp(h, 'torso_deform_net', 3)
        xd = (x + dx).clamp(-1, 1).float()
        feat = self.grid(xd, 1, 'torso_embedder', self.pls_torso, 1, 0)
        h2 = self.mlp(torch.cat([feat, h], -1), 'torso_canonicial_net', 3)
        return torch.sigmoid(h2[..., :1]), torch.sigmoid(h2[..., 1:]), dx

    def torso_bg(self, bg_coords, poses, bg_color, mean_density_torso=0.0):
        N = bg_coords.shape[0]
        dev = bg_coords.device
        thresh = min(float(self.hp['density_thresh_torso']), mean_density_torso)
        occ = F.grid_sample(self.sd['density_grid_torso'].view(1, 1, self.H, self.H), bg_coords.view(1, -1, 1, 2), align_corners=True).view(-1)
        mask = occ > thresh
        alpha = torch.zeros(N, 1, device=dev); color = torch.zeros(N, 3, device=dev)
        if mask.any():
            code = self.sd['torso_individual_codes'][0] if 'torso_individual_codes' in self.sd else None
            a, c, _ = self.forward_torso(bg_coords[mask], poses, code)
            alpha[mask] = a; color[mask] = c
        return color * alpha + bg_color * (1 - alpha), alpha, mask
