"""numpy restatement of the reference's RAD-NeRF field, host render loop and torso
(TEST INFRASTRUCTURE ONLY -- see oracle/gf_oracle.c header).

All parameters come in as a dict of numpy arrays keyed by the reference's
state_dict names (SURVEY.md section 8a "Parameter/buffer inventory").

Reference files restated here (paths under /root/reference):
  modules/radnerfs/encoders/gridencoder/grid.py:96-161   GridEncoder geometry + forward
  modules/radnerfs/cond_encoder.py:7-111                 AudioNet / AudioAttNet / MLP
  modules/radnerfs/radnerf.py:61-105                     cal_cond_feat / forward
  modules/radnerfs/renderer.py:263-367                   NeRFRenderer.render (eval branch)
  modules/radnerfs/radnerf_torso.py:51-84,155-196        forward_torso / torso mix
  modules/radnerfs/utils.py:263-363                      convert_poses / get_bg_coords / get_rays
Dense layers use float64 accumulation and round to float32 at every layer
boundary (the reference keeps fp32 tensors between cuBLAS fp32 GEMMs).
"""
import math

import numpy as np

from . import cpu_ops as ops


# ------------------------------------------------------------------ grid geometry
def grid_offsets(input_dim, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16,
                 desired_resolution=2048, align_corners=False):
    """grid.py:100-128 -> (offsets int32 [L+1], per_level_scale float64)."""
    per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    max_params = 2 ** log2_hashmap_size
    offsets, offset = [], 0
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale


def grid_encode(x, bound, embeddings, offsets, per_level_scale, base_resolution=16, gridtype=1, interp=0):
    """GridEncoder.forward (grid.py:145-161): x in [-bound,bound] -> [B, L*C]."""
    x = np.asarray(x, np.float32)
    inputs = ((x + np.float32(bound)) / np.float32(2 * bound)).astype(np.float32)
    S = np.float32(np.log2(per_level_scale))
    out, _ = ops.grid_encode_forward(inputs, embeddings, offsets, S, base_resolution, False, gridtype, False, interp)
    L, B, C = out.shape
    return np.ascontiguousarray(out.transpose(1, 0, 2).reshape(B, L * C))


# ------------------------------------------------------------------ dense helpers
# float64 accumulation for parity checks; bench.py's CPU-baseline leg switches to float32 (what the reference's
# CPU tensors would use) so that the timing is not penalised by the checker's extra precision.
MATMUL_DTYPE = np.float64


def linear(x, W, b=None):
    y = x.astype(MATMUL_DTYPE) @ W.astype(MATMUL_DTYPE).T
    if b is not None:
        y = y + b.astype(MATMUL_DTYPE)
    return y.astype(np.float32)


def mlp(x, weights):
    """cond_encoder.MLP (bias-free, ReLU between layers)."""
    for l, W in enumerate(weights):
        x = linear(x, W)
        if l != len(weights) - 1:
            x = np.maximum(x, 0)
    return x


def leaky(x, s=0.02):
    return np.where(x >= 0, x, x * np.float32(s)).astype(np.float32)


def conv1d(x, W, b, stride=1, padding=1):
    """x [B,Cin,T], W [Cout,Cin,k] -> [B,Cout,T'] (float64 accumulate)."""
    B, Cin, T = x.shape
    Cout, _, k = W.shape
    xp = np.zeros((B, Cin, T + 2 * padding), np.float64)
    xp[:, :, padding:padding + T] = x
    To = (T + 2 * padding - k) // stride + 1
    out = np.zeros((B, Cout, To), np.float64)
    for t in range(To):
        seg = xp[:, :, t * stride:t * stride + k]            # [B,Cin,k]
        out[:, :, t] = np.einsum('bck,ock->bo', seg, W.astype(np.float64))
    out += b.astype(np.float64)[None, :, None]
    return out.astype(np.float32)


def _sd(sd, key):
    return np.asarray(sd[key], np.float32)


def cal_cond_feat(sd, cond, win_size=1, with_att=True):
    """radnerf.py:61-71 with AudioNet (cond_encoder.py:7-52) and AudioAttNet (:55-89)."""
    strides = {1: [1, 1, 1, 1], 2: [2, 1, 1, 1], 3: [2, 2, 1, 1], 4: [2, 2, 1, 1], 16: [2, 2, 2, 2]}[win_size]
    x = np.asarray(cond, np.float32).transpose(0, 2, 1)        # [b,c,t]
    for i, li in enumerate([0, 2, 4, 6]):
        x = leaky(conv1d(x, _sd(sd, f'cond_prenet.encoder_conv.{li}.weight'), _sd(sd, f'cond_prenet.encoder_conv.{li}.bias'),
                         stride=strides[i], padding=1))
    x = x[:, :, 0]                                             # squeeze(-1)
    x = leaky(linear(x, _sd(sd, 'cond_prenet.encoder_fc1.0.weight'), _sd(sd, 'cond_prenet.encoder_fc1.0.bias')))
    x = linear(x, _sd(sd, 'cond_prenet.encoder_fc1.2.weight'), _sd(sd, 'cond_prenet.encoder_fc1.2.bias'))   # [b,64]
    if not with_att:
        return x
    seq = x.shape[0]
    y = x.T[None]                                              # [1,c,b]
    for li in [0, 2, 4, 6, 8]:
        y = leaky(conv1d(y, _sd(sd, f'cond_att_net.attentionConvNet.{li}.weight'),
                         _sd(sd, f'cond_att_net.attentionConvNet.{li}.bias'), stride=1, padding=1))
    y = linear(y.reshape(1, seq), _sd(sd, 'cond_att_net.attentionNet.0.weight'), _sd(sd, 'cond_att_net.attentionNet.0.bias'))
    y = y.astype(np.float64)
    y = np.exp(y - y.max(axis=1, keepdims=True))
    y = (y / y.sum(axis=1, keepdims=True)).reshape(seq, 1)
    return (y * x.astype(np.float64)).sum(axis=0).astype(np.float32)   # [64]


# ------------------------------------------------------------------ the head field
class FieldOracle:
    """RADNeRF.forward restated (radnerf.py:73-105)."""

    def __init__(self, sd, bound=1.0, desired_resolution=2048, log2_hashmap_size=16, gridtype=1, interp=0):
        self.sd = sd
        self.bound = float(bound)
        self.gridtype, self.interp = gridtype, interp
        self.pos_offsets, self.pos_pls = grid_offsets(3, log2_hashmap_size=log2_hashmap_size,
                                                      desired_resolution=desired_resolution * bound)
        self.amb_offsets, self.amb_pls = grid_offsets(2, log2_hashmap_size=log2_hashmap_size,
                                                      desired_resolution=desired_resolution)
        assert self.pos_offsets[-1] == sd['position_embedder.embeddings'].shape[0]
        assert self.amb_offsets[-1] == sd['ambient_embedder.embeddings'].shape[0]
        self.amb_w = [_sd(sd, f'ambient_net.net.{i}.weight') for i in range(3)]
        self.sig_w = [_sd(sd, f'sigma_net.net.{i}.weight') for i in range(3)]
        self.col_w = [_sd(sd, f'color_net.net.{i}.weight') for i in range(2)]

    def forward(self, position, direction, cond_feat, ind_code):
        position = np.asarray(position, np.float32)
        M = position.shape[0]
        cond = np.broadcast_to(np.asarray(cond_feat, np.float32).reshape(1, -1), (M, cond_feat.size))
        pos_feat = grid_encode(position, self.bound, _sd(self.sd, 'position_embedder.embeddings'), self.pos_offsets,
                               self.pos_pls, gridtype=self.gridtype, interp=self.interp)
        amb_logit = mlp(np.concatenate([pos_feat, cond], 1), self.amb_w)
        ambient_pos = np.tanh(amb_logit.astype(np.float64)).astype(np.float32)
        amb_feat = grid_encode(ambient_pos, 1, _sd(self.sd, 'ambient_embedder.embeddings'), self.amb_offsets,
                               self.amb_pls, gridtype=self.gridtype, interp=self.interp)
        h = mlp(np.concatenate([pos_feat, amb_feat], 1), self.sig_w)
        sigma = np.exp(h[:, 0].astype(np.float64)).astype(np.float32)     # trunc_exp fwd (utils.py:36-42)
        geo = h[:, 1:]
        sh, _ = ops.sh_encode_forward(direction, 4)
        parts = [sh, geo]
        if ind_code is not None:
            parts.append(np.broadcast_to(np.asarray(ind_code, np.float32).reshape(1, -1), (M, ind_code.size)))
        logit = mlp(np.concatenate(parts, 1), self.col_w)
        color = (1.0 / (1.0 + np.exp(-logit.astype(np.float64)))).astype(np.float32)
        return sigma, color, ambient_pos


# ------------------------------------------------------------------ rays / poses
def get_rays(pose, intrinsics, H, W):
    """utils.py:282-363 with N=-1 (all pixels, row-major over (y,x), pixel centres +0.5)."""
    fx, fy, cx, cy = [np.float32(v) for v in intrinsics]
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    i = i.reshape(-1) + np.float32(0.5)
    j = j.reshape(-1) + np.float32(0.5)
    xs = ((i - cx) / fx).astype(np.float32)
    ys = ((j - cy) / fy).astype(np.float32)
    zs = np.ones_like(xs)
    d = np.stack([xs, ys, zs], -1).astype(np.float64)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    R = np.asarray(pose, np.float64)[:3, :3]
    rays_d = (d @ R.T).astype(np.float32)
    rays_o = np.broadcast_to(np.asarray(pose, np.float32)[:3, 3], rays_d.shape).copy()
    return rays_o, rays_d


def get_bg_coords(H, W):
    """utils.py:273-278 (note: first coordinate runs over H, 'ij' meshgrid)."""
    X = (np.arange(H, dtype=np.float32) / np.float32(H - 1) * 2 - 1).astype(np.float32)
    Y = (np.arange(W, dtype=np.float32) / np.float32(W - 1) * 2 - 1).astype(np.float32)
    xs, ys = np.meshgrid(X, Y, indexing='ij')
    return np.stack([xs.reshape(-1), ys.reshape(-1)], -1).astype(np.float32)


def convert_poses(pose):
    """utils.py:263-269 + matrix_to_euler_angles('XYZ') (:160-199): [euler(3), trans(3)]."""
    M = np.asarray(pose, np.float64)
    R = M[:3, :3]
    central = math.asin(R[0, 2])
    ax = math.atan2(-R[1, 2], R[2, 2])
    az = math.atan2(-R[0, 1], R[0, 0])
    return np.array([ax, central, az, M[0, 3], M[1, 3], M[2, 3]], np.float32)


# ------------------------------------------------------------------ head render loop
def render_head(field, sd, rays_o, rays_d, cond_feat, bitfield, cascade, grid_size, aabb, min_near,
                dt_gamma, max_steps, T_thresh=1e-4, trace=None, term_iter=None):
    """renderer.py:314-351 (eval branch).  Returns weights_sum, depth, image, nears, fars, n_samples[N].
    term_iter: optional int32[N] (pre-filled with -1) receiving the loop iteration in which composite_rays marked the ray dead
    (rays_alive[n] = -1, raymarching.cu:1017)."""
    N = rays_o.shape[0]
    nears, fars = ops.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    ind_code = _sd(sd, 'individual_embeddings')[0] if 'individual_embeddings' in sd else None
    weights_sum = np.zeros(N, np.float32)
    depth = np.zeros(N, np.float32)
    image = np.zeros((N, 3), np.float32)
    rays_alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()
    n_samples = np.zeros(N, np.int32)
    step = 0
    while step < max_steps:
        n_alive = rays_alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = ops.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, field.bound, bitfield,
                                            cascade, grid_size, nears, fars, 128, None, dt_gamma, max_steps)
        sigmas, rgbs, _ = field.forward(xyzs, dirs, cond_feat, ind_code)
        if trace is not None:
            trace.append((n_alive, n_step))
        ids = rays_alive.copy()
        composited = ops.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
        n_samples[ids] += composited
        if term_iter is not None:
            term_iter[ids[rays_alive < 0]] = len(trace) - 1 if trace is not None else -2
        rays_alive = np.ascontiguousarray(rays_alive[rays_alive >= 0])
        step += n_step
    return weights_sum, depth, image, nears, fars, n_samples


def finish(image, weights_sum, depth, nears, fars, bg_color):
    """renderer.py:354-362."""
    image = image + (1 - weights_sum)[:, None] * bg_color
    image = np.clip(image, 0, 1).astype(np.float32)
    with np.errstate(invalid='ignore', divide='ignore'):
        depth = (np.maximum(depth - nears, 0) / (fars - nears)).astype(np.float32)
    return image, depth


# ------------------------------------------------------------------ torso
def grid_sample_2d(grid2d, coords):
    """F.grid_sample(grid.view(1,1,H,W), coords.view(1,-1,1,2), align_corners=True), bilinear, zero padding."""
    Hh, Ww = grid2d.shape
    x = (coords[:, 0].astype(np.float32) + 1) / 2 * np.float32(Ww - 1)
    y = (coords[:, 1].astype(np.float32) + 1) / 2 * np.float32(Hh - 1)
    x0 = np.floor(x).astype(np.int64); y0 = np.floor(y).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    wx1 = x - x0; wx0 = 1 - wx1
    wy1 = y - y0; wy0 = 1 - wy1

    def at(yy, xx):
        ok = (xx >= 0) & (xx < Ww) & (yy >= 0) & (yy < Hh)
        v = grid2d[np.clip(yy, 0, Hh - 1), np.clip(xx, 0, Ww - 1)]
        return np.where(ok, v, 0).astype(np.float32)

    return (at(y0, x0) * wx0 * wy0 + at(y0, x1) * wx1 * wy0 + at(y1, x0) * wx0 * wy1 + at(y1, x1) * wx1 * wy1).astype(np.float32)


class TorsoOracle:
    """RADNeRFTorso.forward_torso (radnerf_torso.py:51-84), torso_head_aware=False."""

    def __init__(self, sd, torso_shrink=0.8):
        self.sd = sd
        self.shrink = np.float32(torso_shrink)
        self.offsets, self.pls = grid_offsets(2, log2_hashmap_size=16, desired_resolution=2048)
        self.deform_w = [_sd(sd, f'torso_deform_net.net.{i}.weight') for i in range(3)]
        self.canon_w = [_sd(sd, f'torso_canonicial_net.net.{i}.weight') for i in range(3)]

    def forward(self, x, poses, c):
        x = (np.asarray(x, np.float32) * self.shrink).astype(np.float32)
        enc_pose = ops.freq_encode_forward(np.asarray(poses, np.float32).reshape(1, 6), 4, 54)
        enc_x = ops.freq_encode_forward(x, 10, 42)
        n = x.shape[0]
        parts = [enc_x, np.broadcast_to(enc_pose, (n, 54))]
        if c is not None:
            parts.append(np.broadcast_to(np.asarray(c, np.float32).reshape(1, -1), (n, c.size)))
        h = np.concatenate(parts, 1).astype(np.float32)
        dx = mlp(h, self.deform_w)
        xd = np.clip(x + dx, -1, 1).astype(np.float32)
        feat = grid_encode(xd, 1, _sd(self.sd, 'torso_embedder.embeddings'), self.offsets, self.pls, gridtype=1, interp=0)
        h2 = mlp(np.concatenate([feat, h], 1), self.canon_w)
        sig = (1.0 / (1.0 + np.exp(-h2.astype(np.float64)))).astype(np.float32)
        return sig[:, :1], sig[:, 1:], dx


def render_torso_mix(torso, sd, bg_coords, poses, bg_color, image, weights_sum, grid_size=128,
                     density_thresh_torso=0.01, mean_density_torso=0.0):
    """radnerf_torso.py:155-196.  Returns (rgb_map_unclamped_prefinish_bg, torso_alpha, torso_rgb_map, deform, mask)."""
    N = bg_coords.shape[0]
    thresh = min(density_thresh_torso, mean_density_torso)
    occ = grid_sample_2d(_sd(sd, 'density_grid_torso').reshape(grid_size, grid_size), bg_coords)
    mask = occ > thresh
    torso_alpha = np.zeros((N, 1), np.float32)
    torso_color = np.zeros((N, 3), np.float32)
    deform = None
    if mask.any():
        code = _sd(sd, 'torso_individual_codes')[0] if 'torso_individual_codes' in sd else None
        a, c, deform = torso.forward(bg_coords[mask], poses, code)
        torso_alpha[mask] = a
        torso_color[mask] = c
    bg = (torso_color * torso_alpha + bg_color * (1 - torso_alpha)).astype(np.float32)
    return bg, torso_alpha, deform, mask
