"""Vanilla AD-NeRF path on the GPU (drop-in for modules/nerfs: SURVEY.md section 8 row a19).

Mirrors, with the reference's names, arguments and state_dict keys:
  modules/nerfs/commons/ray_samplers.py:11-44        get_rays
  modules/nerfs/commons/embedders.py:5-45            FreqEmbedder
  modules/nerfs/adnerf/backbone.py:6-135             AudioNet, AudioAttNet, NeRFBackbone
  modules/nerfs/adnerf/adnerf.py:9-44                ADNeRF
  modules/nerfs/commons/volume_rendering.py:9-282    raw2outputs, sample_pdf, render_rays, batchify_render_rays, render_dynamic_face

Every operator of the path is a libgfrender kernel reached through the C ABI: rays, frequency embedding, alpha compositing with the
background-colour last sample, inverse-CDF importance sampling + merge (csrc/adnerf_ops.cu), and the 8 x hid / 3 x hid/2 backbone
itself on tcgen05 tensor cores (csrc/adnerf_mlp_tc.cu, `gf_adnerf_mlp_forward`: fp16 operands, fp32 accumulation).  When the network is
this module's ADNeRF, render_rays evaluates the backbone in a FOLDED form that is algebraically identical to backbone.py:107-135 but never
materialises the per-sample copies the reference concatenates: the per-frame audio feature becomes a bias of layers 0 and 5, the view
embedding an extra K-chunk of the first colour layer, and the position embedding is produced straight from (rays, z) in the tensor-core
operand layout.  `NeRFBackbone.forward` / `forward_folded` (torch, fp32) remain as the reference-form definitions used by the CPU tests
and for inputs outside the fused envelope (per-sample conditions).

Inference only (the C ABI operators have no backward): calling these functions with gradients enabled on inputs that require
grad raises.  CUDA tensors only -- there is no CPU fallback.
"""
import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr, stream_ptr


def require_cuda(t):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError("geneface_b200.adnerf operators need CUDA tensors (sm_100a); there is no CPU fallback")


def _f32c(t):
    return t.detach().float().contiguous()


def _no_grad_inputs(*ts):
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in ts):
        raise NotImplementedError("geneface_b200.adnerf is inference-only: wrap the call in torch.no_grad()")


# ------------------------------------------------------------------------------------------------------ rays / embedding
def get_rays(H, W, focal, c2w, cx=None, cy=None):
    """ray_samplers.py:11-44: OpenGL-convention rays of a full image -> rays_o, rays_d [H, W, 3] (un-normalised directions)."""
    require_cuda(c2w)
    cx = W * 0.5 if cx is None else cx
    cy = H * 0.5 if cy is None else cy
    m = _f32c(c2w[:3, :4])
    rays_o = torch.empty(H, W, 3, device=c2w.device)
    rays_d = torch.empty(H, W, 3, device=c2w.device)
    check(_lib.lib().gf_adnerf_get_rays(H, W, float(focal), float(cx), float(cy), ptr(m), ptr(rays_o), ptr(rays_d), None, stream_ptr()))
    return rays_o, rays_d


class FreqEmbedder(nn.Module):
    """embedders.py:5-45 (log bands, include_input): [x, sin(2^k x), cos(2^k x)] for k < multi_res."""

    def __init__(self, in_dim=3, multi_res=10, use_log_bands=True, include_input=True):
        super().__init__()
        if not (use_log_bands and include_input):
            raise NotImplementedError("only the configuration the reference instantiates (log bands, include_input) is implemented")
        self.in_dim, self.num_freqs = in_dim, multi_res
        self.out_dim = in_dim * (1 + 2 * multi_res)

    def forward(self, x):
        require_cuda(x)
        _no_grad_inputs(x)
        xc = _f32c(x).view(-1, self.in_dim)
        out = torch.empty(xc.shape[0], self.out_dim, device=x.device)
        check(_lib.lib().gf_adnerf_embed(ptr(xc), xc.shape[0], self.in_dim, self.num_freqs, ptr(out), self.out_dim, stream_ptr()))
        return out.view(*x.shape[:-1], self.out_dim)


# ------------------------------------------------------------------------------------------------------ networks
class GfAdnerfDesc(ctypes.Structure):
    """include/gfrender.h: GfAdnerfDesc"""
    _fields_ = [("hid", ctypes.c_uint32), ("cond_dim", ctypes.c_uint32), ("pos_multires", ctypes.c_uint32), ("view_multires", ctypes.c_uint32),
                ("dens_w", ctypes.c_void_p * 8), ("dens_b", ctypes.c_void_p * 8), ("dens_out_w", ctypes.c_void_p), ("dens_out_b", ctypes.c_void_p),
                ("col_w", ctypes.c_void_p * 3), ("col_b", ctypes.c_void_p * 3), ("col_out_w", ctypes.c_void_p), ("col_out_b", ctypes.c_void_p)]


class AudioNet(nn.Module):
    """backbone.py:6-42: deepspeech window [B, 16, 29] -> conv1d x4 (stride 2) -> fc -> [B, out_dim]."""

    def __init__(self, in_dim=29, out_dim=64, win_size=16):
        super().__init__()
        self.win_size, self.out_dim = win_size, out_dim
        self.encoder_conv = nn.Sequential(
            nn.Conv1d(in_dim, 32, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True),
            nn.Conv1d(32, 32, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True),
            nn.Conv1d(32, 64, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True),
            nn.Conv1d(64, 64, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True))
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, out_dim))

    def forward(self, x):
        half = self.win_size // 2
        x = x[:, 8 - half:8 + half, :].permute(0, 2, 1)
        x = self.encoder_conv(x).squeeze(-1)
        return self.encoder_fc1(x).squeeze()


class AudioAttNet(nn.Module):
    """backbone.py:45-79: attention over the smoothing window -> one feature vector."""

    def __init__(self, in_out_dim=64, seq_len=8):
        super().__init__()
        self.seq_len, self.in_out_dim = seq_len, in_out_dim
        chans = (in_out_dim, 16, 8, 4, 2, 1)
        layers = []
        for i in range(5):
            layers += [nn.Conv1d(chans[i], chans[i + 1], kernel_size=3, stride=1, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.attentionConvNet = nn.Sequential(*layers)
        self.attentionNet = nn.Sequential(nn.Linear(seq_len, seq_len, bias=True), nn.Softmax(dim=1))

    def forward(self, x):
        y = x[..., :self.in_out_dim].permute(1, 0).unsqueeze(0)
        y = self.attentionConvNet(y)
        y = self.attentionNet(y.view(1, self.seq_len)).view(self.seq_len, 1)
        return torch.sum(y * x, dim=0)


class NeRFBackbone(nn.Module):
    """backbone.py:82-135: density trunk (8 x hid, the input re-injected after layer 4) + colour head (3 x hid/2)."""

    def __init__(self, pos_dim=3, cond_dim=64, view_dim=3, hid_dim=128, num_density_linears=8, num_color_linears=3, skip_layer_indices=(4,)):
        super().__init__()
        self.pos_dim, self.cond_dim, self.view_dim, self.hid_dim = pos_dim, cond_dim, view_dim, hid_dim
        self.skip_layer_indices = list(skip_layer_indices)
        din = pos_dim + cond_dim
        dens = [nn.Linear(din, hid_dim)]
        for i in range(num_density_linears - 1):
            dens.append(nn.Linear(hid_dim + din if i in self.skip_layer_indices else hid_dim, hid_dim))
        self.density_linears = nn.ModuleList(dens)
        self.density_out_linear = nn.Linear(hid_dim, 1)
        cols = [nn.Linear(view_dim + hid_dim, hid_dim // 2)] + [nn.Linear(hid_dim // 2, hid_dim // 2) for _ in range(num_color_linears - 1)]
        self.color_linears = nn.ModuleList(cols)
        self.color_out_linear = nn.Linear(hid_dim // 2, 3)

    def forward(self, pos, cond, view):
        """Reference form: pos [B, N, pos_dim] embedded, cond [cond_dim] or [B, cond_dim], view [B, view_dim] -> [B, N, 4] (rgb, sigma)."""
        bs, n = pos.shape[0], pos.shape[1]
        cond = cond.view(1, 1, -1).expand(bs, n, -1) if cond.dim() == 1 else cond[:, None, :].expand(bs, n, -1)
        view = view[:, None, :].expand(bs, n, -1)
        inp = torch.cat([pos, cond], dim=-1)
        h = inp
        for i, lin in enumerate(self.density_linears):
            h = F.relu(lin(h))
            if i in self.skip_layer_indices:
                h = torch.cat([inp, h], dim=-1)
        sigma = self.density_out_linear(h)
        h = torch.cat([h, view], dim=-1)
        for lin in self.color_linears:
            h = F.relu(lin(h))
        return torch.cat([self.color_out_linear(h), sigma], dim=-1)

    # -- tensor-core evaluation (csrc/adnerf_mlp_tc.cu) ---------------------------------------------------------------------------
    def tc_supported(self):
        return (self.hid_dim in (128, 256) and len(self.density_linears) == 8 and len(self.color_linears) == 3 and self.skip_layer_indices == [4]
                and (self.pos_dim - 3) % 6 == 0 and (self.view_dim - 3) % 6 == 0 and self.pos_dim <= 63 and self.view_dim <= 63)

    def _tc_handle(self):
        """Packed fp16 weight images of this network (rebuilt when a parameter changes: keyed on data_ptr + version)."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, '_tc', None) is not None and self._tc_key == key:
            return self._tc
        self._tc_free()
        d = GfAdnerfDesc()
        keep = []

        def dev(t):
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()
        d.hid, d.cond_dim, d.pos_multires, d.view_multires = self.hid_dim, self.cond_dim, (self.pos_dim - 3) // 6, (self.view_dim - 3) // 6
        for i, lin in enumerate(self.density_linears):
            d.dens_w[i], d.dens_b[i] = dev(lin.weight), dev(lin.bias)
        d.dens_out_w, d.dens_out_b = dev(self.density_out_linear.weight), dev(self.density_out_linear.bias)
        for i, lin in enumerate(self.color_linears):
            d.col_w[i], d.col_b[i] = dev(lin.weight), dev(lin.bias)
        d.col_out_w, d.col_out_b = dev(self.color_out_linear.weight), dev(self.color_out_linear.bias)
        h = ctypes.c_void_p()
        check(_lib.lib().gf_adnerf_mlp_create(ctypes.byref(d), ctypes.byref(h), stream_ptr()), "gf_adnerf_mlp_create")
        self._tc, self._tc_key = h, key
        return h

    def _tc_free(self):
        if getattr(self, '_tc', None) is not None:
            _lib.lib().gf_adnerf_mlp_destroy(self._tc)
            self._tc = None

    def __del__(self):
        try:
            self._tc_free()
        except Exception:  # noqa: BLE001
            pass

    def forward_tc(self, rays_o, rays_d, z_vals, viewdirs, cond):
        """raw [R, S, 4] at the points rays_o + rays_d * z_vals: embeddings + the whole backbone in libgfrender (one frame's cond [cond_dim])."""
        R, S = z_vals.shape
        h = self._tc_handle()
        dev = z_vals.device
        raw = torch.empty(R, S, 4, dtype=torch.float32, device=dev)
        need = _lib.lib().gf_adnerf_mlp_workspace_bytes(h, R * S)
        ws = getattr(self, '_tc_ws', None)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = self._tc_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        ro, rd, z, vd, c = _f32c(rays_o), _f32c(rays_d), _f32c(z_vals), _f32c(viewdirs), _f32c(cond).view(-1)
        check(_lib.lib().gf_adnerf_mlp_forward(h, ptr(ro), ptr(rd), ptr(z), ptr(vd), ptr(c), R, S, ptr(raw), ptr(ws), need, stream_ptr()),
              "gf_adnerf_mlp_forward")
        return raw

    def forward_folded(self, pos_embed, cond, view_embed, S):
        """Same function for samples of R rays x S depths: pos_embed [R*S, pos_dim], cond [cond_dim] (one frame), view_embed [R, view_dim].
        cond enters layers 0 and skip+1 as a bias, the view embedding enters the first colour layer as a per-ray bias."""
        pd, cd = self.pos_dim, self.cond_dim
        R = view_embed.shape[0]
        h = pos_embed
        n_dens = len(self.density_linears)
        for i, lin in enumerate(self.density_linears):
            W, b = lin.weight, lin.bias
            if i == 0:
                h = F.linear(pos_embed, W[:, :pd], b + W[:, pd:pd + cd] @ cond)
            elif (i - 1) in self.skip_layer_indices:
                # input was cat([pos_embed, cond, h_prev]) in the reference
                y = F.linear(h, W[:, pd + cd:], b + W[:, pd:pd + cd] @ cond)
                h = y.addmm_(pos_embed, W[:, :pd].t())
            else:
                h = F.linear(h, W, b)
            h = F.relu_(h)
        assert n_dens - 1 not in self.skip_layer_indices, "a skip after the last density layer is not supported by the folded form"
        sigma = self.density_out_linear(h)
        W0, b0 = self.color_linears[0].weight, self.color_linears[0].bias
        hd = self.hid_dim
        per_ray = F.linear(view_embed, W0[:, hd:], b0)                                 # [R, hid/2]
        c = F.linear(h, W0[:, :hd]).view(R, S, -1).add_(per_ray[:, None, :]).view(R * S, -1)
        c = F.relu_(c)
        for lin in list(self.color_linears)[1:]:
            c = F.relu_(lin(c))
        return torch.cat([self.color_out_linear(c), sigma], dim=-1)                    # [R*S, 4]


class ADNeRF(nn.Module):
    """adnerf.py:9-44."""

    def __init__(self, hparams=None):
        super().__init__()
        self.hparams = hparams
        self.pos_embedder = FreqEmbedder(in_dim=3, multi_res=10, use_log_bands=True, include_input=True)
        self.view_embedder = FreqEmbedder(in_dim=3, multi_res=4, use_log_bands=True, include_input=True)
        self.cond_dim = hparams['cond_dim']
        kw = dict(pos_dim=self.pos_embedder.out_dim, cond_dim=self.cond_dim, view_dim=self.view_embedder.out_dim, hid_dim=hparams['hidden_size'],
                  num_density_linears=8, num_color_linears=3, skip_layer_indices=[4])
        self.model_coarse = NeRFBackbone(**kw)
        self.model_fine = NeRFBackbone(**kw)
        self.deepspeech_win_size = 16
        self.smo_win_size = 8
        self.aud_net = AudioNet(in_dim=29, out_dim=self.cond_dim, win_size=self.deepspeech_win_size)
        self.audatt_net = AudioAttNet(in_out_dim=self.cond_dim, seq_len=self.smo_win_size)

    def forward(self, pos, cond_feat, view, run_model_fine=True, **kwargs):
        net = self.model_fine if run_model_fine else self.model_coarse
        return {'rgb_sigma': net(self.pos_embedder(pos), cond_feat, self.view_embedder(view))}

    def cal_cond_feat(self, cond, with_att=False):
        cond_feat = self.aud_net(cond)
        if with_att:
            cond_feat = self.audatt_net(cond_feat)
        return cond_feat


# ------------------------------------------------------------------------------------------------------ volume rendering
def raw2outputs(raw, z_vals, rays_d, bc_rgb, raw_noise_std=0, white_bkgd=False):
    """volume_rendering.py:9-59 -> rgb_map, disp_map, acc_map, weights, depth_map, rgb_map_fg."""
    require_cuda(raw)
    _no_grad_inputs(raw, z_vals)
    R, S = z_vals.shape
    rawc = _f32c(raw).view(R, S, 4)
    if raw_noise_std > 0.:
        rawc = rawc.clone()
        rawc[..., 3] += torch.randn(R, S, device=raw.device) * raw_noise_std
    zc, dc, bc = _f32c(z_vals), _f32c(rays_d).view(R, 3), _f32c(bc_rgb).view(R, 3)
    dev = raw.device
    rgb_map, rgb_fg = torch.empty(R, 3, device=dev), torch.empty(R, 3, device=dev)
    disp, acc, depth = torch.empty(R, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev)
    weights = torch.empty(R, S, device=dev)
    check(_lib.lib().gf_adnerf_raw2outputs(ptr(rawc), ptr(zc), ptr(dc), ptr(bc), R, S, int(bool(white_bkgd)), ptr(rgb_map), ptr(disp), ptr(acc),
                                           ptr(weights), ptr(depth), ptr(rgb_fg), stream_ptr()))
    return rgb_map, disp, acc, weights, depth, rgb_fg


def sample_pdf(bins, weights, N_samples, det=False):
    """volume_rendering.py:62-96: bins [R, B], weights [R, B-1] -> samples [R, N_samples]."""
    require_cuda(bins)
    _no_grad_inputs(bins, weights)
    R, B = bins.shape
    bc, wc = _f32c(bins), _f32c(weights)
    assert wc.shape == (R, B - 1), "weights must have one entry fewer than bins"
    u = None if det else torch.rand(R, N_samples, device=bins.device)
    out = torch.empty(R, N_samples, device=bins.device)
    check(_lib.lib().gf_adnerf_sample_pdf(ptr(bc), ptr(wc), ptr(u) if u is not None else None, R, B, N_samples, 0, ptr(out), None, stream_ptr()))
    return out


def _importance_depths(z_vals, weights, N_importance, det):
    """sample_pdf on (z_mid, weights[1:-1]) + concatenate + sort (volume_rendering.py:177-182) in one kernel."""
    R, S = z_vals.shape
    u = None if det else torch.rand(R, N_importance, device=z_vals.device)
    z_out = torch.empty(R, S + N_importance, device=z_vals.device)
    z_samples = torch.empty(R, N_importance, device=z_vals.device)
    check(_lib.lib().gf_adnerf_sample_pdf(ptr(z_vals), ptr(weights), ptr(u) if u is not None else None, R, S, N_importance, 1, ptr(z_out),
                                          ptr(z_samples), stream_ptr()))
    return z_out, z_samples


def _query(network_fn, rays_o, rays_d, z_vals, cond, viewdirs, fine, **kwargs):
    """raw [R, S, 4] of the coarse or fine network at the depths z_vals."""
    R, S = z_vals.shape
    if isinstance(network_fn, ADNeRF) and cond.dim() == 1 and viewdirs is not None:
        net = network_fn.model_fine if fine else network_fn.model_coarse
        if net.tc_supported():
            return net.forward_tc(rays_o, rays_d, z_vals, viewdirs, cond)
        L = network_fn.pos_embedder.num_freqs
        pe = torch.empty(R * S, network_fn.pos_embedder.out_dim, device=z_vals.device)
        check(_lib.lib().gf_adnerf_embed_points(ptr(rays_o), ptr(rays_d), ptr(z_vals), R, S, L, ptr(pe), pe.shape[1], stream_ptr()))
        ve = network_fn.view_embedder(viewdirs)
        return net.forward_folded(pe, cond, ve, S).view(R, S, 4)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    return network_fn.forward(pts, cond, viewdirs, run_model_fine=fine, **kwargs)['rgb_sigma']


def render_rays(ray_batch, bc_rgb, cond, network_fn, N_samples, return_raw=False, linear_disp=False, perturb=1., N_importance=0,
                white_bkgd=False, raw_noise_std=0., **kwargs):
    """volume_rendering.py:98-210.  ray_batch [R, 8 or 11] = rays_o, rays_d, near, far (, viewdirs)."""
    require_cuda(ray_batch)
    _no_grad_inputs(ray_batch, cond)
    with torch.no_grad():
        dev = ray_batch.device
        R = ray_batch.shape[0]
        rays_o, rays_d = _f32c(ray_batch[:, 0:3]), _f32c(ray_batch[:, 3:6])
        viewdirs = _f32c(ray_batch[:, -3:]) if ray_batch.shape[-1] > 8 else None
        near, far = ray_batch[:, 6:7].float(), ray_batch[:, 7:8].float()
        t_vals = torch.linspace(0., 1., steps=N_samples, device=dev)
        z_vals = near * (1. - t_vals) + far * t_vals if not linear_disp else 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)
        z_vals = z_vals.expand(R, N_samples)
        if perturb > 0.:
            mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper, lower = torch.cat([mids, z_vals[..., -1:]], -1), torch.cat([z_vals[..., :1], mids], -1)
            t_rand = torch.rand(R, N_samples, device=dev)
            t_rand[..., -1] = 1.0
            z_vals = lower + (upper - lower) * t_rand
        z_vals = z_vals.contiguous()
        bc = _f32c(bc_rgb).view(R, 3)
        raw = _query(network_fn, rays_o, rays_d, z_vals, cond, viewdirs, False, **kwargs)
        rgb_map, disp_map, acc_map, weights, depth_map, rgb_map_fg = raw2outputs(raw, z_vals, rays_d, bc, raw_noise_std, white_bkgd)
        if N_importance > 0:
            rgb_map_0, disp_map_0, acc_map_0, last_weight_0, rgb_map_fg_0 = rgb_map, disp_map, acc_map, weights[..., -1], rgb_map_fg
            z_vals, z_samples = _importance_depths(z_vals, weights, N_importance, det=(perturb == 0.))
            raw = _query(network_fn, rays_o, rays_d, z_vals, cond, viewdirs, True, **kwargs)
            rgb_map, disp_map, acc_map, weights, depth_map, rgb_map_fg = raw2outputs(raw, z_vals, rays_d, bc, raw_noise_std, white_bkgd)
        ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'rgb_map_fg': rgb_map_fg}
        if return_raw:
            ret['raw'] = raw
        if N_importance > 0:
            ret['rgb_map_coarse'], ret['disp_map_coarse'], ret['accu_map_coarse'] = rgb_map_0, disp_map_0, acc_map_0
            ret['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)
            ret['last_weight'], ret['last_weight0'], ret['rgb_map_fg0'] = weights[..., -1], last_weight_0, rgb_map_fg_0
        return ret


def batchify_render_rays(rays_flat, bc_rgb, cond, chunk, network_fn, N_samples, N_importance, **kwargs):
    """volume_rendering.py:213-231."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        c = cond if cond.squeeze().ndim == 1 else cond[i:i + chunk]
        ret = render_rays(rays_flat[i:i + chunk], bc_rgb[i:i + chunk], c.squeeze() if c.squeeze().ndim == 1 else c, network_fn, N_samples,
                          N_importance=N_importance, **kwargs)
        for k, v in ret.items():
            all_ret.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in all_ret.items()}


def render_dynamic_face(H, W, focal, cx, cy, chunk=1024, rays_o=None, rays_d=None, bc_rgb=None, cond=None, c2w=None, near=0., far=1.,
                        use_viewdirs=True, c2w_staticcam=None, network_fn=None, N_samples=None, N_importance=None, **kwargs):
    """volume_rendering.py:234-282 -> [rgb_map, disp_map, acc_map, last_weight, rgb_map_fg, {everything else}]."""
    if N_importance is None or N_importance <= 0:
        raise KeyError('last_weight')        # the reference indexes all_ret['last_weight'], which only exists with importance sampling
    if bc_rgb is not None:
        bc_rgb = bc_rgb.reshape(-1, 3)
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w, cx, cy)
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam, cx, cy)
        viewdirs = (viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)).reshape(-1, 3).float()
    sh = rays_d.shape
    rays_o, rays_d = rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float()
    near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    rays = torch.cat([rays_o, rays_d, near_t, far_t], -1)
    if use_viewdirs:
        rays = torch.cat([rays, viewdirs], -1)
    all_ret = batchify_render_rays(rays, bc_rgb, cond, chunk, network_fn=network_fn, N_samples=N_samples, N_importance=N_importance, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map', 'last_weight', 'rgb_map_fg']
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


__all__ = ['get_rays', 'FreqEmbedder', 'AudioNet', 'AudioAttNet', 'NeRFBackbone', 'ADNeRF', 'raw2outputs', 'sample_pdf', 'render_rays',
           'batchify_render_rays', 'render_dynamic_face']
_ = math
