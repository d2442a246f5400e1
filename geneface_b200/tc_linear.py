"""Tensor-core MLP for the training step: forward, grad_input and grad_weight of the reference's bias-free MLP
(modules/radnerfs/cond_encoder.py:92-111: Linear(bias=False) -> ReLU -> ... -> Linear; used by radnerf.py:73-105) on the
`gf_tl_*` tcgen05 operators of libgfrender (csrc/train_linear_tc.cu) instead of library GEMMs under autograd.

Arithmetic: fp16 operands, fp32 accumulation, fp32 weights and gradients -- what `torch.autocast(float16)` gives the reference under its
default `amp: true`; the incoming gradient is scaled by a power of two chosen on the device (largest |dy| -> ~2^8) before it is rounded to
fp16 and the factor is divided out of grad_input / grad_weight in their fp32 epilogues, so the result does not depend on an outer GradScaler.

Envelope: hidden width <= 128 (narrower hidden layers -- the 64 / 32-wide torso nets -- run zero-padded to the 128-row tensor-core tile of the
weight-gradient product), 2 or more layers, input / output widths <= 256 / 144.  `supported(dims)` says whether an `MLP` is inside it; outside it
`MLP.forward` keeps the library path.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _chunks(k):
    return (k + 63) // 64


def _pad16(n):
    return (n + 15) // 16 * 16


def supported(dims):
    """dims = [in, hidden, ..., hidden, out]"""
    if len(dims) < 3:
        return False
    return all(16 <= h <= 128 for h in dims[1:-1]) and dims[0] <= 256 and dims[-1] <= 144


def _tiles(M, chunks, device):
    return torch.empty(int(_lib.lib().gf_tl_tiles_bytes(M, chunks)), dtype=torch.uint8, device=device)


def _pack(src, K, chunks, scale=None):
    M = src.shape[0]
    out = _tiles(M, chunks, src.device)
    check(_lib.lib().gf_tl_pack(ptr(src), 1 if src.dtype == torch.float16 else 0, src.stride(0), K, M, chunks, 0, 0, ptr(scale), ptr(out), stream_ptr()), "gf_tl_pack")
    return out


def _prep(x):
    """a part of the input as the pack kernel reads it: fp32 / fp16, unit column stride; an expanded row (stride(0) == 0) stays a broadcast"""
    x = x.detach()
    if x.dtype not in (torch.float32, torch.float16):
        x = x.float()
    if x.dim() != 2 or (x.shape[1] > 1 and x.stride(1) != 1):
        x = x.contiguous()
    return x


def _pack_parts(parts, chunks):
    """torch.cat(parts, dim=1) straight into tile layout: one gf_tl_pack per part over its own column range (column offsets are multiples of 8)"""
    M = parts[0].shape[0]
    out = _tiles(M, chunks, parts[0].device)
    col = 0
    for i, p in enumerate(parts):
        K = p.shape[1]
        last = i == len(parts) - 1
        col1 = 64 * chunks if last else col + K
        check(_lib.lib().gf_tl_pack(ptr(p), 1 if p.dtype == torch.float16 else 0, p.stride(0), K, M, chunks, col, col1, None, ptr(out), stream_ptr()), "gf_tl_pack")
        col += K
    return out


def _image(W, rows=None, chunks=None):
    """fp16 image of W [N, K]: `chunks` blocks of [rows x 128 B]; defaults: rows = ceil16(N), chunks = ceil(K / 64) (zero padded)"""
    N, K = W.shape
    rows, chunks = rows or _pad16(N), chunks or _chunks(K)
    img = torch.empty(rows * chunks * 128, dtype=torch.uint8, device=W.device)
    check(_lib.lib().gf_tl_weight_image(ptr(W), N, K, rows, chunks, ptr(img), stream_ptr()), "gf_tl_weight_image")
    return img, rows, chunks


class TcMLPFunction(torch.autograd.Function):
    """y = W_L relu(... relu(W_0 x)); x = cat(parts, dim=1) with parts [M, K_i] fp32 or fp16 (an expanded row stays a broadcast), W_l fp32
    [N_l, K_l]; y [M, N_L] fp32.  apply(n_parts, *parts, *weights)."""

    @staticmethod
    def forward(ctx, n_parts, *args):
        _lib.require_cuda()
        L = _lib.lib()
        parts = [_prep(p) for p in args[:n_parts]]
        weights = args[n_parts:]
        ws = [w.detach().float().contiguous() for w in weights]
        M, K0 = parts[0].shape[0], sum(p.shape[1] for p in parts)
        dims = [K0] + [w.shape[0] for w in ws]
        assert supported(dims), "TcMLPFunction: layer widths outside the tensor-core envelope"
        assert all(p.shape[0] == M for p in parts)
        ctx.part_cols = [p.shape[1] for p in parts]
        ctx.part_dtypes = [p.dtype for p in parts]
        dev = parts[0].device
        acts = [_pack_parts(parts, _chunks(K0))]               # tile-major fp16 inputs of every layer (saved for backward)
        imgs = []
        # fp32 rows leave the kernels with 16-byte stores: the pitch of an odd-width output is padded to a multiple of 4 (the caller sees a view)
        n_out, ld_y = dims[-1], (dims[-1] + 3) // 4 * 4
        y_full = torch.empty(M, ld_y, dtype=torch.float32, device=dev)
        y = y_full[:, :n_out]
        for l, w in enumerate(ws):
            last = l == len(ws) - 1
            # hidden layers are held 128 wide (2 chunks): a narrower layer's extra rows / columns are zero in the images, so its padded activations and
            # gradients are exactly zero
            img, rows, chunks = _image(w, rows=None if last else 128, chunks=None if l == 0 else 2)
            imgs.append((img, rows, chunks))
            if last:
                check(L.gf_tl_gemm(ptr(acts[l]), chunks, ptr(img), rows, chunks, 0, M, None, 0, 0, None, 0, ptr(y_full), ld_y, min(ld_y, rows), None, stream_ptr()),
                      "gf_tl_gemm(forward, output layer)")
            else:
                h = _tiles(M, 2, dev)
                check(L.gf_tl_gemm(ptr(acts[l]), chunks, ptr(img), rows, chunks, 0, M, ptr(h), 2, 1, None, 0, None, 0, 0, None, stream_ptr()), "gf_tl_gemm(forward)")
                acts.append(h)
        ctx.acts, ctx.imgs, ctx.dims, ctx.M = acts, imgs, dims, M
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _lib.lib()
        acts, imgs, dims, M = ctx.acts, ctx.imgs, ctx.dims, ctx.M
        nl = len(imgs)
        dev = dy.device
        dy = dy.detach()
        if dy.dtype not in (torch.float32, torch.float16):
            dy = dy.float()
        if dy.stride(-1) != 1:
            dy = dy.contiguous()
        # power-of-two scale: largest |dy| -> about 2^8 (head-room for the growth through the layers before fp16 overflows at 2^16); on the device
        amax = dy.abs().amax().float().clamp_min(1e-30)
        scale = torch.exp2(torch.floor(8.0 - torch.log2(amax))).clamp(2.0 ** -20, 2.0 ** 40).reshape(1).contiguous()
        inv = (1.0 / scale).contiguous()
        n_out = dims[-1]
        g = _pack(dy, n_out, _chunks(_pad16(n_out)), scale)              # dY tiles (scaled)
        g_chunks = _chunks(_pad16(n_out))
        grads_w = [None] * nl
        dx = None
        for l in range(nl - 1, -1, -1):
            img, rows, chunks = imgs[l]
            N_l, K_l = dims[l + 1], dims[l]
            dw = torch.zeros(N_l, K_l, dtype=torch.float32, device=dev)
            if l == nl - 1:
                # output layer: M side = its 128-wide input activation, N side = dY (padded to 16): D = dW^T
                check(L.gf_tl_wgrad(ptr(acts[l]), 2, 0, ptr(g), g_chunks, rows, M, ptr(dw), K_l, K_l, N_l, 1, ptr(inv), stream_ptr()), "gf_tl_wgrad(output layer)")
            else:
                # M side = this layer's (128-wide) output gradient, N side = its input (padded to whole chunks): D = dW
                check(L.gf_tl_wgrad(ptr(g), 2, 0, ptr(acts[l]), chunks, min(_pad16(K_l), 64 * chunks), M, ptr(dw), K_l, N_l, K_l, 0, ptr(inv), stream_ptr()), "gf_tl_wgrad")
            grads_w[l] = dw
            if l > 0:
                gn = _tiles(M, 2, dev)
                check(L.gf_tl_gemm(ptr(g), g_chunks, ptr(img), rows, chunks, 1, M, ptr(gn), 2, 0, ptr(acts[l]), 2, None, 0, 0, None, stream_ptr()), "gf_tl_gemm(dgrad)")
                g, g_chunks = gn, 2
            elif any(ctx.needs_input_grad[1:1 + len(ctx.part_cols)]):
                ld_x = (K_l + 3) // 4 * 4
                dx_full = torch.empty(M, ld_x, dtype=torch.float32, device=dev)
                check(L.gf_tl_gemm(ptr(g), g_chunks, ptr(img), rows, chunks, 1, M, None, 0, 0, None, 0, ptr(dx_full), ld_x, min(ld_x, 64 * chunks), ptr(inv), stream_ptr()),
                      "gf_tl_gemm(grad_input)")
                dx = dx_full[:, :K_l]
        # the gradient of the concatenated input goes back to its parts as column slices (a broadcast part's expand() sums it in autograd)
        gparts, col = [], 0
        for i, k in enumerate(ctx.part_cols):
            gp = None
            if dx is not None and ctx.needs_input_grad[1 + i]:
                gp = dx[:, col:col + k]
                if ctx.part_dtypes[i] == torch.float16:
                    gp = gp.half()
            gparts.append(gp)
            col += k
        return (None, *gparts, *grads_w)


def tc_mlp(x, weights):
    """x: one [M, K] tensor or a list of parts to be concatenated along dim 1 (parts other than the last must be a multiple of 8 columns wide)"""
    parts = list(x) if isinstance(x, (list, tuple)) else [x]
    if any(p.shape[1] % 8 for p in parts[:-1]):
        parts = [torch.cat(parts, dim=1)]
    if parts[0].shape[0] == 0:               # no samples: keep the graph, nothing to launch
        return parts[0].new_zeros(0, weights[-1].shape[0], dtype=torch.float32) + 0.0 * sum(w.sum() for w in weights)
    return TcMLPFunction.apply(len(parts), *parts, *weights)
