#!/usr/bin/env python
"""Diagnostics of the gf_tl_* operators one product at a time (integer data -> exact expected results):  forward (K-major operands), grad_input
(MN-major weight image), grad_weight (both operands MN-major).  GF_TL_MN_SWAP=1 exchanges the LBO / SBO fields of the MN-major descriptors."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geneface_b200 import _lib, tc_linear as T
from geneface_b200._lib import check, ptr, stream_ptr

L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)


def unpack(tiles, M, chunks):
    """tile layout -> [M, 64 * chunks] float (host-side inverse of the swizzle)"""
    nt = (M + 127) // 128
    raw = tiles.view(nt, chunks, 128, 8, 16)                 # [tile][chunk][row][unit][16 B]
    rows = torch.arange(128, device=tiles.device)
    units = torch.arange(8, device=tiles.device)
    src_unit = units[None, :] ^ (rows[:, None] & 7)           # logical unit u lives at physical unit u ^ (row & 7)
    out = raw[:, :, rows[:, None], src_unit]                  # [tile][chunk][row][unit][16]
    h = out.contiguous().view(torch.float16).view(nt, chunks, 128, 64)
    return h.permute(0, 2, 1, 3).reshape(nt * 128, chunks * 64)[:M].float()


for (M, K, N) in [(300, 64, 128), (1000, 96, 128), (129, 128, 16), (4096, 128, 144), (555, 148, 128)]:
    x = torch.randint(-2, 3, (M, K), device="cuda", generator=g).float()
    W = torch.randint(-1, 2, (N, K), device="cuda", generator=g).float() * (torch.rand(N, K, device="cuda", generator=g) < 0.2)
    ck, cn, rows = T._chunks(K), T._chunks(T._pad16(N)), T._pad16(N)
    xt = T._pack(x, K, ck)
    assert torch.equal(unpack(xt, M, ck)[:, :K], x), "pack/unpack"
    img, r2, c2 = T._image(W)
    # forward
    y = torch.empty(M, N, device="cuda")
    yt = T._tiles(M, cn, "cuda")
    check(L.gf_tl_gemm(ptr(xt), ck, ptr(img), rows, ck, 0, M, ptr(yt), cn, 0, None, 0, ptr(y), N, N, None, stream_ptr()))
    torch.cuda.synchronize()
    ref = x @ W.t()
    e_f = (y - ref).abs().max().item()
    e_ft = (unpack(yt, M, cn)[:, :N] - ref).abs().max().item() if ref.abs().max() < 2048 else float('nan')
    # grad_input: dX = dY W
    dy = torch.randint(-1, 2, (M, N), device="cuda", generator=g).float()
    dyt = T._pack(dy, N, cn)
    dx = torch.empty(M, K, device="cuda")
    check(L.gf_tl_gemm(ptr(dyt), cn, ptr(img), rows, ck, 1, M, None, 0, 0, None, 0, ptr(dx), K, K, None, stream_ptr()))
    torch.cuda.synchronize()
    e_d = (dx - dy @ W).abs().max().item()
    # grad_weight
    res = {}
    if K >= 128:      # M side = x (first 128 features), N side = dy: D = (x^T dy) -> written transposed into dw [N][K]
        dw = torch.zeros(N, K, device="cuda")
        check(L.gf_tl_wgrad(ptr(xt), ck, 0, ptr(dyt), cn, rows, M, ptr(dw), K, min(K, 128), N, 1, None, stream_ptr()))
        torch.cuda.synchronize()
        res['wgrad(T)'] = (dw[:, :128] - (dy.t() @ x)[:, :128]).abs().max().item()
    if N >= 128:      # M side = dy (first 128 outputs), N side = x
        dw = torch.zeros(N, K, device="cuda")
        check(L.gf_tl_wgrad(ptr(dyt), cn, 0, ptr(xt), ck, min(T._pad16(K), 64 * ck), M, ptr(dw), K, 128, K, 0, None, stream_ptr()))
        torch.cuda.synchronize()
        res['wgrad'] = (dw[:128] - (dy.t() @ x)[:128]).abs().max().item()
    print("M %5d K %3d N %3d  swap=%s  forward %.3g (tiles %.3g)  dgrad %.3g  %s   |ref| fwd %.0f dgrad %.0f wgrad %.0f" % (
        M, K, N, os.environ.get("GF_TL_MN_SWAP", "0"), e_f, e_ft, e_d, res, ref.abs().max().item(), (dy @ W).abs().max().item(), (dy.t() @ x).abs().max().item()), flush=True)
