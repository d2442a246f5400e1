"""GPU tests of the tensor-core training MLP (geneface_b200/tc_linear.py over the gf_tl_* operators, csrc/train_linear_tc.cu) against plain
PyTorch fp32 references of the same op (cond_encoder.py:92-111 of the reference: bias-free Linear stack with ReLU between layers)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_mlp(x, ws):
    h = x
    for l, w in enumerate(ws):
        h = h @ w.t()
        if l != len(ws) - 1:
            h = torch.relu(h)
    return h


def _mk(dims, M, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, dims[0], device="cuda", generator=g)
    ws = [torch.randn(dims[l + 1], dims[l], device="cuda", generator=g) / np.sqrt(dims[l]) for l in range(len(dims) - 1)]
    return x, ws


# the three MLPs of the head field (radnerf.py:73-105): ambient (32 + 64 -> 2), sigma (64 -> 1 + 128), colour (16 + 128 + 4 -> 3); M ragged and whole
# + the torso nets' shapes (radnerf_torso.py:25-44: 64- and 32-wide, run zero-padded to the 128-row tile)
@pytest.mark.parametrize("dims,M", [([96, 128, 128, 2], 1000), ([64, 128, 128, 129], 27531), ([148, 128, 3], 4096), ([32, 128, 128, 128, 16], 129),
                                    ([58, 64, 64, 2], 3000), ([48, 32, 32, 4], 515)])
def test_tc_mlp_forward_and_gradients_vs_fp32(dims, M):
    from geneface_b200 import tc_linear
    assert tc_linear.supported(dims)
    x, ws = _mk(dims, M, seed=sum(dims) + M)
    x1 = x.clone().requires_grad_(True)
    w1 = [w.clone().requires_grad_(True) for w in ws]
    y = tc_linear.tc_mlp(x1, w1)
    assert y.shape == (M, dims[-1]) and y.dtype == torch.float32
    g = torch.Generator(device="cuda").manual_seed(7)
    dy = torch.randn(M, dims[-1], device="cuda", generator=g) * 1e-4          # small gradients: exercises the device-side power-of-two scaling
    y.backward(dy)
    x2 = x.clone().double().requires_grad_(True)
    w2 = [w.clone().double().requires_grad_(True) for w in ws]
    yr = _ref_mlp(x2, w2)
    yr.backward(dy.double())
    torch.cuda.synchronize()

    def rel(a, b):          # relative Frobenius error
        return ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()

    # the library path under autocast (the reference's `amp: true`) on the same data: the yardstick for fp16-operand accuracy.  (A max-norm bar is
    # not meaningful for the gradients: a hidden unit whose pre-activation is within fp16 rounding of zero flips its ReLU mask in ANY fp16 forward.)
    x3 = x.clone().requires_grad_(True)
    w3 = [w.clone().requires_grad_(True) for w in ws]
    with torch.autocast("cuda", dtype=torch.float16):
        ya = _ref_mlp(x3, w3)
    ya.float().backward(dy * 4096.0)                 # a GradScaler's job: keep the fp16 gradients of the library path out of the subnormals
    for t in [x3] + w3:
        t.grad /= 4096.0
    assert rel(y, yr) < 2e-3, f"forward {rel(y, yr):.2e}"
    assert rel(y, yr) < 3 * max(rel(ya.float(), yr), 2e-4), f"forward {rel(y, yr):.2e} vs autocast {rel(ya.float(), yr):.2e}"
    assert rel(x1.grad, x2.grad) < 3 * max(rel(x3.grad, x2.grad), 1e-3), f"grad_input {rel(x1.grad, x2.grad):.2e} vs autocast {rel(x3.grad, x2.grad):.2e}"
    for l, (a, b, c) in enumerate(zip(w1, w2, w3)):
        assert a.grad.shape == b.grad.shape
        assert rel(a.grad, b.grad) < 3 * max(rel(c.grad, b.grad), 1e-3), f"grad_weight[{l}] {rel(a.grad, b.grad):.2e} vs autocast {rel(c.grad, b.grad):.2e}"
    print("dims %s M %d: forward %.1e (autocast %.1e), grad_input %.1e (%.1e), grad_weight %s" % (
        dims, M, rel(y, yr), rel(ya.float(), yr), rel(x1.grad, x2.grad), rel(x3.grad, x2.grad),
        ["%.1e (%.1e)" % (rel(a.grad, b.grad), rel(c.grad, b.grad)) for a, b, c in zip(w1, w2, w3)]))


def test_tc_mlp_exact_on_small_integers():
    """integer-valued operands are exact in fp16 and every partial sum is exact in fp32: the three products must agree with fp64 to the last bit"""
    from geneface_b200 import tc_linear
    g = torch.Generator(device="cuda").manual_seed(1)
    M, dims = 777, [64, 128, 128, 16]
    x = torch.randint(-2, 3, (M, dims[0]), device="cuda", generator=g).float()
    ws = [torch.randint(-1, 2, (dims[l + 1], dims[l]), device="cuda", generator=g).float() for l in range(3)]
    # keep activations small enough to stay exactly representable in fp16 (|h| <= 2048): sparse weights
    ws = [w * (torch.rand(w.shape, device="cuda", generator=g) < 0.1) for w in ws]
    x1 = x.clone().requires_grad_(True)
    w1 = [w.clone().requires_grad_(True) for w in ws]
    y = tc_linear.tc_mlp(x1, w1)
    dy = torch.randint(-1, 2, (M, dims[-1]), device="cuda", generator=g).float()
    y.backward(dy)
    x2 = x.clone().double().requires_grad_(True)
    w2 = [w.clone().double().requires_grad_(True) for w in ws]
    yr = _ref_mlp(x2, w2)
    yr.backward(dy.double())
    h_max = max(torch.relu(x.double() @ ws[0].double().t()).abs().max().item(), 1)
    assert h_max <= 2048
    assert torch.equal(y.double(), yr)
    if x2.grad.abs().max() <= 2048:          # hidden gradients exactly representable as well
        assert torch.equal(x1.grad.double(), x2.grad)
        for a, b in zip(w1, w2):
            assert torch.equal(a.grad.double(), b.grad)


def test_train_step_with_tc_backend_matches_the_library_backend():
    """one training step of the head model (march_rays_train + field + composite + backward) with hparams['train_mlp_backend'] = 'tc' against the
    same step on library GEMMs in fp32, with the library path under fp16 autocast (the reference's `amp: true`, loss scaled as its GradScaler does) as
    the yardstick: the tensor-core backend must be as close to the fp32 gradients as autocast is."""
    from geneface_b200 import synthetic, utils
    H = W = 64
    grads = {}
    for name, backend, amp in (("fp32", "torch", False), ("autocast", "torch", True), ("tc", "tc", False)):
        model, hp = synthetic.build_model(torso=False, bitfield='S', seed=0, train_mlp_backend=backend)
        model.train()
        fi = synthetic.frame_inputs(H, W)
        g = torch.Generator(device="cuda").manual_seed(3)
        inds = torch.randint(0, H * W, [1024], device="cuda", generator=g)
        rays = utils.get_rays(fi['pose'], fi['intrinsics'], H, W)
        rays_o, rays_d = rays['rays_o'][:, inds], rays['rays_d'][:, inds]
        bgc = utils.get_bg_coords(H, W, "cuda")[:, inds]
        target = torch.rand(1, 1024, 3, device="cuda", generator=g)
        torch.manual_seed(4)
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            out = model.render(rays_o, rays_d, fi['cond'], bgc, fi['poses6'], index=0, dt_gamma=hp['dt_gamma'], bg_color=fi['bg_color'][:, inds], perturb=True,
                               force_all_rays=False, max_steps=hp['max_steps'])
            loss = ((out['rgb_map'].float() - target) ** 2).mean()
        ls = 1024.0 if amp else 1.0
        (loss * ls).backward()
        grads[name] = (loss.item(), {n: p.grad.detach().double() / ls for n, p in model.named_parameters() if p.grad is not None})
    (l0, g0), (la, ga), (l1, g1) = grads["fp32"], grads["autocast"], grads["tc"]
    assert abs(l0 - l1) <= 2e-3 * abs(l0), (l0, l1)
    assert g0.keys() == g1.keys()
    # tensors whose gradient vanishes analytically hold rounding noise only: errors are measured against max(|reference|, 1e-4 of the largest gradient norm)
    floor = 1e-4 * max(g.norm().item() for g in g0.values())
    worst, worst_a = 0.0, 0.0
    for n in g0:
        ref = g0[n]
        den = max(ref.norm().item(), floor)
        err, err_a = (g1[n] - ref).norm().item() / den, (ga[n] - ref).norm().item() / den       # relative Frobenius errors
        worst, worst_a = max(worst, err), max(worst_a, err_a)
        # as close to fp32 as the reference's own amp arithmetic is (+ head-room for run-to-run differences: the grid gradients are reductions in
        # arbitrary order), and never far off in absolute terms
        assert err < max(2.0 * err_a, 0.05), f"{n}: {err:.2e} of its norm (library path under autocast: {err_a:.2e})"
        assert err < 0.4, f"{n}: {err:.2e}"
    print("whole-step gradients vs fp32: tensor-core backend worst %.2e, library path under autocast worst %.2e (relative Frobenius); loss %.6f / %.6f / %.6f"
          % (worst, worst_a, l0, la, l1))


def test_tc_mlp_takes_the_pieces_of_a_concatenated_input():
    """tc_mlp([a, b.expand(M, -1), c], W) == tc_mlp(torch.cat([a, b.repeat(M, 1), c], 1), W) bit for bit (the pieces are packed into the same tiles),
    gradients included; the broadcast piece's gradient arrives summed over the samples"""
    from geneface_b200 import tc_linear
    g = torch.Generator(device="cuda").manual_seed(5)
    M = 2000
    a = torch.randn(M, 16, device="cuda", generator=g)
    b = torch.randn(1, 128, device="cuda", generator=g)
    c = torch.randn(1, 4, device="cuda", generator=g)
    ws = [torch.randn(128, 148, device="cuda", generator=g) / 12, torch.randn(3, 128, device="cuda", generator=g) / 11]
    dy = torch.randn(M, 3, device="cuda", generator=g)
    res = []
    for mode in ("parts", "cat"):
        a1, b1, c1 = (t.clone().requires_grad_(True) for t in (a, b, c))
        w1 = [w.clone().requires_grad_(True) for w in ws]
        if mode == "parts":
            y = tc_linear.tc_mlp([a1, b1.expand(M, -1), c1.expand(M, -1)], w1)
        else:
            y = tc_linear.tc_mlp(torch.cat([a1, b1.repeat(M, 1), c1.repeat(M, 1)], dim=1), w1)
        y.backward(dy)
        res.append((y.detach(), a1.grad, b1.grad, c1.grad, w1[0].grad, w1[1].grad))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for i in (2, 3):                      # sums over M samples in different orders
        assert torch.allclose(res[0][i], res[1][i], rtol=1e-4, atol=1e-5 * res[1][i].abs().max().item())
    for i in (4, 5):                      # reductions in arbitrary order
        assert torch.allclose(res[0][i], res[1][i], rtol=1e-4, atol=1e-5 * res[1][i].abs().max().item())
