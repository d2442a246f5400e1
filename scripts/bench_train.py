#!/usr/bin/env python
"""Train-step context benchmark (BASELINE.json configs[4], SURVEY.md section 8d "Config 5"): 4096 rays through
march_rays_train -> field (grid encoders + torch MLPs) -> composite_rays_train, MSE loss, backward (composite backward,
hash-grid backward, SH/freq backward) on the fine-grained ops of libgfrender.  Prints one JSON line with the step time and
the per-kernel share (torch profiler).  Not part of bench.py's headline: training is the reference's secondary path.

usage: python scripts/bench_train.py [--rays 4096] [--steps 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--amp", action="store_true", help="fp16 autocast + GradScaler on both arms (the reference trains with amp: true)")
    ap.add_argument("--graph-only", action="store_true", help="(internal) only the CUDA-graph-captured step, in its own process")
    ap.add_argument("--mlp", default="torch", choices=["torch", "tc"],
                    help="MLP backend of OUR arm: library GEMMs under autograd, or the gf_tl_* tcgen05 operators (fp16 operands, fp32 accumulation = amp arithmetic)")
    ap.add_argument("--no-ref", action="store_true", help="skip the reference's own step")
    args = ap.parse_args()
    print(json.dumps(run(args.steps, args.warmup, 0, args.rays, args.amp, args.graph_only, args.mlp, args.no_ref)), flush=True)


def run(steps=20, warmup=5, local=0, rays=4096, amp=False, graph_only=False, mlp="torch", no_ref=False):
    args = argparse.Namespace(steps=steps, warmup=warmup, rays=rays, amp=amp, mlp=mlp)
    import torch
    from geneface_b200 import synthetic, utils
    assert torch.cuda.is_available(), "needs a GPU"
    dev = torch.device("cuda", local)
    H = W = 512
    model, hp = synthetic.build_model(torso=False, bitfield='S', seed=0, device=dev, train_mlp_backend=mlp)
    model.train()
    fi = synthetic.frame_inputs(H, W, device=dev)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(3)
    inds = torch.randint(0, H * W, [args.rays], device=dev, generator=g)
    rays = utils.get_rays(fi['pose'], fi['intrinsics'], H, W)
    rays_o, rays_d = rays['rays_o'][:, inds], rays['rays_d'][:, inds]
    bgc = utils.get_bg_coords(H, W, dev)[:, inds]
    target = torch.rand(1, args.rays, 3, device=dev, generator=g)
    bg_color = fi['bg_color'][:, inds]

    scaler = torch.amp.GradScaler('cuda', enabled=args.amp)

    def step():
        torch.manual_seed(4)                                   # perturb noise
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.float16, enabled=args.amp):
            out = model.render(rays_o, rays_d, fi['cond'], bgc, fi['poses6'], index=0, dt_gamma=hp['dt_gamma'], bg_color=bg_color, perturb=True,
                               force_all_rays=False, max_steps=hp['max_steps'])
            loss = ((out['rgb_map'].float() - target) ** 2).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss

    step()
    if graph_only:
        model.mean_count = int(model.step_counter[(model.local_step - 1) % 16, 0].item() * 1.1)
        return graphed_step_ms(model, hp, fi, rays_o, rays_d, bgc, bg_color, target, args)
    # steady state: the sample budget M comes from the running mean of the previous steps (update_extra_state, renderer.py:255-258), so
    # march_rays_train does not synchronise on the sample count; the mean of this (static) batch + 10 % plays that role here
    model.mean_count = int(model.step_counter[(model.local_step - 1) % 16, 0].item() * 1.1)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # kernel shares
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:12]
    tot = sum(r.device_time_total for r in prof.key_averages()) or 1.0
    ref_line = {"skipped": True} if no_ref else reference_train_step(model, hp, fi, rays_o, rays_d, bgc, bg_color, target, args)
    # the graph-captured step runs in its own process: a failed capture must not leave this process's RNG / context in capture mode
    import subprocess
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--graph-only", "--rays", str(args.rays), "--steps", str(args.steps), "--warmup", str(args.warmup), "--mlp", mlp]
        r = subprocess.run(cmd + (["--amp"] if args.amp else []), capture_output=True, text=True, timeout=240)
        graph_ms = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else {"unavailable": (r.stderr or r.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        graph_ms = {"unavailable": repr(e)[:300]}
    line = {"metric": "train step, %d rays (march_rays_train + field + composite + backward + Adam)" % args.rays, "ms_per_step": ms, "amp": bool(args.amp), "mlp": mlp,
            "reference_cuda": ref_line, "mean_count": int(model.mean_count), "cuda_graph": graph_ms, "grid_backward": os.environ.get("GF_GRID_BWD", "b200 (privatised small levels)"),
            "rays_per_s": args.rays / (ms / 1000.0), "loss": float(loss), "grads_finite": bool(all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)),
            "top_kernels": [{"name": r.key[:70], "share": r.device_time_total / tot, "calls": r.count} for r in rows]}
    return line


def graphed_step_ms(model, hp, fi, rays_o, rays_d, bgc, bg_color, target, args):
    """The same training step captured ONCE into a CUDA graph (forward through the libgfrender operators, loss, backward, Adam with
    capturable state) and replayed: with a fixed sample budget (mean_count > 0) the step has no host synchronisation, so the ~300
    launches the eager step pays for one by one (the 4096-ray step is host-bound in both implementations) collapse into one replay."""
    import torch
    try:
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-3, capturable=True)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.float16, enabled=args.amp):
                out = model.render(rays_o, rays_d, fi['cond'], bgc, fi['poses6'], index=0, dt_gamma=hp['dt_gamma'], bg_color=bg_color, perturb=True,
                                   force_all_rays=False, max_steps=hp['max_steps'])
                loss = ((out['rgb_map'].float() - target) ** 2).mean()
            loss.backward()
            opt.step()
            return loss
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = step()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        return {"ms_per_step": ms, "rays_per_s": args.rays / (ms / 1000.0), "loss": float(loss), "finite": bool(torch.isfinite(loss))}
    except Exception as e:  # noqa: BLE001
        import traceback
        return {"unavailable": repr(e)[:600], "trace": traceback.format_exc()[-1500:]}


def reference_train_step(model, hp, fi, rays_o, rays_d, bgc, bg_color, target, args):
    """The reference's OWN training step on the same GPU / weights / rays (BASELINE.md B-REF-TRAIN): unmodified RADNeRF.render() in
    train mode (its march_rays_train / composite_rays_train / grid backward extensions, torch fp32 MLPs) + MSE + backward + Adam."""
    import torch
    try:
        from oracle import ref_model
        if not ref_model.available():
            return {"unavailable": "oracle/_ref not built"}
        ref = ref_model.build(model.state_dict(), hp, torso=False, device=rays_o.device)
        ref.train()
        ref.mean_count = int(model.mean_count)
        opt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=1e-3)
        scaler = torch.amp.GradScaler('cuda', enabled=args.amp)

        def step():
            torch.manual_seed(4)
            opt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.float16, enabled=args.amp):
                out = ref.render(rays_o, rays_d, fi['cond'], bgc, fi['poses6'], index=0, dt_gamma=hp['dt_gamma'], bg_color=bg_color, perturb=True,
                                 force_all_rays=False, max_steps=hp['max_steps'])
                loss = ((out['rgb_map'].float() - target) ** 2).mean()
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            return loss
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        return {"ms_per_step": ms, "rays_per_s": args.rays / (ms / 1000.0), "loss": float(loss),
                "what": "the reference's own train step (unmodified render() in train mode + autograd on its compiled extensions, fp32), same GPU"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:300]}


if __name__ == "__main__":
    main()
