"""Generate tests/golden/frame_{may,may_torso,b4}.npz with the REAL reference: the unmodified
`modules.radnerfs.radnerf(_torso).RADNeRF(Torso).render()` (oracle/ref_model.py -> oracle/_ref/pyref + oracle/_ref/*.so)
running in eval / fp32 / perturb=False on a GPU, on rays from the reference's own `get_rays`, `get_bg_coords`, `convert_poses`.

    gpurun -- python oracle/gen_golden_frames.py gpurun_out/golden      # then copy frame_*.npz to tests/golden/ and commit

Weights are the seeded synthetic models of geneface_b200/synthetic.py (regenerated identically by the tests; a checksum of
the state_dict is stored so that a drifting generator is detected rather than mis-reported as a parity failure).  Stored:
the rays, the reference result dict (rgb_map, depth_map, torso maps), weights_sum (captured from the in-place buffer of the
reference's `composite_rays` calls), the host loop's (n_alive, n_step) trace, the per-ray termination iteration (the iteration in
which the reference's composite_rays set rays_alive = -1) and the per-ray marched-sample counts.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SCENES = {
    #  name        torso  bitfield seed sigma bound  H    dt_gamma  max_steps
    "may":        (False, 'S',     4,   4.0,  1,    128, 1 / 256,  16),
    "may_torso":  (True,  'S',     4,   4.0,  1,    128, 1 / 256,  16),
    "b4":         (True,  'F',     0,   0.25, 4,    64,  0.0,      128),
}


def state_checksum(sd):
    """order-independent float64 checksum of a state_dict (detects a different random stream, not a parity metric)."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].detach().double().cpu()
        tot += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64).view(v.shape).remainder(7.0).add(1.0)).sum())
    return tot


def scene_model(name, device="cuda"):
    from geneface_b200 import synthetic
    torso, bf, seed, sigma, bound, H, dt_gamma, max_steps = SCENES[name]
    model, hp = synthetic.build_model(torso=torso, bitfield=bf, seed=seed, sigma_scale=sigma, bound=bound, device=device)
    fi = synthetic.frame_inputs(H, H, device=device)
    return model, hp, fi, dict(torso=torso, H=H, dt_gamma=dt_gamma, max_steps=max_steps, bound=bound)


def main(out_dir):
    from oracle import ref_model
    os.makedirs(out_dir, exist_ok=True)
    ns = ref_model.load()
    for name in SCENES:
        model, hp, fi, cfg = scene_model(name)
        H = cfg["H"]
        ref = ref_model.build(model.state_dict(), hp, torso=cfg["torso"])
        rays = ns.utils.get_rays(fi["pose"], fi["intrinsics"], H, H, -1)
        bgc = ns.utils.get_bg_coords(H, H, "cuda")
        poses6 = ns.utils.convert_poses(fi["pose"])
        res = ref_model.render(ref, rays["rays_o"], rays["rays_d"], fi["cond"], bgc, poses6, fi["bg_color"], cfg["dt_gamma"], cfg["max_steps"])
        torch.cuda.synchronize()
        out = dict(
            rays_o=rays["rays_o"][0].cpu().numpy(), rays_d=rays["rays_d"][0].cpu().numpy(), bg_coords=bgc[0].cpu().numpy(),
            poses6=poses6.cpu().numpy(),                          # cond and bg_color are regenerated from their seeds by the tests
            rgb_map=res["rgb_map"][0].cpu().numpy(), depth_map=res["depth_map"][0].cpu().numpy(),
            weights_sum=res["weights_sum"].cpu().numpy(), trace=np.asarray(res["trace"], np.int32), term_iter=res["term_iter"].cpu().numpy().astype(np.int16),
            n_marched=res["n_marched"].cpu().numpy().astype(np.int16),
            state_checksum=np.float64(state_checksum(model.state_dict())), H=np.int32(H), dt_gamma=np.float32(cfg["dt_gamma"]),
            max_steps=np.int32(cfg["max_steps"]), bound=np.float32(cfg["bound"]), torso=np.int32(cfg["torso"]),
        )
        if np.all(out["rays_o"] == out["rays_o"][:1]):
            out["rays_o"] = out["rays_o"][:1]                     # one camera origin
        if cfg["torso"]:
            out["torso_alpha_map"] = res["torso_alpha_map"][:, 0].cpu().numpy()
            out["torso_rgb_map"] = res["torso_rgb_map"].view(-1, 3).cpu().numpy()
            out["deform"] = res["deform"].cpu().numpy() if "deform" in res else np.zeros((0, 2), np.float32)
        np.savez_compressed(os.path.join(out_dir, f"frame_{name}.npz"), **out)
        print(f"frame_{name}: {H}x{H}, loop iterations {len(res['trace'])}, S_total {sum(s for _, s in res['trace'])}, "
              f"marched samples {int(res['n_marched'].sum())}, rgb mean {float(res['rgb_map'].mean()):.4f}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
