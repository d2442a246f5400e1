"""Generate tests/golden/adnerf.npz by IMPORTING the reference's vanilla AD-NeRF path from /root/reference (pure PyTorch,
runs on CPU in the build container) on the BASELINE.json configs[0] inputs at reduced size (16x16 px so the fixture is
small), with the weights of oracle.adnerf_port.init_state(seed=0) loaded into the reference model.
Run once here:  python oracle/gen_golden_adnerf.py     (the GPU box has no /root/reference).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import adnerf_port  # noqa: E402


def main():
    from utils.commons.hparams import hparams
    hparams.update(dict(cond_dim=64, hidden_size=256, infer_scale_factor=1.0))
    from modules.nerfs.adnerf.adnerf import ADNeRF
    from modules.nerfs.commons.volume_rendering import render_dynamic_face
    torch.set_num_threads(8)
    sd = adnerf_port.init_state(seed=0)
    model = ADNeRF(hparams)
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    model.eval()
    H = W = 16
    focal = 1200.0 * H / 450.0
    c2w = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0.6]])
    cond = torch.randn(8, 16, 29, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        cond_feat = model.cal_cond_feat(cond, with_att=True)
        out = render_dynamic_face(H, W, focal, W / 2, H / 2, chunk=2048, c2w=c2w, cond=cond_feat, near=0.3, far=0.9, network_fn=model,
                                  N_samples=64, N_importance=128, perturb=0., bc_rgb=torch.ones(H, W, 3))
    rgb, disp, acc, last_w, rgb_fg, extras = out
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "adnerf.npz"), cond_feat=cond_feat.numpy(), rgb=rgb.numpy(), acc=acc.numpy(),
                        last_weight=last_w.numpy())
    print("written", rgb.shape, float(rgb.mean()), float(acc.mean()))


if __name__ == "__main__":
    main()
