#!/usr/bin/env python
"""Phase timeline of the two tcgen05 field kernels (experiment build -DGF_TC_TIMING=1 of libgfrender.so):

    python scripts/build_variants.py timing=-DGF_TC_TIMING=1
    GF_LIBGFRENDER=geneface_b200/variants/libgfrender_timing.so python scripts/tc_timeline.py [out.json]

CTA 0 stamps clock64() at every phase boundary of 64 steady-state tiles (field_tc_split.cu, TT_STAMP); this script renders the
benchmark frame (512x512x128, bound 4), reads the stamps of the last full round and prints mean cycles per phase for the consumer
streams and the producer halves of both kernels, plus the tile period (time between consecutive tiles of one stream / producer)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from geneface_b200 import _lib, synthetic

NJ = 64
H = W = 512


def main():
    model, hp = synthetic.build_model(torso=True, bitfield='F', seed=0, sigma_scale=0.25, bound=4)
    fi = synthetic.frame_inputs(H, W)
    timing = "timing" in os.environ.get("GF_LIBGFRENDER", "")
    buf = torch.zeros(2 * 4 * NJ * 8, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        cf = model.cal_cond_feat(fi['cond'])
        handle = model.gf_model()
        if timing:                       # only the -DGF_TC_TIMING build reads the debug pointer as the stamp buffer
            _lib.check(_lib.lib().gf_tc_debug(handle, _lib.ptr(buf)))
        for _ in range(3):
            model.render_fused(cf, H, W, pose=fi['pose'][0], intrinsics=fi['intrinsics'], bg_color=fi['bg_color'], torso_pose=fi['poses6'],
                               dt_gamma=0.0, max_steps=128, precision='fp16')
        torch.cuda.synchronize()
        _lib.lib().gf_tc_debug(handle, None)
        # whole-frame time (eager launches, stamps off) and a sanity check of the fp16 frame against the fp32 field of the same library
        kw = dict(pose=fi['pose'][0], intrinsics=fi['intrinsics'], bg_color=fi['bg_color'], torso_pose=fi['poses6'], dt_gamma=0.0, max_steps=128)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            model.render_fused(cf, H, W, precision='fp16', **kw)
        e0.record()
        for _ in range(40):
            out16 = model.render_fused(cf, H, W, precision='fp16', **kw)
        e1.record()
        torch.cuda.synchronize()
        rgb16 = out16['rgb_map'].clone()
        rgb32 = model.render_fused(cf, H, W, precision='fp32', **kw)['rgb_map']
        print("frame %.3f ms (eager, 40 frames)   fp16 vs fp32 field: max |d rgb| %.2e" % (e0.elapsed_time(e1) / 40, (rgb16 - rgb32).abs().max().item()))
    T = buf.cpu().numpy().reshape(2, 4, NJ, 8)
    res = {}
    names = {
        (0, 'cons'): ["wait_full", "L0 mma", "epi0 (split)", "bar", "L1 mma", "out layer 128->2 + store", "-> next tile"],
        (1, 'cons'): ["wait_full", "sig0 mma", "epi 0", "bar + sig1 mma (+SH)", "epi 1", "bar + merged mma", "epi + bar + issue col1", "-> next tile (col1 mma, out)"],
        (0, 'prod'): ["wait_empty", "batch 0", "batch 1", "fence + arrive", "-> next tile (pos prefetch)"],
        (1, 'prod'): ["wait_empty", "stores + gather", "-> next tile (prefetch)"],
    }
    for kern in (0, 1):
        for who in range(4):
            role = 'cons' if who < 2 else 'prod'
            t = T[kern, who]                       # [NJ][8]
            # a stream handles every second tile j; a producer every tile
            js = [j for j in range(NJ) if t[j, 0] != 0]
            if len(js) < 4:
                continue
            nst = len(names[(kern, role)])
            rows = np.array([t[j, :nst] for j in js], dtype=np.float64)
            d = np.diff(rows, axis=1)              # phase durations within a tile
            nxt = np.array([t[js[i + 1], 0] - t[js[i], nst - 1] for i in range(len(js) - 1)], dtype=np.float64)
            period = np.array([t[js[i + 1], 0] - t[js[i], 0] for i in range(len(js) - 1)], dtype=np.float64)
            key = "k%s %s%d" % ("AB"[kern], role, who if who < 2 else who - 2)
            res[key] = {"tiles": len(js), "period_mean": float(period.mean()), "period_p50": float(np.median(period)),
                        "phases": {names[(kern, role)][i]: float(d[:, i].mean()) for i in range(nst - 1)}}
            res[key]["phases"][names[(kern, role)][nst - 1]] = float(nxt.mean())
            print("%-10s tiles %2d  period mean %7.0f  p50 %7.0f" % (key, len(js), period.mean(), np.median(period)))
            for k, v in res[key]["phases"].items():
                print("      %-34s %7.0f" % (k, v))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
