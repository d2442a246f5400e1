#!/usr/bin/env python
"""Render a talking-head sequence end to end on this rank's GPU(s): the replacement of the frame loop of
inference/nerfs/base_nerf_infer.py:131-179 (BASELINE.json configs[3]: 300 frames 512x512, frames sharded across the GPUs).

  ingress   geneface_b200.ingress      dataset file + landmark sequence -> poses, condition windows, background (host)
  render    geneface_b200.sequence     one NCCL parameter broadcast, rank-block frame partition, pipelined fused frames
  egress    geneface_b200.egress       PNG encoder pool writing <out>/00000.png ... (the reference's naming)

Single GPU:   python scripts/render_sequence.py --synthetic --frames 300 --out /tmp/frames
Multi GPU:    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/render_sequence.py --synthetic --frames 300 --out /tmp/frames
With data:    ... --data <binary_data_dir containing trainval_dataset.npy> --lm3d <pred_lm3d.npy> --ckpt <checkpoint.ckpt | state_dict.pt>
              [--smooth-kernel 7] [--bg white|black|<image>]   (the reference's infer_* defaults, egs/egs_bases/radnerf/base.yaml:110-116)
Prints one JSON line per run (rank 0): frames, seconds, frames/s including PNG encoding, bytes written.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def unwrap_checkpoint(ckpt, model_name="model"):
    """A reference trainer checkpoint -> the model's state_dict (utils/commons/ckpt_utils.py:26-60): {'state_dict': {...}} with either
    'model.'-prefixed flat keys or a nested {'model': state_dict}; a bare state_dict passes through."""
    sd = ckpt.get("state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if isinstance(sd, dict) and model_name in sd and isinstance(sd[model_name], dict):
        return sd[model_name]
    if isinstance(sd, dict) and any(k.startswith(model_name + ".") for k in sd):
        return {k[len(model_name) + 1:]: v for k, v in sd.items() if k.startswith(model_name + ".")}
    return sd


def background_image(spec, dataset_bg, H, W):
    """infer_bg_img_fname semantics (tasks/radnerfs/dataset_utils.py:62-78): '' = dataset background, 'white', 'black', else an image file
    (BGR(A) -> RGB, resized to W x H, /255).  Returns float32 [H, W, 3] in [0, 1]."""
    import numpy as np
    if spec == "":
        return np.asarray(dataset_bg, np.float32)
    if spec == "white":
        return np.ones((H, W, 3), np.float32)
    if spec == "black":
        return np.zeros((H, W, 3), np.float32)
    import cv2
    img = cv2.imread(spec, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise SystemExit("cannot read background image %s" % spec)
    if img.shape[0] != H or img.shape[1] != W:
        img = cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA)
    img = cv2.cvtColor(img, cv2.COLOR_BGRA2RGB if img.ndim == 3 and img.shape[2] == 4 else cv2.COLOR_BGR2RGB)
    return img.astype(np.float32) / 255.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--synthetic", action="store_true", help="random-weight May-configuration model and a synthetic pose/landmark sequence")
    ap.add_argument("--data", default=None, help="directory with trainval_dataset.npy")
    ap.add_argument("--lm3d", default=None, help=".npy with the predicted idexp_lm3d sequence [1, T, 204]")
    ap.add_argument("--ckpt", default=None, help="RADNeRFTorso weights: a bare state_dict or a reference trainer checkpoint "
                    "({'state_dict': {'model': ...}} / 'model.'-prefixed keys, utils/commons/ckpt_utils.py:26-60); loaded strictly")
    ap.add_argument("--smooth-kernel", type=int, default=7, help="camera-path smoothing window (infer_smooth_camera_path_kernel_size, "
                    "egs/egs_bases/radnerf/base.yaml:115-116: on, 7); 0 disables")
    ap.add_argument("--bg", default="", help="infer_bg_img_fname (base.yaml:114): '' = the dataset background, 'white', 'black', or an image file")
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--out", default=None, help="directory for the PNG frames (omit to skip encoding)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--clamp-std", type=float, default=2.5)
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    from geneface_b200 import egress, ingress, sequence, synthetic
    from geneface_b200.utils import orbit_pose

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    # ---- ingress (host) ----
    if args.synthetic or args.data is None:
        H = W = args.size
        model, hp = synthetic.build_model(torso=True, bitfield='S', seed=0 if rank == 0 else 100 + rank, device=dev)
        fi = synthetic.frame_inputs(H, W, device=dev)
        intr, bg = fi['intrinsics'], fi['bg_color']
        rng = np.random.default_rng(0)
        lm = np.cumsum(rng.standard_normal((args.frames, 68, 3)).astype(np.float32) * 0.05, axis=0)      # a smooth-ish landmark walk
        cond = ingress.regularize_lm3d(lm, np.zeros((1, 68, 3), np.float32), np.ones((1, 68, 3), np.float32), args.clamp_std)
        _, wins = ingress.cond_windows(cond, hp['cond_win_size'], hp['smo_win_size'])
        poses = np.stack([orbit_pose(3.35, 10.0 * np.sin(2 * np.pi * f / 100.0)) for f in range(args.frames)])
    else:
        from geneface_b200.renderer import RADNeRFTorso
        inp = ingress.SequenceInputs.load(args.data, prefix="val", smooth_kernel=args.smooth_kernel)
        H, W, intr = inp.H, inp.W, inp.intrinsics
        hp = synthetic.may_hparams()
        model = RADNeRFTorso(hp).to(dev).eval()
        if args.ckpt is None:
            raise SystemExit("--data needs --ckpt (refusing to render a real sequence with random weights)")
        if rank == 0:
            model.load_state_dict(unwrap_checkpoint(torch.load(args.ckpt, map_location=dev)), strict=True)   # missing / unexpected keys raise
        lm = np.load(args.lm3d)[0]
        poses, wins = inp.sequence(lm[:args.frames], args.clamp_std, hp['cond_win_size'], hp['smo_win_size'])
        bg = torch.from_numpy(background_image(args.bg, inp.bg_img, H, W)).view(1, -1, 3).to(dev)
    sequence.broadcast_model_(model, src=0)                 # the only collective: parameters, once
    T = min(args.frames, wins.shape[0])
    start, end = sequence.partition_frames(T, world, rank)
    poses_t = torch.from_numpy(np.ascontiguousarray(poses, dtype=np.float32))
    conds_t = torch.from_numpy(np.ascontiguousarray(wins)).pin_memory()

    # ---- render + egress ----
    seq = sequence.SequenceRenderer(model, H, W, intr, precision=args.precision, max_steps=hp['max_steps'], dt_gamma=hp['dt_gamma'], torso=True)
    writer = egress.PngSequenceWriter(args.out) if args.out else None
    seq.render(poses_t, conds_t, bg, start, min(start + 2, end))          # warm-up: kernels, allocator, weight image
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    seq.render(poses_t, conds_t, bg, start, end, sink=writer.submit if writer else None)
    t_render = time.perf_counter() - t0
    if writer:
        writer.close()
    t_total = time.perf_counter() - t0
    stats = torch.tensor([t_render, t_total, float(end - start), float(writer.bytes_written if writer else 0)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        t_render, t_total = float(mx[0]), float(mx[1])
    frames, nbytes = int(stats[2]), int(stats[3])
    if rank == 0:
        print(json.dumps({"frames": frames, "n_gpus": world, "size": [H, W], "render_s": t_render, "total_s_incl_png": t_total,
                          "frames_per_s_render": frames / t_render, "frames_per_s_incl_png": frames / t_total, "png_bytes": nbytes,
                          "out": args.out, "precision": args.precision}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
