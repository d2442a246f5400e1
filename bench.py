#!/usr/bin/env python
"""bench.py -- frames/sec of the volumetric renderer hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is ONE 512x512 frame of the north-star synthetic workload (SURVEY.md section 8d "512x512x128"): head+torso model of the
May configs with bound=4 (3 cascades), all-ones occupancy bitfield, dt_gamma=0, max_steps=128 -> every one of the 262,144
rays composites exactly 128 samples (33,554,432 field evaluations per frame; asserted), plus the 2-D torso field on the lower
image half.  Random-init weights of the exact architecture, synthetic pose/cond (no datasets or checkpoints here).

value  : frames/s, whole job (all ranks), inputs (pose, cond feature window, background) resident in HBM, device-timed with
         CUDA events over exactly K steps between barrier + synchronize; max over ranks.
e2e    : the same metric through the public API with HOST inputs: per step pinned pose+cond -> H2D, cond encoder, fused
         frame, RGB8 frame -> D2H into pinned memory, all inside the timed region.
roofline: dominant kernel = the field kernel (grid gathers + MLPs); achieved = samples/frame x 1536 B (SURVEY.md 8d algorithmic
         gather bytes per head sample) / its CUDA-event time per frame (events recorded inside gf_render_frame on the launch
         stream); peak = MEASURED_PEAKS.json hbm_gbs.  The tables (16 MB) are L2-resident, so this is an "HBM-equivalent"
         gather rate as the metric asks; tensor-pipe fraction is reported beside it.
cpu_baseline: the CPU oracle (port of the reference path, OpenMP + numpy, all host threads) on a bounded sample of the SAME
         frame (rays x 128 samples), extrapolated to frames/s.
--impl reference: the reference's CPU implementation timed on the host cores (rank 0 only): the oracle port of the RAD-NeRF
         path on a bounded sample per step; the vanilla AD-NeRF (modules/nerfs) port is reported in `adnerf_cpu`.
"""
import os as _os
# The CPU legs alternate OpenMP regions (C oracle) and OpenBLAS GEMMs (numpy) ~10^3 times per frame sample.  With libgomp's default
# ACTIVE wait policy its idle threads spin between regions and starve the BLAS threads (measured: 6 ms instead of 29 us per GEMM).
# Must be set before libgomp is loaded, i.e. before torch/numpy are imported.
_os.environ.setdefault("OMP_WAIT_POLICY", "passive")
_os.environ.setdefault("GOMP_SPINCOUNT", "0")
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "512x512 frames/sec (head+torso), 128 samples/ray"
HEAD_SAMPLE_BYTES = 1536          # 16 lvl x 8 corners x 8 B + 16 lvl x 4 corners x 8 B
HEAD_SAMPLE_FLOP = 178688
TORSO_PIXEL_BYTES = 512
H = W = 512
MAX_STEPS = 128


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi sampler (B200_PROFILING.md clocks line).  Started before the warm-up (nvidia-smi takes ~0.5 s to come up);
    only samples whose timestamp falls inside [mark_begin, mark_end] -- the timed region -- are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines, self.t0, self.t1 = gpu_index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        import datetime
        sm, mx, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 - 0.03 <= ts <= self.t1 + 0.03):
                    continue
                sm.append(float(f[2])); mx = float(f[3]); power.append(float(f[4]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------ CPU legs
def host_threads():
    """Threads for the CPU legs.  The port's host loop issues ~128 small numpy/OpenMP regions per frame sample; beyond a few dozen
    threads their fork/join cost dominates (measured on the 128-core GPU host: 2048 rays took 280 s with 128 threads), so the legs
    use min(cores, 32) and report that number."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("GF_CPU_THREADS", "32"))))


def _cpu_port_once(n_rays, threads):
    import numpy as np
    import torch
    from threadpoolctl import threadpool_limits
    torch.set_num_threads(threads)
    from geneface_b200 import synthetic
    from oracle import field as OF
    model, hp = synthetic.build_model(torso=True, bitfield='F', seed=0, sigma_scale=0.25, bound=4, device='cpu')
    sd = synthetic.state_to_numpy(model)
    fi = synthetic.frame_inputs(H, W, device='cpu')
    ro, rd = OF.get_rays(fi['pose'][0].numpy(), fi['intrinsics'], H, W)
    sel = np.linspace(0, H * W - 1, n_rays).astype(np.int64)
    ro, rd = np.ascontiguousarray(ro[sel]), np.ascontiguousarray(rd[sel])
    fo = OF.FieldOracle(sd, bound=4.0)
    OF.MATMUL_DTYPE = np.float32          # timing leg: fp32 GEMMs like the reference's CPU tensors
    with threadpool_limits(limits=threads):
        t0 = time.perf_counter()
        cf = OF.cal_cond_feat(sd, fi['cond'].numpy())
        ws, depth, img, nears, fars, ns = OF.render_head(fo, sd, ro, rd, cf, sd['density_bitfield'], 3, 128, sd['aabb_infer'], hp['min_near'], 0.0, MAX_STEPS)
        bgc = OF.get_bg_coords(H, W)[sel]
        bg, _, _, _ = OF.render_torso_mix(OF.TorsoOracle(sd), sd, bgc, fi['poses6'].numpy(), fi['bg_color'][0].numpy()[sel], img, ws)
        OF.finish(img, ws, depth, nears, fars, bg)
        dt = time.perf_counter() - t0
    OF.MATMUL_DTYPE = np.float64
    assert int(ns.min()) == MAX_STEPS == int(ns.max())
    return dt


def cpu_port_fps(n_rays=None, threads=None, budget_s=12.0):
    """The CPU oracle (port of the reference RAD-NeRF path) on a BOUNDED sample of the benchmark frame: `n_rays` rays x 128
    samples (head + torso + mix).  n_rays=None: a 256-ray probe sizes the sample for about `budget_s` seconds of CPU work."""
    threads = threads or host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    if n_rays is None:
        _cpu_port_once(256, threads)                      # cold: library load, page faults
        probe = _cpu_port_once(256, threads)
        n_rays = int(min(16384, max(256, 256 * budget_s / max(probe, 1e-3))))
        n_rays = 1 << (n_rays.bit_length() - 1)          # power of two
    dt = _cpu_port_once(n_rays, threads)
    fps = 1.0 / (dt * (H * W) / n_rays)
    return fps, dt, threads, f"{n_rays} of 262144 rays x 128 samples of the same frame (head+torso), {dt:.1f} s of CPU work on {threads} threads, extrapolated"


def adnerf_cpu_fps(threads=None):
    """Vanilla AD-NeRF port (BASELINE.json configs[0]: 64x64, 64 coarse + 128 fine samples, 1 frame)."""
    try:
        from oracle import adnerf_port
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)}
    return adnerf_port.time_frame(threads or host_threads())


def run_reference_arm(args, rank):
    """--impl reference: the reference path on the host CPU (rank 0 only)."""
    if rank != 0:
        return
    steps, warm = max(1, args.steps), args.warmup
    t0 = time.perf_counter()
    fps, dt, threads, sample = cpu_port_fps(None, budget_s=10.0)      # warm-up step; also sizes the bounded sample
    n_rays = int(sample.split()[0])
    fps_list = []
    for _ in range(steps):
        fps, dt, threads, sample = cpu_port_fps(n_rays)
        fps_list.append(fps)
        if time.perf_counter() - t0 > 120:
            break
    k = len(fps_list)
    value = k / sum(1.0 / f for f in fps_list)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": k, "warmup": 1,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RAD-NeRF head+torso 512x512, 128 samples/ray (bound=4, all-ones bitfield); CPU port of the reference path",
                       "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "adnerf_cpu": adnerf_cpu_fps(), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GF_BENCH_PRECISION", "fp16"), choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-may", action="store_true", help="skip the May-configuration context measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from geneface_b200 import _lib, sequence, synthetic
    from geneface_b200.utils import convert_poses, orbit_pose

    assert torch.cuda.is_available(), "bench.py (ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    model, hp = synthetic.build_model(torso=True, bitfield='F', seed=0 if rank == 0 else 100 + rank, sigma_scale=0.25, bound=4, device=dev)
    bcast_bytes = sequence.broadcast_model_(model, src=0)      # parameters reach every rank through ONE NCCL broadcast
    fi = synthetic.frame_inputs(H, W, device=dev)
    N = H * W
    n_frames = args.steps + args.warmup
    first = rank * n_frames                                    # weak scaling: every rank renders its own K frames of the sequence
    poses = torch.stack([torch.from_numpy(orbit_pose(3.35, 10.0 * np.sin(2 * np.pi * (first + f) / 100.0))) for f in range(n_frames)])
    g = torch.Generator().manual_seed(1234)
    conds_all = torch.randn(300 * world + 8, 1, 204, generator=g)
    from geneface_b200.utils import get_audio_features
    conds = torch.stack([get_audio_features(conds_all, 2, first + f, 5) for f in range(n_frames)])       # [F,5,1,204]
    poses_h, conds_h = poses.pin_memory(), conds.pin_memory()
    poses_d, conds_d = poses.to(dev), conds.to(dev)
    pose6_h = convert_poses(poses)           # host: the 6-vector travels by value in the GfFrame struct
    bg = fi['bg_color']
    handle = model.gf_model()
    L = _lib.lib()
    rgb8 = torch.empty(N, 3, dtype=torch.uint8, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    outbuf = {'rgb8': rgb8, 'counters': counters}

    def frame_resident(f):
        with torch.no_grad():
            cf = model.cal_cond_feat(conds_d[f])
            model.render_fused(cf, H, W, pose=poses[f], intrinsics=fi['intrinsics'], bg_color=bg, torso_pose=pose6_h[f], dt_gamma=0.0,
                               max_steps=MAX_STEPS, precision=args.precision, want=('rgb8', 'counters'), out=outbuf)

    # end to end = the public sequence API (geneface_b200.sequence.SequenceRenderer, the replacement of the frame loop of
    # inference/nerfs/base_nerf_infer.py:131-179): per frame the condition window goes host->device from pinned memory, pose and
    # pose6 travel by value in the launch, and the finished RGB8 frame goes device->host into a pinned ring; frames are pipelined
    # (frame k+1 renders while frame k drains) and the call returns when every frame is resident in host memory.
    seq = sequence.SequenceRenderer(model, H, W, fi['intrinsics'], precision=args.precision, max_steps=MAX_STEPS, dt_gamma=0.0, torso=True)
    host_ring = torch.empty(args.steps, H, W, 3, dtype=torch.uint8).pin_memory()

    def sequence_e2e():
        seq.render(poses_h, conds_h, bg, args.warmup, args.warmup + args.steps, out_rgb8=host_ring)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + workload assertions ----
    sampler = ClockSampler(local)
    sampler.start()
    for f in range(args.warmup):
        frame_resident(f)
    torch.cuda.synchronize()
    c = counters.cpu().numpy()
    samples_per_frame, torso_px, s_total, launches = int(c[0]), int(c[1]), int(c[2]), int(c[3])
    assert samples_per_frame == N * MAX_STEPS == 33554432, f"workload must evaluate 262144 x 128 samples, got {samples_per_frame}"
    assert s_total == MAX_STEPS and torso_px > 0

    # ---- timed: resident inputs ----
    barrier()
    sampler.mark_begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        frame_resident(args.warmup + k)
    ev1.record()
    barrier()
    ms_res = ev0.elapsed_time(ev1)
    # ---- timed: end to end (host inputs, host result) ----
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sequence_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    sampler.mark_end()
    clocks = sampler.stop()
    # ---- dominant-kernel time (events inside gf_render_frame) ----
    _lib.check(L.gf_profile_enable(handle, 1))
    field_ms = []
    for k in range(min(5, args.steps)):
        frame_resident(args.warmup + k)
        torch.cuda.synchronize()
        import ctypes
        tot, n = ctypes.c_float(0), ctypes.c_int(0)
        _lib.check(L.gf_profile_field_ms(handle, ctypes.byref(tot), ctypes.byref(n)))
        field_ms.append(tot.value)
    _lib.check(L.gf_profile_enable(handle, 0))
    field_ms_per_frame = float(np.median(field_ms))

    t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    value = world * args.steps / (ms_res / 1000.0)
    e2e = world * args.steps / (ms_e2e / 1000.0)
    achieved = samples_per_frame * HEAD_SAMPLE_BYTES / (field_ms_per_frame / 1000.0) / 1e9
    tflops = samples_per_frame * HEAD_SAMPLE_FLOP / (field_ms_per_frame / 1000.0) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": "RAD-NeRF head+torso (May cfg architecture), 512x512 rays x 128 samples = 33,554,432 field evaluations/frame "
                               "(bound=4, 3 cascades, all-ones bitfield, dt_gamma=0, max_steps=128) + torso field on the lower image half",
                   "frames_per_rank": args.steps, "sharding": "frames, rank-block; one NCCL parameter broadcast (%d B), no per-frame communication" % bcast_bytes,
                   "l2": "per-round sample lists (335 MB) exceed L2; the 16 MB grid tables are the algorithm's own hot set; no explicit flush",
                   "precision": args.precision, "samples_per_frame": samples_per_frame, "torso_pixels": torso_px, "s_total": s_total},
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(conds_h[0].numel() * 4 + 16 * 4 + 6 * 4),
                "d2h_bytes_per_step": int(host_ring[0].numel()), "api": "geneface_b200.sequence.SequenceRenderer.render (pipelined frames)"},
        "gpu_launches": launches * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": 1551318528 if args.precision == "fp16" else None,   # dram read+write bytes per 8.4 M-sample round (k_tc_amb + k_tc_sigcol), ncu capture in profiles/r01_summary.md
                     "algorithmic_bytes_per_launch": samples_per_frame // 4 * HEAD_SAMPLE_BYTES,
                     "kernel": "k_tc_amb+k_tc_sigcol" if args.precision == "fp16" else "k_field_fp32", "kernel_ms_per_frame": field_ms_per_frame,
                     "kernel_share_of_step": field_ms_per_frame / (ms_res / args.steps), "peak_source": peaks["source"],
                     "tensor_tflops": tflops, "tensor_frac_of_bf16_peak": tflops / peaks["bf16_tflops"],
                     "note": "algorithmic gather bytes (1536 B/sample: 16 levels x 8 corners x 8 B + 16 x 4 x 8 B); the 16 MB tables are L2-resident so DRAM traffic "
                             "(sample lists + the 72 B/sample hand-off between the two field kernels) is far below this; one launch = one 8.4 M-sample round"},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            fps, dt, threads, sample = cpu_port_fps(None, budget_s=12.0)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample}
        except Exception as e:  # noqa: BLE001  (the GPU line must still be printed)
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": host_threads(), "kind": "port", "sample": "failed: " + repr(e)[:200]}
    if not args.no_ref_cuda and world == 1:
        try:
            line["reference_cuda"] = reference_cuda_fps(model, hp, fi, dev)
        except Exception as e:  # noqa: BLE001
            line["reference_cuda"] = {"unavailable": repr(e)[:200]}
    if not args.no_may and world == 1:
        try:
            line["may_cfg"] = may_cfg_fps(dev, args.precision, not args.no_ref_cuda)
        except Exception as e:  # noqa: BLE001
            line["may_cfg"] = {"unavailable": repr(e)[:200]}
        try:
            line["adnerf_gpu"] = adnerf_gpu_fps(dev)
        except Exception as e:  # noqa: BLE001
            line["adnerf_gpu"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def adnerf_gpu_fps(dev):
    """Context number (BASELINE.json configs[0] on the GPU): the vanilla AD-NeRF frame -- 64x64 rays, 64 coarse + 128 fine samples,
    8x256 backbone -- through geneface_b200.adnerf (C-ABI operators + library GEMMs); `adnerf_cpu` in the reference arm is its CPU twin."""
    import torch
    from geneface_b200 import adnerf
    torch.manual_seed(0)
    m = adnerf.ADNeRF(dict(cond_dim=64, hidden_size=256)).to(dev).eval()
    Hh = Ww = 64
    focal = 1200.0 * Hh / 450.0
    c2w = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0.6]], device=dev)
    cond = torch.randn(8, 16, 29, generator=torch.Generator().manual_seed(1)).to(dev)
    bc = torch.ones(Hh, Ww, 3, device=dev)

    def frame():
        with torch.no_grad():
            cf = m.cal_cond_feat(cond, with_att=True)
            return adnerf.render_dynamic_face(Hh, Ww, focal, Ww / 2, Hh / 2, chunk=4096, c2w=c2w, cond=cf, near=0.3, far=0.9, network_fn=m,
                                              N_samples=64, N_importance=128, perturb=0., bc_rgb=bc)
    for _ in range(3):
        frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        out = frame()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_frame": ms, "finite": bool(torch.isfinite(out[0]).all()),
            "workload": "vanilla AD-NeRF 64x64, 64 coarse + 128 fine samples/ray, fp32 GEMMs, random weights (BASELINE.json configs[0])"}


def may_cfg_fps(dev, precision, with_ref):
    """Context number (BASELINE.json configs[2]): the reference's own deployment configuration -- May head+torso, bound=1,
    max_steps=16, dt_gamma=1/256, sphere occupancy -- rendered by the same fused path, beside the compiled reference loop."""
    import torch
    from geneface_b200 import synthetic
    model, hp = synthetic.build_model(torso=True, bitfield='S', seed=0, device=dev)
    fi = synthetic.frame_inputs(H, W, device=dev)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    out = {'rgb8': torch.empty(H * W, 3, dtype=torch.uint8, device=dev), 'counters': counters}
    pose6 = fi['poses6']

    def frame():
        with torch.no_grad():
            cf = model.cal_cond_feat(fi['cond'])
            model.render_fused(cf, H, W, pose=fi['pose'][0], intrinsics=fi['intrinsics'], bg_color=fi['bg_color'], torso_pose=pose6,
                               dt_gamma=hp['dt_gamma'], max_steps=hp['max_steps'], precision=precision, want=('rgb8', 'counters'), out=out)
    for _ in range(5):
        frame()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for _ in range(n):
        frame()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    c = counters.cpu().numpy()
    res = {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_frame": ms, "samples_per_frame": int(c[0]), "torso_pixels": int(c[1]),
           "workload": "May cfg head+torso 512x512: bound=1, max_steps=%d, dt_gamma=1/256, sphere bitfield" % hp['max_steps']}
    if with_ref:
        try:
            res["reference_cuda"] = reference_cuda_fps(model, hp, fi, dev, dt_gamma=hp['dt_gamma'], max_steps=hp['max_steps'])
        except Exception as e:  # noqa: BLE001
            res["reference_cuda"] = {"unavailable": repr(e)[:200]}
    return res


def reference_cuda_fps(model, hp, fi, dev, dt_gamma=0.0, max_steps=None):
    """Context number: the reference renderer assembled from the compiled UNMODIFIED reference kernels (oracle/_ref) on the
    same GPU and workload (the 'kernel to beat', BASELINE.md B-REF-CUDA).  Not part of the product path."""
    import torch
    from oracle import ref_gpu
    if not ref_gpu.available():
        return {"unavailable": "oracle/_ref not built"}
    from geneface_b200 import utils
    ref = ref_gpu.RefRenderer(model.state_dict(), hp, torso=True)
    rays = utils.get_rays(fi['pose'], fi['intrinsics'], H, W)
    bgc = utils.get_bg_coords(H, W, dev)
    times = []
    with torch.no_grad():
        cf = model.cal_cond_feat(fi['cond'])
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ws, depth, img, nears, fars, _ = ref.render_head(rays['rays_o'][0], rays['rays_d'][0], cf, dt_gamma, max_steps or MAX_STEPS)
            bg, _, _ = ref.torso_bg(bgc[0], fi['poses6'], fi['bg_color'][0])
            ref.finish(img, ws, depth, nears, fars, bg)
            e1.record()
            torch.cuda.synchronize()
            if it:
                times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    return {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_frame": ms, "what": "reference host loop on its own compiled kernels + torch fp32 GEMMs, same frame"}


if __name__ == "__main__":
    main()
