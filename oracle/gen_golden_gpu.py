"""Generate tests/golden/*.npz by running the UNMODIFIED reference kernels (oracle/_ref/*.so, compiled from
/root/reference by oracle/build_ref.py) on a GPU.  Run on the B200 box:

    gpurun -- python oracle/gen_golden_gpu.py gpurun_out/golden

then copy gpurun_out/golden/*.npz to tests/golden/ and commit them.  Inputs are regenerated from numpy seeds by
tests/scenes.py on both sides, so the fixtures hold only the reference's OUTPUTS (small).  This is how the CPU
oracle (oracle/gf_oracle.c) is pinned: `pytest -m "not gpu"` checks it against these files.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import scenes  # noqa: E402


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    import _raymarching_face as RM
    import _gridencoder as GE
    import _shencoder as SH
    import _freqencoder as FQ

    # ---------------- raymarching: near/far, march (3 bitfields x 2 bounds), composite ----------------
    for tag, bound, C in (("b1", 1.0, 1), ("b4", 4.0, 3)):
        H = 128
        N = 256
        o, d = scenes.camera_rays(N, seed=11)
        if tag == "b4":
            o2, d2 = scenes.inside_rays(N // 2, seed=12, bound=bound)
            o[: N // 2], d[: N // 2] = o2, d2
        aabb = scenes.aabb_of(bound)
        ro, rd = T(o), T(d)
        nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
        RM.near_far_from_aabb(ro, rd, T(aabb), N, 0.05, nears, fars)
        res = dict(nears=nears.cpu().numpy(), fars=fars.cpu().numpy())
        for bf_name, bf in (("R", scenes.random_bitfield(C, H, 0.3, seed=1)), ("F", scenes.full_bitfield(C, H)),
                            ("R05", scenes.random_bitfield(C, H, 0.05, seed=2))):
            for (n_step, dt_gamma, max_steps) in ((4, 1 / 256, 16), (8, 0.0, 128), (3, 1 / 128, 1024)):
                M = N * n_step
                M += 128 - (M % 128)
                xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
                alive = torch.arange(N, dtype=torch.int32, device="cuda")
                rays_t = nears.clone()
                noises = torch.zeros(N, device="cuda")
                RM.march_rays(N, n_step, alive, rays_t, ro, rd, bound, dt_gamma, max_steps, C, H, T(bf), nears, fars, xyzs, dirs, deltas, noises)
                key = f"march_{bf_name}_{n_step}_{max_steps}"
                res[key + "_xyzs"] = xyzs.cpu().numpy(); res[key + "_deltas"] = deltas.cpu().numpy()
                # with perturbation noise (exercises the FFMA at raymarching.cu:873)
                noises = T(np.random.RandomState(5).rand(N).astype(np.float32))
                xyzs.zero_(); dirs.zero_(); deltas.zero_()
                RM.march_rays(N, n_step, alive, rays_t, ro, rd, bound, dt_gamma, max_steps, C, H, T(bf), nears, fars, xyzs, dirs, deltas, noises)
                res[key + "_noise_deltas"] = deltas.cpu().numpy()
                # composite on synthetic sigmas/rgbs
                rs = np.random.RandomState(7)
                sig = T(np.exp(rs.randn(M) * 2.0 + 2.0).astype(np.float32))
                rgb = T(rs.rand(M, 3).astype(np.float32))
                ws = torch.zeros(N, device="cuda"); dep = torch.zeros(N, device="cuda"); img = torch.zeros(N, 3, device="cuda")
                al = alive.clone(); rt = rays_t.clone()
                RM.composite_rays(N, n_step, 1e-4, al, rt, sig, rgb, deltas, ws, dep, img)
                res[key + "_comp_alive"] = al.cpu().numpy(); res[key + "_comp_t"] = rt.cpu().numpy()
                res[key + "_comp_ws"] = ws.cpu().numpy(); res[key + "_comp_depth"] = dep.cpu().numpy(); res[key + "_comp_img"] = img.cpu().numpy()
        # training march + composite (per-ray compare: layout order is the reference's atomics')
        bf = scenes.random_bitfield(C, H, 0.3, seed=1)
        for max_steps, dt_gamma in ((16, 1 / 256), (64, 0.0)):
            M = N * max_steps
            xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); deltas = torch.zeros(M, 2, device="cuda")
            rays = torch.empty(N, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
            noises = T(np.random.RandomState(6).rand(N).astype(np.float32))
            RM.march_rays_train(ro, rd, T(bf), bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
            r = rays.cpu().numpy(); x = xyzs.cpu().numpy(); dl = deltas.cpu().numpy()
            order = np.argsort(r[:, 0])
            r = r[order]
            # re-pack per ray in ray order so the fixture is layout independent
            counts = r[:, 2]
            offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
            px = np.zeros((counts.sum(), 3), np.float32); pd = np.zeros((counts.sum(), 2), np.float32)
            for i in range(N):
                px[offs[i]:offs[i] + counts[i]] = x[r[i, 1]:r[i, 1] + counts[i]]
                pd[offs[i]:offs[i] + counts[i]] = dl[r[i, 1]:r[i, 1] + counts[i]]
            key = f"train_{max_steps}"
            res[key + "_counts"] = counts.astype(np.int32); res[key + "_xyzs"] = px; res[key + "_deltas"] = pd
            res[key + "_counter"] = counter.cpu().numpy()
            # composite train fwd/bwd on the reference's own layout
            Mtot = int(counter[0].item())
            rs = np.random.RandomState(8)
            sig = T(np.exp(rs.randn(M) * 1.5 + 1.0).astype(np.float32)); rgb = T(rs.rand(M, 3).astype(np.float32)); amb = T(rs.rand(M).astype(np.float32))
            ws = torch.empty(N, device="cuda"); ambs = torch.empty(N, device="cuda"); dep = torch.empty(N, device="cuda"); img = torch.empty(N, 3, device="cuda")
            RM.composite_rays_train_forward(sig, rgb, amb, deltas, rays, M, N, 1e-4, ws, ambs, dep, img)
            gws = T(rs.randn(N).astype(np.float32)); gamb = T(rs.randn(N).astype(np.float32)); gimg = T(rs.randn(N, 3).astype(np.float32))
            gs = torch.zeros(M, device="cuda"); gr = torch.zeros(M, 3, device="cuda"); ga = torch.zeros(M, device="cuda")
            RM.composite_rays_train_backward(gws, gamb, gimg, sig, rgb, amb, deltas, rays, ws, ambs, img, M, N, 1e-4, gs, gr, ga)
            res[key + "_rays"] = rays.cpu().numpy()
            res[key + "_layout_deltas"] = deltas.cpu().numpy()[:Mtot]
            for nm, t in (("ws", ws), ("ambs", ambs), ("dep", dep), ("img", img), ("gs", gs[:Mtot]), ("gr", gr[:Mtot]), ("ga", ga[:Mtot])):
                res[key + "_ct_" + nm] = t.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, f"raymarch_{tag}.npz"), **res)

    # ---------------- utils: morton, packbits, dilation ----------------
    rs = np.random.RandomState(3)
    coords = rs.randint(0, 128, size=(1000, 3)).astype(np.int32)
    ind = torch.empty(1000, dtype=torch.int32, device="cuda")
    RM.morton3D(T(coords), 1000, ind)
    inv = torch.empty(1000, 3, dtype=torch.int32, device="cuda")
    RM.morton3D_invert(ind, 1000, inv)
    grid = rs.rand(2, 16 ** 3).astype(np.float32)
    bits = torch.empty(2 * 16 ** 3 // 8, dtype=torch.uint8, device="cuda")
    RM.packbits(T(grid), 2 * 16 ** 3 // 8, 0.5, bits)
    dil = torch.empty(2, 16 ** 3, device="cuda")
    RM.morton3D_dilation(T(grid), 2, 16, dil)
    o, d = scenes.inside_rays(256, seed=4, bound=0.3)
    sph = torch.empty(256, 2, device="cuda")
    RM.sph_from_ray(T(o), T(d), 1.5, 256, sph)
    np.savez_compressed(os.path.join(out_dir, "utils.npz"), morton=ind.cpu().numpy(), morton_inv=inv.cpu().numpy(),
                        packbits=bits.cpu().numpy(), dilation=dil.cpu().numpy(), sph=sph.cpu().numpy())

    # ---------------- grid encoder ----------------
    res = {}
    for D in (2, 3):
        for gridtype in (0, 1):
            for interp in (0, 1):
                offsets, S, emb = scenes.grid_setup(D, seed=20 + D)
                B = 256
                x = scenes.unit_points(B, D, seed=30 + D)
                L, C = 16, 2
                out = torch.empty(L, B, C, device="cuda"); dy = torch.empty(B, L * D * C, device="cuda")
                GE.grid_encode_forward(T(x), T(emb), T(offsets), out, B, D, C, L, S, 16, dy, gridtype, False, interp)
                key = f"D{D}_g{gridtype}_i{interp}"
                res[key + "_out"] = out.cpu().numpy(); res[key + "_dydx"] = dy.cpu().numpy()
                grad = np.random.RandomState(40).randn(L, B, C).astype(np.float32)
                gg = torch.zeros_like(T(emb)); gi = torch.zeros(B, D, device="cuda")
                GE.grid_encode_backward(T(grad), T(x), T(emb), T(offsets), gg, B, D, C, L, S, 16, dy, gi, gridtype, False, interp)
                nz = gg.abs().sum(1).nonzero().view(-1)
                res[key + "_gg_idx"] = nz.cpu().numpy().astype(np.int32); res[key + "_gg_val"] = gg[nz].cpu().numpy(); res[key + "_gi"] = gi.cpu().numpy()
    # bound=4 geometry (desired_resolution 8192) and the C=4 / C=8 vector paths
    offsets, S, emb = scenes.grid_setup(3, desired=8192, seed=50)
    x = scenes.unit_points(256, 3, seed=51)
    out = torch.empty(16, 256, 2, device="cuda")
    GE.grid_encode_forward(T(x), T(emb), T(offsets), out, 256, 3, 2, 16, S, 16, None, 1, False, 0)
    res["D3_res8192_out"] = out.cpu().numpy()
    for C in (1, 4, 8):
        offsets, S, emb = scenes.grid_setup(3, L=4, C=C, log2_hash=12, desired=128, seed=60 + C)
        x = scenes.unit_points(64, 3, seed=61)
        out = torch.empty(4, 64, C, device="cuda")
        GE.grid_encode_forward(T(x), T(emb), T(offsets), out, 64, 3, C, 4, S, 16, None, 0, False, 0)
        res[f"D3_C{C}_out"] = out.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "gridencoder.npz"), **res)

    # ---------------- SH + freq ----------------
    res = {}
    _, d = scenes.field_samples(128, seed=70)
    d[0] = [0, 0, 1]; d[1] = [1, 0, 0]; d[2] = [0, -1, 0]; d[3] *= 0.5
    for deg in (1, 2, 3, 4, 5, 6, 7, 8):
        out = torch.empty(128, deg * deg, device="cuda"); dy = torch.empty(128, 3 * deg * deg, device="cuda")
        SH.sh_encode_forward(T(d), out, 128, 3, deg, dy)
        res[f"sh{deg}_out"] = out.cpu().numpy(); res[f"sh{deg}_dydx"] = dy.cpu().numpy()
    rs = np.random.RandomState(80)
    x6 = (rs.randn(64, 6) * 1.5).astype(np.float32); x2 = (rs.rand(256, 2) * 2 - 1).astype(np.float32)
    o6 = torch.empty(64, 54, device="cuda"); o2 = torch.empty(256, 42, device="cuda")
    FQ.freq_encode_forward(T(x6), 64, 6, 4, 54, o6)
    FQ.freq_encode_forward(T(x2), 256, 2, 10, 42, o2)
    g2 = rs.randn(256, 42).astype(np.float32)
    gi = torch.zeros(256, 2, device="cuda")
    FQ.freq_encode_backward(T(g2), o2, 256, 2, 10, 42, gi)
    res.update(freq6=o6.cpu().numpy(), freq2=o2.cpu().numpy(), freq2_gi=gi.cpu().numpy())
    np.savez_compressed(os.path.join(out_dir, "encoders.npz"), **res)
    torch.cuda.synchronize()
    print("golden fixtures written to", out_dir, {f: os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
