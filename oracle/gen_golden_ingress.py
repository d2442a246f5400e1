"""Generate tests/golden/ingress.npz by RUNNING the reference's own ingress code from /root/reference in the build container:
  data_gen/nerf/binarizer.py:get_win_conds, tasks/radnerfs/dataset_utils.py:smooth_camera_path,
  inference/nerfs/lm3d_radnerf_infer.py:LM3d_RADNeRFInfer.get_cond_from_input (normalise / clamp / smooth / window),
  modules/radnerfs/utils.py:nerf_matrix_to_ngp.
Absent third-party modules the reference imports at module scope (imageio, trimesh, ...) are stubbed; none is used by these
functions.  Run once here:  python oracle/gen_golden_ingress.py     (the GPU box has no /root/reference).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")


def _stub_missing():
    for m in ("imageio", "trimesh", "mcubes", "lpips", "tensorboardX", "matplotlib", "matplotlib.pyplot", "face_alignment", "librosa",
              "python_speech_features", "resampy", "pyloudnorm", "webrtcvad", "skimage", "skimage.transform", "mediapipe"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:  # noqa: BLE001
                sys.modules[m] = types.ModuleType(m)


def ref_function(path, name, ns, class_name=None):
    """Compile ONE function straight from a reference source file (modules whose import needs absent assets -- binarizer.py builds a
    Face3DHelper from BFM files at import -- cannot be imported, but the functions themselves only need numpy/torch)."""
    import ast
    tree = ast.parse(open(path).read())
    scope = tree.body if class_name is None else next(c for c in tree.body if isinstance(c, ast.ClassDef) and c.name == class_name).body
    fn = next(n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def main():
    _stub_missing()
    get_win_conds = ref_function("/root/reference/data_gen/nerf/binarizer.py", "get_win_conds", {"np": np})
    fake_binarizer = types.ModuleType("data_gen.nerf.binarizer")
    fake_binarizer.get_win_conds = get_win_conds
    sys.modules["data_gen.nerf.binarizer"] = fake_binarizer           # resolved by the import inside get_cond_from_input
    from modules.radnerfs.utils import nerf_matrix_to_ngp
    from tasks.radnerfs.dataset_utils import smooth_camera_path
    from utils.commons.hparams import hparams
    rng = np.random.default_rng(0)
    out = {}
    conds = rng.standard_normal((11, 2, 5)).astype(np.float32)
    out["win_in"] = conds
    for pad in ("zero", "edge"):
        for win in (5, 8, 1):
            out[f"win_{pad}_{win}"] = np.stack([get_win_conds(conds, i, smo_win_size=win, pad_option=pad) for i in range(-1, 12)])
    # camera path
    from scipy.spatial.transform import Rotation
    N = 12
    poses = np.tile(np.eye(4, dtype=np.float64), (N, 1, 1))
    poses[:, :3, :3] = Rotation.from_rotvec(rng.standard_normal((N, 3)) * 0.2).as_matrix()
    poses[:, :3, 3] = rng.standard_normal((N, 3))
    out["poses_in"] = poses.copy()
    out["poses_smooth7"] = smooth_camera_path(poses.copy(), kernel_size=7)
    out["poses_smooth3"] = smooth_camera_path(poses.copy(), kernel_size=3)
    out["ngp"] = np.stack([nerf_matrix_to_ngp(p, scale=4, offset=[0.1, -0.2, 0.3]) for p in poses])
    # landmark conditioning through the reference's inference class method (no model needed for this method)
    T = 9
    lm = (rng.standard_normal((1, T, 204)) * 3).astype(np.float32)
    mean = rng.standard_normal((1, 68, 3)).astype(np.float32) * 0.1
    std = (rng.random((1, 68, 3)).astype(np.float32) + 0.5)
    hparams.update(dict(infer_lm3d_clamp_std=2.5, cond_win_size=1, smo_win_size=5, use_window_cond=True))
    get_cond = ref_function("/root/reference/inference/nerfs/lm3d_radnerf_infer.py", "get_cond_from_input",
                            {"np": np, "torch": torch, "hparams": hparams}, class_name="LM3d_RADNeRFInfer")
    LM3d_RADNeRFInfer = types.SimpleNamespace(get_cond_from_input=get_cond)
    fake = types.SimpleNamespace(dataset=types.SimpleNamespace(idexp_lm3d_mean=torch.from_numpy(mean), idexp_lm3d_std=torch.from_numpy(std)),
                                 save_wav16k=lambda inp: None)
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "lm.npy")
        np.save(fn, lm)
        samples = LM3d_RADNeRFInfer.get_cond_from_input(fake, {"cond_name": fn, "audio_source_name": "x.wav"})
    out["lm_in"], out["lm_mean"], out["lm_std"] = lm[0], mean, std
    out["lm_cond"] = torch.stack([s["cond"][0] for s in samples]).numpy()
    out["lm_cond_win"] = torch.stack([s["cond_win"] for s in samples]).numpy()
    out["lm_cond_wins"] = torch.stack([s["cond_wins"] for s in samples]).numpy()
    path = os.path.join(ROOT, "tests", "golden", "ingress.npz")
    np.savez_compressed(path, **out)
    print("written", path, {k: v.shape for k, v in out.items() if k.startswith("lm_")})


if __name__ == "__main__":
    main()
