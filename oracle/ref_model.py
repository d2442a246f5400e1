"""The REAL reference model -- `modules.radnerfs.radnerf_torso.RADNeRFTorso` / `radnerf.RADNeRF` and their own `render()` --
imported, unmodified, from the byte-for-byte mirror `oracle/_ref/pyref/` (written by oracle/build_ref.py where /root/reference
exists; git-ignored; travels to the GPU box) on top of the compiled unmodified reference extensions `oracle/_ref/*.so`.

TEST INFRASTRUCTURE ONLY: frame-level parity oracle (tests/, oracle/gen_golden_frames.py) and the "reference on the same GPU"
context timing of bench.py.  Nothing in geneface_b200/ imports this.

Six third-party packages that `modules/radnerfs/{utils,renderer}.py` import at module level but never touch on the render path
(trimesh, mcubes, lpips, tensorboardX, imageio, matplotlib) are absent in this image and are stubbed in sys.modules (SURVEY.md
section 7 step 0); everything the path executes is the reference's own code.
"""
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_PYREF = os.path.join(_REF, "pyref")
_STUBS = ("trimesh", "mcubes", "lpips", "tensorboardX", "imageio", "matplotlib", "matplotlib.pyplot")
_EXTS = ("_raymarching_face", "_gridencoder", "_shencoder", "_freqencoder")


def available():
    return (os.path.isfile(os.path.join(_PYREF, "modules", "radnerfs", "radnerf_torso.py")) and
            all(os.path.exists(os.path.join(_REF, n + ".so")) for n in _EXTS))


_ns = None


def load():
    """Import the mirrored reference modules; returns a namespace with RADNeRF, RADNeRFTorso, raymarching, hparams, utils."""
    global _ns
    if _ns is not None:
        return _ns
    if not available():
        raise RuntimeError("oracle/_ref (compiled reference extensions + pyref mirror) is not built; run oracle/build_ref.py where /root/reference exists")
    for name in _STUBS:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []          # lets `import matplotlib.pyplot` resolve through sys.modules
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for p in (_REF, _PYREF):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference's batch inference is plain fp32 (SURVEY.md section 8 header)
    torch.backends.cudnn.allow_tf32 = False
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from utils.commons.hparams import hparams as ref_hparams
        import modules.radnerfs.raymarching as raymarching
        from modules.radnerfs.radnerf import RADNeRF
        from modules.radnerfs.radnerf_torso import RADNeRFTorso
        import modules.radnerfs.utils as ref_utils
    _ns = types.SimpleNamespace(RADNeRF=RADNeRF, RADNeRFTorso=RADNeRFTorso, raymarching=raymarching, hparams=ref_hparams, utils=ref_utils)
    return _ns


def build(state_dict, hp, torso=True, device="cuda", mean_density_torso=0.0):
    """Reference model with the reference's constructor, our synthetic weights loaded through its own load_state_dict (strict)."""
    ns = load()
    ns.hparams.clear()
    ns.hparams.update(hp)                       # radnerf_torso.py reads the GLOBAL hparams dict (torso_shrink, torso_head_aware)
    hp = dict(hp)
    hp.setdefault("cuda_ray", True)
    model = (ns.RADNeRFTorso if torso else ns.RADNeRF)(hp)
    missing, unexpected = model.load_state_dict(state_dict, strict=True)
    assert not missing and not unexpected
    if torso:
        model.mean_density_torso = mean_density_torso
    return model.to(device).eval()


class LoopTrace:
    """Observes the reference's own host loop from outside by wrapping the module-level `raymarching.march_rays` /
    `raymarching.composite_rays` it calls (renderer.py:340-345, radnerf_torso.py:146-149).  Nothing is recomputed; recorded are
      trace      (n_alive, n_step) of every iteration,
      term_iter  int32[N]: the iteration in which composite_rays marked the ray dead (rays_alive[i] = -1), -1 = survived the loop,
      weights_sum the in-place accumulator of the last composite call (eval render() does not return it),
      n_marched  per-ray number of non-terminator samples marched."""

    def __init__(self, N, device):
        self.ns = load()
        self.trace = []
        self.n_marched = torch.zeros(N, dtype=torch.int32, device=device)
        self.term_iter = torch.full((N,), -1, dtype=torch.int32, device=device)
        self.weights_sum = None

    def __enter__(self):
        rm = self.ns.raymarching
        self._march, self._comp = rm.march_rays, rm.composite_rays

        def march(n_alive, n_step, rays_alive, rays_t, *a, **k):
            ids = rays_alive[:n_alive].long().clone()
            out = self._march(n_alive, n_step, rays_alive, rays_t, *a, **k)
            self.trace.append((int(n_alive), int(n_step)))
            self.n_marched[ids] += (out[2][:n_alive * n_step, 0].view(n_alive, n_step) != 0).sum(1).int()
            return out

        def comp(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, *a, **k):
            ids = rays_alive[:n_alive].long().clone()
            out = self._comp(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, *a, **k)
            self.term_iter[ids[rays_alive[:n_alive] < 0]] = len(self.trace) - 1
            self.weights_sum = weights_sum
            return out
        rm.march_rays, rm.composite_rays = march, comp
        return self

    def __exit__(self, *exc):
        self.ns.raymarching.march_rays, self.ns.raymarching.composite_rays = self._march, self._comp
        return False


@torch.no_grad()
def render(model, rays_o, rays_d, cond, bg_coords, poses6, bg_color, dt_gamma, max_steps, T_thresh=1e-4, trace=True):
    """`model.render(...)` of the reference exactly as tasks/radnerfs/radnerf(_torso).py:run_model calls it in eval
    (perturb=False, force_all_rays irrelevant in eval).  Returns the reference's result dict, plus the LoopTrace observations
    ('trace', 'term_iter', 'weights_sum', 'n_marched')."""
    N = rays_o.reshape(-1, 3).shape[0]
    ro, rd = rays_o.reshape(1, N, 3), rays_d.reshape(1, N, 3)
    bgc = bg_coords.reshape(1, N, 2) if bg_coords is not None else None
    if trace:
        with LoopTrace(N, rays_o.device) as lt:
            res = model.render(ro, rd, cond, bgc, poses6, index=0, dt_gamma=dt_gamma, bg_color=bg_color, perturb=False,
                               force_all_rays=False, max_steps=max_steps, T_thresh=T_thresh)
        res['trace'], res['n_marched'], res['term_iter'], res['weights_sum'] = lt.trace, lt.n_marched, lt.term_iter, lt.weights_sum
        return res
    return model.render(ro, rd, cond, bgc, poses6, index=0, dt_gamma=dt_gamma, bg_color=bg_color, perturb=False, force_all_rays=False,
                        max_steps=max_steps, T_thresh=T_thresh)
