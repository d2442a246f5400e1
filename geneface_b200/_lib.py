"""Build + load libgfrender.so (the C-ABI sm_100a library) and bind its entry points.

The product path has NO fallback: if the shared library is missing or a symbol is absent the
import fails loudly (RuntimeError), and every op raises on a non-zero return code with
gf_last_error() as the message (mirrors TORCH_CHECK -> RuntimeError in the reference,
gridencoder.cu:448-464).
"""
import ctypes
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
# GF_LIBGFRENDER selects another build of the same library (experiment variants built by `build(out=..., extra=...)`,
# e.g. scripts/build_variants.py); the default is the in-tree geneface_b200/libgfrender.so.
_SO = os.environ.get("GF_LIBGFRENDER") or os.path.join(_PKG, "libgfrender.so")
_VARIANT = bool(os.environ.get("GF_LIBGFRENDER"))
_INCLUDE = os.path.join(os.path.dirname(_PKG), "include")

SOURCES = ["api.cu", "raymarch_ops.cu", "encoders.cu", "render_fused.cu", "field_tc_split.cu", "adnerf_ops.cu", "adnerf_mlp_tc.cu", "train_linear_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if _VARIANT:
        if not os.path.exists(_SO):
            raise RuntimeError("GF_LIBGFRENDER=%s does not exist" % _SO)
        return False
    if not os.path.exists(_SO):
        return True
    t = os.path.getmtime(_SO)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(_INCLUDE, "gfrender.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, extra=None):
    """nvcc-compile every .cu under csrc/ for sm_100a into geneface_b200/libgfrender.so (in-tree).
    out / extra: build an experiment variant (extra nvcc flags such as -DGF_...=1) into another .so with its own object directory."""
    so = out or os.path.join(_PKG, "libgfrender.so")
    if out is None and not force and not _stale():
        return so
    nvcc = _nvcc()
    objs = []
    procs = []
    bdir = os.path.join(_PKG, "build") if out is None else os.path.splitext(out)[0] + "_obj"
    os.makedirs(bdir, exist_ok=True)
    force = force or out is not None
    extra = list(extra or []) + os.environ.get("GF_NVCC_EXTRA", "").split()        # GF_NVCC_EXTRA: experiment switches (-D...)
    for src in SOURCES:
        path = os.path.join(_CSRC, src)
        if not os.path.exists(path):
            raise RuntimeError("missing CUDA source %s" % path)
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(
                os.path.getmtime(path), os.path.getmtime(os.path.join(_CSRC, "gf_common.cuh")),
                os.path.getmtime(os.path.join(_INCLUDE, "gfrender.h")),
                *[os.path.getmtime(os.path.join(_CSRC, h)) for h in os.listdir(_CSRC) if h.endswith(".cuh")]):
            continue
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-I", _INCLUDE, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
    cmd = [nvcc, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return so


_lib = None

c_u32, c_f32, c_int, c_vp, c_u64 = ctypes.c_uint32, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64

# name -> argtypes (all return int unless listed in _RESTYPE)
_SIGS = {
    "gf_near_far_from_aabb": [c_vp, c_vp, c_vp, c_u32, c_f32, c_vp, c_vp, c_vp],
    "gf_sph_from_ray": [c_vp, c_vp, c_f32, c_u32, c_vp, c_vp],
    "gf_morton3D": [c_vp, c_u32, c_vp, c_vp],
    "gf_morton3D_invert": [c_vp, c_u32, c_vp, c_vp],
    "gf_packbits": [c_vp, c_u32, c_f32, c_vp, c_vp],
    "gf_morton3D_dilation": [c_vp, c_u32, c_u32, c_vp, c_vp],
    "gf_march_rays_train": [c_vp, c_vp, c_vp, c_f32, c_f32, c_u32, c_u32, c_u32, c_u32, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp,
                            c_vp, c_vp, c_vp, c_vp],
    "gf_march_rays_train_backward": [c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_vp, c_vp, c_vp],
    "gf_composite_rays_train_forward": [c_vp, c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gf_composite_rays_train_backward": [c_vp] * 11 + [c_u32, c_u32, c_f32, c_vp, c_vp, c_vp, c_vp],
    "gf_march_rays": [c_u32, c_u32, c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, c_u32, c_u32, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp,
                      c_vp, c_vp, c_vp],
    "gf_composite_rays": [c_u32, c_u32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gf_grid_encode_forward": [c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_u32, c_u32, c_f32, c_u32, c_vp, c_u32, c_int, c_u32,
                               c_int, c_vp],
    "gf_grid_encode_backward": [c_vp, c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_u32, c_u32, c_f32, c_u32, c_vp, c_vp, c_u32,
                                c_int, c_u32, c_int, c_vp],
    "gf_grad_total_variation": [c_vp, c_vp, c_vp, c_vp, c_f32, c_u32, c_u32, c_u32, c_u32, c_f32, c_u32, c_u32, c_int, c_vp],
    "gf_sh_encode_forward": [c_vp, c_vp, c_u32, c_u32, c_u32, c_vp, c_vp],
    "gf_sh_encode_backward": [c_vp, c_vp, c_u32, c_u32, c_u32, c_vp, c_vp, c_vp],
    "gf_freq_encode_forward": [c_vp, c_u32, c_u32, c_u32, c_u32, c_vp, c_vp],
    "gf_freq_encode_backward": [c_vp, c_vp, c_u32, c_u32, c_u32, c_u32, c_vp, c_vp],
    "gf_adnerf_get_rays": [c_u32, c_u32, c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gf_adnerf_embed": [c_vp, c_u32, c_u32, c_u32, c_vp, c_u32, c_vp],
    "gf_adnerf_embed_points": [c_vp, c_vp, c_vp, c_u32, c_u32, c_u32, c_vp, c_u32, c_vp],
    "gf_adnerf_raw2outputs": [c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gf_adnerf_sample_pdf": [c_vp, c_vp, c_vp, c_u32, c_u32, c_u32, c_int, c_vp, c_vp, c_vp],
    "gf_get_rays": [c_vp, c_u32, c_f32, c_f32, c_f32, c_f32, c_u32, c_u32, c_vp, c_u32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "gf_adnerf_mlp_create": [c_vp, c_vp, c_vp],
    "gf_adnerf_mlp_destroy": [c_vp],
    "gf_adnerf_mlp_workspace_bytes": [c_vp, c_u32],
    "gf_adnerf_mlp_forward": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_u32, c_u32, c_vp, c_vp, c_u64, c_vp],
    "gf_tl_tiles_bytes": [c_u32, c_u32],
    "gf_tl_pack": [c_vp, c_int, c_u32, c_u32, c_u32, c_u32, c_u32, c_u32, c_vp, c_vp, c_vp],
    "gf_tl_weight_image": [c_vp, c_u32, c_u32, c_u32, c_u32, c_vp, c_vp],
    "gf_tl_gemm": [c_vp, c_u32, c_vp, c_u32, c_u32, c_int, c_u32, c_vp, c_u32, c_int, c_vp, c_u32, c_vp, c_u32, c_u32, c_vp, c_vp],
    "gf_tl_wgrad": [c_vp, c_u32, c_u32, c_vp, c_u32, c_u32, c_u32, c_vp, c_u32, c_u32, c_u32, c_int, c_vp, c_vp],
    "gf_model_create": [c_vp, c_vp, c_vp],
    "gf_model_destroy": [c_vp],
    "gf_model_packed_bytes": [c_vp],
    "gf_render_workspace_bytes": [c_u32],
    "gf_render_frame": [c_vp, c_vp, c_vp, c_vp, c_u64, c_vp],
    "gf_field_forward": [c_vp, c_vp, c_vp, c_vp, c_u32, c_vp, c_vp, c_vp, c_u32, c_vp, c_u64, c_vp],
    "gf_field_workspace_bytes": [c_u32, c_u32],
    "gf_tc_debug": [c_vp, c_vp],
    "gf_gather_probe": [c_vp, c_vp, c_vp, c_u32, c_vp, c_vp],
    "gf_profile_enable": [c_vp, c_int],
    "gf_profile_field_ms": [c_vp, c_vp, c_vp],
    "gf_last_error": [],
    "gf_version": [],
    "gf_device_ok": [],
}
_RESTYPE = {"gf_last_error": ctypes.c_char_p, "gf_model_destroy": None, "gf_model_packed_bytes": c_u64,
            "gf_render_workspace_bytes": c_u64, "gf_field_workspace_bytes": c_u64, "gf_adnerf_mlp_workspace_bytes": c_u64,
            "gf_adnerf_mlp_destroy": None, "gf_tl_tiles_bytes": ctypes.c_size_t}

EXPORTS = sorted(_SIGS)


def lib():
    """Load (building if stale and nvcc is available) the in-tree shared library."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        try:
            build()
        except Exception as e:  # noqa: BLE001
            # never run an old binary against newer ctypes signatures: a stale library that cannot be rebuilt is an error
            raise RuntimeError("libgfrender.so is %s and could not be rebuilt: %s" % ("stale" if os.path.exists(_SO) else "not built", e))
    try:
        L = ctypes.CDLL(_SO)
    except OSError as e:
        raise RuntimeError("cannot load %s: %s (the CUDA extension is mandatory; there is no fallback)" % (_SO, e))
    for name, argtypes in _SIGS.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            if _VARIANT:          # an experiment build of an older source state (A/B runs): the missing operator simply cannot be called
                continue
            raise RuntimeError("libgfrender.so does not export %s" % name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, ctypes.c_int)
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().gf_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what or "libgfrender call", rc, msg.decode() if msg else "?"))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("geneface_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a (contiguous) tensor, or NULL for None"""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(_SO)
