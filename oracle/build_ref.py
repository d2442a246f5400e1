"""Build the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Compiles the four reference extensions from the
sources where they lie under /root/reference (never copied into this repo):

    modules/radnerfs/raymarching/src/{raymarching.cu,bindings.cpp}   -> _raymarching_face
    modules/radnerfs/encoders/gridencoder/src/{gridencoder.cu,bindings.cpp} -> _gridencoder
    modules/radnerfs/encoders/shencoder/src/{shencoder.cu,bindings.cpp}   -> _shencoder
    modules/radnerfs/encoders/freqencoder/src/{freqencoder.cu,bindings.cpp} -> _freqencoder

The only deviation from the reference's own JIT recipe (raymarching/backend.py:6-12)
is -std=c++17 instead of c++14 (torch 2.11 headers need C++17) and the explicit
sm_100a arch.  Outputs go to oracle/_ref/<name>/<name>.so (git-ignored, shipped
to the GPU box by gpurun).  The compiled reference is the GPU-side parity oracle
and the "kernel to beat"; nothing in the product path imports it.
"""
import os
import sys
import shutil

REF = os.environ.get("GF_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "_raymarching_face": "modules/radnerfs/raymarching/src",
    "_gridencoder": "modules/radnerfs/encoders/gridencoder/src",
    "_shencoder": "modules/radnerfs/encoders/shencoder/src",
    "_freqencoder": "modules/radnerfs/encoders/freqencoder/src",
}
CU = {
    "_raymarching_face": "raymarching.cu",
    "_gridencoder": "gridencoder.cu",
    "_shencoder": "shencoder.cu",
    "_freqencoder": "freqencoder.cu",
}


def built(name):
    return os.path.exists(os.path.join(OUT, name + ".so"))


def build_one(name, verbose=False):
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    src = os.path.join(REF, EXTS[name])
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    load(
        name=name,
        sources=[os.path.join(src, CU[name]), os.path.join(src, "bindings.cpp")],
        extra_cflags=["-O3", "-std=c++17"],
        extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__",
                           "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__"],
        build_directory=bdir,
        verbose=verbose,
        is_python_module=False,   # do not import here; tests import from oracle/_ref
    )
    shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))


def main():
    if not os.path.isdir(REF):
        print("reference tree not present (%s): skipping oracle/_ref build" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(EXTS)
    for n in names:
        if built(n):
            print("[oracle/_ref] %s already built" % n)
            continue
        print("[oracle/_ref] building %s ..." % n, flush=True)
        build_one(n, verbose=False)
        print("[oracle/_ref] %s done" % n, flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
