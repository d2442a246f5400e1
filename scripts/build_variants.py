#!/usr/bin/env python
"""Build experiment variants of libgfrender.so (extra -D switches) next to the default library, for A/B runs on the GPU box:

    python scripts/build_variants.py spec=-DGF_SPECIALIZE_GRID=1 bias=-DGF_BIAS_IN_ACC=1 both=-DGF_SPECIALIZE_GRID=1,-DGF_BIAS_IN_ACC=1
    GF_LIBGFRENDER=geneface_b200/variants/libgfrender_spec.so python bench.py ...

Variants are git-ignored (*.so) and travel to the GPU box with the gpurun snapshot."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geneface_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    out_dir = os.path.join(ROOT, "geneface_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        tag, flags = spec.split("=", 1)
        so = os.path.join(out_dir, "libgfrender_%s.so" % tag)
        _lib.build(out=so, extra=[f for f in flags.split(",") if f])
        print("built", so, flags)
