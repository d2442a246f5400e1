#!/bin/bash
# A/B of libgfrender experiment builds (scripts/build_variants.py) on the benchmark frame, interleaved twice: usage gpu_ab.sh "v1 v2 ..." ("" = default build)
mkdir -p gpurun_out
for rep in 1 2; do
  for v in ${1:-default}; do
    if [ "$v" != "default" ]; then export GF_LIBGFRENDER=geneface_b200/variants/libgfrender_$v.so; else unset GF_LIBGFRENDER; fi
    echo "== variant '$v'"; timeout 150 python scripts/tc_timeline.py 2>&1 | grep -E "^frame|rror" | tail -2
  done
done
unset GF_LIBGFRENDER
