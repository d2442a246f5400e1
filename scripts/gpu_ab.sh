#!/bin/bash
# A/B of libgfrender builds on the benchmark frame: base (round-2 checkpoint), current default, timing build (phase timeline).
mkdir -p gpurun_out
for v in c1 "" c1 ""; do
  if [ -n "$v" ]; then export GF_LIBGFRENDER=geneface_b200/variants/libgfrender_$v.so; else unset GF_LIBGFRENDER; fi
  echo "== variant '${v:-default}'"; timeout 150 python scripts/tc_timeline.py 2>&1 | grep -E "^frame|Error|error" | tail -3
done
export GF_LIBGFRENDER=geneface_b200/variants/libgfrender_timing.so
echo "== timing build"; timeout 150 python scripts/tc_timeline.py gpurun_out/tc_timeline_${1:-x}.json 2>&1 | tail -70
unset GF_LIBGFRENDER
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -k "tc or fused or field" 2>&1 | tail -5
