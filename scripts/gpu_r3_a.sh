#!/bin/bash
# round 2, session 2, call A: tensor-core training MLP tests + train-step timings; A/B of the deferred feat_hi stores
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_linear_gpu.py -q -s 2>&1 | grep -E "dims|tc backend|passed|failed|Error|assert" | tail -14
for cfg in "65536 tc" "65536 tc --amp" "65536 torch --amp" "4096 tc" "4096 tc --amp"; do
  set -- $cfg
  timeout 300 python scripts/bench_train.py --rays $1 --steps 10 --warmup 3 --mlp $2 $3 --no-ref > gpurun_out/train_$1_$2$3.json 2> gpurun_out/train_$1_$2$3.err || tail -3 gpurun_out/train_$1_$2$3.err
  python - "$1 $2 $3" gpurun_out/train_$1_$2$3.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step %.3f" % d['ms_per_step'], "graph", d['cuda_graph'], "loss %.5f finite %s" % (d['loss'], d['grads_finite']))
    for k in d['top_kernels'][:6]: print("     %5.1f%% x%-4d %s" % (100 * k['share'], k['calls'], k['name'][:70]))
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
for v in c1 "" c1 ""; do
  if [ -n "$v" ]; then export GF_LIBGFRENDER=geneface_b200/variants/libgfrender_$v.so; else unset GF_LIBGFRENDER; fi
  echo "== variant '${v:-default}'"; timeout 150 python scripts/tc_timeline.py 2>&1 | grep -E "^frame|rror" | tail -2
done
