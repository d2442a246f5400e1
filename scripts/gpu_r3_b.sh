#!/bin/bash
# round 2, session 2, call B: tests of the tensor-core MLP and the cached 2-D grid backward; train-step timings with / without the cache
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_linear_gpu.py "tests/test_parity_gpu.py::test_grid_backward_kernels_and_fp16_path" tests/test_parity_gpu.py::test_train_step_gradients_vs_the_reference_train_step -q -s 2>&1 | grep -E "tc backend|passed|failed|Error|assert|max err" | tail -24
for cfg in "65536 tc x" "65536 tc x nocache" "65536 tc --amp" "65536 torch x" "4096 tc x"; do
  set -- $cfg
  A=""; [ "$3" = "--amp" ] && A="--amp"
  NOREF="--no-ref"; [ "$2$3$4" = "tcx" ] && [ "$1" = "65536" ] && NOREF=""
  GF_GRID_BWD=${4:-auto} timeout 400 python scripts/bench_train.py --rays $1 --steps 10 --warmup 3 --mlp $2 $A $NOREF > gpurun_out/train_$1_$2$3$4.json 2> gpurun_out/train_$1_$2$3$4.err || tail -3 gpurun_out/train_$1_$2$3$4.err
  python - "$1 $2 $3 $4" gpurun_out/train_$1_$2$3$4.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step %.3f" % d['ms_per_step'], "graph", d['cuda_graph'].get('ms_per_step'), "ref", d['reference_cuda'].get('ms_per_step'), "loss %.5f finite %s" % (d['loss'], d['grads_finite']))
    for k in d['top_kernels'][:5]: print("     %5.1f%% x%-4d %s" % (100 * k['share'], k['calls'], k['name'][:70]))
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
