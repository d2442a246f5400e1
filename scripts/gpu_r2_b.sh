#!/bin/bash
# round 2, GPU call B: smoke, full GPU test suite, bench (default + variants + eager), train-step A/B of the grid backward
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR|adnerf tc|worst scaled" gpurun_out/pytest_gpu.log | tail -40
echo "== bench default"; timeout 600 python bench.py --steps 30 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -c 4000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
for v in spec bias both; do
  echo "== bench variant $v"
  GF_LIBGFRENDER=geneface_b200/variants/libgfrender_$v.so timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_$v.json'));print('$v', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
done
echo "== bench eager"; timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_eager.json'));print('eager', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
echo "== train step (new grid backward)"; timeout 300 python scripts/bench_train.py > gpurun_out/train_new.json 2> gpurun_out/train_new.err; echo "rc=$?"; tail -c 1500 gpurun_out/train_new.json
echo "== train step (legacy grid backward)"; GF_GRID_BWD=legacy timeout 300 python scripts/bench_train.py > gpurun_out/train_legacy.json 2> gpurun_out/train_legacy.err; echo "rc=$?"; tail -c 1500 gpurun_out/train_legacy.json
