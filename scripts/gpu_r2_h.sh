#!/bin/bash
# round 2, GPU call H: verification after the torso re-tiling: tests, smoke, final bench lines, train graph arm diagnostics
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
echo "== bench default"; timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
r=d['roofline']
print('default', d['value'], d['e2e']['value'], r['kernel_ms_per_frame'], r['frac'], r.get('frac_of_gather_ceiling'), r.get('traffic'))
for k in ('may_head','may_torso','adnerf_gpu','parity','reference_cuda','cpu_baseline'): print(k, {kk: vv for kk, vv in d[k].items() if kk in ('value','ms_per_frame','eager_ms_per_frame','rgb_worst','ok','launches_per_frame')})
PY
echo "== train graph arm"; timeout 300 python scripts/bench_train.py --graph-only --rays 4096 --steps 20 > gpurun_out/train_graph.json 2> gpurun_out/train_graph.err; echo "rc=$?"; tail -c 2500 gpurun_out/train_graph.json; tail -5 gpurun_out/train_graph.err
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_ref.json
