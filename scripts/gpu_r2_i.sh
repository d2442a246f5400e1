#!/bin/bash
# round 2, GPU call I: verification after the encoder rewrites (grid forward single-gather, unified backward), bench --config train
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
echo "== bench --config train"; timeout 900 python bench.py --config train --steps 20 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1])
print('fp32', d['ms_per_step'], d['cuda_graph'], d['reference_cuda'].get('ms_per_step'))
a=d['amp']; print('amp ', a['ms_per_step'], a['cuda_graph'], a['reference_cuda'].get('ms_per_step'))
PY
tail -3 gpurun_out/bench_train.err
echo "== train 65536"; timeout 300 python scripts/bench_train.py --rays 65536 --steps 20 > gpurun_out/train_65536.json 2> gpurun_out/train_65536.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/train_65536.json'));print(d['ms_per_step'], d['cuda_graph'], d['reference_cuda'].get('ms_per_step')); [print('   ', k['name'][:60], round(k['share'],3)) for k in d['top_kernels'][:4]]" 2>&1 | tail -6
echo "== bench short"; timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-may > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_short.json'));print(d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'], d['parity'], d['clocks'])" 2>&1 | tail -1
