#!/bin/bash
# round 2, session 2, final call: all GPU tests, smoke, the bench lines of every configuration, ncu launch list + full capture of the field pair
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
echo "== bench default"; timeout 700 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    r=d['roofline']
    print('default', d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel_ms_per_frame'], r['frac'], r.get('frac_of_gather_ceiling'), d['clocks'])
    for k in ('may_head','may_torso','adnerf_gpu','parity','reference_cuda','cpu_baseline'): print(k, {kk: vv for kk, vv in d.get(k, {}).items() if kk in ('value','ms_per_frame','eager_ms_per_frame','rgb_worst','ok','launches_per_frame')})
except Exception as e: print('bench default parse failed', e)
PY
for C in may_head may_torso seq300; do
  echo "== bench --config $C"; timeout 400 python bench.py --config $C > gpurun_out/bench_$C.json 2> gpurun_out/bench_$C.err; echo "rc=$?"; tail -c 700 gpurun_out/bench_$C.json | head -c 700; echo
done
echo "== bench --config train"; timeout 900 python bench.py --config train --steps 20 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_train.json').read().strip().splitlines()[-1])
    for k, v in (('fp32', d), ('amp', d['amp']), ('tc', d['tc_mlp']), ('tc 65536', d['tc_mlp_65536'])):
        print(k, v['ms_per_step'], v['cuda_graph'].get('ms_per_step'), v['reference_cuda'].get('ms_per_step'))
except Exception as e: print('train parse failed', e)
PY
echo "== ncu"; bash scripts/gpu_profile.sh r02c > gpurun_out/gpu_profile_r02c.log 2>&1; tail -4 gpurun_out/gpu_profile_r02c.log
echo "== train 65536 with the reference arms"
timeout 240 python scripts/bench_train.py --rays 65536 --steps 10 --warmup 3 --mlp tc > gpurun_out/train_65536_tc_ref.json 2> gpurun_out/train_65536_tc_ref.err; echo "rc=$?"
timeout 240 python scripts/bench_train.py --rays 65536 --steps 10 --warmup 3 --mlp tc --amp > gpurun_out/train_65536_tc_amp_ref.json 2> gpurun_out/train_65536_tc_amp_ref.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('train_65536_tc_ref', 'train_65536_tc_amp_ref'):
    try:
        d=json.load(open('gpurun_out/%s.json' % f)); print(f, d['ms_per_step'], d['cuda_graph'].get('ms_per_step'), d['reference_cuda'])
    except Exception as e: print(f, 'failed', e)
PY
ls -la gpurun_out | tail -30
