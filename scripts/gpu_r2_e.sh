#!/bin/bash
# round 2, GPU call E: N-split + position prefetch (new default) vs the previous kernels (variant 'old'): tests, bench A/B, launch lists
mkdir -p gpurun_out
echo "== pytest gpu (tc tests first)"; timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
for rep in 1 2; do
echo "== bench new ($rep)"
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_new$rep.json 2> gpurun_out/bench_new$rep.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_new$rep.json'));print('new', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
echo "== bench old ($rep)"
GF_LIBGFRENDER=geneface_b200/variants/libgfrender_old.so timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_old$rep.json 2> gpurun_out/bench_old$rep.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_old$rep.json'));print('old', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
done
echo "== launch list (new)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/ncu_list_new.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_new.csv | head -8
echo "== May torso launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_may_torso.csv python bench.py --config may_torso --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/ncu_list_may.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_may_torso.csv | head -24
echo "== may configs standalone"
for c in may_head may_torso; do timeout 300 python bench.py --config $c --steps 100 --no-ref-cuda > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$c.json'));print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1; done
