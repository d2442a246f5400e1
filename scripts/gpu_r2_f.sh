#!/bin/bash
# round 2, GPU call F: tests; position-prefetch (default) vs previous kernels ('old'); May configs after the histogram fix; train step with CUDA graph
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
for rep in 1 2; do
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_new$rep.json 2> gpurun_out/bench_new$rep.err
python -c "
import json;d=json.load(open('gpurun_out/bench_new$rep.json'));print('new', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
GF_LIBGFRENDER=geneface_b200/variants/libgfrender_old.so timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_old$rep.json 2> gpurun_out/bench_old$rep.err
python -c "
import json;d=json.load(open('gpurun_out/bench_old$rep.json'));print('old', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
done
echo "== may configs standalone"
for c in may_head may_torso; do timeout 300 python bench.py --config $c --steps 100 --no-ref-cuda > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; python -c "
import json;d=json.load(open('gpurun_out/bench_$c.json'));print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1; done
echo "== May torso launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_may_torso.csv python bench.py --config may_torso --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/ncu_list_may.log 2>&1; echo "rc=$?"
python scripts/launch_summary.py gpurun_out/launches_may_torso.csv | head -8
for A in "" "--amp"; do
echo "== train 4096 $A (with CUDA graph arm)"
timeout 300 python scripts/bench_train.py --rays 4096 --steps 30 $A > gpurun_out/train_4096g$A.json 2> gpurun_out/train_4096g$A.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/train_4096g$A.json'));print(d['ms_per_step'], d['cuda_graph'], d['reference_cuda'])" 2>&1 | tail -2
done
echo "== train 65536 (graph arm)"
timeout 300 python scripts/bench_train.py --rays 65536 --steps 20 > gpurun_out/train_65536g.json 2> gpurun_out/train_65536g.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/train_65536g.json'));print(d['ms_per_step'], d['cuda_graph'], d['reference_cuda'])" 2>&1 | tail -2
