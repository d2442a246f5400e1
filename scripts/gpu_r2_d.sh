#!/bin/bash
# round 2, GPU call D: new tests, L0-staging variant (bench + ncu), train-step (fp32 + amp) at 4096 / 65536 rays, May-config launch list, ncu profile
mkdir -p gpurun_out
echo "== pytest gpu (grid backward modes, full suite)"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
echo "== bench variant l0"
GF_LIBGFRENDER=geneface_b200/variants/libgfrender_l0.so timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_l0.json 2> gpurun_out/bench_l0.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_l0.json'));print('l0', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_base.json'));print('base', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
for R in 4096 65536; do
  for A in "" "--amp"; do
    echo "== train step rays=$R $A"
    timeout 300 python scripts/bench_train.py --rays $R --steps 20 $A > gpurun_out/train_${R}${A}.json 2> gpurun_out/train_${R}${A}.err; echo "rc=$?"
    python -c "
import json;d=json.load(open('gpurun_out/train_${R}${A}.json'));print(d['ms_per_step'], d['mean_count'], d['reference_cuda']); [print('   ', k['name'][:60], round(k['share'],3)) for k in d['top_kernels'][:5]]" 2>&1 | tail -7
  done
done
echo "== May torso launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_may_torso.csv python bench.py --config may_torso --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/ncu_list_may.log 2>&1; echo "rc=$?"
echo "== ncu profile"; bash scripts/gpu_profile.sh r02 > gpurun_out/profile_r02.log 2>&1; tail -3 gpurun_out/profile_r02.log | cut -c1-1500
echo "== ncu L0 variant (k_tc_amb only)"
GF_LIBGFRENDER=geneface_b200/variants/libgfrender_l0.so timeout 600 ncu --set full --clock-control none -k regex:k_tc_amb -s 10 -c 2 -f -o gpurun_out/prof_l0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/ncu_full_l0.log 2>&1; echo "rc=$?"
python scripts/ncu_field_json.py gpurun_out/prof_l0.ncu-rep gpurun_out/field_ncu_l0.json | cut -c1-700
