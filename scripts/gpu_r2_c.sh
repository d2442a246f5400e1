#!/bin/bash
# round 2, GPU call C: tests, bench with gather probe, L0-staging variant, seq300 (1 GPU), train-step A/B at 4096 / 65536 rays, ncu profile
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
echo "== bench default"; timeout 600 python bench.py --steps 30 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
r=d['roofline']
print('default', d['value'], d['e2e']['value'], r['kernel_ms_per_frame'], r['frac'], r.get('frac_of_gather_ceiling'), r.get('gather_probe'))
PY
echo "== bench variant l0"
GF_LIBGFRENDER=geneface_b200/variants/libgfrender_l0.so timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_l0.json 2> gpurun_out/bench_l0.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_l0.json'));print('l0', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
echo "== seq300 (1 GPU)"; timeout 300 python bench.py --config seq300 > gpurun_out/seq300_n1.json 2> gpurun_out/seq300_n1.err; echo "rc=$?"; tail -c 1200 gpurun_out/seq300_n1.json; tail -2 gpurun_out/seq300_n1.err
for R in 4096 65536; do
  for M in default legacy priv; do
    echo "== train step rays=$R grid_bwd=$M"
    if [ $M = default ]; then unset GF_GRID_BWD; else export GF_GRID_BWD=$M; fi
    timeout 300 python scripts/bench_train.py --rays $R --steps 20 > gpurun_out/train_${R}_$M.json 2> gpurun_out/train_${R}_$M.err; echo "rc=$?"
    python -c "
import json;d=json.load(open('gpurun_out/train_${R}_$M.json'));print(d['ms_per_step'], d['mean_count'], d['reference_cuda']); [print('   ', k['name'][:60], round(k['share'],3)) for k in d['top_kernels'][:5]]" 2>&1 | tail -7
  done
done
unset GF_GRID_BWD
echo "== ncu profile"; bash scripts/gpu_profile.sh r02 > gpurun_out/profile_r02.log 2>&1; tail -5 gpurun_out/profile_r02.log
echo "== ncu L0 variant (k_tc_amb only)"
GF_LIBGFRENDER=geneface_b200/variants/libgfrender_l0.so timeout 600 ncu --set full --clock-control none -k regex:k_tc_amb -s 10 -c 1 -f -o gpurun_out/prof_l0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/ncu_full_l0.log 2>&1; echo "rc=$?"
python scripts/ncu_field_json.py gpurun_out/prof_l0.ncu-rep gpurun_out/field_ncu_l0.json | cut -c1-700
