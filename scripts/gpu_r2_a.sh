#!/bin/bash
# round 2, GPU call A: golden frames from the REAL reference render(), full GPU test suite, gather ubench, smoke, bench (default + variants)
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
echo "== golden frames"; timeout 600 python oracle/gen_golden_frames.py gpurun_out/golden > gpurun_out/golden_frames.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/golden_frames.log
cp gpurun_out/golden/frame_*.npz tests/golden/ 2>/dev/null
echo "== ubench"; ./scripts/ubench/l2_gather_bw > gpurun_out/l2_gather_bw.json 2>&1; cat gpurun_out/l2_gather_bw.json
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 600 python bench.py --steps 30 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
for v in spec bias both; do
  echo "== bench variant $v"
  GF_LIBGFRENDER=geneface_b200/variants/libgfrender_$v.so timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_$v.json'));print('$v', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
done
echo "== bench eager"; timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-ref-cuda --no-may --eager > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_eager.json'));print('eager', d['value'], d['e2e']['value'], d['roofline']['kernel_ms_per_frame'])" 2>&1 | tail -1
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_ref.json
