#!/bin/bash
# GPU pass: parity tests (all, no -x), smoke, a short bench.  Usage: gpu_check.sh [pytest -k expr]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
if [ -n "$1" ]; then KEXPR=(-k "$1"); else KEXPR=(); fi
timeout 1700 python -m pytest tests -m gpu -q --timeout=900 -s "${KEXPR[@]}" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|tc sigma|tc frame|rc=" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
for P in fp32 fp16; do
  timeout 900 python bench.py --steps 5 --warmup 3 --precision $P --no-cpu-baseline > gpurun_out/bench_$P.json 2> gpurun_out/bench_$P.err; echo "bench $P rc=$?"; tail -c 1800 gpurun_out/bench_$P.json; tail -5 gpurun_out/bench_$P.err
done
