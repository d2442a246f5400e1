#!/bin/bash
# First GPU pass: golden fixtures from the compiled reference, parity tests, a quick fp32 frame timing.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python oracle/gen_golden_gpu.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?" >> gpurun_out/golden.log
tail -3 gpurun_out/golden.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
