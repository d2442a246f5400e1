#!/bin/bash
# ncu launch list only (per-kernel durations of a short bench run).  Usage: gpu_list.sh [precision]
mkdir -p gpurun_out
P=${1:-fp16}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$P.csv \
    python bench.py --steps 2 --warmup 3 --precision $P --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/ncu_list_$P.log 2>&1
echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/launches_$P.csv
