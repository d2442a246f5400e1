#!/bin/bash
# ncu launch list + one full capture of the dominant kernel (B200_PROFILING.md recipe).  1 GPU.
set -x
mkdir -p gpurun_out
P=${1:-fp16}
K=${2:-k_field_tc}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$P.csv \
    python bench.py --steps 2 --warmup 3 --precision $P --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/ncu_list_$P.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_list_$P.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${3:-10} -c ${4:-1} -f -o gpurun_out/prof_$P \
    python bench.py --steps 2 --warmup 3 --precision $P --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/ncu_full_$P.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full_$P.log; ls -la gpurun_out/*.ncu-rep
