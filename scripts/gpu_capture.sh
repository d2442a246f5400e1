#!/bin/bash
# one `ncu --set full` capture.  Usage: gpu_capture.sh <kernel-regex> <skip> <count> [precision]
mkdir -p gpurun_out
P=${4:-fp16}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-10} -c ${3:-1} -f -o gpurun_out/prof_$P \
    python bench.py --steps 2 --warmup 3 --precision $P --no-cpu-baseline --no-ref-cuda --no-may > gpurun_out/ncu_full_$P.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full_$P.log; ls -la gpurun_out/*.ncu-rep
