#!/bin/bash
# ncu launch list + one full capture of the dominant kernel pair (B200_PROFILING.md recipe).  1 GPU.  Never a bench value.
#   bash scripts/gpu_profile.sh [tag]   ->  gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep, gpurun_out/field_ncu_<tag>.json
set -x
mkdir -p gpurun_out
T=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --precision fp16 --no-cpu-baseline --no-ref-cuda --no-may --eager"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_list_$T.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_list_$T.log
# four consecutive k_tc_* launches inside the warm-up frames (10 per frame: 4 full rounds of 8,388,608 samples + the empty extra round);
# scripts/ncu_field_json.py keeps the longest launch of each kernel = one full round
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 20 -c 4 -f -o gpurun_out/prof_$T $CMD > gpurun_out/ncu_full_$T.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full_$T.log; ls -la gpurun_out/*.ncu-rep
python scripts/ncu_field_json.py gpurun_out/prof_$T.ncu-rep gpurun_out/field_ncu_$T.json
