#!/bin/bash
# round 2, GPU call G (8 GPUs of one box): weak-scaling bench at N = 8, strong-scaled 300-frame sequence at N = 1, 2, 4, 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() {  # run N args...
  N=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N "$@"
}
for N in 1 2 4 8; do
  echo "== seq300 N=$N"
  if [ $N = 1 ]; then timeout 300 python bench.py --config seq300 > gpurun_out/seq300_n$N.json 2> gpurun_out/seq300_n$N.err
  else timeout 300 bash -c "$(declare -f run); run $N --config seq300" > gpurun_out/seq300_n$N.json 2> gpurun_out/seq300_n$N.err; fi
  echo "rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/seq300_n$N.json').read().strip().splitlines()[-1]);print($N, d['value'], d['seconds'], d['repeats_s'], d['png_egress'])" 2>&1 | tail -1
done
for N in 8 2; do
  echo "== bench weak N=$N"
  timeout 400 bash -c "$(declare -f run); run $N --steps 50 --warmup 3" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print($N, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])" 2>&1 | tail -1
done
