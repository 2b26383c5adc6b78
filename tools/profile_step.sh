#!/bin/bash
# Profiles the bench step under ncu on the GPU box and leaves only small CSV exports in gpurun_out/
# (the .ncu-rep files embed the whole cubin and exceed the transfer limit).
set -u
TAG=${1:-vX}
mkdir -p gpurun_out /tmp/prof
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
for K in garble eval; do
  ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 3 -c 1 -f -o /tmp/prof/$K $B > /dev/null 2>&1
  ncu -i /tmp/prof/$K.ncu-rep --page raw --csv > gpurun_out/${K}_${TAG}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$K.ncu-rep --page source --csv > gpurun_out/${K}_${TAG}_src.csv 2>/dev/null
done
ncu --set full --clock-control none -k regex:iknp_kernel -c 2 -f -o /tmp/prof/iknp python tools/bench_paths.py iknp > /dev/null 2>&1
ncu -i /tmp/prof/iknp.ncu-rep --page raw --csv > gpurun_out/iknp_${TAG}_raw.csv 2>/dev/null
python tools/bench_paths.py > gpurun_out/paths_$TAG.jsonl 2>/dev/null
./tools/_build/microbench_lds > gpurun_out/microbench_$TAG.log 2>&1
ls -la gpurun_out | grep $TAG
