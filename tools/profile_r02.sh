#!/bin/bash
# Round-2 profile pass (1 GPU): launch list of the bench step, ncu --set full of the headline kernels, the sha256 kernels
# and the IKNP receiver; raw and source pages exported to CSV on the box (the .ncu-rep files embed the cubin: too large).
set -u
mkdir -p gpurun_out /tmp/prof
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches.csv $B > /dev/null 2>&1
for K in garble eval; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 3 -c 1 -f -o /tmp/prof/aes_$K $B > /dev/null 2>&1
  ncu -i /tmp/prof/aes_$K.ncu-rep --page raw --csv > gpurun_out/p_aes128_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/aes_$K.ncu-rep --page source --csv > gpurun_out/p_aes128_${K}_src.csv 2>/dev/null
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/sha_$K python tools/run_one.py sha256 1184 2 > /dev/null 2>&1
  ncu -i /tmp/prof/sha_$K.ncu-rep --page raw --csv > gpurun_out/p_sha256_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/sha_$K.ncu-rep --page source --csv > gpurun_out/p_sha256_${K}_src.csv 2>/dev/null
done
timeout 300 ncu --set full --clock-control none -k regex:iknp_kernel -c 2 -f -o /tmp/prof/iknp python tools/bench_paths.py iknp > /dev/null 2>&1
ncu -i /tmp/prof/iknp.ncu-rep --page raw --csv > gpurun_out/p_iknp_raw.csv 2>/dev/null
ls -la gpurun_out | grep " p_"; head -4 gpurun_out/p_launches.csv | cut -c1-200
