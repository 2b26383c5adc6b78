#!/bin/bash
# Round-2 GPU pass C (1 GPU): hot / cold plans, e2e pipelined over steps, hybrid microbench.
set -u
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
for T in 0 8 16; do GCB_HOT_TEAMS=$T timeout 200 python tools/time_circuit.py sha256 2368; done > gpurun_out/c_hot.txt 2>&1
for T in 0 4 6 8 12 16; do GCB_HOT_TEAMS=$T timeout 200 python tools/time_circuit.py sha512 1184; done >> gpurun_out/c_hot.txt 2>&1
for T in 0 16; do GCB_HOT_TEAMS=$T timeout 200 python tools/time_circuit.py sha256xor 2368 32; GCB_HOT_TEAMS=$T timeout 200 python tools/time_circuit.py chacha20block 2368; done >> gpurun_out/c_hot.txt 2>&1
GCB_HOT_TEAMS=16 GCB_TEAM_THREADS=64 timeout 200 python tools/time_circuit.py sha256 2368 >> gpurun_out/c_hot.txt 2>&1
timeout 200 python tools/time_circuit.py sha256 1184 >> gpurun_out/c_hot.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_bench.err
timeout 300 ./tools/_build/microbench_bs > gpurun_out/c_microbench_bs.log 2>&1
for K in garble eval; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/sha_$K python tools/run_one.py sha256 2368 2 > /dev/null 2>&1
  ncu -i /tmp/prof/sha_$K.ncu-rep --page raw --csv > gpurun_out/c_sha256_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/sha_$K.ncu-rep --page source --csv > gpurun_out/c_sha256_${K}_src.csv 2>/dev/null
done
tail -3 gpurun_out/c_pytest.log; cat gpurun_out/c_hot.txt; tail -3 gpurun_out/c_bench.err; tail -5 gpurun_out/c_microbench_bs.log
