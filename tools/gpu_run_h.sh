#!/bin/bash
# Round-2 GPU pass H (1 GPU): twin (lock-step) teams.
set -u
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests/test_gpu_garble.py tests/test_gpu_stream.py -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
{
timeout 200 python tools/time_circuit.py sha256 1184
GCB_TWIN=1 timeout 200 python tools/time_circuit.py sha256 1184
GCB_TWIN=1 GCB_STAGGER=0 timeout 200 python tools/time_circuit.py sha256 1184
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=16 GCB_TWIN=0 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=16 GCB_STAGGER=0 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha512 2368
GCB_HOT_TEAMS=8 timeout 200 python tools/time_circuit.py sha512 1184
GCB_HOT_TEAMS=8 GCB_TWIN=1 timeout 200 python tools/time_circuit.py sha512 1184
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py chacha20block 2368
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha256xor 2368 32
} > gpurun_out/h_times.txt 2>&1
GCB_HOT_TEAMS=16 timeout 400 ncu --set full --clock-control none --import-source on -k regex:garble_kernel -s 1 -c 1 -f -o /tmp/prof/tw python tools/run_one.py sha256 2368 2 > /dev/null 2>&1
ncu -i /tmp/prof/tw.ncu-rep --page raw --csv > gpurun_out/h_sha256tw_garble_raw.csv 2>/dev/null
ncu -i /tmp/prof/tw.ncu-rep --page source --csv > gpurun_out/h_sha256tw_garble_src.csv 2>/dev/null
tail -3 gpurun_out/h_pytest.log; cat gpurun_out/h_times.txt
