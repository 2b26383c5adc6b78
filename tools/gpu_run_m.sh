#!/bin/bash
# Round-2 GPU pass M (1 GPU): hot / cold by live-range splitting (asynchronous reloads).
set -u
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests/test_gpu_garble.py tests/test_gpu_stream.py tests/test_gpu_fullsize.py -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
{
timeout 200 python tools/time_circuit.py sha256 1184
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=16 GCB_TWIN=0 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=12 timeout 200 python tools/time_circuit.py sha256 1776
GCB_HOT_TEAMS=16 GCB_HOT_MODE=1 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha512 2368
GCB_HOT_TEAMS=16 GCB_HOT_MODE=1 timeout 200 python tools/time_circuit.py sha512 2368
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha256xor 2368 32
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py chacha20block 2368
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py mul64 4736
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py aes_128 4736
} > gpurun_out/m_times.txt 2>&1
GCB_HOT_TEAMS=16 timeout 400 ncu --set full --clock-control none --import-source on -k regex:garble_kernel -s 1 -c 1 -f -o /tmp/prof/sp python tools/run_one.py sha256 2368 2 > /dev/null 2>&1
ncu -i /tmp/prof/sp.ncu-rep --page raw --csv > gpurun_out/m_sha256split_garble_raw.csv 2>/dev/null
ncu -i /tmp/prof/sp.ncu-rep --page source --csv > gpurun_out/m_sha256split_garble_src.csv 2>/dev/null
tail -4 gpurun_out/m_pytest.log; cat gpurun_out/m_times.txt
