#!/bin/bash
# Round-2 GPU pass O (1 GPU): reference-transcript pin on the CUDA path, sanitizer runs of the split-plan kernels,
# geometry experiments (two resident tables with more, smaller teams on aes_128; stagger of 16 one-warp teams).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_transcript.py -q > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
GCB_HOT_TEAMS=16 timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python tools/run_one.py sha256 20 1 > gpurun_out/o_race_sha256_split.txt 2>&1
GCB_HOT_TEAMS=16 timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python tools/run_one.py sha512 20 1 > gpurun_out/o_memcheck_sha512_split.txt 2>&1
{
GCB_NT=2 GCB_TEAMS=8 timeout 200 python tools/time_circuit.py aes_128 4096
GCB_NT=2 GCB_TEAMS=7 timeout 200 python tools/time_circuit.py aes_128 4096
GCB_NT=2 GCB_TEAMS=5 timeout 200 python tools/time_circuit.py aes_128 4096
GCB_NT=2 GCB_TEAMS=9 GCB_ILP=1 timeout 200 python tools/time_circuit.py aes_128 4096
GCB_HOT_TEAMS=10 timeout 200 python tools/time_circuit.py aes_128 4096
GCB_STAGGER=0 timeout 200 python tools/time_circuit.py sha256 2368
GCB_STAGGER=30000 timeout 200 python tools/time_circuit.py sha256 2368
GCB_STAGGER=300000 timeout 200 python tools/time_circuit.py sha256 2368
GCB_STAGGER=0 timeout 200 python tools/time_circuit.py aes_128 4096
GCB_STAGGER=300000 timeout 200 python tools/time_circuit.py aes_128 4096
} > gpurun_out/o_times.txt 2>&1
tail -3 gpurun_out/o_pytest.log; tail -3 gpurun_out/o_race_sha256_split.txt; tail -3 gpurun_out/o_memcheck_sha512_split.txt; cat gpurun_out/o_times.txt
