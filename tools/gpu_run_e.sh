#!/bin/bash
# Round-2 GPU pass E (1 GPU): split team headers + balanced schedule, rolled single-block AES, e2e trace.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
{
timeout 200 python tools/time_circuit.py sha256 1184
GCB_HOT_TEAMS=16 timeout 200 python tools/time_circuit.py sha256 2368
GCB_HOT_TEAMS=12 timeout 200 python tools/time_circuit.py sha256 1776
timeout 200 python tools/time_circuit.py sha256xor 1332 32
timeout 200 python tools/time_circuit.py sha512 444
GCB_HOT_TEAMS=8 timeout 200 python tools/time_circuit.py sha512 1184
timeout 200 python tools/time_circuit.py mul64 4096
timeout 200 python tools/time_circuit.py chacha20block 1480
timeout 200 python tools/time_circuit.py aes_128 4096
timeout 200 python tools/time_circuit.py add64 8192
timeout 200 python tools/time_circuit.py div64 2368
} > gpurun_out/e_times.txt 2>&1
GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
GCB_E2E_TRACE=1 GCB_E2E_AHEAD=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/e_bench_noahead.json 2>> gpurun_out/e_bench.err
tail -3 gpurun_out/e_pytest.log; cat gpurun_out/e_times.txt; tail -6 gpurun_out/e_bench.err
