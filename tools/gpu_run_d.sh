#!/bin/bash
# Round-2 GPU pass D (1 GPU): balanced schedule with the minimal cold set, e2e with the garbler one step ahead.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
{
timeout 200 python tools/time_circuit.py sha256 1184
GCB_ILP=2 timeout 200 python tools/time_circuit.py sha256 1184
GCB_TEAM_THREADS=64 timeout 200 python tools/time_circuit.py sha256 1184
timeout 200 python tools/time_circuit.py sha256 2368
timeout 200 python tools/time_circuit.py sha256xor 1480 32
timeout 200 python tools/time_circuit.py sha512 444
GCB_ILP=2 timeout 200 python tools/time_circuit.py sha512 444
GCB_HOT_TEAMS=8 timeout 200 python tools/time_circuit.py sha512 1184
timeout 200 python tools/time_circuit.py mul64 4096
timeout 200 python tools/time_circuit.py chacha20block 1480
timeout 200 python tools/time_circuit.py aes_256 4096 32
} > gpurun_out/d_times.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?" >> gpurun_out/d_bench.err
tail -3 gpurun_out/d_pytest.log; cat gpurun_out/d_times.txt; tail -3 gpurun_out/d_bench.err
