#!/bin/bash
# Round-2 GPU pass X (1 GPU): whole suite + timings after a change.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
{
timeout 200 python tools/time_circuit.py sha256xor 2368 32
timeout 200 python tools/time_circuit.py chacha20block 2368
timeout 200 python tools/time_circuit.py sha256 2368
timeout 200 python tools/time_circuit.py sha512 2368
} > gpurun_out/x_times.txt 2>&1
tail -4 gpurun_out/x_pytest.log; cat gpurun_out/x_times.txt
