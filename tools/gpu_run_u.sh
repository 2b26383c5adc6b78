#!/bin/bash
# Round-2 GPU pass U (1 GPU): experiment build (-DGC_ROUNDS_UNROLL=3): the two-block AES rounds three per loop iteration.
set -u
mkdir -p gpurun_out
{
timeout 200 python tools/time_circuit.py aes_128 4096
timeout 200 python tools/time_circuit.py aes_128 4096 32
timeout 200 python tools/time_circuit.py aes_256 4096 32
timeout 200 python tools/time_circuit.py aes_128 148
} > gpurun_out/u_times.txt 2>&1
cat gpurun_out/u_times.txt
