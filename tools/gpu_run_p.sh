#!/bin/bash
# Round-2 GPU pass P (1 GPU): stagger sweep (start offset between the teams of an SM) for the headline and the split-plan kernels.
set -u
mkdir -p gpurun_out
{
for S in 0 3000 10000 30000; do GCB_STAGGER=$S timeout 200 python tools/time_circuit.py aes_128 4096; done
for S in 0 10000 30000; do GCB_NT=2 GCB_TEAMS=8 GCB_STAGGER=$S timeout 200 python tools/time_circuit.py aes_128 4096; done
for S in 5000 10000 20000 40000 60000; do GCB_STAGGER=$S timeout 200 python tools/time_circuit.py sha256 2368; done
for S in 10000 30000 100000; do GCB_STAGGER=$S timeout 200 python tools/time_circuit.py sha256 4736; done
for S in 10000 30000 100000; do GCB_STAGGER=$S timeout 200 python tools/time_circuit.py sha256 1184; done
for S in 10000 30000; do GCB_STAGGER=$S timeout 200 python tools/time_circuit.py sha512 2368; done
for S in 0 10000 30000; do GCB_STAGGER=$S timeout 200 python tools/time_circuit.py aes_128 4096 32; done
} > gpurun_out/p_times.txt 2>&1
cat gpurun_out/p_times.txt
