#!/bin/bash
# Round-2 GPU pass Z (1 GPU): geometry of mul64-like circuits (wide by average level, 10 instances beside four tables, 16 beside two).
set -u
mkdir -p gpurun_out
{
timeout 200 python tools/time_circuit.py mul64 4736
GCB_NT=2 timeout 200 python tools/time_circuit.py mul64 4736
GCB_NT=2 GCB_TEAMS=8 timeout 200 python tools/time_circuit.py mul64 4736
GCB_NT=2 GCB_ILP=1 timeout 200 python tools/time_circuit.py mul64 4736
GCB_NT=4 GCB_TEAMS=8 GCB_TEAM_THREADS=64 timeout 200 python tools/time_circuit.py mul64 4736
timeout 200 python tools/time_circuit.py div64 2368
GCB_NT=2 timeout 200 python tools/time_circuit.py div64 2368
} > gpurun_out/z_times.txt 2>&1
cat gpurun_out/z_times.txt
