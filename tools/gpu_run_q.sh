#!/bin/bash
# Round-2 GPU pass Q (1 GPU): new geometry defaults (wide circuits: 8 two-warp teams on two tables; no start offset for
# multi-warp teams): timings against the previous choice, the whole GPU suite, the default bench line + launch list.
set -u
mkdir -p gpurun_out
{
timeout 200 python tools/time_circuit.py aes_128 4096
GCB_NT=4 timeout 200 python tools/time_circuit.py aes_128 4096
timeout 200 python tools/time_circuit.py aes_128 4096 32
GCB_NT=4 timeout 200 python tools/time_circuit.py aes_128 4096 32
timeout 200 python tools/time_circuit.py aes_256 4096 32
GCB_NT=4 timeout 200 python tools/time_circuit.py aes_256 4096 32
timeout 200 python tools/time_circuit.py aes_128 4736
timeout 200 python tools/time_circuit.py sha512 444
timeout 200 python tools/time_circuit.py sha256 2368
timeout 200 python tools/time_circuit.py mul64 4736
} > gpurun_out/q_times.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?" >> gpurun_out/q_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
cat gpurun_out/q_times.txt; tail -4 gpurun_out/q_pytest.log; tail -2 gpurun_out/q_bench.err; head -c 300 gpurun_out/q_bench.json
