#!/bin/bash
# Round-2 GPU pass F (1 GPU): e2e schedules, sha512 default plan, stream program.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
{
timeout 200 python tools/time_circuit.py sha512 1184
GCB_HOT_TEAMS=6 timeout 200 python tools/time_circuit.py sha512 888
GCB_HOT_TEAMS=0 timeout 200 python tools/time_circuit.py sha512 444
} > gpurun_out/f_times.txt 2>&1
for P in 16 32; do for A in 1 0; do
GCB_E2E_TRACE=1 GCB_E2E_AHEAD=$A GCB_E2E_PARTS=$P timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/f_bench_a${A}_p${P}.json 2>> gpurun_out/f_e2e.err
done; done
GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> gpurun_out/f_bench.err
tail -3 gpurun_out/f_pytest.log; cat gpurun_out/f_times.txt; cat gpurun_out/f_e2e.err; tail -4 gpurun_out/f_bench.err
