#!/bin/bash
# Round-2 GPU pass I (1 GPU): full parity suite, smoke, final bench lines (both arms), launch list, sha512 many-plan timing.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/i_smoke.log
{
timeout 200 python tools/time_circuit.py sha512 2368
timeout 200 python tools/time_circuit.py sha512 444
timeout 200 python tools/time_circuit.py sha256 1184
timeout 200 python tools/time_circuit.py aes_128 4096
} > gpurun_out/i_times.txt 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench.err
GCB_E2E_TRACE=1 timeout 900 python bench.py > gpurun_out/i_bench.json 2>> gpurun_out/i_bench.err; echo "bench rc=$?" >> gpurun_out/i_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
tail -3 gpurun_out/i_pytest.log; cat gpurun_out/i_smoke.log | tail -2; cat gpurun_out/i_times.txt; tail -4 gpurun_out/i_bench.err
