#!/bin/bash
# Round-2 GPU pass V (1 GPU): the round's final state: whole GPU suite, smoke, both bench arms.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/v_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench.err
GCB_E2E_TRACE=1 timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/v_bench.json 2>> gpurun_out/v_bench.err; echo "bench rc=$?" >> gpurun_out/v_bench.err
tail -3 gpurun_out/v_pytest.log; tail -2 gpurun_out/v_smoke.log; tail -4 gpurun_out/v_bench.err; head -c 400 gpurun_out/v_bench.json; echo; cat gpurun_out/v_bench_ref.json | head -c 300
