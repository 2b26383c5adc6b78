#!/bin/bash
set -u
mkdir -p gpurun_out
for A in 0 2 3 5; do
GCB_E2E_AHEAD=$A GCB_E2E_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/l2_bench_a${A}.json 2>> gpurun_out/l2_e2e.err
done
cat gpurun_out/l2_e2e.err
