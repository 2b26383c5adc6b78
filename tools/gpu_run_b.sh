#!/bin/bash
# Round-2 GPU pass B (1 GPU): parity with the balanced plans, bench with the shared stream set, bitsliced microbench,
# sha256 geometry sweep, ncu of the headline kernels.
set -u
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> gpurun_out/b_bench.err
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/b_bench_conn8.json 2>> gpurun_out/b_bench.err
timeout 300 ./tools/_build/microbench_bs > gpurun_out/b_microbench_bs.log 2>&1
timeout 400 python tools/tune_geometry.py sha256 1184 16 "1,32,100000,2" "1,32,0,2" "1,64,100000,2" "2,32,100000,2" "2,64,100000,2" "1,96,100000,2" > gpurun_out/b_tune_sha256.txt 2>&1
timeout 300 python tools/tune_geometry.py aes_128 4096 16 "2,96,100000,4" "2,64,100000,4" "2,128,100000,4" "1,96,100000,4" > gpurun_out/b_tune_aes.txt 2>&1
for K in garble eval; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 3 -c 1 -f -o /tmp/prof/aes_$K python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
  ncu -i /tmp/prof/aes_$K.ncu-rep --page raw --csv > gpurun_out/b_aes128_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/aes_$K.ncu-rep --page source --csv > gpurun_out/b_aes128_${K}_src.csv 2>/dev/null
done
ls -la gpurun_out | grep " b_"
tail -3 gpurun_out/b_pytest.log; tail -3 gpurun_out/b_bench.err; cat gpurun_out/b_microbench_bs.log | tail -12; cat gpurun_out/b_tune_sha256.txt
