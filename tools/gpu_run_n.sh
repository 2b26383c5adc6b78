#!/bin/bash
# Round-2 GPU pass N (1 GPU): split live ranges as the default plan for batches that overflow narrow circuits:
# whole GPU test suite, timings with the library's own choice, ncu of the sha256 x 2368 kernels, the default bench line.
set -u
mkdir -p gpurun_out /tmp/prof
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
{
timeout 200 python tools/time_circuit.py sha256 1184
timeout 200 python tools/time_circuit.py sha256 1300
timeout 200 python tools/time_circuit.py sha256 2368
timeout 200 python tools/time_circuit.py sha256 4736
timeout 200 python tools/time_circuit.py sha512 2368
timeout 200 python tools/time_circuit.py sha256xor 2368 32
timeout 200 python tools/time_circuit.py chacha20block 2368
timeout 200 python tools/time_circuit.py mul64 4736
timeout 200 python tools/time_circuit.py aes_128 4096
} > gpurun_out/n_times.txt 2>&1
for K in garble eval; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/sha_$K python tools/run_one.py sha256 2368 2 > /dev/null 2>&1
  ncu -i /tmp/prof/sha_$K.ncu-rep --page raw --csv > gpurun_out/n_sha256_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/sha_$K.ncu-rep --page source --csv > gpurun_out/n_sha256_${K}_src.csv 2>/dev/null
done
timeout 900 python bench.py --steps 30 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?" >> gpurun_out/n_bench.err
tail -5 gpurun_out/n_pytest.log; cat gpurun_out/n_times.txt; tail -3 gpurun_out/n_bench.err
