#!/bin/bash
# Round-2 GPU pass T (1 GPU): parity, timings, then the profile set of the final kernels (launch list, ncu --set full of the headline and sha256 kernels).
set -u
mkdir -p gpurun_out /tmp/prof
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_pytest.log
{
timeout 200 python tools/time_circuit.py aes_128 4096
timeout 200 python tools/time_circuit.py aes_128 4096 32
timeout 200 python tools/time_circuit.py aes_256 4096 32
timeout 200 python tools/time_circuit.py sha256 1184
timeout 200 python tools/time_circuit.py sha256 2368
timeout 200 python tools/time_circuit.py sha256 4736
timeout 200 python tools/time_circuit.py sha512 2368
timeout 200 python tools/time_circuit.py sha256xor 2368 32
timeout 200 python tools/time_circuit.py chacha20block 2368
timeout 200 python tools/time_circuit.py mul64 4736
timeout 200 python tools/time_circuit.py aes_128 148
timeout 200 python tools/time_circuit.py aes_128 1
} > gpurun_out/t_times.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t_launches.csv $B > /dev/null 2>&1
for K in garble eval; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 3 -c 1 -f -o /tmp/prof/aes_$K $B > /dev/null 2>&1
  ncu -i /tmp/prof/aes_$K.ncu-rep --page raw --csv > gpurun_out/t_aes128_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/aes_$K.ncu-rep --page source --csv > gpurun_out/t_aes128_${K}_src.csv 2>/dev/null
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/sha_$K python tools/run_one.py sha256 2368 2 > /dev/null 2>&1
  ncu -i /tmp/prof/sha_$K.ncu-rep --page raw --csv > gpurun_out/t_sha256_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/sha_$K.ncu-rep --page source --csv > gpurun_out/t_sha256_${K}_src.csv 2>/dev/null
done
tail -4 gpurun_out/t_pytest.log; cat gpurun_out/t_times.txt; head -6 gpurun_out/t_launches.csv | cut -c1-200
