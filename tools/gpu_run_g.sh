#!/bin/bash
# Round-2 GPU pass G (1 GPU): parity after the on-demand many-instance plan, launch list, profile of the 16-team hot/cold sha256 kernels.
set -u
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g1_pytest.log
{
timeout 200 python tools/time_circuit.py sha512 1184
timeout 200 python tools/time_circuit.py sha512 444
timeout 200 python tools/time_circuit.py sha256 1184
} > gpurun_out/g1_times.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/g1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
for K in garble eval; do
  GCB_HOT_TEAMS=16 timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/hc_$K python tools/run_one.py sha256 2368 2 > /dev/null 2>&1
  ncu -i /tmp/prof/hc_$K.ncu-rep --page raw --csv > gpurun_out/g1_sha256hc_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/hc_$K.ncu-rep --page source --csv > gpurun_out/g1_sha256hc_${K}_src.csv 2>/dev/null
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/s8_$K python tools/run_one.py sha256 1184 2 > /dev/null 2>&1
  ncu -i /tmp/prof/s8_$K.ncu-rep --page raw --csv > gpurun_out/g1_sha256_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/s8_$K.ncu-rep --page source --csv > gpurun_out/g1_sha256_${K}_src.csv 2>/dev/null
done
tail -3 gpurun_out/g1_pytest.log; cat gpurun_out/g1_times.txt; head -5 gpurun_out/g1_launches.csv
