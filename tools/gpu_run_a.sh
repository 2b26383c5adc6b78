#!/bin/bash
# Round-2 GPU pass A: parity tests, bench (both arms), racecheck artefacts, ncu of the sha256 kernels.
set -u
mkdir -p gpurun_out /tmp/prof
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1; nproc >> gpurun_out/a_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2>> gpurun_out/a_bench.err
for C in "aes_128 10" "sha256 16" "mul64 40"; do
  set -- $C
  timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python tools/run_one.py $1 $2 1 > gpurun_out/a_race_$1.txt 2>&1
  tail -3 gpurun_out/a_race_$1.txt > gpurun_out/a_race_$1.tail
done
for K in garble eval; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${K}_kernel -s 1 -c 1 -f -o /tmp/prof/sha_$K python tools/run_one.py sha256 1184 2 > /dev/null 2>&1
  ncu -i /tmp/prof/sha_$K.ncu-rep --page raw --csv > gpurun_out/a_sha256_${K}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/sha_$K.ncu-rep --page source --csv > gpurun_out/a_sha256_${K}_src.csv 2>/dev/null
done
ls -la gpurun_out | grep " a_"
tail -5 gpurun_out/a_pytest.log; tail -3 gpurun_out/a_bench.err; head -c 600 gpurun_out/a_bench.json
