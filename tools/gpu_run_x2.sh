#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/x2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x2_pytest.log
timeout 45 python tools/fuzz_stream_eval.py 100 > gpurun_out/x2_fuzz_plain.txt 2>&1; echo "rc=$?" >> gpurun_out/x2_fuzz_plain.txt
tail -3 gpurun_out/x2_pytest.log; tail -8 gpurun_out/x2_fuzz_plain.txt
