#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_stream.py tests/test_gpu_fullsize.py::test_stream_program_over_1e8_gates -x -q > gpurun_out/x2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x2_pytest.log
timeout 15 python tools/fuzz_stream_eval.py 40 > gpurun_out/x2_fuzz_plain.txt 2>&1; echo "rc=$?" >> gpurun_out/x2_fuzz_plain.txt
tail -3 gpurun_out/x2_pytest.log; tail -4 gpurun_out/x2_fuzz_plain.txt
