#!/bin/bash
# Round-2 GPU pass X (1 GPU): new small-batch tests, whole suite once more.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
tail -4 gpurun_out/x_pytest.log
