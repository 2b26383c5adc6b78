#!/bin/bash
# Round-2 GPU pass on TWO devices: the multi-device tests, the in-library fan-out, a 2-rank bench line.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/w_gpus.txt; nvidia-smi topo -m >> gpurun_out/g_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/w_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest_multi.log
timeout 400 python tools/fanout_bench.py > gpurun_out/w_fanout.jsonl 2> gpurun_out/w_fanout.err; echo "fanout rc=$?" >> gpurun_out/w_fanout.err
GCB_E2E_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/w_bench_2gpu.json 2> gpurun_out/w_bench_2gpu.err; echo "bench rc=$?" >> gpurun_out/w_bench_2gpu.err
tail -4 gpurun_out/w_pytest_multi.log; cat gpurun_out/w_fanout.jsonl; tail -3 gpurun_out/w_fanout.err; tail -5 gpurun_out/w_bench_2gpu.err; head -c 400 gpurun_out/w_bench_2gpu.json
