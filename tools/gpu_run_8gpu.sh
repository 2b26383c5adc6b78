#!/bin/bash
# Round-2 sanity pass on EIGHT devices: the bench line with every extra at N = 8 (what the driver's scaling run executes).
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/y_gpus.txt; free -g >> gpurun_out/y_gpus.txt; nproc >> gpurun_out/y_gpus.txt
GCB_E2E_TRACE=1 timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/y_bench_8gpu.json 2> gpurun_out/y_bench_8gpu.err; echo "bench rc=$?" >> gpurun_out/y_bench_8gpu.err
tail -5 gpurun_out/y_bench_8gpu.err; cat gpurun_out/y_gpus.txt; tail -c 1500 gpurun_out/y_bench_8gpu.json
