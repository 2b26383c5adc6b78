#!/bin/bash
# Round-2 GPU pass S (1 GPU): wider teams for small batches, kernel streams per job class (e2e), whole GPU suite.
set -u
mkdir -p gpurun_out
{
for B in 1 64 148 256 512 900 1024; do timeout 200 python tools/time_circuit.py aes_128 $B; done
for B in 1 148 256 592 1000; do timeout 200 python tools/time_circuit.py sha256 $B; done
for B in 148 300; do timeout 200 python tools/time_circuit.py sha512 $B; done
timeout 200 python tools/time_circuit.py mul64 300
} > gpurun_out/s_times.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
for A in 3 16; do
GCB_E2E_AHEAD=$A GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/s_bench_a$A.json 2>> gpurun_out/s_bench.err
GCB_KERNEL_STREAMS=shared GCB_E2E_AHEAD=$A GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/s_bench_shared_a$A.json 2>> gpurun_out/s_bench.err
done
GCB_E2E_PARTS=8 GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/s_bench_p8.json 2>> gpurun_out/s_bench.err
GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench.json 2>> gpurun_out/s_bench.err; echo "bench rc=$?" >> gpurun_out/s_bench.err
cat gpurun_out/s_times.txt; tail -3 gpurun_out/s_pytest.log; cat gpurun_out/s_bench.err
python - <<'PY'
import json
for f in ['s_bench_a3','s_bench_shared_a3','s_bench_a16','s_bench_shared_a16','s_bench_p8','s_bench']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='how'})
d=json.load(open('gpurun_out/s_bench.json'))
print(d['extra']['latency_batch1'], d['extra']['stream_program']['m_gates_per_s'], d['extra']['stream_sha256_step'])
PY
