#!/bin/bash
# Round-2 GPU pass R (1 GPU): small batches -- spread over the SMs (new default) against filling a few SMs, and wider
# teams for batches below one instance per team slot; e2e with the spread.
set -u
mkdir -p gpurun_out
{
for B in 1 64 148 256 512 1024; do
  timeout 200 python tools/time_circuit.py aes_128 $B
  GCB_SPREAD=0 timeout 200 python tools/time_circuit.py aes_128 $B
done
for TT in 96 128 192 256; do GCB_TEAMS=2 GCB_TEAM_THREADS=$TT timeout 200 python tools/time_circuit.py aes_128 256; done
for TT in 96 128 256 480; do GCB_TEAMS=1 GCB_TEAM_THREADS=$TT timeout 200 python tools/time_circuit.py aes_128 148; done
for TT in 64 128 256; do GCB_NT=2 GCB_TEAMS=2 GCB_TEAM_THREADS=$TT timeout 200 python tools/time_circuit.py aes_128 256; done
for B in 1 148 256 592; do
  timeout 200 python tools/time_circuit.py sha256 $B
  GCB_SPREAD=0 timeout 200 python tools/time_circuit.py sha256 $B
done
for TT in 32 64 96; do GCB_TEAMS=1 GCB_TEAM_THREADS=$TT timeout 200 python tools/time_circuit.py sha256 148; done
for TT in 32 64 96; do GCB_TEAMS=1 GCB_TEAM_THREADS=$TT timeout 200 python tools/time_circuit.py sha512 148; done
} > gpurun_out/r_times.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_garble.py tests/test_gpu_stream.py -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r_pytest.log
GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?" >> gpurun_out/r_bench.err
cat gpurun_out/r_times.txt; tail -3 gpurun_out/r_pytest.log; tail -4 gpurun_out/r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r_bench.json'))
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='how'})
print(d['extra']['latency_batch1'], d['extra']['stream_program']['m_gates_per_s'], d['extra']['stream_sha256_step'])
PY
