#!/bin/bash
# Round-2 GPU pass K (1 GPU): e2e against the pattern-matched copy floor.
set -u
mkdir -p gpurun_out
GCB_E2E_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?" >> gpurun_out/k_bench.err
tail -4 gpurun_out/k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/k_bench.json'))
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='how'})
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d['extra']['pcie_probe'].items() if k!='how'})
PY
