#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
python tools/quick_bench.py --dims 64,1024,63,101,4 --frames 128 --iters 4 2>&1 | tail -1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ("value","ms_per_step","roofline","e2e","clocks","sustained")})[:1800])
PY
