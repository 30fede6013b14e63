#!/bin/bash
for cfg in c2 c3 c4; do
for v in 0 4096; do
echo "== $cfg LSMRC_H2D_STRIP_MIN_ROW=$v"
LSMRC_H2D_STRIP_MIN_ROW=$v python bench.py --config $cfg --no-cpu-baseline --no-extras --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(json.dumps(d['e2e']))"
done; done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
