#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py -m gpu -x -q 2>&1 | tail -2
for c in c1 c5; do python tools/quick_bench.py --config $c --frames 16384 --iters 3 2>&1 | tail -1; done
python tools/quick_bench.py --config c2 --frames 256 --iters 3 2>&1 | tail -1
