#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -4
for c in c1 c5; do python tools/quick_bench.py --config $c --frames 16384 --iters 3 2>&1 | tail -1; done
