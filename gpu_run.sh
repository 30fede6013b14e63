#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_bench.py --config c4 --frames 48 --iters 4 2>&1 | tail -1
python tools/quick_bench.py --config c4 --frames 192 --iters 4 2>&1 | tail -1
python tools/quick_bench.py --config c3 --frames 384 --iters 4 2>&1 | tail -1
python tools/quick_bench.py --config c5 --frames 1 --iters 4 2>&1 | tail -1
python tools/quick_bench.py --config c2 --frames 256 --iters 4 2>&1 | tail -1
