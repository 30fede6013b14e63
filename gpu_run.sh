#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_bench.py --config c2 --frames 256 --iters 4 2>&1 | tail -1
