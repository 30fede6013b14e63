#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -3
python tools/quick_bench.py --dims 32,256,32,16,4 --frames 2048 --iters 4 2>&1 | tail -1
python tools/quick_bench.py --dims 32,512,32,16,4 --frames 1024 --iters 4 2>&1 | tail -1
