#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
FRAMES=256 ./gpu_variants.sh 2>&1 | grep -E "==|iter 3"
FRAMES=64 ./gpu_variants.sh 2>&1 | grep -E "==|iter 3"
