#!/bin/bash
# ring streaming (c3) on one GPU: producer threads sweep, and the consumer's ceiling with a producer that only publishes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nproc
timeout 900 python -m pytest tests/test_gpu_golden_and_host.py -m gpu -x -q 2>&1 | tail -3
python tools/ring_probe.py 8x 4 8 12 16 8 2>&1 | tee gpurun_out/ring_sweep.txt
