#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for v in gpu-accel-ofdm-ls-mrc_b200/variants/lib_*.so; do
  echo "== $v"
  LSMRC_LIB=$v python tools/quick_bench.py --config c3 --frames 128 --iters 4 2>&1 | tail -1
  LSMRC_LIB=$v python tools/quick_bench.py --config c4 --frames 32 --iters 4 2>&1 | tail -1
  LSMRC_LIB=$v python tools/quick_bench.py --config c2 --frames 256 --iters 4 2>&1 | tail -1
done
