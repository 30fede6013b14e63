#!/bin/bash
for v in gpu-accel-ofdm-ls-mrc_b200/variants/lib_*.so; do
  echo "== $v"
  LSMRC_LIB=$v python tools/quick_bench.py --config c3 --frames 384 --iters 4 2>&1 | tail -1
  LSMRC_LIB=$v python tools/quick_bench.py --config c4 --frames 192 --iters 4 2>&1 | tail -1
done
