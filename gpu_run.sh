#!/bin/bash
for v in base twrec base twrec; do
  echo "== $v"; LSMRC_LIB=gpu-accel-ofdm-ls-mrc_b200/variants/lib_$v.so timeout 200 python tools/quick_bench.py --config c2 --frames 256 --iters 120 2>&1 | awk 'NR==6||NR==60||NR==120' | cut -c1-90
done
LSMRC_LIB=gpu-accel-ofdm-ls-mrc_b200/variants/lib_twrec.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c2 or 1024 or odd_ant or one_ant or syms_not" 2>&1 | tail -2
