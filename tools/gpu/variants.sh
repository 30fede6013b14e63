#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 100 > gpurun_out/clk.txt &
SMI=$!
for v in gpu-accel-ofdm-ls-mrc_b200/variants/lib_*.so; do
  echo "== $v"
  LSMRC_LIB=$v python tools/quick_bench.py --config c2 --frames ${FRAMES:-128} --iters ${ITERS:-4} 2>&1 | tail -${TAIL:-3}
done 2>&1 | tee gpurun_out/variants.txt
kill $SMI
sort -n gpurun_out/clk.txt | tail -3
