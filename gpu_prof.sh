#!/bin/bash
mkdir -p gpurun_out
export LSMRC_LIB=gpu-accel-ofdm-ls-mrc_b200/variants/lib_xtma0.so
CMD="python tools/quick_bench.py --config c2 --frames 256 --iters 2"
$CMD > gpurun_out/plainx.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 3 -c 1 -o gpurun_out/profx -f $CMD > gpurun_out/ncux.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/plainx.log; tail -2 gpurun_out/ncux.log
