#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/quick_bench.py --config c5 --frames 1 --iters 3"
$CMD > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 4 -c 2 -o gpurun_out/prof5 -f $CMD > gpurun_out/ncu5.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu5.log
