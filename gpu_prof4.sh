#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/quick_bench.py --config c4 --frames 192 --iters 2"
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 2 -c 2 -o gpurun_out/prof4 -f $CMD > gpurun_out/ncu4.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain4.log
