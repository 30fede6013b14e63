#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/quick_bench.py --config c1 --frames 16384 --iters 2"
$CMD > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 3 -c 1 -o gpurun_out/prof1 -f $CMD > gpurun_out/ncu1.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain1.log; tail -2 gpurun_out/ncu1.log
