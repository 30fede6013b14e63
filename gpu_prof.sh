#!/bin/bash
# ncu capture of the data kernel (plain run first, as the profiling recipe requires)
mkdir -p gpurun_out
CMD="python tools/quick_bench.py --config c2 --frames 64 --iters 2"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 2 -c 2 -o gpurun_out/prof -f $CMD > gpurun_out/ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu.log
