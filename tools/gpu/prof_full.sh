#!/bin/bash
# full-set capture of the pilot and data kernels of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 8 -c 2 -o gpurun_out/prof_bench -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full.log
