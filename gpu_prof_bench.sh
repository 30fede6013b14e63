#!/bin/bash
# profiles for the round: launch list of bench.py + full-set capture of the data kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 8 -c 2 -o gpurun_out/prof_bench -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full.log
