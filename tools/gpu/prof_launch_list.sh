#!/bin/bash
# launch list of bench.py (plain run first, as the profiling recipe requires)
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -c 400 gpurun_out/bench_plain.json
