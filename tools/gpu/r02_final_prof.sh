#!/bin/bash
# closing profiling session of round 2 (one GPU): ncu --set full of the shipped kernels of c2 / c3 / c4 (each after the
# plain run of the same command), then the launch list of the bench command
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
for c in c2 c3 c4; do
  F=256; S=2; N=2; [ $c = c3 ] && F=384 && S=1 && N=1; [ $c = c4 ] && F=192 && S=1 && N=1
  CMD="python tools/quick_bench.py --config $c --frames $F --iters 2"
  $CMD > gpurun_out/plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:lsmrc_ -s $S -c $N -o gpurun_out/r02_final_$c -f $CMD > gpurun_out/ncu_$c.log 2>&1
  echo "ncu $c rc=$?"; tail -1 gpurun_out/ncu_$c.log
done
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
