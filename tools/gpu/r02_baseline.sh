#!/bin/bash
# round-2 first GPU session: micro-benchmarks, the GPU test suite, baseline timings and ncu captures of c3/c4
mkdir -p gpurun_out tools/bin
cd "$(dirname "$0")/../.."
./tools/bin/ubench_lsu > gpurun_out/ubench_lsu.txt 2>&1; echo "ubench rc=$?"; tail -40 gpurun_out/ubench_lsu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.txt
for c in c2 c3 c4; do
  F=256; [ $c = c3 ] && F=384; [ $c = c4 ] && F=192
  python tools/quick_bench.py --config $c --frames $F --iters 6 > gpurun_out/qb_$c.txt 2>&1; tail -2 gpurun_out/qb_$c.txt
done
for c in c3 c4; do
  F=384; [ $c = c4 ] && F=192
  CMD="python tools/quick_bench.py --config $c --frames $F --iters 2"
  $CMD > gpurun_out/plain_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 2 -c 2 -o gpurun_out/r02_before_$c -f $CMD > gpurun_out/ncu_$c.log 2>&1
  echo "ncu $c rc=$?"; tail -2 gpurun_out/ncu_$c.log
done
