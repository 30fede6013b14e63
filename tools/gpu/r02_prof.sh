#!/bin/bash
# one ncu --set full capture of the kernels of a config (after the plain run of the same command)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
c=${1:-c4}; tag=${2:-r02}
F=256; [ $c = c3 ] && F=384; [ $c = c4 ] && F=192
CMD="python tools/quick_bench.py --config $c --frames $F --iters 2"
$CMD > gpurun_out/plain_$c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_ -s 2 -c 2 -o gpurun_out/${tag}_$c -f $CMD > gpurun_out/ncu_$c.log 2>&1
echo "ncu $c rc=$?"; tail -2 gpurun_out/ncu_$c.log
