#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/latency_probe.py --config c5 --launches 100"
$CMD > gpurun_out/plain_lat.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_kernel -s 598 -c 2 -o gpurun_out/prof_lat -f $CMD > gpurun_out/ncu_lat.log 2>&1
echo "rc=$?"; gpu-accel-ofdm-ls-mrc_b200/host/bin/latency_main --launches 3000; tail -3 gpurun_out/ncu_lat.log
