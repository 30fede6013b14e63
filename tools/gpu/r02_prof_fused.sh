#!/bin/bash
# ncu --set full of the single-launch frames kernel (default lead, interleaved) for c4, 192 frames
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
CMD="python tools/quick_bench.py --config c4 --frames 192 --iters 2"
$CMD > gpurun_out/plain_fused.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lsmrc_frames -s 1 -c 1 -o gpurun_out/r02_fused_c4 -f $CMD > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fused.log
LSMRC_FUSED_LEAD=100000 ncu --set full --clock-control none --import-source on -k regex:lsmrc_frames -s 1 -c 1 -o gpurun_out/r02_fusedseq_c4 -f $CMD > gpurun_out/ncu_fusedseq.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fusedseq.log
