#!/bin/bash
# single-launch kernel of the 2048/4096-point plans: parity (under a short timeout), then A/B timing against the kernel pair
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_launch" 2>&1 | tail -3
timeout 120 python tools/fused_probe.py 32,4096,288,4,6,150 10 2>&1 | tail -2
for c in c4 c3; do
  for fu in 1 0 1 0; do
    echo "== $c one_launch=$fu"
    LSMRC_ONE_LAUNCH=$fu timeout 120 python tools/quick_bench.py --config $c --frames $([ $c = c4 ] && echo 192 || echo 384) --iters 5 2>&1 | tail -1
  done
done
