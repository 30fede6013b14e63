#!/bin/bash
# parity of the big-FFT paths + timings of c2/c3/c4 (development loop); extra libraries under variants/ are timed too
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_golden_and_host.py -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.txt
for c in ${CONFIGS:-c2 c3 c4}; do
  F=256; [ $c = c3 ] && F=384; [ $c = c4 ] && F=192
  python tools/quick_bench.py --config $c --frames $F --iters 6 > gpurun_out/qb_$c.txt 2>&1; tail -2 gpurun_out/qb_$c.txt
  for v in gpu-accel-ofdm-ls-mrc_b200/variants/lib_*.so; do
    [ -e "$v" ] || continue
    echo "== $v"; LSMRC_LIB=$v python tools/quick_bench.py --config $c --frames $F --iters 5 2>&1 | tail -2
  done
done
