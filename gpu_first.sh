#!/bin/bash
# first GPU session: microbench + parity tests + quick timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
./tools/ubench_fp32x2 > gpurun_out/ubench.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
python tools/quick_bench.py --config c2 --frames 64 > gpurun_out/qb_c2.txt 2>&1
python tools/quick_bench.py --config c4 --frames 32 > gpurun_out/qb_c4.txt 2>&1
python tools/quick_bench.py --config c3 --frames 64 > gpurun_out/qb_c3.txt 2>&1
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/ubench.txt; cat gpurun_out/qb_c2.txt
