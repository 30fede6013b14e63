#!/bin/bash
# full GPU test suite, then the default bench line (one GPU)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.txt
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python tools/show_bench.py gpurun_out/bench.json
