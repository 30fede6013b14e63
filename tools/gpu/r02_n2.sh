#!/bin/bash
# two GPUs: the multi-GPU tests (skipped on one-GPU boxes) and the bench line at N=2
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gpus or multi or two" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n2.err
python tools/show_bench.py gpurun_out/bench_n2.json | cut -c1-1500
