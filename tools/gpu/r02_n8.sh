#!/bin/bash
# the bench line on all GPUs of the box, as the driver launches it
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
python tools/show_bench.py gpurun_out/bench_n$N.json | cut -c1-1200
