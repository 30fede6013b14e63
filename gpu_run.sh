#!/bin/bash
python -m pytest tests/test_gpu_golden_and_host.py -m gpu -x -q -k "ring or stream" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['ring_stream'])"
