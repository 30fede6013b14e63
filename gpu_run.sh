#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -2 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['frontend']); print(d['latency']['p50_us'], d['roofline']['frac'])"
