#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -12
gpu-accel-ofdm-ls-mrc_b200/host/bin/latency_main --launches 3000
