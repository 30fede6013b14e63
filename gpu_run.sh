#!/bin/bash
python -m pytest tests/test_gpu_golden_and_host.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -12
