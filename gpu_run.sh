#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -2
