#!/bin/bash
python -m pytest tests/test_zero_forcing.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -5
