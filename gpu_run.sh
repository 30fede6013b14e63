#!/bin/bash
python -m pytest tests/test_soft_demap.py -m gpu -x -q 2>&1 | grep -v "^$" | tail -12
