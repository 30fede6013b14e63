#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^$" | tail -2
