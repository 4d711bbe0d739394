#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_distributed.py -k "not 2x2x2 and not mesh2" 2>&1 | tail -4
