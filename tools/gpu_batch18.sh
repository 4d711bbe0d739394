#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest -q -x -s -m gpu tests/test_frontend.py -k "weibel_growth" 2>&1 | grep -v "^$" | tail -12 | cut -c1-1500 | tee gpurun_out/r02_weibel_test.log
timeout 600 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "float32" 2>&1 | tail -4 | tee -a gpurun_out/r02_weibel_test.log
