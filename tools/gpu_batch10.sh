#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_golden.py -k "yee or resident or golden or step or electrodynamic" 2>&1 | tail -8 | tee gpurun_out/r02_ystream_test.log
timeout 900 python tools/ab.py run ystream ytile -- --steps 20 --warmup 5 2>&1 | tee gpurun_out/r02_ystream_run.log
