#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_golden.py -k "resident or tile or golden or supercell" 2>&1 | tail -2 | tee gpurun_out/r02_sortruns.log
timeout 200 python tools/ab.py run sortruns sort0 -- --steps 20 --warmup 5 2>&1 | tee -a gpurun_out/r02_sortruns.log
