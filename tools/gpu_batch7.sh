#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_golden.py -k "tile or variant or golden or supercell or resident" 2>&1 | tail -5 | tee gpurun_out/r02_v11d_test.log
timeout 900 python tools/ab.py run v11d v11d@PIC_K9_GROUPRED=0 -- --steps 20 --warmup 5 2>&1 | tee gpurun_out/r02_v11e_run.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pair -c 44 --csv --log-file gpurun_out/r02_launches_v11d.csv python bench.py --steps 10 --warmup 1 --no-cpu-baseline --no-e2e --no-check --no-second-leg > gpurun_out/r02_launches_v11d.log 2>&1
