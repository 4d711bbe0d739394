#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python tools/ab.py run v11e str1 str4 ev str1ev str16ev -- --steps 20 --warmup 5 2>&1 | tee gpurun_out/r02_v11f_run.log
for v in str1ev; do PIC_B200_LIB=build_ab/$v/libpic_b200.so timeout 600 python -m pytest -q -x -m gpu tests/test_gpu_parity.py -k "tile or supercell or resident" 2>&1 | tail -3 | tee -a gpurun_out/r02_v11f_run.log; done
