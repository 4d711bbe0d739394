#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_golden.py tests/test_frontend.py tests/test_gpu_electrostatic.py 2>&1 | tail -6 | tee gpurun_out/r02_graph_test.log
timeout 600 python tools/demo_bench.py --steps 400 --oracle-steps 2 > gpurun_out/r02_demo_bench_graph.json 2> gpurun_out/r02_demo_bench_graph.err; tail -3 gpurun_out/r02_demo_bench_graph.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['workload'], 'dev ms', round(d['ms_per_step_device'],4), 'wall', round(d['ms_per_step_wall'],4), 'x oracle', round(d['speedup_vs_oracle'],1))
    except Exception as e: print(l[:200])"
PIC_GRAPH=0 timeout 600 python tools/demo_bench.py --steps 400 --oracle-steps 2 2>&1 >/dev/null | tail -3 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('nograph', d['workload'], 'dev ms', round(d['ms_per_step_device'],4))
    except Exception as e: print(l[:200])"
