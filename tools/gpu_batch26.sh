#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -1 gpurun_out/r02_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1 ms', d['ms_per_step'], 'value', '%.4g'%d['value'], 'frac', d['roofline']['frac'], 'K1', d['roofline']['avg_launch_ms_by_species']); print('e2e', d['e2e']); print('cpu', d['cpu_baseline']); print('check', d['check']['continuity_relative'], d['check']['gauss_drift_relative']); print('f64', d['legs']['f64']['value'], d['legs']['f64']['ms_per_step'], d['legs']['f64']['roofline']['frac'], d['legs']['f64']['check']['continuity_relative']); print(d['legs']['f32']['phases_ms'])"
