#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-second-leg"
timeout 600 $B --shape-factor 2 > gpurun_out/r02_bench_n1_tsc_f32.json 2>/dev/null
timeout 600 $B --shape-factor 2 --deposition j_from_rhov --current-filter bilinear > gpurun_out/r02_bench_n1_tsc_direct_bilinear_f32.json 2>/dev/null
timeout 600 $B --shape-factor 1 --deposition j_from_rhov --current-filter bilinear > gpurun_out/r02_bench_n1_cic_direct_bilinear_f32.json 2>/dev/null
for f in r02_bench_n1_tsc_f32 r02_bench_n1_tsc_direct_bilinear_f32 r02_bench_n1_cic_direct_bilinear_f32; do tail -1 gpurun_out/$f.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$f', 'ms', round(d['ms_per_step'],2), 'value', '%.3g'%d['value'], 'K1', r['avg_launch_ms_by_species'] if r else None, 'frac', round(r['frac'],3) if r else None, d['k1_variant'], 'check', (d.get('check') or {}).get('gauss_drift_relative'))"; done
