#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/dbg_check.py 128 8 f32 2>&1 | grep -v "^\*\|OMP\|NCCL" | tee gpurun_out/r02_dbg_check4.log
