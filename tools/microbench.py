"""Runs the design-evidence microbenchmarks of csrc/microbench.cu on cuda:0 and prints one JSON line per case."""
import ctypes
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pypic3d_b200 import _lib

NAMES = {0: "smem atomicAdd f32, spread", 1: "smem atomicAdd f32, 8-lane runs", 2: "smem atomicAdd f32, warp-uniform",
         3: "smem atomicAdd f64, spread", 4: "smem atomicAdd f64, 8-lane runs", 5: "smem atomicAdd f64, warp-uniform",
         6: "gmem RED f32, spread", 7: "gmem RED f32, 8-lane runs", 8: "gmem RED f32, warp-uniform",
         9: "gmem RED f64, spread", 10: "gmem RED f64, 8-lane runs", 11: "gmem RED f64, warp-uniform",
         12: "shfl-reduce 8 lanes then gmem RED f32", 13: "L1 gather f32 (ldg), 8-lane runs"}

if __name__ == "__main__":
    torch.cuda.init()
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    lane_ops = sms * 16 * 256 * 64
    iters = 20
    for which, name in NAMES.items():
        ms = ctypes.c_float(0)
        rc = L.pic_microbench(which, iters, ctypes.byref(ms))
        per_launch_ms = ms.value / iters
        print(json.dumps({"case": which, "name": name, "rc": rc, "ms_per_launch": per_launch_ms,
                          "lane_ops_per_s": lane_ops / (per_launch_ms * 1e-3) if per_launch_ms > 0 else None,
                          "lane_ops_per_clk_per_sm_at_1.9GHz": lane_ops / (per_launch_ms * 1e-3) / sms / 1.9e9 if per_launch_ms > 0 else None}))
