"""Dev-time: the output-format kernels alone (bench.py's `also_measured.output_kernels`), for ncu captures."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench

peak, _ = bench.measured_peak()
print(json.dumps(bench.output_kernels(bench.WORKLOADS["1080p420_ipb"], 0, peak), indent=1))
