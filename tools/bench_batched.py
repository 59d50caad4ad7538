"""Batched 32x32 DGETRF timing only (the `batched_dgetrf_32x32` leg of bench.py)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lapack_b200 as lb
import bench
print(json.dumps(bench.bench_batched(lb, torch, torch.device("cuda", 0), int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20)))
