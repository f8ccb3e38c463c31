"""Developer tool: the coarse-tracker side line of bench.py on its own (also the workload of its ncu capture)."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-eds_b200"))
import bench  # noqa: E402
import edsgpu  # noqa: E402

s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx = edsgpu.Context(0, s.cuda_stream)
print(json.dumps(bench.bench_coarse(ctx, s, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 5), indent=1))
