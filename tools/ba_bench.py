import sys, time, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'slam-eds_b200'))
import torch, edsgpu, bench
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = edsgpu.Context(0, stream.cuda_stream)
print(bench.bench_ba(ctx, stream, reps=int(sys.argv[1]) if len(sys.argv) > 1 else 50))
