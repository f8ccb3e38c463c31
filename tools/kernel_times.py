"""Developer tool: warm per-kernel durations (CUPTI through torch.profiler, no replay, caches as the workload leaves them) of
one of the side-line workloads: `python tools/kernel_times.py ba|coarse [reps]`.  ncu's launch lists are cold-cache and
serialised; this is the number to compare a kernel's share of a host-synchronous call against."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-eds_b200"))
import bench  # noqa: E402
import edsgpu  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "ba"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
ctx = edsgpu.Context(0, s.cuda_stream)
fn = {"ba": bench.bench_ba, "coarse": bench.bench_coarse}[what]
fn(ctx, s, reps=2)  # warm-up outside the trace
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn(ctx, s, reps=reps)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if e.device_time_total > 0]
for name, count, us in sorted(rows, key=lambda r: -r[1] * r[2]):
    print("%8.2f us x %5d  %s" % (us, count, name[:110]))
