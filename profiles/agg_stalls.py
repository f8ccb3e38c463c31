"""Aggregate the ncu source page (--import-source on) of a kernel by CUDA source line: instructions executed, stall samples
and the stall reasons per line, plus totals per line range.  usage: python profiles/agg_stalls.py rep [top_n] [lo-hi:name ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ranges = []
for a in sys.argv[3:]:
    r, name = a.split(":")
    lo, hi = r.split("-")
    ranges.append((int(lo), int(hi), name))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Line No" and "# Samples" in r:
        hdr = r
        ix = {c: i for i, c in enumerate(hdr)}
        stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0]); n = int(r[ix["# Samples"]]); ie = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    data.append((line, r[1].strip()[:90], n, ie, {c: int(r[ix[c]] or 0) for c in stalls}))
ts = sum(d[2] for d in data) or 1
ti = sum(d[3] for d in data) or 1
tot = {}
for d in data:
    for k, v in d[4].items():
        tot[k] = tot.get(k, 0) + v
print("samples %d, instructions %d" % (ts, ti))
print("stall totals:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / ts) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
for lo, hi, name in ranges:
    sel = [d for d in data if lo <= d[0] <= hi]
    s = sum(d[2] for d in sel); i = sum(d[3] for d in sel)
    t = {}
    for d in sel:
        for k, v in d[4].items():
            t[k] = t.get(k, 0) + v
    print("%-22s inst %5.1f%% samples %5.1f%% | %s" % (name, 100.0 * i / ti, 100.0 * s / ts, ", ".join("%s %.1f" % (k.replace("stall_", ""), 100.0 * v / ts) for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:5])))
print("\nhottest lines by samples")
for d in sorted(data, key=lambda d: -d[2])[:top_n]:
    top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in sorted(d[4].items(), key=lambda kv: -kv[1])[:3] if v)
    print("%5.1f%% smp %5.1f%% inst  L%-5d %-90s | %s" % (100.0 * d[2] / ts, 100.0 * d[3] / ti, d[0], d[1], top))
