"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        li, si = 0, 1
        wi = hdr.index("Warp Stall Sampling (All Samples)")
        ie = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= wi or not r[li]:
        continue
    try:
        data.append((int(r[wi]), int(r[ie]), int(r[li]), r[si][:120]))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "total warp-instructions", sum(d[1] for d in data))
for d in sorted(data, reverse=True)[:top]:
    print("%7d %5.1f%% inst=%9d L%4d %s" % (d[0], 100 * d[0] / tot, d[1], d[2], d[3]))
