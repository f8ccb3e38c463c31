"""Turn an .ncu-rep (ncu --set full --import-source on) into the markdown summary committed here.
usage: python profiles/summarize.py gpurun_out/<rep>.ncu-rep <kernel label> > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

rep, label = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, v = rows[0], rows[1], rows[2]
m = {name: (v[i], units[i]) for i, name in enumerate(h)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print("# ncu summary: %s\n\nsource: `%s` (ncu --set full --clock-control none --import-source on; replayed, cold caches:\ncompare shares, not absolutes)\n" % (label, rep))
print("| metric | value | unit |\n|---|---|---|")
for k in want:
    if k in m:
        print("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
stalls = sorted(((float(val[0]), k) for k, val in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
print("\n## warp states per issue-active cycle (top)\n\n| stall | warps |\n|---|---|")
for val, k in stalls[:8]:
    print("| %s | %.3f |" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), val))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
hdr, data = None, []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Line No":
        hdr = r
        wi, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= wi or not r[0]:
        continue
    try:
        data.append((int(r[ie]), int(r[wi]), int(r[0]), r[1].strip()[:110]))
    except ValueError:
        pass
ti, ts = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
print("\n## hottest source lines (by stall samples)\n\n| %inst | %stall | line | source |\n|---|---|---|---|")
for d in sorted(data, key=lambda d: -d[1])[:16]:
    print("| %.1f | %.1f | %d | `%s` |" % (100 * d[0] / ti, 100 * d[1] / ts, d[2], d[3].replace("|", "\\|")))
