"""Summarise an .ncu-rep (raw page + hot source lines) into a small text file for profiles/."""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor"]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]} (one column per captured launch)\n")
    for i, h in enumerate(hdr):
        if any(h == k or h.startswith(k) for k in KEYS) and "peak_sustained_elapsed.per" not in h:
            if all(r[i] in ("0", "0.000000", "") for r in rows[2:]):
                continue  # a pipe this kernel does not touch
            f.write(f"{h} [{units[i]}]: " + " | ".join(r[i] for r in rows[2:]) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    cur, agg, stall = None, [], collections.Counter()
    for r in srows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 8 and r[0].strip().isdigit():
            agg.append((cur, int(r[0]), r[1].strip()[:110], int(r[6]) if r[6].isdigit() else 0, int(r[7]) if r[7].isdigit() else 0))
    tot_s, tot_i = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
    f.write(f"\n# hot source lines (first captured launch): % of warp-stall samples, % of executed warp instructions\n")
    for a in sorted(agg, key=lambda a: -a[3])[:25]:
        f.write(f"{a[0]:16s}:{a[1]:<4d} {100 * a[3] / tot_s:5.1f}% {100 * a[4] / tot_i:5.1f}%  {a[2]}\n")
    # warp stall reasons per captured launch, from the raw page's PC-sampling counters (smsp__pcsamp_warps_issue_stalled_*)
    pre = "smsp__pcsamp_warps_issue_stalled_"
    cols = [(i, h[len(pre):]) for i, h in enumerate(hdr) if h.startswith(pre) and not h.endswith("_not_issued")]
    f.write("\n# warp stall reasons, % of PC samples, one line per captured launch\n")
    for r in rows[2:]:
        vals = []
        for i, name in cols:
            try:
                vals.append((float(r[i]), name))
            except ValueError:
                pass
        tot = sum(v for v, _ in vals) or 1.0
        name_i = hdr.index("Kernel Name")
        f.write(r[name_i].split("(")[0][-40:] + ": " + "  ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(vals, reverse=True)[:8]) + "\n")
print("wrote", out)
