"""Hot CUDA source lines of one kernel from an .ncu-rep: % of warp-stall samples, % of executed warp instructions, instructions per warp-step.
usage: src_hot.py <rep> <kernel regex> <warp-steps of the launch> [n lines]"""
import csv
import io
import subprocess
import sys

rep, kern, div = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
cur, agg = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 8 and r[0].strip().isdigit():
        agg.append((cur, int(r[0]), r[1].strip()[:118], int(r[6]) if r[6].isdigit() else 0, int(r[7]) if r[7].isdigit() else 0))
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print(f"# {kern}: {ti / div:.1f} warp instructions per warp-step; columns: % stall samples, % instructions, instructions per warp-step")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{a[0]:14s}:{a[1]:<5d} {100 * a[3] / ts:5.1f}% {100 * a[4] / ti:5.1f}% {a[4] / div:7.1f}  {a[2]}")
