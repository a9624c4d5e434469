"""Executed-instruction profile of one kernel from an .ncu-rep: per SASS opcode class and per region of the hot loop.
usage: sass_hot.py <rep> <kernel regex> [steps*warps of the launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
warp_steps = float(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
by_op, by_op_samp = collections.Counter(), collections.Counter()
tot = 0
lines = []
for r in rows[2:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    op = m.group(1) if m else "?"
    n = int(r[iE])
    by_op[op] += n
    by_op_samp[op] += int(r[iSamp]) if r[iSamp].isdigit() else 0
    tot += n
    lines.append((r[0], r[iS].strip(), n, int(r[iSamp]) if r[iSamp].isdigit() else 0))
div = warp_steps or 1.0
print(f"total executed warp instructions {tot}" + (f" = {tot / div:.1f} per warp-step" if warp_steps else ""))
f64 = sum(v for k, v in by_op.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "F2F", "I2F", "F2I"))
print(f"fp64-pipe instructions {f64} ({100.0 * f64 / tot:.1f} %)" + (f" = {f64 / div:.1f} per warp-step" if warp_steps else ""))
tots = sum(by_op_samp.values()) or 1
for k, v in by_op.most_common(28):
    print(f"  {k:10s} {v / div:9.2f} {100.0 * v / tot:5.1f} %   stall samples {100.0 * by_op_samp[k] / tots:5.1f} %")
if "--dump" in sys.argv:
    for a, s, n, sm in lines:
        print(f"{n / div:8.3f} {sm:6d}  {s}")
