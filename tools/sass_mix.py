"""Executed-instruction mix of one kernel from an .ncu-rep: warp instructions per warp-step by SASS opcode.
usage: sass_mix.py <rep> <kernel regex> <warp-steps of the captured launch> [n opcodes]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep, kern, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if "Instructions Executed" in r)
i_src, i_exe = hdr.index("Source"), hdr.index("Instructions Executed")
mix, total = Counter(), 0.0
for r in rows:
    if len(r) <= i_exe or not r[i_exe].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
    mix[m.group(2) if m else "?"] += int(r[i_exe]) / steps
    total += int(r[i_exe]) / steps
fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
uni = sum(v for k, v in mix.items() if k in ("LDCU", "LDC", "UMOV", "S2UR", "ULEA", "UIADD3", "R2UR", "ULOP3", "UISETP", "UIMAD", "USHF", "USEL"))
print(f"# {kern}: {total:.1f} warp instructions per warp-step, fp64 {fp64:.1f} ({100 * fp64 / total:.0f} %), constant / uniform-path loads and moves {uni:.1f} ({100 * uni / total:.0f} %)")
print("  ".join(f"{k} {v:.1f}" for k, v in mix.most_common(top)))
