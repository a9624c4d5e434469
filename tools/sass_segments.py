"""Executed-instruction breakdown of one kernel of an .ncu-rep by called subroutine (segments between CALL targets).
usage: python tools/sass_segments.py report.ncu-rep warp_steps [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, ws = sys.argv[1], float(sys.argv[2])
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')[1:]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[which])))
print(rows[0][1][:90])
data = [(int(r[0], 16), r[1].strip(), int(r[2]), int(r[5]), float(r[8]) if r[8] not in ("", "-") else 0.0) for r in rows[2:] if r and r[0].startswith("0x")]
base = data[0][0]
tot, tots = sum(d[3] for d in data), sum(d[2] for d in data)
print(f"instructions per warp-step {tot / ws:.1f}")
targets = collections.Counter()
for a, src, smp, ins, thr in data:
    m = re.search(r"CALL\.\S+\s+(0x[0-9a-f]+)", src)
    if m and ins > 0:
        targets[int(m.group(1), 16)] += ins
ents = [base] + sorted(set(targets)) + [data[-1][0] + 16]
for k in range(len(ents) - 1):
    seg = [d for d in data if ents[k] <= d[0] < ents[k + 1]]
    ins, smp = sum(d[3] for d in seg), sum(d[2] for d in seg)
    if ins == 0:
        continue
    op = lambda s: (s.split()[1] if s.startswith("@") else s.split()[0])
    fp = sum(d[3] for d in seg if op(d[1])[0] == "D")
    div = sum(d[3] for d in seg if "MUFU.RCP64H" in d[1])
    thr = sum(d[4] * d[3] for d in seg) / ins
    print(f"segment @{ents[k] - base:6x} static {len(seg):5d}  calls/ws {targets.get(ents[k], 0) / ws:6.3f}  instr/ws {ins / ws:7.1f} ({100 * ins / tot:4.1f} %)  "
          f"stall samples {100 * smp / tots:4.1f} %  fp64/ws {fp / ws:6.1f}  div/ws {div / ws:5.2f}  active threads {thr:4.1f}")
