"""The numbers bench.py's `roofline` block quotes from an ncu capture of the three step kernels (one chunk of the shipped configuration):
per kernel the launch time, DRAM bytes, pipe utilisation, executed warp instructions and -- from the SASS page -- executed fp64-pipe
instructions, all also per cell-step.  usage: ncu_pipeline_json.py <report.ncu-rep> <n_cells> <chunk_steps> <out.json>"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

rep, n_cells, chunk, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
cell_steps = float(n_cells) * chunk
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, key, unit_scale=True):
    v = float(r[col[key]].replace(",", ""))
    u = units[col[key]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0) if unit_scale else 1.0
    return v * scale


FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
kernels = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("sb2::", "")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + re.escape(short.split("<")[0])],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    sh = srows[1]
    iS, iE = sh.index("Source"), sh.index("Instructions Executed")
    ops = collections.Counter()
    for q in srows[2:]:
        if len(q) > iE and q[iE].isdigit():
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", q[iS])
            ops[m.group(1) if m else "?"] += int(q[iE])
    tot_src = sum(ops.values()) or 1
    inst = num(r, "smsp__inst_executed.sum", False)
    fp64 = sum(ops[k] for k in FP64) * inst / tot_src   # the source page may count a launch twice: scale to the raw page's total
    kernels.append({
        "kernel": short, "ms": num(r, "gpu__time_duration.sum"), "registers": int(num(r, "launch__registers_per_thread", False)),
        "dram_read_bytes": num(r, "dram__bytes_read.sum"), "dram_write_bytes": num(r, "dram__bytes_write.sum"),
        "fp64_pipe_active_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", False),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
        "threads_per_instruction": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio", False),
        "warp_instructions": inst, "fp64_warp_instructions": fp64,
        "instructions_per_cell_step": inst * 32.0 / cell_steps, "fp64_instructions_per_cell_step": fp64 * 32.0 / cell_steps,
        "dram_bytes_per_cell_step": (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / cell_steps,
    })
doc = {"report": rep.split("/")[-1], "n_cells": n_cells, "chunk_steps": chunk, "cell_steps": cell_steps, "kernels": kernels,
       "dram_bytes_per_cell_step": sum(k["dram_bytes_per_cell_step"] for k in kernels),
       "note": "ncu --set full --clock-control none; instructions_per_cell_step counts thread-level instructions (warp instructions x 32 / cell-steps); "
               "times under ncu are cold-cache and serialised -- bench.py uses the shares and the byte / instruction counts, never these times"}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps({k["kernel"]: [round(k["ms"], 3), round(k["dram_bytes_per_cell_step"], 1), round(k["fp64_instructions_per_cell_step"], 1)] for k in kernels}))
