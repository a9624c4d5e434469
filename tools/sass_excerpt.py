"""SASS evidence for profiles/: per kernel the instruction mix (tensor / TMA / barrier / fp64 mnemonics) and the first lines of each kind.
usage: python tools/sass_excerpt.py [library.so] > profiles/sass_excerpt_r02.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "shyft_b200/libshyft_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KERNELS = ["dense_apply_dmma_kernelILi6ELi2ELi0ELb1ELi16ELi32E", "dense_apply_dmma_kernelILi16ELi2ELi1ELb0ELi16ELi0E", "ptgsk_forcing_terms_kernelILb1E", "ptgsk_snow_kernelILi0ELb1E",
           "ptgsk_response_kernelILi1ELb1E", "hbv_run_kernelILb0E", "hbv_run_kernelILb1E", "route_local_inflow_kernel", "route_river_level_kernel", "goal_kernel"]
WATCH = ["DMMA", "UBLKCP", "SYNCS", "UTMA", "LDS", "STS", "SHFL", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "BAR", "ATOM", "RED", "LDG", "STG", "CALL"]
print(f"# cuobjdump -sass {lib} (sm_100a), per kernel: static instruction counts by mnemonic and the first occurrences of the tensor-core / TMA / barrier ones")
for k in KERNELS:
    m = re.search(r"Function : (\S*" + re.escape(k) + r"\S*)\n(.*?)(?=\n\s*Function :|\Z)", txt, re.S)
    if not m:
        print(f"\n## {k}: not found")
        continue
    lines = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in m.group(2).split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if mm:
            ops[mm.group(1).split(".")[0]] += 1
    print(f"\n## {m.group(1)}\n   {len(lines)} instructions: " + ", ".join(f"{w} {ops[w]}" for w in WATCH if ops[w]))
    for w in ("DMMA", "UBLKCP", "SYNCS", "UTMA"):
        hits = [l.strip() for l in lines if re.search(r"\b" + w, l)]
        for h in hits[:3]:
            print("   " + h)
