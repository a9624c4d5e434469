"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    agg.setdefault(r[4][:64], []).append(float(r[-1]) / 1e6)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:66s} n={len(v):4d} total={sum(v):9.2f} ms ({100 * sum(v) / tot:5.1f}%)  avg={sum(v) / len(v):8.3f}  min={min(v):.3f} max={max(v):.3f}")
print(f"all kernels: {tot:.2f} ms")
