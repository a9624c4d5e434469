"""Operation counts per cell-step of the pt_gs_k stack on the synthetic region (uses a -DSHO_COUNT build of the CPU oracle).

usage: python tools/cost_model.py [n_cells] [years]   -- single-threaded; prints counts per cell-step, per year of the run."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
so = os.path.join(ROOT, "build", "libsho_oracle_count.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-mfma", "-DSHO_COUNT", "-shared", "-o", so,
                       os.path.join(ROOT, "oracle", "capi.cpp")])
from oracle import oracle as O  # noqa: E402

O._LIB = None
O.build = lambda force=False: so
from shyft_b200 import synthetic as S  # noqa: E402
import bench  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 500
years = float(sys.argv[2]) if len(sys.argv) > 2 else 3
NAMES = ["exp", "log", "lgamma", "gser_calls", "gser_iter", "gcf_calls", "gcf_iter", "brent_calls", "brent_eval", "kir_try", "kir_reject", "snow_state",
         "gs_active", "cell_steps"]
T = int(8760 * years)
big = 100000
geo, ta, env = S.make_region(big, T, 64, config_index=1)[:3]
idx = np.linspace(0, big - 1, n_cells).astype(np.int64)
gm = O.geo_matrix(geo[idx])
dt_us = ta.delta_t * 10**6
f = {}
for name in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum"):
    xyz, vals = getattr(env, name)
    vals = O.average_accessor_same_axis(vals[:T], dt_us)
    if name == "temperature":
        f[name] = O.btk_run(xyz, vals, gm[:, :3], ta.start * 10**6, dt_us)
    else:
        f[name] = O.idw_run(name, xyz, vals, gm[:, :3], O.idw_par(max_members=20 if name == "precipitation" else 10), dst_slope=gm[:, 5], ncore=1)
st = np.tile(np.array([0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, 0.8]), (n_cells, 1))
out = (C.c_longlong * 16)()
O.lib().sho_counters(out, 1)
chunk = 8760 // 4
print("quarter " + " ".join(f"{n:>10s}" for n in NAMES[:-1]))
for k in range(T // chunk):
    O.ptgsk_run_cells(gm, bench.PTGSK_DEFAULT, f, st, ta.start * 10**6, dt_us, start_step=k * chunk, n_steps=chunk, collect_response=False, ncore=1)
    O.lib().sho_counters(out, 1)
    cs = out[13]
    print(f"{k:7d} " + " ".join(f"{out[i] / cs:10.3f}" for i in range(13)))
