"""Debug helper (GPU box): first (step, cell) where the GPU pt_gs_k state series leave the oracle's, with the step's inputs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import shyft_b200 as sb
from shyft_b200 import synthetic
from oracle import oracle as O
from fixtures import FORCING, PTGSK_DEFAULT

n, T, S = 256, 8760, 16
geo, ta, env = synthetic.make_region(n, T, S, config_index=0, cells_per_catchment=100)
st0 = synthetic.default_state(0, n)
m = sb.PTGSKModel(geo, PTGSK_DEFAULT)
m.run_interpolation(sb.InterpolationParameter(use_idw_for_temperature=1), ta, env)
f = {k: m.cell_forcing(k) for k in FORCING}
m.set_states(st0)
m.set_state_collection(-1, True)
m.run_cells()
want = O.ptgsk_run_cells(O.geo_matrix(geo), PTGSK_DEFAULT, f, st0, ta.start * 10**6, ta.delta_t * 10**6, collect_state=True, collect_substeps=True, ncore=8)
names = sb.capi.STATE_SERIES_NAMES[0]
got = {k: m.state_series(k) for k in names}
rel = np.zeros((T + 1, n))
for k in names:
    d = np.abs(got[k] - want[k]) / np.maximum(np.maximum(np.abs(got[k]), np.abs(want[k])), 1e-300)
    d[np.abs(got[k] - want[k]) < 1e-14 * np.nanmax(np.abs(want[k]))] = 0
    rel = np.maximum(rel, d)
print("cells with any divergence > 1e-12:", int((rel.max(axis=0) > 1e-12).sum()), "of", n)
for thr in (1e-13, 1e-12, 1e-10, 1e-9, 1e-7):
    bad = rel > thr
    print(f"thr {thr:g}: {int(bad.sum())} values, {int(bad.any(axis=0).sum())} cells")
events = 0
for c in np.argsort(-(rel.max(axis=0)))[:6]:
    steps = np.nonzero(rel[:, c] > 1e-12)[0]
    if steps.size == 0:
        continue
    i = steps[0]  # state at the BEGINNING of step i differs -> step i-1 produced it
    print(f"\ncell {c}: first diverging state point {i} (produced by step {i - 1}); max rel {rel[:, c].max():.3e}")
    for k in names:
        print(f"   {k:22s} before: gpu {got[k][i - 1, c]!r} cpu {want[k][i - 1, c]!r} | after: gpu {got[k][i, c]!r} cpu {want[k][i, c]!r}")
    print("   forcing step", i - 1, {k: f[k][i - 1, c] for k in FORCING}, "substeps", want["kirchner_substeps"][i - 1, c])
    print("   z", geo["z"][c], "doy", O.day_of_year((ta.start + (i - 1) * 3600) * 10**6))
    # growth of the divergence afterwards
    print("   rel at +1,+10,+100,+1000:", [float(rel[min(T, i + d), c]) for d in (1, 10, 100, 1000)])
