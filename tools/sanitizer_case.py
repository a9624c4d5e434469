"""Workloads for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), run by tools/sanitize.sh on the GPU box:
  ptgsk    pt_gs_k run_windowed, <cells> x <steps>: multi-wave, time-sliced by ticket (the global-memory progress protocol of
           ptgsk_snow_kernel / ptgsk_response_kernel), dense DMMA interpolation with TMA staging, catchment reduction
  routing  pt_hs_k windowed run with a river network: hbv_run_kernel, route_local_inflow_kernel, route_river_level_kernel
  goal     calibration: goal_kernel over a single evaluation and a batched ensemble (members as grid layers)
usage: python tools/sanitizer_case.py <case> [cells] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import shyft_b200 as sb  # noqa: E402
from fixtures import PTGSK_DEFAULT, PTHSK_DEFAULT  # noqa: E402
from shyft_b200 import synthetic  # noqa: E402

case = sys.argv[1]
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if case == "ptgsk":
    n, T = cells or 60000, steps or 640
    geo, ta, env = synthetic.make_region(n, T, 64, config_index=1, start=1414800000)
    m = sb.PTGSKOptModel(geo, PTGSK_DEFAULT)
    m.initialize_cell_environment(ta)
    m._set_sources(env)
    m.set_states(synthetic.default_state(0, n))
    m.run_windowed(sb.InterpolationParameter(), window_steps=320)
    q = m.catchment_discharges()
    assert np.all(np.isfinite(q))
    print("ptgsk ok", n, T, m.kernel_launches(), float(q.sum()))
elif case == "routing":
    n, T = cells or 4000, steps or 500
    geo, ta, env = synthetic.make_region(n, T, 16, config_index=2, cells_per_catchment=40, with_routing=True, start=1414800000)
    m = sb.PTHSKOptModel(geo, PTHSK_DEFAULT)
    m.initialize_cell_environment(ta)
    m.set_states(synthetic.default_state(1, n))
    m.set_river_network(synthetic.river_chain(n // 40, depth=8))
    m.run_windowed(sb.InterpolationParameter(), env=env, window_steps=111)
    out = m.river_output_flow_m3s(8)
    assert np.all(np.isfinite(out))
    print("routing ok", n, T, m.kernel_launches(), float(out.sum()))
elif case == "goal":
    n, T = cells or 2000, steps or 24 * 20
    geo, ta, env = synthetic.make_region(n, T, 9, config_index=4, cells_per_catchment=500, start=1425168000)
    m = sb.PTGSKOptModel(geo, PTGSK_DEFAULT)
    m.run_interpolation(sb.InterpolationParameter(), ta, env)
    m.set_states(synthetic.default_state(0, n))
    m.run_cells()
    obs = m.catchment_discharges()[:, :2].sum(axis=1).reshape(-1, 24).mean(axis=1)
    opt = sb.Optimizer(m, [sb.TargetSpecification(obs, ta.start, 86400, [1, 2], 1.0, 0)])
    g0 = opt.calculate_goal_function(PTGSK_DEFAULT)
    P = np.tile(PTGSK_DEFAULT, (6, 1))
    P[:, 0] = np.linspace(-3.0, -2.0, 6)
    g = opt.calculate_goal_function_batch(P)
    assert np.all(np.isfinite(g))
    print("goal ok", n, T, g0, g.tolist())
else:
    raise SystemExit("unknown case")
