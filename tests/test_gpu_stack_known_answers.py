"""The reference's stack-level known answers (test/pt_gs_k_test.cpp:174-354, test/pt_hs_k_test.cpp:93-153) through the C ABI on the
device: the same cases and asserts as tests/test_oracle_stack_known_answers.py (tests/stack_cases.py)."""
import numpy as np
import pytest

import stack_cases as sc
from fixtures import FORCING

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run():
    import shyft_b200 as sb
    models = {}
    last = {}

    def f(stack, geo, par, forcing, state, t0_us, T):
        key = (stack, geo.tobytes(), T)
        if key not in models:
            g = geo[0]
            cells = sb.geo_cell_data_vector([g[0]], [g[1]], [g[2]], area=g[3], catchment_id=np.array([int(g[4])]), radiation_slope_factor=g[5],
                                            glacier=g[6], lake=g[7], reservoir=g[8], forest=g[9])
            m = (sb.PTGSKModel if stack == 0 else sb.PTHSKModel)(cells, par)
            m.initialize_cell_environment(sb.TimeAxis(t0_us // 10**6, 3600, T))
            m.set_state_collection(-1, True)
            models[key] = m
            last[key] = {}
        m, seen = models[key], last[key]
        if seen.get("par") is None or not np.array_equal(seen["par"], par):
            m.set_region_parameter(par)
            seen["par"] = np.array(par)
        for k in FORCING:
            if seen.get(k) is None or not np.array_equal(seen[k], forcing[k]):
                m.set_cell_forcing(k, forcing[k])
                seen[k] = forcing[k].copy()
        m.set_states(state)
        m.run_cells()
        out = {name: m.response(name) for name in ("avg_discharge", "ae_output", "snow_outflow", "snow_sca", "snow_swe", "glacier_melt")}
        out["state"] = m.get_states()
        if stack == 1:
            out["state_snow_swe"] = m.state_series("snow_swe")
        return out
    return f


def test_pt_gs_k_mass_balance_and_land_type_routing(run):
    st = sc.ptgsk_mass_balance(run)
    sc.ptgsk_direct_response_on_reservoir_only(run, st)
    sc.ptgsk_glacier_and_reservoir_direct_response(run, st)


def test_pt_gs_k_lake_reservoir_response(run):
    sc.ptgsk_lake_reservoir_response(run)


def test_pt_hs_k_lake_reservoir_response(run):
    sc.pthsk_lake_reservoir_response(run)
