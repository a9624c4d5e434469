"""adjust_state_to_target_flow through the C ABI: the reference test's own calls and asserts
(shyft/tests/api/test_region_model_stacks.py:311-333), then every number against the oracle's restatement of
adjust_state_model::tune_flow (core/model_state_tuning.h:38-118)."""
import numpy as np
import pytest

from fixtures import FORCING, HBV_DEFAULT, geo_matrix, py_region_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import shyft_b200
    return shyft_b200


def _py_fixture_model(sb, catchment_id=None):
    fx = py_region_fixture()
    g = fx["geo"].copy()
    if catchment_id is not None:
        g[:, 4] = catchment_id
    geo = sb.geo_cell_data_vector(g[:, 0], g[:, 1], g[:, 2], area=g[:, 3], catchment_id=g[:, 4].astype(np.int64), radiation_slope_factor=g[:, 5],
                                  glacier=g[:, 6], lake=g[:, 7], reservoir=g[:, 8], forest=g[:, 9])
    m = sb.PTGSKModel(geo, fx["par"])
    ta = sb.TimeAxis(fx["t0"], fx["dt"], fx["T"])
    env = sb.RegionEnvironment(**{k: (fx["station"][None, :], np.full((fx["T"], 1), v)) for k, v in fx["consts"].items()})
    assert m.run_interpolation(sb.InterpolationParameter(), ta, env)
    m.set_states(fx["state"])
    return fx, g, m


def _oracle_runner(oracle, fx, g, m):
    f = {k: m.cell_forcing(k) for k in FORCING}

    def run_cells(state, start_step, n_steps, mask):
        return oracle.ptgsk_run_cells(g, fx["par"], f, state, fx["t0"] * 10**6, fx["dt"] * 10**6, start_step=start_step, n_steps=n_steps,
                                      cell_mask=mask)
    return run_cells


def test_reference_python_test_sequence(sb, oracle):
    fx, g, m = _py_fixture_model(sb)
    cids = []
    q_0 = m.get_states()[0, 8]
    m.adjust_q(2.0, cids)
    assert m.get_states()[0, 8] == q_0 * 2.0
    m.revert_to_initial_state()
    m.run_cells(0, 10, 2)
    q_avg = (m.statistics.discharge_value(cids, 10) + m.statistics.discharge_value(cids, 11)) / 2.0
    x = 0.7
    m.revert_to_initial_state()
    r = m.adjust_state_to_target_flow(x * q_avg, cids, start_step=10, scale_range=3.0, scale_eps=1e-3, max_iter=350, n_steps=2)
    assert len(r.diagnostics) == 0
    assert r.q_r == pytest.approx(q_avg * x, abs=0.005)
    assert r.q_0 == pytest.approx(q_avg, abs=0.005)
    # against the oracle: same evaluations, same result
    want = oracle.adjust_state_to_target_flow(_oracle_runner(oracle, fx, g, m), fx["state"], [8], g[:, 4].astype(np.int64), x * q_avg, cids=cids,
                                              start_step=10, scale_range=3.0, scale_eps=1e-3, max_iter=350, n_steps=2)
    assert r.q_0 == pytest.approx(want["q_0"], rel=1e-9) and r.q_r == pytest.approx(want["q_r"], rel=1e-9)
    got_state = m.get_states()
    assert np.array_equal(got_state[:, :8], fx["state"][:, :8])
    np.testing.assert_allclose(got_state[:, 8], want["state"][:, 8], rtol=1e-9)
    assert np.array_equal(m.initial_state, fx["state"])          # "region-model initial state is not changed during the process"
    # running on from the adjusted state gives the tuned flow
    m.run_cells(0, 10, 2)
    assert (m.statistics.discharge_value(cids, 10) + m.statistics.discharge_value(cids, 11)) / 2.0 == pytest.approx(r.q_r, rel=1e-12)
    # bad observed value, then bad simulated values
    m.revert_to_initial_state()
    r = m.adjust_state_to_target_flow(float("nan"), cids, start_step=10, n_steps=2)
    assert len(r.diagnostics) > 0
    t = m.cell_forcing("temperature")
    t[10, 0] = float("nan")
    m.set_cell_forcing("temperature", t)
    m.revert_to_initial_state()
    r = m.adjust_state_to_target_flow(30.0, cids, start_step=10, n_steps=2)
    assert len(r.diagnostics) > 0


def test_subset_of_catchments_keeps_filter_and_other_cells(sb, oracle):
    cat = np.where(np.arange(20) < 8, 1, 2)
    fx, g, m = _py_fixture_model(sb, catchment_id=cat)
    m.run_cells(0, 0, 2)
    q2 = m.statistics.discharge([2])[:2].sum() / 2.0
    m.revert_to_initial_state()
    m.set_catchment_calculation_filter([1])
    r = m.adjust_state_to_target_flow(1.6 * q2, [2], start_step=0, scale_range=10.0, n_steps=2)
    assert r.diagnostics == "" and r.q_r == pytest.approx(1.6 * q2, abs=0.1)
    want = oracle.adjust_state_to_target_flow(_oracle_runner(oracle, fx, g, m), fx["state"], [8], cat.astype(np.int64), 1.6 * q2, cids=[2],
                                              start_step=0, scale_range=10.0, n_steps=2)
    assert r.q_r == pytest.approx(want["q_r"], rel=1e-9)
    s = m.get_states()
    assert np.array_equal(s[cat == 1], fx["state"][cat == 1])
    np.testing.assert_allclose(s[cat == 2, 8], want["state"][cat == 2, 8], rtol=1e-9)
    # the caller's filter ([1]) is back: a run now leaves catchment 2 untouched
    before = m.get_states()
    m.run_cells(0, 0, 4)
    after = m.get_states()
    assert np.array_equal(after[cat == 2], before[cat == 2]) and not np.array_equal(after[cat == 1], before[cat == 1])
    with pytest.raises(RuntimeError, match="no cells have supplied cid"):
        m.adjust_state_to_target_flow(10.0, [3])


def test_hbv_stack_scales_soil_and_tank_storages(sb, oracle):
    from shyft_b200 import synthetic
    n, T = 96, 240
    geo, ta, env = synthetic.make_region(n, T, 9, config_index=2, cells_per_catchment=32, start=1430438400)  # 2015-05-01
    m = sb.HbvStackModel(geo, HBV_DEFAULT)
    m.run_interpolation(sb.InterpolationParameter(), ta, env)
    st0 = synthetic.default_state(2, n)
    st0[:, 12] = 120.0  # soil moisture
    m.set_states(st0)
    f = {k: m.cell_forcing(k) for k in FORCING}
    G = geo_matrix(geo)

    def run_cells(state, start_step, n_steps, mask):
        # the oracle's hbv_stack entry has no cell mask: cells outside the filter are not read by the tuning (cids select them out)
        return oracle.hbv_stack_run_cells(G, HBV_DEFAULT, f, state, ta.start * 10**6, ta.delta_t * 10**6, start_step=start_step, n_steps=n_steps)
    m.run_cells(0, 5, 3)
    q = m.statistics.discharge([])[5:8].mean()
    m.revert_to_initial_state()
    r = m.adjust_state_to_target_flow(1.3 * q, [], start_step=5, n_steps=3)
    want = oracle.adjust_state_to_target_flow(run_cells, st0, [12, 13, 14], G[:, 4].astype(np.int64), 1.3 * q, start_step=5, n_steps=3)
    assert r.diagnostics == want["diagnostics"] == ""
    assert r.q_0 == pytest.approx(want["q_0"], rel=1e-9) and r.q_r == pytest.approx(want["q_r"], rel=1e-9)
    assert r.q_r == pytest.approx(1.3 * q, rel=2e-3)
    np.testing.assert_allclose(m.get_states(), want["state"], rtol=1e-9)
