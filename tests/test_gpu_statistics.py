"""Statistics readers through the C ABI against the reference's own literals and the oracle.

model.statistics / gamma_snow_response / ... mirror api/api.h:178-1600 (shyft/api/pt_gs_k/__init__.py:14-20); the reductions run
on the device over the resident [step][cell] series."""
import json
import os

import numpy as np
import pytest

from fixtures import py_region_fixture

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))["region_pt_gs_k_20x240"]


@pytest.fixture(scope="module")
def sb():
    import shyft_b200
    return shyft_b200


def _py_fixture_model(sb):
    fx = py_region_fixture()
    g = fx["geo"]
    geo = sb.geo_cell_data_vector(g[:, 0], g[:, 1], g[:, 2], area=g[:, 3], catchment_id=g[:, 4].astype(np.int64), radiation_slope_factor=g[:, 5],
                                  glacier=g[:, 6], lake=g[:, 7], reservoir=g[:, 8], forest=g[:, 9])
    m = sb.PTGSKModel(geo, fx["par"])
    ta = sb.TimeAxis(fx["t0"], fx["dt"], fx["T"])
    env = sb.RegionEnvironment(**{k: (fx["station"][None, :], np.full((fx["T"], 1), v)) for k, v in fx["consts"].items()})
    assert m.run_interpolation(sb.InterpolationParameter(), ta, env)
    m.set_states(fx["state"])
    m.set_state_collection(-1, True)   # as the reference test does (test_region_model_stacks.py:206)
    m.run_cells()
    return fx, m


def test_reference_literals_through_the_statistics_api(sb):
    """shyft/tests/api/test_region_model_stacks.py:219-241, same calls, same literals"""
    from shyft_b200.statistics import CELL_IX
    fx, m = _py_fixture_model(sb)
    cids = []
    assert m.statistics.charge_value(cids, 0) == pytest.approx(GOLD["charge_sum_step0"]["value"], abs=1e-4)
    assert m.statistics.charge_value([0, 1, 3], 0, ix_type=CELL_IX) == pytest.approx(GOLD["charge_cells_0_1_3_step0"]["value"], abs=1e-4)
    assert m.statistics.charge([1, 2, 6], ix_type=CELL_IX).sum() == pytest.approx(GOLD["charge_sum_cells_1_2_6_all_steps"]["value"], abs=2e-4)
    ae_output = m.actual_evaptranspiration_response.output(cids)
    assert ae_output.max() == pytest.approx(GOLD["ae_output_max"]["value"], abs=1e-13)
    pot_ratio = m.actual_evaptranspiration_response.pot_ratio(cids)
    assert pot_ratio.size == fx["T"] + 1
    assert pot_ratio.min() == pytest.approx(GOLD["ae_pot_ratio_min"]["value"], abs=1e-13)
    assert pot_ratio.max() == pytest.approx(1.0, abs=1e-7)
    assert m.statistics.discharge_value(cids, 0) >= GOLD["discharge_step0_min"]
    assert m.statistics.discharge(cids).size == fx["T"]
    with pytest.raises(RuntimeError, match="one or more supplied catchment_indexes does not exist:3"):
        m.statistics.discharge([1, 3])
    with pytest.raises(RuntimeError, match="Supplied cell index reference 100 is ouside valid range"):
        m.statistics.discharge_value([100], 0, ix_type=CELL_IX)


def test_every_reader_against_the_oracle_restatement(sb, oracle):
    from shyft_b200 import synthetic
    from shyft_b200.statistics import CATCHMENT_IX, CELL_IX
    n, T = 1500, 96
    geo, ta, env = synthetic.make_region(n, T, 9, config_index=11, cells_per_catchment=256, start=1420070400)[:3]
    geo["area"] = np.random.default_rng(3).uniform(0.5e6, 2.0e6, n)   # unequal weights
    m = sb.PTGSKModel(geo)
    assert m.run_interpolation(sb.InterpolationParameter(), ta, env)
    m.set_states(synthetic.default_state(0, n))
    m.set_state_collection(-1, True)
    m.run_cells()
    cids, area = geo["catchment_id"], geo["area"]
    sel = [int(cids[0]), int(cids[-1])]
    some_cells = [3, 700, 701, 1499]
    tol = dict(rtol=1e-12, atol=1e-300)

    for name in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum"):
        s = m.cell_forcing(name)
        assert np.allclose(getattr(m.statistics, name)(sel), oracle.average_catchment_feature(s, area, cids, sel), **tol)
        assert np.allclose(getattr(m.statistics, name)([]), oracle.average_catchment_feature(s, area, cids, []), **tol)
        assert getattr(m.statistics, name + "_value")(some_cells, 17, ix_type=CELL_IX) == pytest.approx(
            oracle.average_catchment_feature_value(s, area, cids, some_cells, 17, oracle.CELL_IX), rel=1e-12)
        assert np.array_equal(getattr(m.statistics, name)(sel, 5), oracle.catchment_feature(s, cids, sel, 5))

    q, ch = m.response("avg_discharge"), m.response("charge_m3s")
    assert np.allclose(m.statistics.discharge(sel), oracle.sum_catchment_feature(q, cids, sel), **tol)
    assert np.allclose(m.statistics.charge(some_cells, ix_type=CELL_IX), oracle.sum_catchment_feature(ch, cids, some_cells, oracle.CELL_IX), **tol)
    assert m.statistics.discharge_value(sel, 40) == pytest.approx(oracle.sum_catchment_feature_value(q, cids, sel, 40), rel=1e-12)
    # the catchment sums the step kernels keep are the same quantity
    k = list(m.catchment_ids).index(sel[0])
    assert np.allclose(m.statistics.discharge([sel[0]]), m.catchment_discharges()[:, k], rtol=1e-12)

    resp = m.gamma_snow_response
    assert np.allclose(resp.sca(sel), oracle.average_catchment_feature(m.response("snow_sca"), area, cids, sel), **tol)
    assert np.allclose(resp.swe([]), oracle.average_catchment_feature(m.response("snow_swe"), area, cids, []), **tol)
    assert np.allclose(resp.outflow(sel), oracle.sum_catchment_feature(m.response("snow_outflow"), cids, sel), **tol)
    assert np.allclose(resp.glacier_melt(sel), oracle.sum_catchment_feature(m.response("glacier_melt"), cids, sel), **tol)
    assert np.allclose(m.priestley_taylor_response.output(sel), oracle.average_catchment_feature(m.response("pe_output"), area, cids, sel), **tol)
    assert np.allclose(m.actual_evaptranspiration_response.output(sel), oracle.average_catchment_feature(m.response("ae_output"), area, cids, sel), **tol)

    st = m.gamma_snow_state
    for field in ("albedo", "lwc", "surface_heat", "alpha", "sdc_melt_mean", "acc_melt", "iso_pot_energy", "temp_swe"):
        s = m.state_series("gs_" + field)
        got = getattr(st, field)(sel)
        assert got.size == T + 1
        assert np.allclose(got, oracle.average_catchment_feature(s, area, cids, sel), **tol)
    kd = m.state_series("kirchner_discharge")
    assert np.allclose(m.kirchner_state.discharge(sel), oracle.sum_catchment_feature(kd, cids, sel), **tol)
    pr = oracle.ae_pot_ratio(kd, area, m.get_region_parameter()[3])
    assert np.allclose(m.actual_evaptranspiration_response.pot_ratio(sel), oracle.average_catchment_feature(pr, area, cids, sel), **tol)
    assert np.array_equal(m.actual_evaptranspiration_response.pot_ratio(some_cells, 9, ix_type=CELL_IX),
                          oracle.catchment_feature(pr, cids, some_cells, 9, oracle.CELL_IX))

    # geo sums (api/api.h:183-288)
    in_sel = np.isin(cids, sel)
    assert m.statistics.total_area([]) == pytest.approx(area.sum(), rel=1e-14)
    assert m.statistics.total_area(sel) == pytest.approx(area[in_sel].sum(), rel=1e-14)
    assert m.statistics.total_area(some_cells, ix_type=CELL_IX) == pytest.approx(area[some_cells].sum(), rel=1e-14)
    assert m.statistics.forest_area(sel) == pytest.approx((area * geo["forest"])[in_sel].sum(), rel=1e-14)
    assert m.statistics.glacier_area([]) == pytest.approx((area * geo["glacier"]).sum(), rel=1e-14)
    assert m.statistics.lake_area(sel) == pytest.approx((area * geo["lake"])[in_sel].sum(), rel=1e-14)
    assert m.statistics.reservoir_area(sel) == pytest.approx((area * geo["reservoir"])[in_sel].sum(), rel=1e-14)
    assert m.statistics.snow_storage_area(sel) == pytest.approx((area * (1 - geo["lake"] - geo["reservoir"]))[in_sel].sum(), rel=1e-14)
    assert m.statistics.unspecified_area([]) == pytest.approx((area * (1 - geo["glacier"] - geo["lake"] - geo["reservoir"] - geo["forest"])).sum(), rel=1e-13)
    assert m.statistics.elevation(sel) == pytest.approx((geo["z"] * area)[in_sel].sum() / area[in_sel].sum(), rel=1e-14)


def test_statistics_of_an_uncollected_series_fail_loudly(sb):
    from shyft_b200 import synthetic
    geo, ta, env = synthetic.make_region(64, 24, 4, config_index=12, cells_per_catchment=32)[:3]
    m = sb.PTGSKOptModel(geo)   # discharge collector only
    assert m.run_interpolation(sb.InterpolationParameter(), ta, env)
    m.set_states(synthetic.default_state(0, 64))
    m.run_cells()
    assert m.statistics.discharge([]).size == 24
    with pytest.raises(RuntimeError, match="not collected"):
        m.statistics._series(1, 6, [], 0, 1)   # ae_output


@pytest.mark.parametrize("stack", [1, 2])
def test_hbv_stack_readers_against_the_oracle_restatement(sb, oracle, stack):
    """shyft/api/pt_hs_k/__init__.py:13-19 and shyft/api/hbv_stack/__init__.py:13-19: the per-method statistics of the HBV stacks"""
    from fixtures import FORCING, HBV_DEFAULT, PTHSK_DEFAULT, geo_matrix
    from shyft_b200 import synthetic
    from shyft_b200.statistics import CELL_IX
    n, T = 900, 120
    geo, ta, env = synthetic.make_region(n, T, 9, config_index=13, cells_per_catchment=128, start=1420070400)[:3]
    geo["area"] = np.random.default_rng(4).uniform(0.5e6, 2.0e6, n)
    par = PTHSK_DEFAULT if stack == 1 else HBV_DEFAULT
    m = (sb.PTHSKModel if stack == 1 else sb.HbvStackModel)(geo, par)
    assert m.run_interpolation(sb.InterpolationParameter(), ta, env)
    st0 = synthetic.default_state(stack, n)
    m.set_states(st0)
    m.set_state_collection(-1, True)
    m.run_cells()
    cids, area = geo["catchment_id"], geo["area"]
    sel = [int(cids[0]), int(cids[-1])]
    some_cells = [0, 450, 899]
    tol = dict(rtol=1e-12, atol=1e-300)
    run = oracle.pthsk_run_cells if stack == 1 else oracle.hbv_stack_run_cells
    want = run(geo_matrix(geo), par, {k: m.cell_forcing(k) for k in FORCING}, st0, ta.start * 10**6, ta.delta_t * 10**6, ncore=4)
    r = 1e-9   # the series themselves carry the step kernels' parity tolerance
    assert np.allclose(m.statistics.discharge(sel), oracle.sum_catchment_feature(want["avg_discharge"], cids, sel), rtol=r)
    assert np.allclose(m.hbv_snow_response.outflow(sel), oracle.sum_catchment_feature(want["snow_outflow"], cids, sel), rtol=r)
    assert np.allclose(m.hbv_snow_response.glacier_melt([]), oracle.sum_catchment_feature(want["glacier_melt"], cids, []), rtol=r, atol=1e-300)
    assert np.allclose(m.priestley_taylor_response.output(sel), oracle.average_catchment_feature(want["pe_output"], area, cids, sel), rtol=r)
    assert m.hbv_snow_response.outflow_value(some_cells, 30, ix_type=CELL_IX) == pytest.approx(
        oracle.sum_catchment_feature_value(want["snow_outflow"], cids, some_cells, 30, oracle.CELL_IX), rel=r)
    swe, sca = m.state_series("snow_swe"), m.state_series("snow_sca")
    assert swe.shape == (T + 1, n) and np.nanmax(swe) > 1.0
    assert np.allclose(m.hbv_snow_state.swe(sel), oracle.average_catchment_feature(swe, area, cids, sel), **tol)
    assert np.allclose(m.hbv_snow_state.sca([]), oracle.average_catchment_feature(sca, area, cids, []), **tol)
    assert np.array_equal(m.hbv_snow_state.swe(sel, 7), oracle.catchment_feature(swe, cids, sel, 7))
    assert m.hbv_snow_state.sca_value(sel, 50) == pytest.approx(oracle.average_catchment_feature_value(sca, area, cids, sel, 50), rel=1e-12)
    assert len(m.hbv_snow_state.sp(sel)) == 5 and len(m.hbv_snow_state.sw_value(sel, 3)) == 5   # per-bin series (tests/test_gpu_hbv_routing.py)
    if stack == 1:
        kd = m.state_series("kirchner_discharge")
        assert np.allclose(m.kirchner_state.discharge(sel), oracle.sum_catchment_feature(kd, cids, sel), **tol)
        assert np.allclose(m.actual_evaptranspiration_response.output(sel), oracle.average_catchment_feature(want["ae_output"], area, cids, sel), rtol=r)
        pr = oracle.ae_pot_ratio(kd, area, par[3])
        assert np.allclose(m.actual_evaptranspiration_response.pot_ratio(sel), oracle.average_catchment_feature(pr, area, cids, sel), **tol)
    else:
        sm, uz = m.state_series("soil_moisture"), m.state_series("tank_uz")
        assert np.allclose(m.soil_state.discharge(sel), oracle.sum_catchment_feature(sm, cids, sel), **tol)
        assert np.allclose(m.tank_state.discharge([]), oracle.sum_catchment_feature(uz, cids, []), **tol)
        assert m.tank_state.discharge_value(sel, 3) == pytest.approx(oracle.sum_catchment_feature_value(uz, cids, sel, 3), rel=1e-12)
        assert np.allclose(m.soil_response.output(sel), oracle.average_catchment_feature(want["soil_outflow"], area, cids, sel), rtol=r)
        assert np.allclose(m.hbv_actual_evaptranspiration_response.output(sel), oracle.average_catchment_feature(want["ae_output"], area, cids, sel), rtol=r)
        with pytest.raises(RuntimeError, match="Kirchner stack"):
            m.statistics._series(3, 0, [], 0, 1, 0, 1)   # pot_ratio
