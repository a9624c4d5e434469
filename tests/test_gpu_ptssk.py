"""GPU parity of the pt_ss_k stack (core/pt_ss_k.h:195-293; Skaugen snow routine core/skaugen.h:24-386) through the C ABI: the snow routine
one step at a time against the oracle (the reference unit-tests it the same way, test/skaugen_test.cpp), the stack over a winter, the
reference's own stack-level asserts (test/pt_ss_k_test.cpp:118-168) on the device."""
import numpy as np
import pytest

import stack_cases as sc
from fixtures import FORCING, PTSSK_DEFAULT, geo_matrix
from parity import assert_parity

pytestmark = pytest.mark.gpu
PAR8 = PTSSK_DEFAULT[4:12]


@pytest.fixture(scope="module")
def sb():
    import shyft_b200
    return shyft_b200


def _bits_equal(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.array_equal(a.view(np.uint64), b.view(np.uint64)) or np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]) and np.array_equal(np.isnan(a), np.isnan(b))


def test_sca_rel_red_is_bit_identical(sb, oracle):
    """statistics::sca_rel_red (skaugen.h:52-79): crossing of two gamma densities by a 2-bit Brent minimum search and a 10-bit bisection"""
    rng = np.random.default_rng(31)
    n = 4000
    nnn = rng.integers(20, 6000, n).astype(np.float64)
    u = np.maximum(2.0, np.floor(nnn * rng.uniform(0.01, 0.6, n)))
    alpha = rng.uniform(2.0, 40.77, n)
    nu = alpha * 0.1 * nnn * rng.uniform(0.6, 1.0, n)          # nu scaled with the number of units, as step() holds it
    got = sb.capi.unit_eval("sca_rel_red", np.stack([u, nnn, nu, alpha], axis=1))
    want = np.array([oracle.skaugen_sca_rel_red(u[i], nnn[i], nu[i], alpha[i]) for i in range(n)])
    ok = ~np.isnan(want)                                       # NaN: the reference (and the oracle) throw "no change of sign"
    assert ok.mean() > 0.8
    assert np.all(got[~ok, 1] == 1.0) and np.all(got[ok, 1] == 0.0)
    assert _bits_equal(got[ok, 0], want[ok])
    assert np.all((want[ok] >= 0.0) & (want[ok] <= 1.0))


def test_skaugen_step_sequences_are_bit_identical(sb, oracle):
    """skaugen::calculator::step chained over random weather for many cells, device step by step against the oracle: accumulation, refreeze,
    partial and complete melt (sca_rel_red, compute_shape_vars), rain on snow, the residual bookkeeping"""
    rng = np.random.default_rng(37)
    n_cells, n_steps = 96, 400
    st = np.tile(np.array([4.077, 40.77, 0.0, 0.0, 0.0, 0.0, 0.0]), (n_cells, 1))
    st_o = st.copy()
    season = np.sin(2 * np.pi * np.arange(n_steps) / n_steps)
    bad_seen = 0
    for i in range(n_steps):
        temp = -6.0 * season[i] + rng.normal(0, 4.0, n_cells)
        prec = rng.exponential(1.5, n_cells) * (rng.random(n_cells) < 0.3)
        dt_h = 24.0 if i % 3 == 0 else 3.0
        rows = np.concatenate([np.tile(PAR8, (n_cells, 1)), st, np.full((n_cells, 1), dt_h), temp[:, None], prec[:, None]], axis=1)
        got = sb.capi.unit_eval("skaugen_step", rows)
        for c in range(n_cells):
            try:
                s1, r = oracle.skaugen_step(st_o[c], temp[c], prec[c], dt_us=int(dt_h * 3600 * 10**6), par=PAR8)
            except RuntimeError:
                assert got[c, 10] == 1.0            # the device raised its flag where the reference throws
                bad_seen += 1
                s1, r = got[c, :7].copy(), got[c, 7:10]
            else:
                assert got[c, 10] == 0.0
                assert _bits_equal(got[c, :7], s1) and _bits_equal(got[c, 7:10], r), (i, c, got[c], s1, r)
            st_o[c] = s1
        st = got[:, :7].copy()
    assert bad_seen < n_cells * n_steps * 0.01


def test_pt_ss_k_stack_parity_through_a_winter(sb, oracle):
    from shyft_b200 import synthetic
    n, T = 320, 6000
    geo, ta, env = synthetic.make_region(n, T, 16, config_index=2, cells_per_catchment=40, start=1414800000)   # 2014-11-01
    m = sb.PTSSKModel(geo, PTSSK_DEFAULT)
    assert m.parameter_size == 21 and m.state_size == 8
    m.run_interpolation(sb.InterpolationParameter(), ta, env)
    st0 = synthetic.default_state(3, n)
    m.set_states(st0)
    m.set_state_collection(-1, True)
    f = {k: m.cell_forcing(k) for k in FORCING}
    m.run_cells()
    want = oracle.ptssk_run_cells(geo_matrix(geo), PTSSK_DEFAULT, f, st0, ta.start * 10**6, ta.delta_t * 10**6, collect_state=True, ncore=8)
    assert np.nanmax(want["snow_swe"]) > 5.0, "the fixture must build a snow pack"
    for name in ("avg_discharge", "charge_m3s", "snow_sca", "snow_swe", "snow_outflow", "glacier_melt", "ae_output", "pe_output"):
        assert_parity(m.response(name), want[name], "pt_ss_k " + name)
        assert np.array_equal(m.response(name), want[name], equal_nan=True), name + " is within 1e-9 but not bit-identical"
    for name in sb.capi.STATE_SERIES_NAMES[sb.PT_SS_K]:
        assert_parity(m.state_series(name), want["state_" + name], "pt_ss_k state " + name)
    assert_parity(m.get_states(), want["state"], "pt_ss_k end state")
    cd = m.catchment_discharges()
    assert_parity(cd[:, 0], want["avg_discharge"][:, :40].sum(axis=1), "catchment discharge", rtol=1e-12)
    # chunked run = one shot; windowed run = resident run
    q, s = m.response("avg_discharge"), m.get_states()
    m.revert_to_initial_state()
    for k in range(4):
        m.run_cells(0, 1500 * k, 1500)
    assert np.array_equal(m.response("avg_discharge"), q) and np.array_equal(m.get_states(), s)
    b = sb.PTSSKOptModel(geo, PTSSK_DEFAULT)
    b.initialize_cell_environment(ta)
    b.set_states(st0)
    b.run_windowed(sb.InterpolationParameter(), env=env, window_steps=777)
    assert np.array_equal(b.catchment_discharges(), cd) and np.array_equal(b.get_states(), s)


def test_pt_ss_k_reference_known_answers_on_the_device(sb):
    """test/pt_ss_k_test.cpp:118-168 (lake / reservoir response) with the reference's own asserts (tests/stack_cases.py)"""
    models = {}

    def run(stack, geo, par, forcing, state, t0_us, T):
        assert stack == 3
        key = (geo.tobytes(), T)
        if key not in models:
            g = geo[0]
            cells = sb.geo_cell_data_vector([g[0]], [g[1]], [g[2]], area=g[3], catchment_id=np.array([int(g[4])]), radiation_slope_factor=g[5],
                                            glacier=g[6], lake=g[7], reservoir=g[8], forest=g[9])
            m = sb.PTSSKModel(cells, par)
            m.initialize_cell_environment(sb.TimeAxis(t0_us // 10**6, 3600, T))
            m.set_state_collection(-1, True)
            models[key] = m
        m = models[key]
        m.set_region_parameter(par)
        for k in FORCING:
            m.set_cell_forcing(k, forcing[k])
        m.set_states(state)
        m.run_cells()
        out = {name: m.response(name) for name in ("avg_discharge", "snow_swe", "snow_sca", "snow_outflow")}
        out["state_snow_swe"] = m.state_series("snow_swe")
        out["state"] = m.get_states()
        return out
    sc.ptssk_lake_reservoir_response(run)


def test_pt_ss_k_through_the_pybind_module_and_statistics(sb):
    from shyft_b200 import _build, synthetic
    _build.build_pybind_module()
    from shyft_b200 import _shyft_b200_cpp as cpp
    n, T = 64, 240
    geo, ta, env = synthetic.make_region(n, T, 9, config_index=2, cells_per_catchment=32, start=1417392000)
    a = cpp.PTSSKModel(geo, list(PTSSK_DEFAULT))
    envd = {k: getattr(env, k) for k in sb.capi.FORCING_NAMES}
    assert a.run_interpolation(ta.start * 10**6, 3600 * 10**6, T, envd)
    a.set_states(synthetic.default_state(3, n))
    a.run_cells()
    b = sb.PTSSKModel(geo, PTSSK_DEFAULT)
    b.run_interpolation(sb.InterpolationParameter(), ta, env)
    b.set_states(synthetic.default_state(3, n))
    b.set_state_collection(-1, True)
    b.run_cells()
    assert np.array_equal(a.catchment_discharges()[0], b.catchment_discharges()[:, 0])
    assert np.array_equal(b.statistics.discharge([1]), b.response("avg_discharge")[:, :32].sum(axis=1)) or \
        np.allclose(b.statistics.discharge([1]), b.response("avg_discharge")[:, :32].sum(axis=1), rtol=1e-12)
    assert b.kirchner_state.discharge([1]).shape == (T + 1,)
