"""The oracle against the reference's stack-level known answers: test/pt_gs_k_test.cpp:174-354 (mass balance, land-type
routing of the response, lake / reservoir), test/pt_hs_k_test.cpp:93-153, test/hbv_snow_test.cpp:78-138 (sca after a snowfall for three
redistribution vectors)."""
import numpy as np
import pytest

import stack_cases as sc


@pytest.fixture(scope="module")
def run(oracle):
    def f(stack, geo, par, forcing, state, t0_us, T):
        if stack == 4:
            return oracle.pthpsk_run_cells(geo, par, forcing, state, t0_us, 3600 * 10**6)
        if stack == 3:
            return oracle.ptssk_run_cells(geo, par, forcing, state, t0_us, 3600 * 10**6, collect_state=True)
        fn = oracle.ptgsk_run_cells if stack == 0 else oracle.pthsk_run_cells
        return fn(geo, par, forcing, state, t0_us, 3600 * 10**6)
    return f


def test_pt_gs_k_mass_balance_and_land_type_routing(run):
    st = sc.ptgsk_mass_balance(run)
    sc.ptgsk_direct_response_on_reservoir_only(run, st)
    sc.ptgsk_glacier_and_reservoir_direct_response(run, st)


def test_pt_gs_k_lake_reservoir_response(run):
    sc.ptgsk_lake_reservoir_response(run)


def test_pt_hs_k_lake_reservoir_response(run):
    sc.pthsk_lake_reservoir_response(run)


def test_pt_ss_k_lake_reservoir_response(run):
    sc.ptssk_lake_reservoir_response(run)


def test_pt_hps_k_lake_reservoir_response(run):
    sc.pthpsk_lake_reservoir_response(run)


def test_pt_hps_k_call_stack(oracle):
    """test/pt_hps_k_test.cpp:43-95: three days from HbvPhysicalSnowState(albedo 0.4, iso 0, surface_heat 30000, swe 10, sca 0.5), q = 5; the state's
    empty bin vectors are filled by distribute() at the top of run() -- here through the oracle's hbv_snow distribute"""
    T = 72
    sp, sw, swe, sca = oracle.hbv_snow_distribute(10.0, 0.5, [1.0] * 5, [0.0, 0.25, 0.5, 0.75, 1.0], lw=0.1)
    st = sc.hps_state(swe=swe, sca=sca, q=5.0, sp=sp, sw=sw)
    h = np.arange(T)
    f = dict(temperature=(-5.0 + 10.0 * np.sin(2 * np.pi * h / 24.0))[:, None], precipitation=np.where(h % 7 == 0, 3.0, 0.0)[:, None].astype(float),
             radiation=np.maximum(0.0, 300.0 * np.sin(2 * np.pi * (h - 6) / 24.0))[:, None], wind_speed=np.full((T, 1), 2.0), rel_hum=np.full((T, 1), 0.7))
    out = oracle.pthpsk_run_cells(sc.geo_cell(), sc.PTHPSK_DEFAULT, f, st, sc.T0, 3600 * 10**6)
    assert np.all(np.isfinite(out["snow_swe"])) and np.all(out["snow_swe"] >= 0.0)
    assert np.all(np.isfinite(out["avg_discharge"])) and out["state"].shape == (1, 24)


@pytest.mark.parametrize("T,prec,swe,sca,rain_no_snow", [(1.0, 0.04, 0.05, 1.0, False), (-1.0, 0.15, 0.2, 0.6, False), (0.0, 0.15, 0.2, 0.6, False),
                                                         (0.0, 0.15, 0.0, 0.0, True), (3.0, 0.0, 10.0, 0.5, False)])
def test_hbv_physical_snow_mass_balance(oracle, T, prec, swe, sca, rain_no_snow):
    """test/hbv_physical_snow_test.cpp:29-176: snow pack reset, build-up (and at T = tx), rain without snow, melt without precipitation"""
    st = np.array([0.0] * 10 + [0.6] * 5 + [1752.56396484375] * 5 + [0.0, swe, sca])
    s1, r = oracle.hps_step(st, T, 10.0, prec, 2.0, 0.70, distribute=True)
    assert s1[21] + r[0] == pytest.approx(prec + swe, abs=1e-8)
    if rain_no_snow:
        assert s1[22] == pytest.approx(0.0, abs=1e-8) and s1[21] == pytest.approx(0.0, abs=1e-8)


def test_skaugen_known_answers(oracle):
    """test/skaugen_test.cpp:9-281: accumulation, melt with mass balance, liquid water, the melt-down regression"""
    alpha_0, unit = 40.77, 0.1
    s0 = np.array([alpha_0 * unit, alpha_0, 0.0, 0.0, 0.0, 0.0, 0.0])
    s = s0.copy()
    for _ in range(10):                                   # test_accumulation: ten hours of 10 mm/h at -10 degC
        s, r = oracle.skaugen_step(s, -10.0, 10.0)
    assert s[3] * s[2] == pytest.approx(100.0, abs=1e-6) and s[2] == pytest.approx(1.0, abs=1e-6) and s[0] < alpha_0 * unit
    day = 24 * 3600 * 10**6
    s = s0.copy()
    for _ in range(10):                                   # test_melt: ten days of 10 mm/day, then +10 degC
        s, r = oracle.skaugen_step(s, -10.0, 10.0 / 24.0, dt_us=day)
    total_water = s[3] * s[2]
    s, r = oracle.skaugen_step(s, 10.0, 0.0, dt_us=day)
    agg = r[0] * 24.0
    assert s[2] * (s[3] + s[4]) < total_water
    assert r[0] * 24.0 + s[4] >= 1.0
    assert r[0] * 24.0 + s[2] * (s[4] + s[3]) == pytest.approx(total_water, abs=1e-6)
    for _ in range(100):
        s, r = oracle.skaugen_step(s, 10.0, 0.0, dt_us=day)
        agg += r[0] * 24.0
    assert s[2] == pytest.approx(0.0, abs=1e-6) and s[3] == pytest.approx(0.0, abs=1e-6)
    assert agg == pytest.approx(total_water, abs=1e-10)
    assert s[1] == pytest.approx(alpha_0, abs=1e-6) and s[0] == pytest.approx(alpha_0 * unit, abs=1e-6)
    s = s0.copy()
    for _ in range(10):                                   # test_lwc
        s, r = oracle.skaugen_step(s, -10.0, 10.0 / 24.0, dt_us=day)
    assert s[4] == pytest.approx(0.0, abs=1e-6)
    s, r = oracle.skaugen_step(s, 10.0, 0.0, dt_us=day)
    assert s[4] <= s[3] * 0.1
    for _ in range(5):
        s, r = oracle.skaugen_step(s, 2.0, 0.0, dt_us=day)
    assert s[4] == pytest.approx(s[3] * 0.1, abs=1e-6)
    # skaugen_meltdown: the state of 2017-07-17T22:00 must step without throwing
    s = np.array([0.012785227731289801, 0.127852277312898, 0.005033599471562574, 32.1, 3.21, 0.0, 321.0])
    s, r = oracle.skaugen_step(s, 4.891358376624782, 0.0010356738461072955, dt_us=3 * 3600 * 10**6)
    assert s[3] == 0.0 and s[2] == 0.0 and np.isfinite(r[0])


@pytest.mark.parametrize("s,sca_after", [([1.0, 1.0, 1.0, 0.0, 0.0], 0.75), ([1.0, 1.0, 1.0, 1.0, 1.0], 1.0), ([1.0, 0.0, 0.0, 0.0, 0.0], 0.25)])
def test_hbv_snow_sca_at_snowpack_buildup(oracle, s, sca_after):
    """test_snow_distr / uniform / skewed _at_snowpack_buildup: the explicit parameter(s, intervals) constructor does not normalise s"""
    iv = [0.0, 0.25, 0.5, 0.75, 1.0]
    sp, sw, swe, sca = oracle.hbv_snow_distribute(10.0, 0.15, s, iv)
    sp, sw, swe, sca, out = oracle.hbv_snow_step(sp, sw, swe, sca, 0.15, -1.0, s=s, intervals=iv)
    assert sca == pytest.approx(sca_after, abs=1e-8)


def test_pt_hps_k_oracle_properties(oracle):
    """Properties of the restated stack that hold by construction of the reference's code (core/pt_hps_k.h:201-303, hbv_physical_snow.h:266-529):
    a run in chunks carries its state exactly; iso_pot_energy is a diagnostic (switching it on changes that state and nothing else); without
    precipitation and without a pack the snow routine passes nothing on; water is conserved over the snow routine step by step"""
    T = 24 * 20
    h = np.arange(T)
    rng = np.random.default_rng(5)
    f = dict(temperature=(-6.0 + 9.0 * np.sin(2 * np.pi * h / (24.0 * 9)) + rng.normal(0, 1.0, T))[:, None],
             precipitation=(rng.exponential(1.2, T) * (rng.random(T) < 0.25))[:, None],
             radiation=np.maximum(0.0, 250.0 * np.sin(2 * np.pi * (h - 6) / 24.0))[:, None], wind_speed=np.full((T, 1), 2.5), rel_hum=np.full((T, 1), 0.75))
    geo = sc.geo_cell(lake=0.1, reservoir=0.1, glacier=0.05)
    st0 = sc.hps_state(q=0.5)
    par = sc.PTHPSK_DEFAULT.copy()
    dt = 3600 * 10**6
    one = oracle.pthpsk_run_cells(geo, par, f, st0, sc.T0, dt)
    assert np.nanmax(one["snow_swe"]) > 1.0
    # chunked = one shot
    st, q = st0.copy(), []
    for k in range(0, T, 96):
        part = oracle.pthpsk_run_cells(geo, par, f, st, sc.T0, dt, start_step=k, n_steps=min(96, T - k))
        st = part["state"]
        q.append(part["avg_discharge"][k:k + 96])
    assert np.array_equal(np.concatenate(q), one["avg_discharge"]) and np.array_equal(st, one["state"])
    # iso_pot_energy: a diagnostic
    par_iso = par.copy()
    par_iso[15] = 1.0
    iso = oracle.pthpsk_run_cells(geo, par_iso, f, st0, sc.T0, dt)
    assert np.array_equal(iso["avg_discharge"], one["avg_discharge"]) and np.array_equal(iso["snow_swe"], one["snow_swe"])
    assert np.any(iso["state"][0, 15:20] != 0.0) and np.all(one["state"][0, 15:20] == 0.0)
    assert np.array_equal(np.delete(iso["state"], np.s_[15:20], axis=1), np.delete(one["state"], np.s_[15:20], axis=1))
    # dry and bare: nothing leaves the snow routine
    f_dry = {k: v.copy() for k, v in f.items()}
    f_dry["precipitation"][:] = 0.0
    dry = oracle.pthpsk_run_cells(geo, par, f_dry, st0, sc.T0, dt)
    assert np.all(dry["snow_outflow"] == 0.0) and np.all(dry["snow_swe"] == 0.0) and np.all(dry["snow_sca"] == 0.0)
    # water balance of the snow routine, step by step: swe_before + prec = swe_after + outflow (response swe and outflow both scaled by the
    # snow storage fraction 0.8; mm per 1 h step)
    frac = 0.8
    swe = np.concatenate([[0.0], one["snow_swe"][:, 0]]) / frac
    bal = swe[:-1] + f["precipitation"][:, 0] * par[17] - swe[1:] - one["snow_outflow"][:, 0] / frac
    assert np.max(np.abs(bal)) < 1e-8


def test_pt_ss_k_oracle_properties(oracle):
    """pt_ss_k restated (core/pt_ss_k.h:195-293): a run in chunks carries its state exactly (num_units travels as a double in the flat state);
    without precipitation and without a pack the Skaugen routine passes nothing on"""
    T = 24 * 20
    h = np.arange(T)
    rng = np.random.default_rng(6)
    f = dict(temperature=(-6.0 + 9.0 * np.sin(2 * np.pi * h / (24.0 * 9)) + rng.normal(0, 1.0, T))[:, None],
             precipitation=(rng.exponential(1.2, T) * (rng.random(T) < 0.25))[:, None],
             radiation=np.maximum(0.0, 250.0 * np.sin(2 * np.pi * (h - 6) / 24.0))[:, None], wind_speed=np.full((T, 1), 2.5), rel_hum=np.full((T, 1), 0.75))
    geo = sc.geo_cell(lake=0.1, reservoir=0.1, glacier=0.05)
    st0 = np.array([[4.077, 40.77, 0.0, 0.0, 0.0, 0.0, 0.0, 0.5]])
    par = sc.PTSSK_DEFAULT.copy()
    dt = 3600 * 10**6
    one = oracle.ptssk_run_cells(geo, par, f, st0, sc.T0, dt)
    assert np.nanmax(one["snow_swe"]) > 1.0
    st, q = st0.copy(), []
    for k in range(0, T, 96):
        part = oracle.ptssk_run_cells(geo, par, f, st, sc.T0, dt, start_step=k, n_steps=min(96, T - k))
        st = part["state"]
        q.append(part["avg_discharge"][k:k + 96])
    assert np.array_equal(np.concatenate(q), one["avg_discharge"]) and np.array_equal(st, one["state"])
    f_dry = {k: v.copy() for k, v in f.items()}
    f_dry["precipitation"][:] = 0.0
    dry = oracle.ptssk_run_cells(geo, par, f_dry, st0, sc.T0, dt)
    assert np.all(dry["snow_outflow"] == 0.0) and np.all(dry["snow_swe"] == 0.0) and np.all(dry["snow_sca"] == 0.0)
