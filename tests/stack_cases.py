"""The reference's stack-level known-answer cases (test/pt_gs_k_test.cpp:174-354, test/pt_hs_k_test.cpp:93-153, test/pt_ss_k_test.cpp:118-168,
test/pt_hps_k_test.cpp:43-156), restated once and
driven through a `run(stack, geo [1][12], params, forcing dict of [T][1], state [1][k], t0_us, T) -> dict` callable, so that the
CPU oracle (tests/test_oracle_stack_known_answers.py) and the CUDA path through the C ABI (tests/test_gpu_stack_known_answers.py)
are held to the same asserts."""
import calendar as pycal

import numpy as np
import pytest

from fixtures import PTGSK_DEFAULT, PTHPSK_DEFAULT, PTHSK_DEFAULT, PTSSK_DEFAULT

T0 = pycal.timegm((2014, 8, 1, 0, 0, 0)) * 10**6
AREA = 1000.0 * 1000.0
DT_S = 3600.0
MMH_TO_M3S = AREA / (3600.0 * 1000.0)   # mmh_to_m3s(1, cell_area)


def approx(value, epsilon):
    """doctest::Approx(value).epsilon(e): |lhs - value| < e * (1 + max(|lhs|, |value|)) -- an absolute floor of e comes with it"""
    return pytest.approx(value, abs=epsilon * (1.0 + abs(value)))


def geo_cell(glacier=0.0, lake=0.0, reservoir=0.0, forest=0.0):
    # geo_cell_data(geo_point(1000, 1000, 100)): area 1e6, catchment id -1, radiation_slope_factor 0.9 (core/geo_cell_data.h:107-115)
    return np.array([[1000.0, 1000.0, 100.0, AREA, -1, 0.9, glacier, lake, reservoir, forest, 0, 0.0]])


def forcing(T, temp, prec, first_prec=None):
    f = dict(temperature=np.full((T, 1), temp), precipitation=np.full((T, 1), prec), radiation=np.full((T, 1), 300.0),
             wind_speed=np.full((T, 1), 2.0), rel_hum=np.full((T, 1), 0.8))
    if first_prec is not None:
        f["precipitation"][0, 0] = first_prec
    return f


def gs_state(lwc=0.1, acc_melt=0.0, temp_swe=0.0, q=5.0):
    # gamma_snow::state defaults (albedo 0.4, lwc 0.1, surface_heat 30000, alpha 1.26, sdc_melt_mean 0, acc_melt 0, iso_pot_energy 0,
    # temp_swe 0; core/gamma_snow.h:114-121) + kirchner.q
    return np.array([[0.4, lwc, 30000.0, 1.26, 0.0, acc_melt, 0.0, temp_swe, q]])


def ptgsk_mass_balance(run):
    """test_mass_balance: 15 degC, 3 mm/h, no snow; the same hour 10 001 times -> discharge + evapotranspiration = precipitation"""
    par = PTGSK_DEFAULT.copy()
    geo = geo_cell()
    st = gs_state(lwc=0.0, acc_melt=-1.0, q=5.0)
    f = forcing(1, 15.0, 3.0)
    out = None
    for _ in range(10001):
        out = run(0, geo, par, f, st, T0, 1)
        st = out["state"]
    assert out["avg_discharge"][0, 0] * DT_S * 1000 / AREA + out["ae_output"][0, 0] == pytest.approx(3.0, abs=1e-7)
    assert out["snow_outflow"][0, 0] * DT_S * 1000 / AREA == pytest.approx(3.0, abs=1e-7)
    return st


def ptgsk_direct_response_on_reservoir_only(run, st):
    par = PTGSK_DEFAULT.copy()
    geo = geo_cell(lake=0.5, reservoir=0.5)
    st = st.copy()
    st[0, 8], st[0, 1], st[0, 5] = 1e-4, 0.0, -1.0
    out = run(0, geo, par, forcing(1, 15.0, 3.0), st, T0, 1)
    assert out["avg_discharge"][0, 0] * DT_S * 1000.0 / AREA == approx(0.5 * 3.0, 0.001)
    st = out["state"]
    st[0, 1], st[0, 5] = 1.0, 300.0
    out = run(0, geo, par, forcing(1, -10.0, 3.0), st, T0, 1)
    assert out["snow_sca"][0, 0] == pytest.approx(0.96, abs=0.01)
    assert out["avg_discharge"][0, 0] * DT_S * 1000.0 / AREA == approx(0.5 * 3.0, 0.05)
    st = out["state"]
    st[0, 5], st[0, 7], st[0, 1] = 5.0, 3.0, 10.0
    f = forcing(1, 10.0, 3.0)
    for _ in range(5000):
        out = run(0, geo, par, f, st, T0, 1)
        st = out["state"]
        if out["snow_sca"][0, 0] < 0.1:
            break
    assert out["snow_sca"][0, 0] == pytest.approx(0.0, abs=0.1)
    assert out["avg_discharge"][0, 0] == approx(0.5 * 0.8333, 0.001)   # "empirical": the reference's own output


def ptgsk_glacier_and_reservoir_direct_response(run, st):
    f = forcing(1, 15.0, 3.0)
    for kind in ("glacier", "reservoir"):
        geo = geo_cell(glacier=0.5) if kind == "glacier" else geo_cell(reservoir=0.5)
        s = st.copy()
        s[0, 8], s[0, 1], s[0, 5] = 1e-4, 0.0, -1.0
        for frac, scale in ((1.0, 1.0), (0.5, 0.5)):
            par = PTGSK_DEFAULT.copy()
            par[29 if kind == "glacier" else 30] = frac      # gm.direct_response / msp.reservoir_direct_response_fraction
            out = run(0, geo, par, f, s, T0, 1)
            s = out["state"]
            gm = out["glacier_melt"][0, 0] if kind == "glacier" else 0.0
            expected = scale * (0.5 * 3.0 * MMH_TO_M3S + gm) if kind == "glacier" else scale * 0.5 * 3.0 * MMH_TO_M3S
            if kind == "glacier" and frac == 1.0:
                expected = 0.5 * 3.0 * MMH_TO_M3S + gm
            assert out["avg_discharge"][0, 0] == approx(expected, 0.001), (kind, frac)
        par = PTGSK_DEFAULT.copy()
        par[29 if kind == "glacier" else 30] = 0.0
        out = run(0, geo, par, f, s, T0, 1)
        assert out["avg_discharge"][0, 0] == approx(2.778e-5, 0.01e-5), kind


def ptgsk_lake_reservoir_response(run):
    """ptgsk_lake_reservoir_response: -15 degC, 3 mm/h from the second hour, 20 % lake, 30 % reservoir"""
    n = 50
    geo = geo_cell(lake=0.2, reservoir=0.3)
    f = forcing(n, -15.0, 3.0, first_prec=0.0)
    st = gs_state(lwc=100.0, acc_melt=100.0, q=1.0)
    par = PTGSK_DEFAULT.copy()
    par[30] = 0.0
    out = run(0, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266, 0.01)
    assert out["avg_discharge"][n - 1, 0] == approx(0.5 * 3.0 * MMH_TO_M3S, 0.01)
    par[30] = 1.0
    out = run(0, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266 * 0.7, 0.01)
    assert out["snow_swe"][0, 0] == pytest.approx(0.0, abs=1e-3)
    assert out["snow_swe"][1, 0] == approx(1.548, 0.01)
    assert out["snow_swe"][2, 0] == approx(3.048, 0.01)
    assert out["avg_discharge"][1, 0] == approx(0.266 + 0.3 * 0.5 * 3.0 * MMH_TO_M3S, 0.05)
    assert out["avg_discharge"][n - 1, 0] == approx(0.2 * 3.0 * MMH_TO_M3S * (1.0 - 0.3) + 0.3 * 3.0 * MMH_TO_M3S, 0.01)


def pthsk_lake_reservoir_response(run):
    """pt_hs_k_lake_reservoir_response: the same story with hbv_snow; swe 0 / 1.5 / 3.0 over the cell"""
    n = 50
    geo = geo_cell(lake=0.2, reservoir=0.3)
    f = forcing(n, -15.0, 3.0, first_prec=0.0)
    st = np.zeros((1, 13))
    st[0, 12] = 1.0
    par = PTHSK_DEFAULT.copy()
    par[17] = 0.0
    out = run(1, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266, 0.01)
    assert out["avg_discharge"][n - 1, 0] == approx(0.5 * 3.0 * MMH_TO_M3S, 0.01)
    par[17] = 1.0
    out = run(1, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266 * 0.7, 0.01)
    assert out["avg_discharge"][1, 0] == approx(0.266 + 0.3 * 0.5 * 3.0 * MMH_TO_M3S, 0.05)
    assert out["snow_swe"][0, 0] == pytest.approx(0.0, abs=1e-4)
    assert out["snow_swe"][1, 0] == approx(1.5, 1e-4)
    assert out["snow_swe"][2, 0] == approx(3.0, 1e-4)
    if "state_snow_swe" in out:   # state collector: instant values at the beginning of each step
        assert out["state_snow_swe"][0, 0] == pytest.approx(0.0, abs=1e-4)
        assert out["state_snow_swe"][1, 0] == pytest.approx(0.0, abs=1e-4)
        assert out["state_snow_swe"][2, 0] == approx(1.5, 1e-4)
    assert out["avg_discharge"][n - 1, 0] == approx(0.2 * 3.0 * MMH_TO_M3S * (1.0 - 0.3) + 0.3 * 3.0 * MMH_TO_M3S, 0.01)


def ptssk_lake_reservoir_response(run):
    """pt_ss_k_lake_reservoir_response (test/pt_ss_k_test.cpp:118-168): the same story with the Skaugen routine; the state collector's
    snow_swe (instant, over the cell) is 0 / 0 / 1.5 / 3.0 at the first four points"""
    n = 50
    geo = geo_cell(lake=0.2, reservoir=0.3)
    f = forcing(n, -15.0, 3.0, first_prec=0.0)
    st = np.array([[4.077, 40.77, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]])   # skaugen::state() + kirchner q = 1
    par = PTSSK_DEFAULT.copy()
    par[20] = 0.0
    out = run(3, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266, 0.01)
    assert out["avg_discharge"][n - 1, 0] == approx(0.5 * 3.0 * MMH_TO_M3S, 0.01)
    par[20] = 1.0
    out = run(3, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266 * 0.7, 0.01)
    assert out["avg_discharge"][1, 0] == approx(0.266 + 0.3 * 0.5 * 3.0 * MMH_TO_M3S, 0.05)
    assert out["state_snow_swe"][0, 0] == approx(0.0, 0.001)
    assert out["state_snow_swe"][1, 0] == approx(0.0, 0.001)
    assert out["state_snow_swe"][2, 0] == approx(1.5, 0.001)
    assert out["state_snow_swe"][3, 0] == approx(3.0, 0.001)
    assert out["avg_discharge"][n - 1, 0] == approx(0.2 * 3.0 * MMH_TO_M3S * (1.0 - 0.3) + 0.3 * 3.0 * MMH_TO_M3S, 0.01)
    assert np.all(np.isfinite(out["snow_swe"])) and np.all(out["snow_swe"] >= 0.0)   # test_call_stack's assert


def hps_state(swe=0.0, sca=0.0, q=1.0, albedo=0.4, iso=0.0, surface_heat=30000.0, sp=None, sw=None):
    """hbv_physical_snow::state flat: sp[5], sw[5], albedo[5], iso_pot_energy[5], surface_heat, swe, sca + kirchner.q"""
    sp = [0.0] * 5 if sp is None else list(sp)
    sw = [0.0] * 5 if sw is None else list(sw)
    return np.array([sp + sw + [albedo] * 5 + [iso] * 5 + [surface_heat, swe, sca, q]])


def pthpsk_lake_reservoir_response(run):
    """pt_hps_k_lake_reservoir_response (test/pt_hps_k_test.cpp:97-156): the same story with hbv_physical_snow; the response swe (stair) is
    0 / 1.5 / 3.0, the state collector's (instant) 0 / 0 / 1.5"""
    n = 50
    geo = geo_cell(lake=0.2, reservoir=0.3)
    f = forcing(n, -15.0, 3.0, first_prec=0.0)
    st = hps_state()
    par = PTHPSK_DEFAULT.copy()
    par[23] = 0.0
    out = run(4, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266, 0.01)
    assert out["avg_discharge"][n - 1, 0] == approx(0.5 * 3.0 * MMH_TO_M3S, 0.01)
    par[23] = 1.0
    out = run(4, geo, par, f, st, T0, n)
    assert out["avg_discharge"][0, 0] == approx(0.266 * 0.7, 0.01)
    assert out["avg_discharge"][1, 0] == approx(0.266 + 0.3 * 0.5 * 3.0 * MMH_TO_M3S, 0.05)
    if "state_snow_swe" in out:
        assert out["state_snow_swe"][0, 0] == pytest.approx(0.0, abs=1e-4)
        assert out["state_snow_swe"][1, 0] == pytest.approx(0.0, abs=1e-4)
        assert out["state_snow_swe"][2, 0] == approx(1.5, 1e-4)
    assert out["snow_swe"][0, 0] == pytest.approx(0.0, abs=1e-4)
    assert out["snow_swe"][1, 0] == approx(1.5, 1e-4)
    assert out["snow_swe"][2, 0] == approx(3.0, 1e-4)
    assert out["avg_discharge"][n - 1, 0] == approx(0.2 * 3.0 * MMH_TO_M3S * (1.0 - 0.3) + 0.3 * 3.0 * MMH_TO_M3S, 0.01)
    assert np.all(np.isfinite(out["snow_swe"])) and np.all(out["snow_swe"] >= 0.0)   # test_call_stack's assert
