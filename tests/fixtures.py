"""Fixtures shared by the CPU (oracle) and GPU (C-ABI) tests."""
import calendar as pycal

import numpy as np

HOUR_US = 3600 * 10**6
PTGSK_DEFAULT = np.array([-2.439, 0.966, -0.10, 1.5, -0.5, 2.0, 0.1, 1.0, 5.0, 5.0, 30.0, 0.9, 0.6, 5.0, 0.4, 0.4, 1.0, 0.0, 0.0, 0.2, 1.26, 0.04, 100.0,
                          0.0, 6.0, 1.0, 7.0, 0.0, 221.0, 0.0, 1.0])
PTHSK_DEFAULT = np.array([-2.439, 0.966, -0.10, 1.5, 0.1, 0.0, 1.0, 0.0, 0.5, 6.0, 1.0, 0.2, 1.26, 1.0, 7.0, 0.0, 0.0, 1.0])
# pt_ss_k::parameter::get order (core/pt_ss_k.h:90-117): kirchner c1 c2 c3, ae scale, ss alpha_0 d_range unit_size max_water_fraction tx cx ts cfr,
# p_corr, pt albedo alpha, gm.dtf, routing velocity alpha beta, gm.direct_response, msp.reservoir_direct_response_fraction
PTSSK_DEFAULT = np.array([-2.439, 0.966, -0.10, 1.5, 40.77, 113.0, 0.1, 0.1, 0.16, 2.5, 0.14, 0.01, 1.0, 0.2, 1.26, 6.0, 1.0, 7.0, 0.0, 0.0, 1.0])
# pt_hps_k::parameter::get order (core/pt_hps_k.h:93-120): kirchner c1 c2 c3, ae scale, hps lw tx cfr wind_scale wind_const surface_magnitude max_albedo
# min_albedo fast_albedo_decay_rate slow_albedo_decay_rate snowfall_reset_depth calculate_iso_pot_energy, gm.dtf, p_corr, pt albedo alpha, routing
# velocity alpha beta, msp.reservoir_direct_response_fraction
PTHPSK_DEFAULT = np.array([-2.439, 0.966, -0.10, 1.5, 0.1, 0.0, 0.5, 2.0, 1.0, 30.0, 0.9, 0.6, 5.0, 5.0, 5.0, 0.0, 6.0, 1.0, 0.2, 1.26, 1.0, 7.0, 0.0, 1.0])
HBV_DEFAULT = np.array([300.0, 2.0, 150.0, 25.0, 0.5, 0.3, 0.8, 0.02, 0.1, 0.0, 1.0, 0.0, 0.5, 1.0, 0.2, 1.26, 6.0, 1.0, 7.0, 0.0, 0.0, 1.0])
FORCING = ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum")
GEO_COLS = ("x", "y", "z", "area", "catchment_id", "radiation_slope_factor", "glacier", "lake", "reservoir", "forest", "routing_id", "routing_distance")


def geo_matrix(geo_records):
    """numpy record array (shyft_b200.capi.GEO_DTYPE) -> the oracle's [n][12] double matrix"""
    return np.stack([geo_records[c].astype(np.float64) for c in GEO_COLS], axis=1)


def py_region_fixture():
    """The region of shyft/tests/api/test_region_model_stacks.py:16-58,145-218: 20 cells x 240 h, one station per variable
    at the mid point of cell 10, constant forcing (P 5, T 10, wind 2, rh 0.7, rad 300), q0 = 40 mm/h."""
    n = 20
    geo = np.zeros((n, 12))
    for i in range(n):
        geo[i] = [500 + 1000.0 * i, 500.0, 500.0 * i / n, 1000.0 * 1000.0, 1, 0.9, 0.01, 0.05, 0.19, 0.3, 0, 0.0]
    par = PTGSK_DEFAULT.copy()
    par[17] = 0.1      # gs.snow_cv_forest_factor
    par[18] = 0.0001   # gs.snow_cv_altitude_factor
    t0 = pycal.timegm((2015, 1, 1, 0, 0, 0))
    T = 240
    station = geo[n // 2, :3].copy()
    consts = dict(temperature=10.0, precipitation=5.0, radiation=300.0, wind_speed=2.0, rel_hum=0.7)
    state = np.tile(np.array([0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, 40.0]), (n, 1))
    return dict(geo=geo, par=par, t0=t0, dt=3600, T=T, station=station, consts=consts, state=state)


def oracle_interpolate_py_fixture(oracle, fx):
    """interpolate() of the fixture through the oracle: single temperature source -> clean copy (region_model.h:470-481),
    IDW with one source for the rest."""
    n, T = fx["geo"].shape[0], fx["T"]
    src = fx["station"][None, :]
    dst = fx["geo"][:, :3]
    f = {}
    f["temperature"] = np.full((T, n), fx["consts"]["temperature"])
    for name, mm in (("precipitation", 20), ("radiation", 10), ("wind_speed", 10), ("rel_hum", 10)):
        vals = oracle.average_accessor_same_axis(np.full((T, 1), fx["consts"][name]), fx["dt"] * 10**6)
        f[name] = oracle.idw_run(name, src, vals, dst, oracle.idw_par(max_members=mm), dst_slope=fx["geo"][:, 5])
    return f
