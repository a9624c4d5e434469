"""GPU parity of region_model::interpolate (IDW for five variables, Bayesian temperature kriging) through the C ABI."""
import numpy as np
import pytest

from fixtures import FORCING, geo_matrix
from parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import shyft_b200
    return shyft_b200


def _region(sb, n, T, S, **kw):
    from shyft_b200 import synthetic
    return synthetic.make_region(n, T, S, **kw)


def test_idw_neighbour_lists_and_values_all_variables(sb, oracle):
    n, T, S = 1500, 96, 64
    geo, ta, env = _region(sb, n, T, S, config_index=11)
    gm = geo_matrix(geo)
    m = sb.PTGSKOptModel(geo)
    ip = sb.InterpolationParameter(use_idw_for_temperature=1)
    assert m.run_interpolation(ip, ta, env)
    for name in FORCING:
        xyz, vals = getattr(env, name)
        mm = 20 if name in ("temperature", "precipitation") else 10
        want = oracle.idw_run(name, xyz, oracle.average_accessor_same_axis(vals, 3600 * 10**6), gm[:, :3], oracle.idw_par(max_members=mm),
                              dst_slope=gm[:, 5], ncore=4)
        assert_parity(m.cell_forcing(name), want, "idw " + name, rtol=1e-12)


def test_idw_with_missing_values_and_gradient_by_equation(sb, oracle):
    n, T, S = 400, 200, 25
    geo, ta, env = _region(sb, n, T, S, config_index=12, nan_fraction=0.05)
    gm = geo_matrix(geo)
    m = sb.PTGSKOptModel(geo)
    ip = sb.InterpolationParameter(use_idw_for_temperature=1)
    ip.temperature_idw.gradient_by_equation = 1
    ip.temperature_idw.max_members = 6
    ip.precipitation.max_members = 4
    ip.precipitation.scale_factor = 1.05
    ip.wind_speed.max_distance = 9000.0   # some cells end up with no station in reach -> 0/0 = NaN like the reference
    ip.rel_hum.distance_measure_factor = 1.0
    ip.radiation.zscale = 5.0
    m.run_interpolation(ip, ta, env)
    pars = dict(temperature=oracle.idw_par(max_members=6, gradient_by_equation=True), precipitation=oracle.idw_par(max_members=4, scale_factor=1.05),
                wind_speed=oracle.idw_par(max_distance=9000.0), rel_hum=oracle.idw_par(distance_measure_factor=1.0), radiation=oracle.idw_par(zscale=5.0))
    for name in FORCING:
        xyz, vals = getattr(env, name)
        want = oracle.idw_run(name, xyz, oracle.average_accessor_same_axis(vals, 3600 * 10**6), gm[:, :3], pars[name], dst_slope=gm[:, 5])
        # the 3x3 gradient solve amplifies round-off by its condition number; everything else is a short weighted mean
        assert_parity(m.cell_forcing(name), want, "idw+nan " + name, rtol=1e-9 if name == "temperature" else 1e-12)
    assert not m.is_cell_env_ts_ok()


def test_btk_full_and_reduced_station_sets(sb, oracle):
    n, T, S = 900, 120, 36
    geo, ta, env = _region(sb, n, T, S, config_index=13)
    gm = geo_matrix(geo)
    xyz, vals = env.temperature
    vals = vals.copy()
    vals[10:30, 3] = np.nan          # one station drops out for a while
    vals[50:55, [0, 7, 20]] = np.nan  # three more later
    env.temperature = (xyz, vals)
    m = sb.PTGSKOptModel(geo)
    ip = sb.InterpolationParameter()
    assert m.run_interpolation(ip, ta, env)
    want = oracle.btk_run(xyz, oracle.average_accessor_same_axis(vals, 3600 * 10**6), gm[:, :3], ta.start * 10**6, 3600 * 10**6)
    assert_parity(m.cell_forcing("temperature"), want, "btk temperature", rtol=1e-9)


def test_btk_all_stations_missing_is_an_error_unless_best_effort(sb):
    n, T, S = 64, 24, 9
    geo, ta, env = _region(sb, n, T, S, config_index=14)
    xyz, vals = env.temperature
    vals = vals.copy()
    vals[5, :] = np.nan
    env.temperature = (xyz, vals)
    m = sb.PTGSKOptModel(geo)
    ip = sb.InterpolationParameter()
    assert m.run_interpolation(ip, ta, env, best_effort=True) is False          # swallowed, reported (region_model.h:517-526)
    assert np.all(np.isfinite(m.cell_forcing("precipitation")))
    with pytest.raises(RuntimeError, match="No valid sources for time period"):
        m.run_interpolation(ip, ta, env, best_effort=False)


def test_single_temperature_source_is_copied_and_missing_variable_stays_nan(sb):
    n, T = 50, 48
    geo, ta, env = _region(sb, n, T, 4, config_index=15)
    xyz, vals = env.temperature
    env.temperature = (xyz[:1], vals[:, :1])
    env.wind_speed = None
    m = sb.PTGSKOptModel(geo)
    m.run_interpolation(sb.InterpolationParameter(), ta, env)
    t = m.cell_forcing("temperature")
    assert np.array_equal(t, np.repeat(vals[:, :1] * 3600.0 / 3600.0, n, axis=1))
    assert np.all(np.isnan(m.cell_forcing("wind_speed")))
    # cell-major readback is the transpose
    assert np.array_equal(m.cell_forcing("temperature", layout=sb.capi.CELL_MAJOR), t.T)


def test_dense_tensor_core_idw_equals_per_neighbour_idw(sb):
    """All station values finite: IDW runs as one dense DMMA contraction; it must agree with the per-neighbour kernel (which is
    bit-identical to the reference's operation order) to round-off, including cells with no station in reach and both
    temperature-gradient regimes (height span above / below 50 m)."""
    n, T, S = 3000, 200, 40
    geo, ta, env = _region(sb, n, T, S, config_index=16)
    xyz = env.temperature[0].copy()
    xyz[: S // 2, 2] = 500.0 + np.arange(S // 2) * 0.5   # a flat cluster: cells near it see < 50 m height span -> default gradient
    env.temperature = (xyz, env.temperature[1])
    ip = sb.InterpolationParameter(use_idw_for_temperature=1)
    ip.temperature_idw.max_members = 5
    ip.wind_speed.max_distance = 12000.0    # some cells have no wind station -> NaN in both paths
    out = {}
    for dense in (True, False):
        m = sb.PTGSKOptModel(geo)
        m.set_idw_dense(dense)
        m.run_interpolation(ip, ta, env)
        out[dense] = {k: m.cell_forcing(k) for k in FORCING}
    assert np.isnan(out[False]["wind_speed"]).any() and not np.isnan(out[False]["wind_speed"]).all()
    for k in FORCING:
        assert_parity(out[True][k], out[False][k], "dense vs per-neighbour idw " + k, rtol=1e-13)


def test_sources_on_their_own_axis_are_projected_on_the_device(sb, oracle):
    """region_environment series at their native resolution.  3-hourly stair-case precipitation with gaps,
    20-minute linear (instant-value) temperature that starts late and ends early, irregular radiation: the device projection equals
    the oracle's average_accessor bit for bit, and interpolation from it equals interpolation from the pre-projected series."""
    from shyft_b200 import synthetic
    n, T, S = 600, 72, 6
    geo, ta, env0 = synthetic.make_region(n, T, S, config_index=14, cells_per_catchment=200)[:3]
    rng = np.random.default_rng(14)
    t0, dt = ta.start, ta.delta_t
    xyz = env0.temperature[0]
    # temperature: 20-minute instant values from t0+2h40 to t0+60h (NaN before the first point, NaN from the total period end on)
    tt = t0 + 9600 + 1200 * np.arange(172)
    tv = 5.0 + rng.normal(0, 2, (tt.size, S)).cumsum(axis=0) * 0.1
    tv[40:43, 2] = np.nan
    temp = sb.GeoPointSources(xyz, tt, tv, t_end=tt[-1] + 1200, point_fx="instant")
    # precipitation: 3-hourly stair-case with a gap
    pt = t0 - 3 * 3600 + 10800 * np.arange(30)
    pv = rng.exponential(1.0, (pt.size, S)) * (rng.random((pt.size, S)) < 0.4)
    pv[7, :] = np.nan
    prec = sb.GeoPointSources(xyz, pt, pv, point_fx="average")
    # radiation: irregular point times
    rt = np.sort(t0 + rng.choice(np.arange(0, 80 * 3600, 600), 90, replace=False))
    rv = rng.uniform(0, 400, (rt.size, S))
    rad = sb.GeoPointSources(xyz, rt, rv, t_end=rt[-1] + 3600, point_fx="average")
    env = sb.RegionEnvironment(temperature=temp, precipitation=prec, radiation=rad, wind_speed=env0.wind_speed, rel_hum=env0.rel_hum)
    m = sb.PTGSKModel(geo)
    ip = sb.InterpolationParameter(use_idw_for_temperature=1)
    m.run_interpolation(ip, ta, env, best_effort=True)
    US = 10**6
    projected = {}
    for name, src in (("temperature", temp), ("precipitation", prec), ("radiation", rad)):
        want = oracle.average_accessor(src.times_us, src.values, src.t_end_us, src.point_fx == "instant", t0 * US, dt * US, T)
        got = m.sources_on_model_axis(name)
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), name     # same operations in the same order
        projected[name] = want
    assert np.isnan(projected["temperature"][:2]).all() and np.isnan(projected["temperature"][60:]).all()
    assert np.isfinite(projected["temperature"][3:60, 0]).all()
    f_native = {k: m.cell_forcing(k) for k in ("temperature", "precipitation", "radiation")}
    # the same environment handed over pre-projected: identical cell forcing
    env_pre = sb.RegionEnvironment(temperature=(xyz, projected["temperature"]), precipitation=(xyz, projected["precipitation"]),
                                   radiation=(xyz, projected["radiation"]), wind_speed=env0.wind_speed, rel_hum=env0.rel_hum)
    m2 = sb.PTGSKModel(geo)
    m2.run_interpolation(ip, ta, env_pre, best_effort=True)
    for k, v in f_native.items():
        w = m2.cell_forcing(k)
        assert np.array_equal(np.isnan(v), np.isnan(w)) and np.allclose(v[~np.isnan(v)], w[~np.isnan(w)], rtol=1e-12, atol=0), k
    with pytest.raises(RuntimeError, match="strictly increasing"):
        bad = sb.GeoPointSources(xyz, [t0, t0], np.zeros((2, S)), t_end=t0 + 10)
        m._set_sources(sb.RegionEnvironment(temperature=bad))


def test_every_station_on_its_own_axis(sb, oracle):
    """vector<geo_point_ts>: each station's series has its own time axis, length, end and point interpretation (api/api.h:137-168);
    sb2_set_sources_on_axes projects them all in one launch, bit-identical to the oracle's average_accessor station by station."""
    from shyft_b200 import synthetic
    n, T, S = 300, 96, 5
    geo, ta, env0 = synthetic.make_region(n, T, S, config_index=15, cells_per_catchment=100)[:3]
    rng = np.random.default_rng(15)
    t0, dt = ta.start, ta.delta_t
    xyz = env0.temperature[0]
    stations = []
    for s in range(S):
        step = (600, 3600, 10800, 1800, 86400)[s]
        start = t0 + (-7200, 0, -10800, 5400, -86400)[s]
        npts = (700, 80, 40, 150, 6)[s]
        times = start + step * np.arange(npts)
        if s == 3:  # irregular
            times = np.sort(t0 + rng.choice(np.arange(0, 90 * 3600, 300), npts, replace=False))
        vals = 2.0 + rng.normal(0, 1.0, npts).cumsum() * 0.2
        if s == 1:
            vals[10:14] = np.nan
        stations.append((xyz[s], times, vals, None if s != 3 else times[-1] + 900, "instant" if s in (0, 3) else "average"))
    src = sb.GeoPointSourceVector(stations)
    env = sb.RegionEnvironment(temperature=src, precipitation=env0.precipitation, radiation=env0.radiation, wind_speed=env0.wind_speed,
                               rel_hum=env0.rel_hum)
    m = sb.PTGSKModel(geo)
    ip = sb.InterpolationParameter(use_idw_for_temperature=1)
    m.run_interpolation(ip, ta, env, best_effort=True)
    got = m.sources_on_model_axis("temperature")
    US = 10**6
    want = np.zeros((T, S))
    off = np.concatenate([[0], np.cumsum(src.n_points)])
    for s in range(S):
        want[:, s] = oracle.average_accessor(src.times_us[off[s]:off[s + 1]], src.values[off[s]:off[s + 1], None], int(src.t_end_us[s]),
                                             bool(src.point_fx[s]), t0 * US, dt * US, T)[:, 0]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    assert np.isnan(want[:, 4]).sum() == 0 and np.isnan(want[:, 1]).any() and np.isnan(want[90:, 3]).all()
    # interpolation from them = interpolation from the projected series
    m2 = sb.PTGSKModel(geo)
    env_pre = sb.RegionEnvironment(temperature=(xyz, want), precipitation=env0.precipitation, radiation=env0.radiation,
                                   wind_speed=env0.wind_speed, rel_hum=env0.rel_hum)
    m2.run_interpolation(ip, ta, env_pre, best_effort=True)
    a, b = m.cell_forcing("temperature"), m2.cell_forcing("temperature")
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=1e-12, atol=0)
    with pytest.raises(RuntimeError, match="strictly increasing"):
        m._set_sources(sb.RegionEnvironment(temperature=sb.GeoPointSourceVector([(xyz[0], [t0, t0], [1.0, 2.0], t0 + 10, "average")])))


def test_a_new_time_axis_drops_everything_laid_out_on_the_old_one():
    """initialize_cell_environment (core/region_model.h:359-364) with another axis: the station series, interpolation plans and calibration
    targets of the previous axis are gone -- a windowed run without new sources sees unset sources (NaN forcing), never the old buffers
    read past their end."""
    import shyft_b200 as sb
    from shyft_b200 import synthetic
    geo, ta, env = synthetic.make_region(200, 96, 9, config_index=5)
    m = sb.PTGSKOptModel(geo)
    ip = sb.InterpolationParameter()
    assert m.run_interpolation(ip, ta, env)
    m.set_states(synthetic.default_state(0, 200))
    m.run_cells()
    assert np.all(np.isfinite(m.catchment_discharges()))
    longer = sb.TimeAxis(ta.start + 7 * 3600, 3600, 4 * 96)          # same dt, other start, four times the length
    m.initialize_cell_environment(longer)
    with pytest.raises(RuntimeError):
        m.sources_on_model_axis("temperature")                       # no sources on this axis
    try:
        m.run_windowed(ip, window_steps=128)
    except RuntimeError:
        pass                                                         # NaN forcing may be reported by the Kirchner stepper
    assert np.all(np.isnan(m.cell_forcing("temperature", 3 * 96, 96)))
    # with sources of the new axis everything works again
    geo2, ta2, env2 = synthetic.make_region(200, 4 * 96, 9, config_index=5, start=longer.start)
    m.revert_to_initial_state()
    m.run_windowed(ip, env=env2, window_steps=128)
    assert np.all(np.isfinite(m.catchment_discharges()))
