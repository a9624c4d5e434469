"""The pybind11 module over the C++ shim: the reference's Python class names
(api/boostpython/expose.h:143-430, shyft/api/<stack>/__init__.py) bound to shyft_b200's region_model<STACK>."""
import numpy as np
import pytest

from fixtures import PTGSK_DEFAULT, py_region_fixture


@pytest.fixture(scope="module")
def cpp():
    from shyft_b200 import _build
    _build.build_pybind_module()
    from shyft_b200 import _shyft_b200_cpp
    return _shyft_b200_cpp


def test_module_exposes_the_reference_names(cpp):
    for name in ("PTGSKModel", "PTGSKOptModel", "PTHSKModel", "PTHSKOptModel", "HbvModel", "HbvOptModel", "FlowAdjustResult"):
        assert hasattr(cpp, name), name
    for method in ("run_interpolation", "interpolate", "initialize_cell_environment", "run_cells", "set_states", "get_states",
                   "revert_to_initial_state", "adjust_q", "adjust_state_to_target_flow", "set_region_parameter", "get_region_parameter",
                   "set_catchment_parameter", "remove_catchment_parameter", "has_catchment_parameter", "get_catchment_parameter",
                   "set_catchment_calculation_filter", "catchment_discharges", "size", "number_of_catchments", "is_cell_env_ts_ok"):
        assert hasattr(cpp.PTGSKModel, method), method   # region_model members bound by expose.h:251-400
    assert "use_ncore" in cpp.PTGSKModel.run_cells.__doc__ and "start_step" in cpp.PTGSKModel.run_cells.__doc__


def test_no_cpu_fallback_through_the_module(cpp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import shyft_b200 as sb
    geo = sb.geo_cell_data_vector(np.arange(4.0), np.zeros(4), np.zeros(4))
    with pytest.raises(RuntimeError, match="no CUDA device available"):
        cpp.PTGSKModel(geo, list(PTGSK_DEFAULT))
    with pytest.raises(RuntimeError, match="96-byte geo_cell_data records"):
        cpp.PTGSKModel(np.zeros((4, 3)), list(PTGSK_DEFAULT))


@pytest.mark.gpu
def test_reference_fixture_through_the_module_equals_the_ctypes_mirror(cpp):
    """shyft/tests/api/test_region_model_stacks.py:145-218 driven through the pybind module; same library underneath, so bit-equal"""
    import shyft_b200 as sb
    fx = py_region_fixture()
    g = fx["geo"]
    geo = sb.geo_cell_data_vector(g[:, 0], g[:, 1], g[:, 2], area=g[:, 3], catchment_id=g[:, 4].astype(np.int64), radiation_slope_factor=g[:, 5],
                                  glacier=g[:, 6], lake=g[:, 7], reservoir=g[:, 8], forest=g[:, 9])
    env = {k: (fx["station"][None, :], np.full((fx["T"], 1), v)) for k, v in fx["consts"].items()}
    a = cpp.PTGSKModel(geo, list(fx["par"]))
    assert a.size() == 20 and a.number_of_catchments() == 1 and a.catchment_ids == [1]
    assert a.run_interpolation(fx["t0"] * 10**6, fx["dt"] * 10**6, fx["T"], env)
    assert a.is_cell_env_ts_ok()
    a.set_states(fx["state"])
    a.run_cells()
    b = sb.PTGSKModel(geo, fx["par"])
    assert b.run_interpolation(sb.InterpolationParameter(), sb.TimeAxis(fx["t0"], fx["dt"], fx["T"]), sb.RegionEnvironment(**env))
    b.set_states(fx["state"])
    b.run_cells()
    assert np.array_equal(a.catchment_discharges()[0], b.catchment_discharges()[:, 0])
    assert np.array_equal(a.response(cpp.R_AVG_DISCHARGE), b.response("avg_discharge"))
    assert np.array_equal(a.get_states(), b.get_states())
    assert a.statistics(cpp.STAT_RESPONSE, cpp.R_CHARGE_M3S, [], cpp.STAT_SUM)[0] == pytest.approx(-110.6998, abs=1e-4)   # the reference's literal
    # state tuning and errors through the module
    a.revert_to_initial_state()
    a.run_cells(0, 10, 2)
    q_avg = sum(a.statistics_value(cpp.STAT_RESPONSE, cpp.R_AVG_DISCHARGE, [], i, cpp.STAT_SUM) for i in (10, 11)) / 2.0
    a.revert_to_initial_state()
    r = a.adjust_state_to_target_flow(0.7 * q_avg, [], start_step=10, scale_range=3.0, scale_eps=1e-3, max_iter=350, n_steps=2)
    assert r.diagnostics == "" and r.q_r == pytest.approx(0.7 * q_avg, abs=0.005) and r.q_0 == pytest.approx(q_avg, abs=0.005)
    with pytest.raises(RuntimeError, match="start_step"):
        a.run_cells(0, 10**6, 1)
    o = cpp.PTGSKOptModel(geo, list(fx["par"]))
    assert o.size() == 20


def test_calendar_axes_expand_to_period_points(cpp):
    """calendar_dt target axes (month / quarter / year steps, core/utctime_utilities.cpp:151-228) spelled as point axes: the C++ shim's
    calendar arithmetic against numpy's datetime64 months, and against the Python mirror's helper"""
    import calendar as pycal

    import shyft_b200 as sb
    t0 = pycal.timegm((2015, 1, 31, 6, 0, 0))            # the 31st: clipped to the month's length on the way
    pts = cpp.calendar_period_points(t0 * 10**6, "month", 14)
    want = [pycal.timegm((2015 + (m // 12), m % 12 + 1, min(31, pycal.monthrange(2015 + m // 12, m % 12 + 1)[1]), 6, 0, 0)) for m in range(15)]
    assert pts == [w * 10**6 for w in want]
    assert [p // 10**6 for p in pts] == sb.calendar_period_points(t0, "month", 14)
    y = cpp.calendar_period_points(pycal.timegm((2012, 2, 29, 0, 0, 0)) * 10**6, "year", 4)
    assert [p // 10**6 for p in y] == [pycal.timegm((2012 + k, 2, 29 if (2012 + k) % 4 == 0 else 28, 0, 0, 0)) for k in range(5)]
    assert [p // 10**6 for p in y] == sb.calendar_period_points(pycal.timegm((2012, 2, 29, 0, 0, 0)), "year", 4)
    q = cpp.calendar_period_points(pycal.timegm((1969, 11, 15, 0, 0, 0)) * 10**6, "quarter", 2)   # across the epoch
    assert [p // 10**6 for p in q] == [pycal.timegm((1969, 11, 15, 0, 0, 0)), pycal.timegm((1970, 2, 15, 0, 0, 0)), pycal.timegm((1970, 5, 15, 0, 0, 0))]
    assert cpp.calendar_period_points(0, "day", 3) == [0, 86400 * 10**6, 2 * 86400 * 10**6, 3 * 86400 * 10**6]
    with pytest.raises(RuntimeError, match="unit"):
        cpp.calendar_period_points(0, "fortnight", 1)
    for name in ("TargetSpecificationPts", "PTGSKOptimizer", "PTHSKOptimizer", "HbvOptimizer"):
        assert hasattr(cpp, name), name
    for method in ("set_target_specification", "set_parameter_ranges", "set_verbose_level", "reset_states", "calculate_goal_function", "optimize",
                   "optimize_dream", "optimize_sceua", "parameter_active", "trace_goal_function_value", "trace_parameter", "calculate_goal_function_batch"):
        assert hasattr(cpp.PTGSKOptimizer, method), method   # model_calibrator<RegionModel>, api/boostpython/expose.h:472-731


@pytest.mark.gpu
def test_optimizer_through_the_module(cpp, oracle):
    """PTGSKOptimizer (api/boostpython/expose.h:472-731): goal function, traces, parameter ranges; a population driver hands every generation
    to the device in one batch, and the batch equals one-at-a-time evaluation; a monthly calendar axis as target axis"""
    import shyft_b200 as sb
    from shyft_b200 import synthetic
    n, T = 240, 24 * 90
    geo, ta, env = synthetic.make_region(n, T, 9, config_index=4, cells_per_catchment=80, start=1425168000)   # 2015-03-01
    envd = {k: getattr(env, k) for k in sb.capi.FORCING_NAMES}
    m = cpp.PTGSKOptModel(geo, list(PTGSK_DEFAULT))
    assert m.run_interpolation(ta.start * 10**6, 3600 * 10**6, T, envd)
    m.set_states(synthetic.default_state(0, n))
    m.run_cells()
    q = m.catchment_discharges()                      # [catchment][step]
    obs_daily = (q[0] + q[1]).reshape(-1, 24).mean(axis=1)
    t_daily = cpp.TargetSpecificationPts(list(obs_daily), ta.start * 10**6, 86400 * 10**6, [1, 2], 1.0, cpp.NASH_SUTCLIFFE)
    lo, hi = list(PTGSK_DEFAULT), list(PTGSK_DEFAULT)
    lo[0], hi[0] = -3.0, -2.0       # kirchner.c1
    lo[4], hi[4] = -1.5, 1.0        # gs.tx
    opt = cpp.PTGSKOptimizer(m, [t_daily], lo, hi)
    assert [i for i in range(31) if opt.parameter_active(i)] == [0, 4]
    assert opt.calculate_goal_function(list(PTGSK_DEFAULT)) == pytest.approx(0.0, abs=1e-9)      # the twin reproduces itself
    rng = np.random.default_rng(2)
    P = np.tile(PTGSK_DEFAULT, (6, 1))
    P[:, 0] = rng.uniform(-3.0, -2.0, 6)
    P[:, 4] = rng.uniform(-1.5, 1.0, 6)
    g_batch = opt.calculate_goal_function_batch(P)
    g_single = [opt.calculate_goal_function(list(p)) for p in P]
    assert g_batch == g_single                         # bit-identical: members are grid layers of the same kernels
    assert opt.trace_size == 1 + 6 + 6 and opt.trace_goal_function_value(1) == g_batch[0] and opt.trace_parameter(1) == list(P[0])
    # a differential-evolution generation = ONE device batch (8 chains: the initial population + 3 generations = 4 batches, 32 evaluations)
    start = list(PTGSK_DEFAULT)
    start[0], start[4] = -2.9, 0.8
    b0, s0, n0 = opt.n_batch_calls, opt.n_single_calls, opt.trace_size
    best = opt.optimize_dream(start, 32)
    assert opt.n_batch_calls - b0 == 4 and opt.n_single_calls == s0 and opt.trace_size - n0 == 32
    assert opt.calculate_goal_function(best) <= opt.calculate_goal_function(start)
    best2 = opt.optimize_sceua(start, 60, 1e-4, 1e-7)
    assert opt.calculate_goal_function(best2) <= opt.calculate_goal_function(start)
    best3 = opt.optimize(start, 40, 0.1, 1e-4)
    assert opt.calculate_goal_function(best3) <= opt.calculate_goal_function(start)
    assert all(lo[i] - 1e-12 <= best3[i] <= hi[i] + 1e-12 for i in range(31))
    # monthly target periods (a calendar_dt axis) as a point axis: March, April, May 2015 (until the model axis ends)
    pts = cpp.calendar_period_points(ta.start * 10**6, "month", 3)
    end_us = (ta.start + T * 3600) * 10**6
    pts[-1] = min(pts[-1], end_us)
    hours = [(pts[i] // 10**6 - ta.start) // 3600 for i in range(4)]
    obs_monthly = [float((q[0] + q[1])[hours[i]:hours[i + 1]].mean()) for i in range(3)]
    t_monthly = cpp.TargetSpecificationPts(obs_monthly, 0, 0, [1, 2], 1.0, cpp.RMSE, period_points_us=pts)
    opt.set_target_specification([t_monthly], lo, hi)
    assert opt.calculate_goal_function(list(PTGSK_DEFAULT)) == pytest.approx(0.0, abs=1e-9)
    assert opt.calculate_goal_function(start) > 1e-6
