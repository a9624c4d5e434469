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
