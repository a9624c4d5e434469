"""State tuning: the oracle's restatement of adjust_state_model::tune_flow (core/model_state_tuning.h:38-118)
and of dlib's find_min_single_variable, checked with the asserts the reference's own tests make
(shyft/tests/api/test_region_model_stacks.py:311-333; test/cell_builder_test.cpp:281-296).  The dlib boundary itself is
parity-unpinned (library absent, the reference pins the reached flow to 2 decimals only)."""
import numpy as np
import pytest

from fixtures import oracle_interpolate_py_fixture, py_region_fixture


def test_minimiser_on_known_functions(oracle):
    x, fx = oracle.find_min_single_variable(lambda x: (x - 2.5) ** 2 + 1.0, 1.0, -10.0, 10.0, 1e-6, 100)
    assert x == pytest.approx(2.5, abs=1e-6) and fx == pytest.approx(1.0, abs=1e-12)
    # minimum outside the bounds: the bound itself is returned once the bracket has shrunk to eps
    x, _ = oracle.find_min_single_variable(lambda x: (x - 20.0) ** 2, 1.0, 0.0, 3.0, 1e-4, 200)
    assert x == pytest.approx(3.0, abs=1e-4)
    x, _ = oracle.find_min_single_variable(lambda x: np.cosh(x - 0.3), 4.0, 0.0, 5.0, 1e-5, 100)
    assert x == pytest.approx(0.3, abs=1e-4)
    with pytest.raises(oracle.MinimiserFailure):   # argument check: start outside [begin, end]; also what a NaN start runs into
        oracle.find_min_single_variable(lambda x: x * x, float("nan"), 0.0, 1.0, 1e-3, 100)
    with pytest.raises(oracle.MinimiserFailure, match="max number of iterations"):
        oracle.find_min_single_variable(lambda x: (x - 2.5) ** 2, 1.0, -10.0, 10.0, 1e-12, 5)
    calls = []
    oracle.find_min_single_variable(lambda x: calls.append(x) or (x - 1.0) ** 2, 0.7, 0.7 / 3, 0.7 * 3, 0.7e-3, 300)
    assert calls[:3] == [0.7 / 3, 1.7, 0.7]   # start - 1 clipped to the lower bound, start + 1, then the start itself


def _py_fixture(oracle):
    fx = py_region_fixture()
    f = oracle_interpolate_py_fixture(oracle, fx)

    def run_cells(state, start_step, n_steps, mask):
        return oracle.ptgsk_run_cells(fx["geo"], fx["par"], f, state, fx["t0"] * 10**6, fx["dt"] * 10**6, start_step=start_step, n_steps=n_steps,
                                      cell_mask=mask)
    return fx, f, run_cells


def test_reference_python_test_asserts(oracle):
    fx, f, run_cells = _py_fixture(oracle)
    cat = fx["geo"][:, 4].astype(np.int64)
    out = run_cells(fx["state"], 10, 2, None)
    q_avg = (out["avg_discharge"][10].sum() + out["avg_discharge"][11].sum()) / 2.0
    x = 0.7
    r = oracle.adjust_state_to_target_flow(run_cells, fx["state"], [8], cat, x * q_avg, cids=[], start_step=10, scale_range=3.0, scale_eps=1e-3,
                                           max_iter=350, n_steps=2)
    assert r["diagnostics"] == ""
    assert r["q_r"] == pytest.approx(q_avg * x, abs=0.005)    # assertAlmostEqual(..., 2)
    assert r["q_0"] == pytest.approx(q_avg, abs=0.005)
    assert np.array_equal(r["state"][:, :8], fx["state"][:, :8])
    assert np.allclose(r["state"][:, 8], fx["state"][:, 8] * r["scale"], rtol=0, atol=0)
    assert 3 < len(r["evaluations"]) < 60
    # bad observed value, then a bad simulated value (NaN temperature at step 10)
    r = oracle.adjust_state_to_target_flow(run_cells, fx["state"], [8], cat, float("nan"), start_step=10, n_steps=2)
    assert len(r["diagnostics"]) > 0
    f["temperature"][10, 0] = float("nan")
    r = oracle.adjust_state_to_target_flow(run_cells, fx["state"], [8], cat, 30.0, start_step=10, n_steps=2)
    assert len(r["diagnostics"]) > 0


def test_subset_of_catchments_and_unknown_id(oracle):
    fx, f, run_cells = _py_fixture(oracle)
    cat = np.where(np.arange(20) < 8, 1, 2).astype(np.int64)
    out = run_cells(fx["state"], 0, 2, None)
    q2 = out["avg_discharge"][:2, cat == 2].sum() / 2.0
    r = oracle.adjust_state_to_target_flow(run_cells, fx["state"], [8], cat, 1.6 * q2, cids=[2], start_step=0, scale_range=10.0, n_steps=2)
    assert r["diagnostics"] == "" and r["q_r"] == pytest.approx(1.6 * q2, abs=0.1)   # TS_ASSERT_DELTA(q_adjusted, q_wanted, 0.1)
    assert np.array_equal(r["state"][cat == 1], fx["state"][cat == 1])            # cells outside cids keep their state
    with pytest.raises(RuntimeError, match="no cells have supplied cid"):
        oracle.adjust_state_to_target_flow(run_cells, fx["state"], [8], cat, 10.0, cids=[3])
