"""The oracle against the reference's own region-level outputs (shyft/tests/api/test_region_model_stacks.py:145-304).

These literals are what real Shyft printed for this fixture; reproducing them pins the Kirchner/odeint restatement, the
stack wiring, the collectors' unit conversions, IDW and the routing convolution.  The fixture is
snow-free (10 degC), so the incomplete-gamma branch of gamma_snow stays "parity unpinned" (oracle/sho_core.hpp header).
"""
import json
import os

import numpy as np
import pytest

from fixtures import oracle_interpolate_py_fixture, py_region_fixture

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))["region_pt_gs_k_20x240"]


@pytest.fixture(scope="module")
def run(oracle):
    fx = py_region_fixture()
    f = oracle_interpolate_py_fixture(oracle, fx)
    out = oracle.ptgsk_run_cells(fx["geo"], fx["par"], f, fx["state"], fx["t0"] * 10**6, fx["dt"] * 10**6, collect_response=True,
                                 collect_state=True, collect_substeps=True, ncore=2)
    return fx, f, out


def test_charge_sums(run):
    fx, f, out = run
    ch = out["charge_m3s"]
    assert ch[0].sum() == pytest.approx(GOLD["charge_sum_step0"]["value"], abs=1.0e-4)          # literal written with 4 decimals (truncated), asserted to 2 in the reference
    assert ch[0, [0, 1, 3]].sum() == pytest.approx(GOLD["charge_cells_0_1_3_step0"]["value"], abs=1.0e-4)
    assert ch[:, [1, 2, 6]].sum() == pytest.approx(GOLD["charge_sum_cells_1_2_6_all_steps"]["value"], abs=2.0e-4)


def test_ae_output_and_pot_ratio_full_precision(run):
    fx, f, out = run
    ae_avg = out["ae_output"].mean(axis=1)  # equal areas: area-weighted mean = mean
    assert ae_avg.max() == pytest.approx(GOLD["ae_output_max"]["value"], abs=5e-15)
    q_mmh = out["kirchner_discharge"] / (fx["geo"][:, 3] / 3.6e6)
    pot_ratio = (1.0 - np.exp(-q_mmh * 3.0 / fx["par"][3])).mean(axis=1)
    assert pot_ratio.min() == pytest.approx(GOLD["ae_pot_ratio_min"]["value"], abs=5e-15)
    assert pot_ratio.max() == pytest.approx(1.0, abs=1e-7)


def test_discharge_first_step(run):
    fx, f, out = run
    d0 = out["avg_discharge"][0].sum()
    assert d0 >= GOLD["discharge_step0_min"]
    assert d0 == pytest.approx(138.3778, abs=1e-3)  # value of the restatement


def test_chunked_equals_one_shot(oracle, run):
    fx, f, out = run
    st = fx["state"].copy()
    q = np.zeros_like(out["avg_discharge"])
    for k in range(10):
        o = oracle.ptgsk_run_cells(fx["geo"], fx["par"], f, st, fx["t0"] * 10**6, fx["dt"] * 10**6, start_step=24 * k, n_steps=24)
        st = o["state"]
        q[24 * k:24 * k + 24] = o["avg_discharge"][24 * k:24 * k + 24]
    assert np.array_equal(q, out["avg_discharge"])  # the reference asserts 1e-4; the oracle is exactly reproducible
    assert np.array_equal(st, out["state"])


def test_routed_river_flow_full_precision(oracle, run):
    fx, f, out = run
    n = fx["geo"].shape[0]
    rivers = [[1, 0, 3000.0, 1 / 3.60, 7.0, 0.0]]
    local, up, outflow = oracle.river_flows(rivers, 1, out["avg_discharge"], np.ones(n, dtype=np.int64), np.zeros(n),
                                            np.tile([fx["par"][25], fx["par"][26], fx["par"][27]], (n, 1)), fx["dt"] * 10**6)
    assert outflow[8] == pytest.approx(GOLD["river_out_value_8"]["value"], abs=2e-14)
    assert np.allclose(local, out["avg_discharge"].sum(axis=1), rtol=1e-14)
    assert np.all(up == 0.0)


def test_kirchner_substep_statistics(run):
    fx, f, out = run
    sub = out["kirchner_substeps"]
    assert sub.min() >= 1 and sub.max() <= 6
    assert 1.0 < sub.mean() < 1.5  # mean 1.13 try_steps per model step on this fixture
