"""GPU parity of the calibration goal-function entry (model_calibration.h:691-699, 830-899) and of its batched form
(BASELINE config 5 shape: parameter-set ensemble x catchment x hourly steps, NSE goal)."""
import numpy as np
import pytest

from fixtures import FORCING, PTGSK_DEFAULT, geo_matrix
from parity import assert_parity

pytestmark = pytest.mark.gpu
DAY = 86400


@pytest.fixture(scope="module")
def sb():
    import shyft_b200
    return shyft_b200


@pytest.fixture(scope="module")
def setup(sb, oracle):
    from shyft_b200 import synthetic
    n, T, S = 240, 24 * 60, 9
    geo, ta, env = synthetic.make_region(n, T, S, config_index=4, cells_per_catchment=80, start=1425168000)  # 2015-03-01: melt season
    st0 = synthetic.default_state(0, n)
    m = sb.PTGSKOptModel(geo, PTGSK_DEFAULT)
    m.run_interpolation(sb.InterpolationParameter(), ta, env)
    m.set_states(st0)
    f = {k: m.cell_forcing(k) for k in FORCING}
    gm = geo_matrix(geo)
    # the "observed" series: the default-parameter run of catchments 1+2, daily means (twin experiment, test/calibration_test.cpp:337-559)
    truth = oracle.ptgsk_run_cells(gm, PTGSK_DEFAULT, f, st0, ta.start * 10**6, 3600 * 10**6, ncore=8)
    sel = np.isin(gm[:, 4], [1, 2])
    obs_daily = oracle.average_to_axis(truth["avg_discharge"][:, sel].sum(axis=1), 3600 * 10**6, 0, 24, 60)
    return m, geo, gm, ta, st0, f, obs_daily, sel


def _oracle_goal(oracle, gm, ta, st0, f, p, sel, obs, mode):
    run = oracle.ptgsk_run_cells(gm, p, f, st0, ta.start * 10**6, 3600 * 10**6, ncore=8)
    sim = oracle.average_to_axis(run["avg_discharge"][:, sel].sum(axis=1), 3600 * 10**6, 0, 24, obs.size)
    return {0: oracle.nash_sutcliffe, 1: oracle.kling_gupta, 2: oracle.abs_diff_sum, 3: oracle.rmse}[mode](obs, sim)


def _perturbed(rng, k):
    P = np.tile(PTGSK_DEFAULT, (k, 1))
    P[:, 0] = rng.uniform(-3.0, -1.9, k)      # kirchner.c1
    P[:, 1] = rng.uniform(0.8, 0.99, k)       # c2
    P[:, 2] = rng.uniform(-0.15, -0.05, k)    # c3
    P[:, 3] = rng.uniform(0.5, 2.5, k)        # ae scale
    P[:, 4] = rng.uniform(-2.0, 2.0, k)       # tx
    P[:, 5] = rng.uniform(1.0, 4.0, k)        # wind scale
    P[:, 14] = rng.uniform(0.2, 0.8, k)       # snow cv
    P[:, 16] = rng.uniform(0.8, 1.4, k)       # p_corr
    return P


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_goal_function_matches_the_reference_formulas(sb, oracle, setup, mode):
    m, geo, gm, ta, st0, f, obs, sel = setup
    opt = sb.Optimizer(m, [sb.TargetSpecification(obs, ta.start, DAY, [1, 2], 1.0, mode)])
    assert opt.calculate_goal_function(PTGSK_DEFAULT) == pytest.approx(0.0, abs=1e-9)  # the twin reproduces itself
    p = _perturbed(np.random.default_rng(mode), 1)[0]
    got = opt.calculate_goal_function(p)
    want = _oracle_goal(oracle, gm, ta, st0, f, p, sel, obs, mode)
    assert got == pytest.approx(want, rel=1e-9)
    # the calculation filter is the union of the target catchments (prepare_optimize, :517-552): catchment 3 is not stepped
    assert np.all(m.catchment_discharges()[:, 2] == 0.0)


def test_weighted_multi_target_goal_with_missing_observations(sb, oracle, setup):
    m, geo, gm, ta, st0, f, obs, sel = setup
    obs2 = obs.copy()
    obs2[7:11] = np.nan
    sel3 = gm[:, 4] == 3
    truth3 = oracle.ptgsk_run_cells(gm, PTGSK_DEFAULT, f, st0, ta.start * 10**6, 3600 * 10**6, ncore=8)["avg_discharge"][:, sel3].sum(axis=1)
    obs3 = oracle.average_to_axis(truth3, 3600 * 10**6, 24 * 10, 6, 100)  # 6-hourly target starting on day 10
    targets = [sb.TargetSpecification(obs2, ta.start, DAY, [1, 2], 2.0, 0),
               sb.TargetSpecification(obs3, ta.start + 10 * DAY, 6 * 3600, [3], 0.5, 1, s_r=1.0, s_a=0.5, s_b=2.0)]
    opt = sb.Optimizer(m, targets)
    p = _perturbed(np.random.default_rng(9), 1)[0]
    got = opt.calculate_goal_function(p)
    run = oracle.ptgsk_run_cells(gm, p, f, st0, ta.start * 10**6, 3600 * 10**6, ncore=8)["avg_discharge"]
    g1 = oracle.nash_sutcliffe(obs2, oracle.average_to_axis(run[:, sel].sum(axis=1), 3600 * 10**6, 0, 24, 60))
    g2 = oracle.kling_gupta(obs3, oracle.average_to_axis(run[:, sel3].sum(axis=1), 3600 * 10**6, 240, 6, 100), 1.0, 0.5, 2.0)
    assert got == pytest.approx((2.0 * g1 + 0.5 * g2) / 2.5, rel=1e-9)
    with pytest.raises(RuntimeError, match="positive delta_t"):
        sb.Optimizer(m, [sb.TargetSpecification(obs, ta.start, 0, [1])])
    with pytest.raises(RuntimeError, match="not found"):
        sb.Optimizer(m, [sb.TargetSpecification(obs, ta.start, DAY, [77])])


def test_batched_ensemble_equals_one_at_a_time(sb, oracle, setup):
    m, geo, gm, ta, st0, f, obs, sel = setup
    opt = sb.Optimizer(m, [sb.TargetSpecification(obs, ta.start, DAY, [1, 2], 1.0, 0)])
    P = _perturbed(np.random.default_rng(21), 37)
    batch = opt.calculate_goal_function_batch(P)
    single = np.array([opt.calculate_goal_function(p) for p in P])
    assert np.array_equal(batch, single)  # same kernels, same order of operations per member
    for i in (0, 17, 36):
        assert batch[i] == pytest.approx(_oracle_goal(oracle, gm, ta, st0, f, P[i], sel, obs, 0), rel=1e-9)
    assert np.argmin(batch) == np.argmin(single)


def test_snow_targets_use_area_weighted_catchment_means(sb, oracle, setup):
    m, geo, gm, ta, st0, f, obs, sel = setup
    truth = oracle.ptgsk_run_cells(gm, PTGSK_DEFAULT, f, st0, ta.start * 10**6, 3600 * 10**6, ncore=8)
    area = gm[:, 3]
    swe_obs = oracle.average_to_axis((truth["snow_swe"][:, sel] * area[sel]).sum(axis=1) / area[sel].sum(), 3600 * 10**6, 0, 24, 60)
    opt = sb.Optimizer(m, [sb.TargetSpecification(swe_obs, ta.start, DAY, [1, 2], 1.0, 3, catchment_property=2)])
    assert opt.calculate_goal_function(PTGSK_DEFAULT) == pytest.approx(0.0, abs=1e-9)
    p = _perturbed(np.random.default_rng(5), 1)[0]
    run = oracle.ptgsk_run_cells(gm, p, f, st0, ta.start * 10**6, 3600 * 10**6, ncore=8)
    sim = oracle.average_to_axis((run["snow_swe"][:, sel] * area[sel]).sum(axis=1) / area[sel].sum(), 3600 * 10**6, 0, 24, 60)
    assert opt.calculate_goal_function(p) == pytest.approx(oracle.rmse(swe_obs, sim), rel=1e-9)


def test_cell_charge_targets_and_the_shared_series_cache(sb, oracle, setup):
    """CELL_CHARGE targets (model_calibration.h:854-856): ABS_DIFF on them is the scaled form over max_abs_average_accessor
    (:870-873; core/time_series.h:2198-2267, 2435-2448).  And the reference's one cache vector for discharge AND charge sums: the first
    DISCHARGE / CELL_CHARGE target of the list decides which series every such target sees."""
    m, geo, gm, ta, st0, f, obs, sel = setup
    H = 3600 * 10**6
    truth = oracle.ptgsk_run_cells(gm, PTGSK_DEFAULT, f, st0, ta.start * 10**6, H, ncore=8)
    charge_obs = oracle.average_to_axis(truth["charge_m3s"][:, sel].sum(axis=1), H, 0, 24, 60) + 0.25
    p = _perturbed(np.random.default_rng(11), 1)[0]
    run = oracle.ptgsk_run_cells(gm, p, f, st0, ta.start * 10**6, H, ncore=8)
    ch = run["charge_m3s"][:, sel].sum(axis=1)
    q = run["avg_discharge"][:, sel].sum(axis=1)
    sim_c, scale_c = oracle.average_to_axis(ch, H, 0, 24, 60), oracle.max_abs_average_to_axis(ch, H, 0, 24, 60)
    assert (ch < 0).any() and (ch > 0).any() and np.all(scale_c >= np.abs(sim_c) - 1e-12)
    # RMSE and (scaled) ABS_DIFF of the charge
    opt = sb.Optimizer(m, [sb.TargetSpecification(charge_obs, ta.start, DAY, [1, 2], 1.0, 3, catchment_property=4)])
    assert opt.calculate_goal_function(p) == pytest.approx(oracle.rmse(charge_obs, sim_c), rel=1e-9)
    opt = sb.Optimizer(m, [sb.TargetSpecification(charge_obs, ta.start, DAY, [1, 2], 1.0, 2, catchment_property=4)])
    assert opt.calculate_goal_function(p) == pytest.approx(oracle.abs_diff_sum_scaled(charge_obs, sim_c, scale_c), rel=1e-9)
    # a DISCHARGE target first: the CELL_CHARGE target after it is evaluated on the cached DISCHARGE sums (and scaled by them)
    sim_q, scale_q = oracle.average_to_axis(q, H, 0, 24, 60), oracle.max_abs_average_to_axis(q, H, 0, 24, 60)
    opt = sb.Optimizer(m, [sb.TargetSpecification(obs, ta.start, DAY, [1, 2], 1.0, 0),
                           sb.TargetSpecification(charge_obs, ta.start, DAY, [1, 2], 3.0, 2, catchment_property=4)])
    want = (1.0 * oracle.nash_sutcliffe(obs, sim_q) + 3.0 * oracle.abs_diff_sum_scaled(charge_obs, sim_q, scale_q)) / 4.0
    assert opt.calculate_goal_function(p) == pytest.approx(want, rel=1e-9)
    # the other way round: the DISCHARGE target sees the cached CHARGE sums
    opt = sb.Optimizer(m, [sb.TargetSpecification(charge_obs, ta.start, DAY, [1, 2], 3.0, 3, catchment_property=4),
                           sb.TargetSpecification(obs, ta.start, DAY, [1, 2], 1.0, 0)])
    want = (3.0 * oracle.rmse(charge_obs, sim_c) + 1.0 * oracle.nash_sutcliffe(obs, sim_c)) / 4.0
    assert opt.calculate_goal_function(p) == pytest.approx(want, rel=1e-9)
    # batched evaluation of charge targets equals one at a time
    P = _perturbed(np.random.default_rng(12), 9)
    assert np.array_equal(opt.calculate_goal_function_batch(P), np.array([opt.calculate_goal_function(x) for x in P]))


def test_target_axes_not_aligned_with_the_model_axis(sb, oracle, setup):
    """average_accessor<pts_t, ta_t>(property_sum, t.ts.time_axis()) (model_calibration.h:859) for ANY fixed_dt target axis: shifted by
    half a step, 90-minute periods, starting before and ending after the model axis -- against the oracle's line-by-line accumulate_value"""
    m, geo, gm, ta, st0, f, obs, sel = setup
    H = 3600 * 10**6
    T = ta.n
    p = _perturbed(np.random.default_rng(31), 1)[0]
    run = oracle.ptgsk_run_cells(gm, p, f, st0, ta.start * 10**6, H, ncore=8)
    q = run["avg_discharge"][:, sel].sum(axis=1)
    ch = run["charge_m3s"][:, sel].sum(axis=1)
    t_us = ta.start * 10**6 + H * np.arange(T, dtype=np.int64)
    t_end = ta.start * 10**6 + H * T
    rng = np.random.default_rng(32)

    def project(series, t0_s, dt_s, n):
        return oracle.average_accessor(t_us, series[:, None], t_end, False, t0_s * 10**6, dt_s * 10**6, n)[:, 0]

    cases = [(ta.start + 1800, DAY, 59),            # shifted by half a model step
             (ta.start + 2 * 3600, 5400, 700),      # 90-minute periods
             (ta.start - 3 * DAY - 900, DAY, 66)]   # starts three days before the model axis, runs past its end
    for t0_s, dt_s, n in cases:
        sim = project(q, t0_s, dt_s, n)
        assert np.isfinite(sim).sum() >= n - 8 and (np.isnan(sim).any() or t0_s > ta.start)
        target = np.where(np.isfinite(sim), sim, 1.0) * rng.uniform(0.8, 1.2, n)
        for mode, fn in ((0, oracle.nash_sutcliffe), (1, oracle.kling_gupta), (2, oracle.abs_diff_sum), (3, oracle.rmse)):
            opt = sb.Optimizer(m, [sb.TargetSpecification(target, t0_s, dt_s, [1, 2], 1.0, mode)])
            assert opt.calculate_goal_function(p) == pytest.approx(fn(target, sim), rel=1e-9), (t0_s, dt_s, mode)
    # scaled ABS_DIFF of the charge on a shifted axis
    t0_s, dt_s, n = ta.start + 1800, 6 * 3600, 200
    sim = project(ch, t0_s, dt_s, n)
    pos = project(np.maximum(0.0, ch), t0_s, dt_s, n)
    neg = project(np.maximum(0.0, -ch), t0_s, dt_s, n)
    scale = np.where(pos < neg, neg, pos)
    target = sim + rng.normal(0.0, 0.3, n)
    opt = sb.Optimizer(m, [sb.TargetSpecification(target, t0_s, dt_s, [1, 2], 1.0, 2, catchment_property=4)])
    assert opt.calculate_goal_function(p) == pytest.approx(oracle.abs_diff_sum_scaled(target, sim, scale), rel=1e-9)
    # the batch entry evaluates such targets one set at a time
    P = _perturbed(rng, 3)
    assert np.array_equal(opt.calculate_goal_function_batch(P), np.array([opt.calculate_goal_function(x) for x in P]))


def test_target_on_a_point_axis(sb, oracle, setup):
    """a target whose periods are given by explicit points (time_axis::point_dt under apoint_ts, api/boostpython/api_target_specification.cpp:48):
    irregular periods -- hours, days, a week, with a gap-free but uneven cut -- against the oracle's accumulate_value period by period"""
    m, geo, gm, ta, st0, f, obs, sel = setup
    H = 3600 * 10**6
    T = ta.n
    p = _perturbed(np.random.default_rng(41), 1)[0]
    q = oracle.ptgsk_run_cells(gm, p, f, st0, ta.start * 10**6, H, ncore=8)["avg_discharge"][:, sel].sum(axis=1)
    t_us = ta.start * 10**6 + H * np.arange(T, dtype=np.int64)
    t_end = ta.start * 10**6 + H * T
    rng = np.random.default_rng(42)
    cuts = np.concatenate([[0], np.cumsum(rng.choice([1800, 3600, 5400, 6 * 3600, 86400, 7 * 86400], 60))])
    cuts = cuts[cuts < (T + 30) * 3600]
    pts = ta.start + 900 + cuts                                  # seconds; the last periods reach past the model axis
    n = pts.size - 1
    sim = np.array([oracle.average_accessor(t_us, q[:, None], t_end, False, int(pts[i]) * 10**6, int(pts[i + 1] - pts[i]) * 10**6, 1)[0, 0]
                    for i in range(n)])
    assert np.isfinite(sim).sum() >= n - 3
    target = np.where(np.isfinite(sim), sim, 1.0) * rng.uniform(0.9, 1.1, n)
    for mode, fn in ((0, oracle.nash_sutcliffe), (2, oracle.abs_diff_sum), (3, oracle.rmse)):
        opt = sb.Optimizer(m, [sb.TargetSpecification(target, 0, 0, [1, 2], 1.0, mode, time_points=pts)])
        assert opt.calculate_goal_function(p) == pytest.approx(fn(target, sim), rel=1e-9), mode
    with pytest.raises(RuntimeError, match="strictly increasing"):
        sb.Optimizer(m, [sb.TargetSpecification([1.0, 2.0], 0, 0, [1], time_points=[ta.start, ta.start + 10, ta.start + 10])])
