"""The oracle's restatement of hint_based_search / accumulate_value / average_value / average_accessor (oracle/sho_ts.hpp,
core/time_series.h:144-310, 2033-2072) against the reference's own known answers: test/time_series_test.cpp:480-640."""
import calendar as pycal

import numpy as np
import pytest

US = 10**6
T0 = pycal.timegm((2000, 1, 1, 0, 0, 0)) * US
DT = 3600 * US
NAN = float("nan")


def _pts(*tv):
    t = np.array([p[0] for p in tv], dtype=np.int64)
    v = np.array([p[1] for p in tv], dtype=np.float64)
    return t, v


def test_average_value_staircase_known_answers(oracle):
    """test_average_value_staircase, time_series_test.cpp:480-571"""
    t, v = _pts((T0, 1.0), (T0 + DT // 2, 2.0), (T0 + 2 * DT, 3.0))
    r, ix = oracle.average_value(t, v, T0, T0 + 3 * DT, 0)
    assert r == pytest.approx((1 * 0.5 + 2 * 1.5 + 3 * 1.0) / 3.0, abs=1e-7)                       # case 1
    r, ix = oracle.average_value(t, v, T0 + 10 * 60 * US, T0 + 11 * 60 * US, -1)
    assert r == pytest.approx(1.0, abs=1e-7) and ix == 1                                            # case 2
    r, ix = oracle.average_value(t, v, T0 + 5 * DT, T0 + 60 * DT, 2)
    assert r == pytest.approx(3.0, abs=1e-7) and ix == 2                                            # case 3: flat after the last point
    r, ix = oracle.average_value(t, v, T0 - 5 * DT, T0 - 4 * DT, 2)
    assert not np.isfinite(r) and ix == 0                                                           # case 4: before the first point
    full = (T0, T0 + 3 * DT)
    t0n, v0n = _pts((T0, NAN), (T0 + DT // 2, 2.0), (T0 + DT, NAN), (T0 + 2 * DT, 3.0))            # case 5: NaN handling
    t1n, v1n = _pts((T0, 1.0), (T0 + DT // 2, 2.0), (T0 + DT, NAN), (T0 + 2 * DT, 3.0))
    t2n, v2n = _pts((T0, 1.0), (T0 + DT // 2, 2.0), (T0 + DT, NAN), (T0 + 2 * DT, NAN))
    assert oracle.average_value(t0n, v0n, *full, 0)[0] == pytest.approx((0.5 * 2 + 3.0) / 1.5, abs=1e-5)
    assert oracle.average_value(t1n, v1n, *full, 0)[0] == pytest.approx((1.0 * 0.5 + 0.5 * 2 + 3.0) / 2.0, abs=1e-5)
    assert oracle.average_value(t2n, v2n, *full, 0)[0] == pytest.approx((1.0 * 0.5 + 0.5 * 2 + 0.0) / 1.0, abs=1e-5)
    assert oracle.average_value(*_pts((T0, 1.0)), *full, 0)[0] == pytest.approx(1.0, abs=1e-5)
    assert not np.isfinite(oracle.average_value(np.zeros(0, dtype=np.int64), np.zeros(0), *full, -1)[0])
    assert not np.isfinite(oracle.average_value(*_pts((T0, 1.0)), T0 - 10 * DT, T0 - 9 * DT, -1)[0])
    t10 = T0 + DT * np.arange(10, dtype=np.int64)
    v10 = np.arange(10.0)
    assert oracle.average_value(t10, v10, *full, 7)[0] == pytest.approx(1.0, abs=1e-5)              # hint far to the right: search downwards
    r, ix = oracle.average_value(t10, v10, T0 + 3 * DT, T0 + 4 * DT, 2)
    assert r == pytest.approx(3.0, abs=1e-5) and ix == 4
    assert oracle.average_value(t10, v10, T0 + 7 * DT, T0 + 8 * DT, 0)[0] == pytest.approx(7.0, abs=1e-5)
    acc = oracle.average_accessor(t, v, T0 + 3 * DT, False, T0, DT, 3)[:, 0]                        # average_accessor over fixed_dt(t0, dt, 3)
    assert acc == pytest.approx([(1 * 0.5 + 2 * 0.5) / 1.0, 2.0, 3.0], abs=1e-6)


def test_average_value_linear_between_points_known_answers(oracle):
    """test_average_value_linear_between_points, time_series_test.cpp:573-640"""
    t, v = _pts((T0, 1.0), (T0 + DT // 2, 2.0), (T0 + 2 * DT, 3.0))
    assert oracle.average_value(t, v, T0, T0 + 3 * DT, 0, linear=True)[0] == pytest.approx(2.25)    # case 1: nothing after the last point
    r, ix = oracle.average_value(t, v, T0 + 10 * 60 * US, T0 + 11 * 60 * US, -1, linear=True)
    assert r == pytest.approx(1.0 + 2 * 10.5 / 60, abs=1e-7) and ix == 1                           # case 2
    r, ix = oracle.average_value(t, v, T0 + 5 * DT, T0 + 60 * DT, 2, linear=True)
    assert not np.isfinite(r) and ix == 2                                                           # case 3
    r, ix = oracle.average_value(t, v, T0 - 5 * DT, T0 - 4 * DT, 2, linear=True)
    assert not np.isfinite(r) and ix == 0                                                           # case 4
    assert not np.isfinite(oracle.average_value(np.zeros(0, dtype=np.int64), np.zeros(0), T0, T0 + 3 * DT, -1, linear=True)[0])


def test_average_accessor_projections(oracle):
    """3-hourly and 20-minute sources onto an hourly axis: means are preserved where the source covers the period, NaN outside"""
    rng = np.random.default_rng(5)
    n3 = 16
    t3 = T0 + 3 * DT * np.arange(n3, dtype=np.int64)
    v3 = rng.normal(size=(n3, 4))
    out = oracle.average_accessor(t3, v3, T0 + 3 * DT * n3, False, T0 - 2 * DT, DT, 3 * n3 + 6)
    assert np.all(np.isnan(out[:2]))                                  # before the first point
    assert np.allclose(out[2:2 + 3 * n3], np.repeat(v3, 3, axis=0), rtol=4e-16, atol=0)  # a stair-case repeats its value: (3600 v) / 3600
    assert np.all(np.isnan(out[2 + 3 * n3:]))                         # USE_NAN at and after total_period().end
    n20 = 90
    t20 = T0 + (DT // 3) * np.arange(n20, dtype=np.int64)
    v20 = rng.normal(size=(n20, 2))
    out = oracle.average_accessor(t20, v20, T0 + (DT // 3) * n20, False, T0, DT, n20 // 3)
    assert np.allclose(out, v20.reshape(n20 // 3, 3, 2).mean(axis=1), rtol=1e-14, atol=1e-15)
    lin = oracle.average_accessor(t20, v20, T0 + (DT // 3) * n20, True, T0, DT, n20 // 3)
    trap = np.stack([(0.5 * v20[3 * k] + v20[3 * k + 1] + v20[3 * k + 2] + 0.5 * v20[3 * k + 3]) / 3.0 for k in range(n20 // 3 - 1)])
    # linear between points = the trapezoid rule; the reference's b = r.v - a * to_seconds(r.t) works on absolute epoch seconds (~1e9),
    # so its own arithmetic carries ~1e-9 of cancellation noise
    assert np.allclose(lin[:-1], trap, rtol=0, atol=2e-8)
