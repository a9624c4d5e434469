"""The library's host-side algorithms against the oracle, on the CPU (sb2_host_eval needs no device): the one-dimensional minimiser of the
state tuning, the calendar slice behind gamma_snow's day tables, the gamma unit hydrographs of the routing plan."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def capi():
    from shyft_b200 import capi as c
    c.lib()
    return c


def test_minimiser_takes_the_same_steps_as_the_oracle_restatement(capi, oracle):
    rng = np.random.default_rng(7)
    for _ in range(200):
        a, c = rng.uniform(-3, 3, 2)
        b = rng.choice([0.0, 0.1, 1.0])
        lo = rng.uniform(-6, 0)
        hi = lo + rng.uniform(0.5, 9)
        start = rng.uniform(lo, hi)
        eps = 10.0 ** rng.uniform(-6, -2)
        max_iter = int(rng.choice([5, 12, 40, 300]))
        calls = []

        def f(x):
            calls.append(x)
            return (x - a) * (x - a) + b * math.cosh(x - c)
        got = capi.host_eval(0, [a, b, c, start, lo, hi, eps, max_iter], 4)
        try:
            x, fx = oracle.find_min_single_variable(f, start, lo, hi, eps, max_iter)
            assert got[3] == 0.0 and got[0] == x and got[1] == fx      # same doubles: the same sequence of evaluation points
        except oracle.MinimiserFailure:
            assert got[3] == 1.0
        assert got[2] == len(calls)
    # argument check: start outside the bounds, NaN start
    assert capi.host_eval(0, [0, 0, 0, 5.0, 0.0, 1.0, 1e-3, 100], 4)[3] == 1.0
    assert capi.host_eval(0, [0, 0, 0, float("nan"), 0.0, 1.0, 1e-3, 100], 4)[3] == 1.0


def test_calendar_tables(capi, oracle):
    rng = np.random.default_rng(8)
    for t in np.concatenate([rng.integers(0, 2_000_000_000, 300), [0, 951782400, 1078012800, 1709164800 + 86399, 1735689599, 1735689600]]):
        t_us = int(t) * 10**6
        doy, soy = capi.host_eval(1, [float(t_us)], 2)
        assert int(doy) == oracle.day_of_year(t_us)
        assert int(soy) * 10**6 == t_us - oracle.trim_year(t_us)


def test_unit_hydrographs(capi, oracle):
    for n, alpha, beta in [(0, 1.0, 1.0), (1, 7.0, 0.0), (3, 7.0, 0.0), (3, 0.6, 0.1), (8, 3.0, 0.0), (24, 7.0, 0.0), (5, 2.0, -0.5)]:
        got = capi.host_eval(2, [n, alpha, beta], 64)
        want = oracle.make_uhg(n, alpha, beta)
        assert int(got[0]) == want.size
        np.testing.assert_allclose(got[1:1 + want.size], want, rtol=1e-12, atol=1e-15)   # two Newton solvers of the 0.99 quantile
        assert got[1:1 + want.size].sum() == pytest.approx(1.0, abs=1e-14)
    for dist, vel, dt in [(1000.0, 1.0, 3600), (3000.0, 1 / 3.6, 3600), (5400.0, 1.0, 3600), (1799.0, 1.0, 3600), (1801.0, 1.0, 3600), (86400.0, 2.0, 10800)]:
        assert int(capi.host_eval(3, [dist, vel, dt * 10**6], 1)[0]) == oracle.uhg_steps(dist, vel, dt * 10**6)
