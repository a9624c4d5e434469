"""The oracle against the reference's stack-level known answers: test/pt_gs_k_test.cpp:174-354 (mass balance, land-type
routing of the response, lake / reservoir), test/pt_hs_k_test.cpp:93-153, test/hbv_snow_test.cpp:78-138 (sca after a snowfall for three
redistribution vectors)."""
import numpy as np
import pytest

import stack_cases as sc


@pytest.fixture(scope="module")
def run(oracle):
    def f(stack, geo, par, forcing, state, t0_us, T):
        fn = oracle.ptgsk_run_cells if stack == 0 else oracle.pthsk_run_cells
        return fn(geo, par, forcing, state, t0_us, 3600 * 10**6)
    return f


def test_pt_gs_k_mass_balance_and_land_type_routing(run):
    st = sc.ptgsk_mass_balance(run)
    sc.ptgsk_direct_response_on_reservoir_only(run, st)
    sc.ptgsk_glacier_and_reservoir_direct_response(run, st)


def test_pt_gs_k_lake_reservoir_response(run):
    sc.ptgsk_lake_reservoir_response(run)


def test_pt_hs_k_lake_reservoir_response(run):
    sc.pthsk_lake_reservoir_response(run)


@pytest.mark.parametrize("s,sca_after", [([1.0, 1.0, 1.0, 0.0, 0.0], 0.75), ([1.0, 1.0, 1.0, 1.0, 1.0], 1.0), ([1.0, 0.0, 0.0, 0.0, 0.0], 0.25)])
def test_hbv_snow_sca_at_snowpack_buildup(oracle, s, sca_after):
    """test_snow_distr / uniform / skewed _at_snowpack_buildup: the explicit parameter(s, intervals) constructor does not normalise s"""
    iv = [0.0, 0.25, 0.5, 0.75, 1.0]
    sp, sw, swe, sca = oracle.hbv_snow_distribute(10.0, 0.15, s, iv)
    sp, sw, swe, sca, out = oracle.hbv_snow_step(sp, sw, swe, sca, 0.15, -1.0, s=s, intervals=iv)
    assert sca == pytest.approx(sca_after, abs=1e-8)
