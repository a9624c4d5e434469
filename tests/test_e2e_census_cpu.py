"""The end-to-end census (tests/e2e_census.py) on the CPU: two ORACLE runs of a BASELINE configs[1] slice whose forcing differs by a
synthetic relative perturbation of 1e-13 -- the size of the summation-order noise of the device's tensor-core interpolation.  It checks
the census itself (a zero perturbation is all within, every divergence gets a cause) and records how sensitive the reference ALGORITHM is
to last-bit noise: Brent's 12-bit search in corr_lwc (core/gamma_snow.h:214-227) flips in a few per cent of the cells per year, with
any perturbation however small.  tests/test_gpu_e2e_parity.py holds the device to the same picture."""
import numpy as np

from e2e_cases import oracle_forcing, region_slice
from e2e_census import census
from fixtures import PTGSK_DEFAULT


def test_census_on_perturbed_oracle_runs(oracle):
    from shyft_b200 import synthetic
    T = 8760
    geo, ta, env = region_slice(100000, T, 64, 16, config_index=1)
    n = geo.shape[0]
    gm, f = oracle_forcing(oracle, geo, ta, env, btk_temperature=True)
    st0 = synthetic.default_state(0, n)
    run = lambda forcing: oracle.ptgsk_run_cells(gm, PTGSK_DEFAULT, forcing, st0, ta.start * 10**6, ta.delta_t * 10**6, collect_response=True,
                                                 collect_state=True, ncore=8)
    want = run(f)
    assert np.nanmax(want["snow_swe"]) > 10.0
    same = census(run(f), want, f, f, PTGSK_DEFAULT[4])
    assert same["cell_steps_within"] == 1.0 and same["cells_with_a_divergence"] == 0
    rng = np.random.default_rng(5)
    f2 = {k: v * (1.0 + 1e-13 * rng.standard_normal(v.shape)) for k, v in f.items()}
    c = census(run(f2), want, f2, f, PTGSK_DEFAULT[4])
    assert all(v == 0 for v in c["forcing_outside_1e-11"].values())
    assert set(c["first_divergence_by_cause"]) <= {"brent_corr_lwc", "pe_cancellation_next_to_zero", "precipitation_phase_T_lt_tx", "sign_of_T",
                                                   "rk_accept_reject"}, c["first_divergence_by_cause"]
    assert c["cell_steps_within"] >= 0.93
    # the decisions that flip touch liquid water only; every other snow state stays within 1e-9 everywhere
    for k in ("gs_albedo", "gs_surface_heat", "gs_alpha", "gs_sdc_melt_mean", "gs_acc_melt", "gs_temp_swe"):
        assert c["series"][k]["outside"] == 0, (k, c["series"][k])
