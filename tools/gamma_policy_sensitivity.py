"""How far can the full-double oracle be from real Shyft in the one branch no reference test pins tightly?

The reference evaluates `gamma_p` / `lgamma` of gamma_snow under boost policies digits10<5> (a >= 2) and digits10<10> (a < 2)
(core/gamma_snow.h:189-201): 18 / 35 binary digits, i.e. series and continued fractions that stop at 2^-17 / 2^-34 relative.  boost's
arithmetic cannot be reproduced offline (boost 1.68 is absent), so the oracle and the kernels evaluate both functions in full double --
"parity unpinned" (DESIGN.md section 2).  This tool puts a NUMBER on that hole: it runs the oracle twice on BASELINE configs[0] (1 000
cells x 1 year hourly, IDW) -- once as it is, once with the incomplete gamma stopped at boost's policy epsilons and lgamma rounded to the
policy's digits (oracle/sho_core.hpp, g_gamma_policy) -- and reports the relative deviation of discharge, swe, sca and liquid water.
An error of that size is what real Shyft carries relative to exact arithmetic; the 1e-9 contract can only hold where it does not reach.

  python tools/gamma_policy_sensitivity.py [--cells 1000] [--steps 8760]      -> JSON on stdout (profiles/gamma_policy_sensitivity_r02.json)
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=8760)
    args = ap.parse_args()
    from e2e_cases import oracle_forcing
    from fixtures import PTGSK_DEFAULT
    from oracle import oracle as O
    from shyft_b200 import synthetic
    geo, ta, env = synthetic.make_region(args.cells, args.steps, 16, config_index=0, cells_per_catchment=100)
    gm, f = oracle_forcing(O, geo, ta, env, btk_temperature=False)
    st0 = synthetic.default_state(0, args.cells)
    run = lambda: O.ptgsk_run_cells(gm, PTGSK_DEFAULT, f, st0, ta.start * 10**6, ta.delta_t * 10**6, collect_response=True, collect_state=True,
                                    ncore=O.hardware_concurrency())
    O.lib().sho_set_gamma_policy(C.c_int(0))
    full = run()
    O.lib().sho_set_gamma_policy(C.c_int(1))
    pol = run()
    O.lib().sho_set_gamma_policy(C.c_int(0))
    out = {"workload": f"BASELINE configs[0]: pt_gs_k {args.cells} cells x {args.steps} hourly steps, IDW interpolation, oracle full double vs policy epsilons",
           "policy": {"a<2": "35 bits: series / fraction stop at 2^-34, lgamma rounded to 35 bits", "a>=2": "18 bits: stop at 2^-17, lgamma rounded to 18 bits"},
           "max_swe_mm": float(np.nanmax(full["snow_swe"])), "series": {}}
    for name in ("avg_discharge", "snow_swe", "snow_sca", "snow_outflow", "gs_lwc", "gs_alpha", "gs_sdc_melt_mean", "gs_acc_melt"):
        a, b = full[name], pol[name]
        scale = np.nanmax(np.abs(a))
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-6 * scale)        # relative to the value, floored at 1e-6 of the series' maximum
        rel = rel[np.isfinite(rel)]
        out["series"][name] = {"max_rel": float(rel.max()), "p99_rel": float(np.percentile(rel, 99)), "p90_rel": float(np.percentile(rel, 90)),
                               "median_rel": float(np.median(rel)), "share_within_1e-9": float((rel <= 1e-9).mean()),
                               "share_within_1e-6": float((rel <= 1e-6).mean()), "share_within_1e-4": float((rel <= 1e-4).mean())}
    # the catchment sums the reference's tests look at (region level)
    qa, qb = full["avg_discharge"].sum(axis=1), pol["avg_discharge"].sum(axis=1)
    out["region_discharge"] = {"max_rel": float(np.max(np.abs(qa - qb) / np.abs(qa))), "annual_volume_rel": float(abs(qa.sum() - qb.sum()) / qa.sum())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
