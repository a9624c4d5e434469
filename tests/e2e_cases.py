"""Workloads of the end-to-end parity census: slices of BASELINE configs[1] / configs[2] small enough for the oracle, cut out of the
full-size synthetic region so that station distances, elevations and climate are those of the benchmark workload."""
import numpy as np

from fixtures import FORCING, geo_matrix


def region_slice(n_region, n_steps, n_stations, n_runs, config_index, run_len=16, **kw):
    """-> (geo of n_runs x run_len cells: runs of neighbouring cells spread evenly over the n_region-cell region, TimeAxis, env).
    Runs of 16 keep the interpolation's 16-cell tiles (station unions, sb2_interp.cuh) as they are in the full region."""
    from shyft_b200 import synthetic
    geo, ta, env = synthetic.make_region(n_region, n_steps, n_stations, config_index=config_index, **kw)
    starts = np.linspace(0, n_region - run_len, n_runs).astype(np.int64) // run_len * run_len
    idx = (starts[:, None] + np.arange(run_len)[None, :]).ravel()
    assert np.unique(idx).size == idx.size
    return geo[idx], ta, env


def oracle_forcing(oracle, geo, ta, env, btk_temperature=True, ncore=8):
    """interpolate() of the reference through the oracle: BTK temperature (or IDW) + IDW for the other four variables."""
    gm = geo_matrix(geo)
    dst = gm[:, :3]
    dt_us = ta.delta_t * 10**6
    f = {}
    for name in FORCING:
        xyz, vals = getattr(env, name)
        vals = oracle.average_accessor_same_axis(vals, dt_us)
        if name == "temperature" and btk_temperature:
            f[name] = oracle.btk_run(xyz, vals, dst, ta.start * 10**6, dt_us)
        else:
            mm = 20 if name in ("temperature", "precipitation") else 10
            f[name] = oracle.idw_run(name, xyz, vals, dst, oracle.idw_par(max_members=mm), dst_slope=gm[:, 5], ncore=ncore)
    return gm, f


def run_device_windows(m, ip, T, window, response_names, state_names, on_window):
    """drive sb2_run_windowed one window at a time (the same interpolate -> step sequence per window as one call over the whole axis)
    and hand each window's per-cell series to on_window(w0, got, forcing_got)"""
    from shyft_b200 import capi
    for w0 in range(0, T, window):
        wn = min(window, T - w0)
        m.run_windowed(ip, start_step=w0, n_steps=wn, window_steps=window)
        got = {k: m.response(k, w0, wn) for k in response_names}
        got.update({k: m.state_series(k, w0, wn + 1) for k in state_names})
        on_window(w0, got, {k: m.cell_forcing(k, w0, wn) for k in capi.FORCING_NAMES})
