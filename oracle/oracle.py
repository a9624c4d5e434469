"""ctypes front end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  It builds oracle/libsho_oracle.so on demand (g++ only) and exposes the
restated reference functions with numpy arrays.  See oracle/sho_core.hpp for the parity status.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)

N_GEO = 12  # x y z area catchment_id radiation_slope_factor glacier lake reservoir forest routing_id routing_distance
PTGSK_RESPONSES = ("avg_discharge", "charge_m3s", "snow_sca", "snow_swe", "snow_outflow", "glacier_melt", "ae_output", "pe_output")
PTGSK_STATES = ("kirchner_discharge", "gs_albedo", "gs_lwc", "gs_surface_heat", "gs_alpha", "gs_sdc_melt_mean", "gs_acc_melt",
                "gs_iso_pot_energy", "gs_temp_swe")
HS_RESPONSES = PTGSK_RESPONSES + ("soil_outflow",)


GEO_COLS = ("x", "y", "z", "area", "catchment_id", "radiation_slope_factor", "glacier", "lake", "reservoir", "forest", "routing_id",
            "routing_distance")


def geo_matrix(geo_records):
    """numpy record array with the fields of sb2_geo_cell -> the oracle's [n][12] double matrix"""
    return np.stack([np.asarray(geo_records[c], dtype=np.float64) for c in GEO_COLS], axis=1)


def build(force=False):
    so = os.path.join(_HERE, "libsho_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "sho_detmath.hpp", "sho_math_tables.inc", "sho_core.hpp", "sho_pt_gs_k.hpp", "sho_hbv.hpp", "sho_skaugen.hpp", "sho_hps.hpp", "sho_region.hpp", "sho_ts.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libsho_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.sho_last_error.restype = C.c_char_p
        for name in ("sho_day_of_year", "sho_trim_year", "sho_calendar_time", "sho_catchment_index"):
            getattr(L, name).restype = C.c_int64
        for name in ("sho_gamma_p", "sho_gamma_quantile", "sho_pt_potential_evapotranspiration", "sho_ae_calculate_step",
                     "sho_glacier_melt_step", "sho_gs_corr_lwc", "sho_gs_calc_q", "sho_hbv_soil_step", "sho_hbv_tank_step",
                     "sho_hbv_ae_step", "sho_btk_prior_gradient", "sho_nash_sutcliffe", "sho_rmse", "sho_kling_gupta", "sho_abs_diff_sum",
                     "sho_average_value", "sho_skaugen_sca_rel_red"):
            getattr(L, name).restype = C.c_double
    return _LIB


def _d(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().sho_last_error().decode())


def hardware_concurrency():
    return int(lib().sho_hardware_concurrency())


DM_FUNCTIONS = dict(exp=0, log=1, pow=2, lgamma=3, gamma_p=4)


def dm_eval(fn, a, b=None):
    """Deterministic elementary functions of oracle/sho_detmath.hpp (and gamma_p built on them), elementwise."""
    a = _f64(a).ravel()
    b = _f64(b).ravel() if b is not None else np.zeros_like(a)
    out = np.zeros_like(a)
    lib().sho_dm_eval(C.c_int(DM_FUNCTIONS[fn]), C.c_int64(a.size), _d(a), _d(b), _d(out))
    return out


# ---- calendar / special ---------------------------------------------------------------------
def day_of_year(t_us):
    return int(lib().sho_day_of_year(C.c_int64(int(t_us))))


def trim_year(t_us):
    return int(lib().sho_trim_year(C.c_int64(int(t_us))))


def calendar_time(y, m, d):
    return int(lib().sho_calendar_time(int(y), int(m), int(d)))


def gamma_p(a, x):
    return float(lib().sho_gamma_p(C.c_double(a), C.c_double(x)))


def gamma_quantile(alpha, p):
    return float(lib().sho_gamma_quantile(C.c_double(alpha), C.c_double(p)))


# ---- method units ------------------------------------------------------------------------------
def pt_potential_evapotranspiration(albedo, alpha, t, rad, rh):
    return float(lib().sho_pt_potential_evapotranspiration(*(C.c_double(v) for v in (albedo, alpha, t, rad, rh))))


def ae_calculate_step(water_level, pot, scale, snow_fraction):
    return float(lib().sho_ae_calculate_step(*(C.c_double(v) for v in (water_level, pot, scale, snow_fraction))))


def glacier_melt_step(dtf, t, sca_m2, glacier_m2):
    return float(lib().sho_glacier_melt_step(*(C.c_double(v) for v in (dtf, t, sca_m2, glacier_m2))))


def kirchner_step(q, p, e, dt_us=3600 * 10**6, c=(-2.439, 0.966, -0.10), abs_err=1e-7, rel_err=1e-8):
    """-> (q_end, q_avg, n_accepted, n_rejected)"""
    qq, qa = C.c_double(q), C.c_double(0.0)
    na, nr = C.c_int(0), C.c_int(0)
    _check(lib().sho_kirchner_step(C.c_double(c[0]), C.c_double(c[1]), C.c_double(c[2]), C.c_double(abs_err), C.c_double(rel_err),
                                   C.c_int64(dt_us), C.byref(qq), C.byref(qa), C.c_double(p), C.c_double(e), C.byref(na), C.byref(nr)))
    return qq.value, qa.value, na.value, nr.value


GS_PARAM_DEFAULT = dict(winter_end_day_of_year=100, initial_bare_ground_fraction=0.04, snow_cv=0.4, tx=-0.5, wind_scale=2.0, wind_const=1.0,
                        max_water=0.1, surface_magnitude=30.0, max_albedo=0.9, min_albedo=0.6, fast_albedo_decay_rate=5.0,
                        slow_albedo_decay_rate=5.0, snowfall_reset_depth=5.0, glacier_albedo=0.4, calculate_iso_pot_energy=0.0,
                        snow_cv_forest_factor=0.0, snow_cv_altitude_factor=0.0, n_winter_days=221)


def gs_param(**kw):
    d = dict(GS_PARAM_DEFAULT)
    d.update(kw)
    return _f64([float(d[k]) for k in GS_PARAM_DEFAULT])


def gs_calc_snow_state(shape, scale, y0, lam, lwd, max_water_frac, temp_swe):
    swe, sca = C.c_double(0), C.c_double(0)
    lib().sho_gs_calc_snow_state(*(C.c_double(v) for v in (shape, scale, y0, lam, lwd, max_water_frac, temp_swe)), C.byref(swe), C.byref(sca))
    return swe.value, sca.value


def gs_corr_lwc(z1, a1, b1, z2, a2, b2):
    n = C.c_int(0)
    r = lib().sho_gs_corr_lwc(*(C.c_double(v) for v in (z1, a1, b1, z2, a2, b2)), C.byref(n))
    return float(r), n.value


def gs_calc_q(a, b, z):
    return float(lib().sho_gs_calc_q(C.c_double(a), C.c_double(b), C.c_double(z)))


def gs_reset_snow_pack(storage, **kw):
    out = np.zeros(6)
    p = gs_param(**kw)
    lib().sho_gs_reset_snow_pack(_d(p), C.c_double(storage), _d(out))
    return dict(zip(("sca", "lwc", "alpha", "sdc_melt_mean", "acc_melt", "temp_swe"), out))


def gs_step(state8, t_us, dt_us, T, rad, prec_mm_h, wind_speed, rel_hum, forest_fraction=0.0, altitude=0.0, **kw):
    """state8: albedo lwc surface_heat alpha sdc_melt_mean acc_melt iso_pot_energy temp_swe -> (state8', (sca, storage, outflow))"""
    s = _f64(state8).copy()
    r = np.zeros(3)
    p = gs_param(**kw)
    _check(lib().sho_gs_step(_d(p), _d(s), _d(r), C.c_int64(t_us), C.c_int64(dt_us),
                             *(C.c_double(v) for v in (T, rad, prec_mm_h, wind_speed, rel_hum, forest_fraction, altitude))))
    return s, r


# ---- stack runs -----------------------------------------------------------------------------------
def _ptrs(arrs):
    P = (c_dp * len(arrs))()
    for i, a in enumerate(arrs):
        P[i] = _d(a) if a is not None else None
    return P


def ptgsk_run_cells(geo, params, forcing, state, t0_us, dt_us, n_axis=None, start_step=0, n_steps=0, pset_of_cell=None,
                    collect_response=True, collect_state=False, collect_substeps=False, cell_mask=None, ncore=1):
    """Reference-equivalent pt_gs_k run_cells on the CPU.

    geo [n][12]; params [n_sets][31]; forcing dict of [T][n] arrays (temperature, precipitation, radiation, wind_speed, rel_hum);
    state [n][9] (8 gamma_snow doubles + kirchner.q).  Returns dict(state=[n][9], <response>=[T][n], <state series>=[T+1][n]).
    """
    geo = _f64(geo)
    n = geo.shape[0]
    params = _f64(params).reshape(-1, 31)
    f = [_f64(forcing[k]) for k in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum")]
    T = f[0].shape[0] if n_axis is None else n_axis
    st = _f64(state).reshape(n, 9).copy()
    out = {}
    resp = [np.full((T, n), np.nan) if collect_response else None for _ in PTGSK_RESPONSES]
    sts = [np.full((T + 1, n), np.nan) if collect_state else None for _ in PTGSK_STATES]
    sub = np.zeros((T, n)) if collect_substeps else None
    ps = np.ascontiguousarray(pset_of_cell, dtype=np.int32) if pset_of_cell is not None else None
    mask = np.ascontiguousarray(cell_mask, dtype=np.uint8) if cell_mask is not None else None
    _check(lib().sho_ptgsk_run_cells(
        C.c_int64(n), _d(geo), C.c_int64(params.shape[0]), _d(params), ps.ctypes.data_as(c_i32p) if ps is not None else None,
        C.c_int64(t0_us), C.c_int64(dt_us), C.c_int64(T), C.c_int(start_step), C.c_int(n_steps),
        _d(f[0]), _d(f[1]), _d(f[2]), _d(f[3]), _d(f[4]), C.c_int64(n), C.c_int64(1), _d(st),
        _ptrs(resp) if collect_response else None, _ptrs(sts) if collect_state else None, _d(sub) if sub is not None else None,
        C.c_int64(n), C.c_int64(1), mask.ctypes.data_as(c_u8p) if mask is not None else None, C.c_int(ncore)))
    out["state"] = st
    if collect_response:
        out.update(dict(zip(PTGSK_RESPONSES, resp)))
    if collect_state:
        out.update(dict(zip(PTGSK_STATES, sts)))
    if collect_substeps:
        out["kirchner_substeps"] = sub
    return out


def _hs_run(fn, n_param, geo, params, forcing, state, t0_us, dt_us, start_step, n_steps, pset_of_cell, ncore):
    geo = _f64(geo)
    n = geo.shape[0]
    params = _f64(params).reshape(-1, n_param)
    f = [_f64(forcing[k]) for k in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum")]
    T = f[0].shape[0]
    st = _f64(state).copy()
    n_state = st.shape[1]
    resp = [np.full((T, n), np.nan) for _ in HS_RESPONSES]
    ps = np.ascontiguousarray(pset_of_cell, dtype=np.int32) if pset_of_cell is not None else None
    _check(fn(C.c_int64(n), _d(geo), C.c_int64(params.shape[0]), _d(params), C.c_int(n_param), ps.ctypes.data_as(c_i32p) if ps is not None else None,
              C.c_int64(t0_us), C.c_int64(dt_us), C.c_int64(T), C.c_int(start_step), C.c_int(n_steps),
              _d(f[0]), _d(f[1]), _d(f[2]), _d(f[3]), _d(f[4]), C.c_int64(n), C.c_int64(1), _d(st), C.c_int(n_state),
              _ptrs(resp), C.c_int64(n), C.c_int64(1), C.c_int(ncore)))
    out = dict(zip(HS_RESPONSES, resp))
    out["state"] = st
    return out


def pthsk_run_cells(geo, params, forcing, state, t0_us, dt_us, start_step=0, n_steps=0, pset_of_cell=None, ncore=1):
    """state [n][3+2*nb] = swe, sca, sp[nb], sw[nb], kirchner.q"""
    return _hs_run(lib().sho_pthsk_run_cells, 18, geo, params, forcing, state, t0_us, dt_us, start_step, n_steps, pset_of_cell, ncore)


def hbv_stack_run_cells(geo, params, forcing, state, t0_us, dt_us, start_step=0, n_steps=0, pset_of_cell=None, ncore=1):
    """state [n][5+2*nb] = swe, sca, sp[nb], sw[nb], soil.sm, tank.uz, tank.lz"""
    return _hs_run(lib().sho_hbv_stack_run_cells, 22, geo, params, forcing, state, t0_us, dt_us, start_step, n_steps, pset_of_cell, ncore)


PTSSK_STATES = ("kirchner_discharge", "snow_swe", "snow_sca", "snow_alpha", "snow_nu", "snow_lwc", "snow_residual")
SKAUGEN_DEFAULT = (40.77, 113.0, 0.1, 0.1, 0.16, 2.5, 0.14, 0.01)   # alpha_0 d_range unit_size max_water_fraction tx cx ts cfr (skaugen.h:92-99)


def ptssk_run_cells(geo, params, forcing, state, t0_us, dt_us, start_step=0, n_steps=0, pset_of_cell=None, collect_state=False, ncore=1):
    """Reference-equivalent pt_ss_k run_cells (core/pt_ss_k.h:195-293).  state [n][8] = snow.nu, alpha, sca, swe, free_water, residual,
    num_units, kirchner.q.  -> dict(state, <response series [T][n]>, <state series [T+1][n]> when collect_state)"""
    geo = _f64(geo)
    n = geo.shape[0]
    params = _f64(params).reshape(-1, 21)
    f = [_f64(forcing[k]) for k in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum")]
    T = f[0].shape[0]
    st = _f64(state).reshape(n, 8).copy()
    resp = [np.full((T, n), np.nan) for _ in HS_RESPONSES]
    sts = [np.full((T + 1, n), np.nan) if collect_state else None for _ in PTSSK_STATES]
    ps = np.ascontiguousarray(pset_of_cell, dtype=np.int32) if pset_of_cell is not None else None
    _check(lib().sho_ptssk_run_cells(C.c_int64(n), _d(geo), C.c_int64(params.shape[0]), _d(params), C.c_int(21), ps.ctypes.data_as(c_i32p) if ps is not None else None,
                                     C.c_int64(t0_us), C.c_int64(dt_us), C.c_int64(T), C.c_int(start_step), C.c_int(n_steps),
                                     _d(f[0]), _d(f[1]), _d(f[2]), _d(f[3]), _d(f[4]), C.c_int64(n), C.c_int64(1), _d(st),
                                     _ptrs(resp), _ptrs(sts) if collect_state else None, C.c_int64(n), C.c_int64(1), C.c_int(ncore)))
    out = dict(zip(HS_RESPONSES, resp))
    out.pop("soil_outflow")
    out["state"] = st
    if collect_state:   # state series share names with response series (snow_swe, snow_sca): keyed "state_<name>"
        out.update({"state_" + k: v for k, v in zip(PTSSK_STATES, sts)})
    return out


def skaugen_step(state7, temp, prec, dt_us=3600 * 10**6, par=SKAUGEN_DEFAULT):
    """skaugen::calculator::step -> (new state [nu, alpha, sca, swe, free_water, residual, num_units], (outflow, sca, swe))"""
    s = _f64(state7).copy()
    out = np.zeros(3)
    _check(lib().sho_skaugen_step(_d(_f64(par)), _d(s), C.c_int64(dt_us), C.c_double(temp), C.c_double(prec), _d(out)))
    return s, out


def skaugen_sca_rel_red(u, n, nu_a, alpha, unit_size=0.1):
    return float(lib().sho_skaugen_sca_rel_red(C.c_double(u), C.c_double(n), C.c_double(unit_size), C.c_double(nu_a), C.c_double(alpha)))


HPS_DEFAULT = (0.0, 0.1, 0.5, 2.0, 1.0, 30.0, 0.9, 0.6, 5.0, 5.0, 5.0)   # tx lw cfr wind_scale wind_const surface_magnitude max/min albedo fast/slow decay reset_depth


def pthpsk_run_cells(geo, params, forcing, state, t0_us, dt_us, start_step=0, n_steps=0, pset_of_cell=None, ncore=1):
    """Reference-equivalent pt_hps_k run_cells (core/pt_hps_k.h:201-303).  state [n][24] = sp[5], sw[5], albedo[5], iso_pot_energy[5],
    surface_heat, swe, sca, kirchner.q"""
    out = _hs_run(lib().sho_pthpsk_run_cells, 24, geo, params, forcing, state, t0_us, dt_us, start_step, n_steps, pset_of_cell, ncore)
    out.pop("soil_outflow")
    return out


def hps_step(state, T, rad, prec, wind_speed, rel_hum, dt_us=3600 * 10**6, s=(1.0,) * 5, intervals=(0.0, 0.25, 0.5, 0.75, 1.0), par=HPS_DEFAULT, iso=False,
             distribute=False):
    """hbv_physical_snow::calculator::step on state [sp.., sw.., albedo.., iso.., surface_heat, swe, sca] -> (state, (outflow, sca, storage))"""
    st = _f64(state).copy()
    out = np.zeros(3)
    _check(lib().sho_hps_step(_d(_f64(s)), _d(_f64(intervals)), C.c_int(len(s)), _d(_f64(par)), C.c_int(int(iso)), C.c_int(int(distribute)), _d(st),
                              C.c_int64(dt_us), _d(_f64([T, rad, prec, wind_speed, rel_hum])), _d(out)))
    return st, out


def hbv_snow_step(sp, sw, swe, sca, prec, temp, dt_us=3600 * 10**6, s=None, intervals=None, tx=0.0, cx=1.0, ts=0.0, lw=0.1, cfr=0.5):
    n = len(sp)
    s = _f64(s if s is not None else [1.0] * n)
    iv = _f64(intervals if intervals is not None else np.linspace(0, 1, n))
    sp, sw = _f64(sp).copy(), _f64(sw).copy()
    cswe, csca, out = C.c_double(swe), C.c_double(sca), C.c_double(0)
    par = _f64([tx, cx, ts, lw, cfr])
    _check(lib().sho_hbv_snow_step(_d(s), _d(iv), C.c_int(n), _d(par), _d(sp), _d(sw), C.byref(cswe), C.byref(csca), C.c_int64(dt_us),
                                   C.c_double(prec), C.c_double(temp), C.byref(out)))
    return sp, sw, cswe.value, csca.value, out.value


def hbv_snow_distribute(swe, sca, s, intervals, lw=0.1):
    """hbv_snow::state::distribute(p): the bins sp / sw of a pack given by swe and sca"""
    s, iv = _f64(s), _f64(intervals)
    sp, sw = np.zeros(s.size), np.zeros(s.size)
    cswe, csca = C.c_double(swe), C.c_double(sca)
    _check(lib().sho_hbv_snow_distribute(_d(s), _d(iv), C.c_int(s.size), C.c_double(lw), _d(sp), _d(sw), C.byref(cswe), C.byref(csca)))
    return sp, sw, cswe.value, csca.value


def hbv_soil_step(sm, insoil, act_evap, fc=300.0, beta=2.0):
    s = C.c_double(sm)
    out = lib().sho_hbv_soil_step(C.c_double(fc), C.c_double(beta), C.byref(s), C.c_double(insoil), C.c_double(act_evap))
    return s.value, float(out)


def hbv_tank_step(uz, lz, soil_outflow, uz1=25.0, kuz2=0.5, kuz1=0.3, perc=0.8, klz=0.02):
    u, l = C.c_double(uz), C.c_double(lz)
    par = _f64([uz1, kuz2, kuz1, perc, klz])
    out = lib().sho_hbv_tank_step(_d(par), C.byref(u), C.byref(l), C.c_double(soil_outflow))
    return u.value, l.value, float(out)


def hbv_ae_step(sm, pot, lp, snow_fraction):
    return float(lib().sho_hbv_ae_step(*(C.c_double(v) for v in (sm, pot, lp, snow_fraction))))


# ---- interpolation -----------------------------------------------------------------------------------
IDW_KINDS = dict(temperature=0, precipitation=1, radiation=2, wind_speed=3, rel_hum=4)


def idw_par(max_members=10, max_distance=200000.0, distance_measure_factor=2.0, zscale=1.0, default_temp_gradient=-0.006,
            gradient_by_equation=False, scale_factor=1.02):
    return _f64([max_members, max_distance, distance_measure_factor, zscale, default_temp_gradient, float(gradient_by_equation), scale_factor])


def idw_run(kind, src_xyz, src_values, dst_xyz, par, dst_slope=None, ncore=1):
    """src_values [T][n_src] -> [T][n_dst]"""
    src_xyz, dst_xyz, v = _f64(src_xyz), _f64(dst_xyz), _f64(src_values)
    T, ns = v.shape
    nd = dst_xyz.shape[0]
    out = np.full((T, nd), np.nan)
    sl = _f64(dst_slope) if dst_slope is not None else None
    _check(lib().sho_idw_run(C.c_int(IDW_KINDS[kind]), C.c_int64(ns), _d(src_xyz), _d(v), C.c_int64(T), C.c_int64(nd), _d(dst_xyz),
                             _d(sl) if sl is not None else None, _d(_f64(par)), _d(out), C.c_int64(nd), C.c_int64(1), C.c_int(ncore)))
    return out


def idw_neighbours(src_xyz, dst_xyz, par):
    src_xyz, dst_xyz = _f64(src_xyz), _f64(dst_xyz)
    nd, mm = dst_xyz.shape[0], int(par[0])
    idx = np.full((nd, mm), -1, dtype=np.int32)
    w = np.zeros((nd, mm))
    cnt = np.zeros(nd, dtype=np.int32)
    _check(lib().sho_idw_neighbours(C.c_int64(src_xyz.shape[0]), _d(src_xyz), C.c_int64(nd), _d(dst_xyz), _d(_f64(par)),
                                    idx.ctypes.data_as(c_i32p), _d(w), cnt.ctypes.data_as(c_i32p)))
    return idx, w, cnt


def btk_par(gradient_sd=0.0025, sill=25.0, nug=0.5, range_=200000.0, zscale=20.0, fixed_gradient=float("nan")):
    return _f64([gradient_sd, sill, nug, range_, zscale, fixed_gradient])


def btk_run(src_xyz, src_values, dst_xyz, t0_us, dt_us, par=None):
    src_xyz, dst_xyz, v = _f64(src_xyz), _f64(dst_xyz), _f64(src_values)
    T, ns = v.shape
    nd = dst_xyz.shape[0]
    out = np.full((T, nd), np.nan)
    _check(lib().sho_btk_run(C.c_int64(ns), _d(src_xyz), _d(v), C.c_int64(t0_us), C.c_int64(dt_us), C.c_int64(T), C.c_int64(nd), _d(dst_xyz),
                             _d(par if par is not None else btk_par()), _d(out), C.c_int64(nd), C.c_int64(1)))
    return out


def btk_covariance(src_xyz, dst_xyz, par=None):
    src_xyz, dst_xyz = _f64(src_xyz), _f64(dst_xyz)
    ns, nd = src_xyz.shape[0], dst_xyz.shape[0]
    K, k = np.zeros((ns, ns)), np.zeros((ns, nd))
    _check(lib().sho_btk_covariance(C.c_int64(ns), _d(src_xyz), C.c_int64(nd), _d(dst_xyz), _d(par if par is not None else btk_par()), _d(K), _d(k)))
    return K, k


def btk_prior_gradient(t_us, dt_us):
    return float(lib().sho_btk_prior_gradient(C.c_int64(t_us), C.c_int64(dt_us)))


def average_accessor_same_axis(values, dt_us):
    v = _f64(values).copy()
    lib().sho_average_accessor_same_axis(_d(v), C.c_int64(v.size), C.c_int64(dt_us))
    return v


# ---- routing -----------------------------------------------------------------------------------------
def make_uhg(n_steps, alpha, beta):
    out = np.zeros(max(1, n_steps))
    n = C.c_int(0)
    _check(lib().sho_make_uhg(C.c_int(n_steps), C.c_double(alpha), C.c_double(beta), _d(out), C.byref(n)))
    return out[: n.value].copy()


def uhg_steps(distance, velocity, dt_us):
    return int(lib().sho_uhg_steps(C.c_double(distance), C.c_double(velocity), C.c_int64(dt_us)))


def river_flows(rivers, rid, cell_discharge, cell_rid, cell_distance, cell_uhg_par, dt_us):
    """rivers [n][6] = id downstream_id distance velocity alpha beta; cell_discharge [T][n_cells] -> (local, upstream, output) each [T]"""
    rivers, q = _f64(rivers).reshape(-1, 6), _f64(cell_discharge)
    T, n = q.shape
    out = np.zeros((3, T))
    crid = np.ascontiguousarray(cell_rid, dtype=np.int64)
    _check(lib().sho_river_flows(C.c_int64(rivers.shape[0]), _d(rivers), C.c_int64(rid), C.c_int64(n), _d(q), C.c_int64(n), C.c_int64(1),
                                 crid.ctypes.data_as(c_i64p), _d(_f64(cell_distance)), _d(_f64(cell_uhg_par)), C.c_int64(T), C.c_int64(dt_us), _d(out)))
    return out[0], out[1], out[2]


def average_to_axis(values, dt_us, first, k, n_out):
    """True average of a stair-case series over aligned coarser periods (average_accessor semantics)."""
    v = _f64(values)
    out = np.zeros(n_out)
    lib().sho_average_to_axis(_d(v), C.c_int64(v.size), C.c_int64(dt_us), C.c_int64(first), C.c_int64(k), C.c_int64(n_out), _d(out))
    return out


# ---- goal functions -----------------------------------------------------------------------------------
def nash_sutcliffe(o, m):
    o, m = _f64(o), _f64(m)
    return float(lib().sho_nash_sutcliffe(_d(o), _d(m), C.c_int64(o.size)))


def rmse(o, m):
    o, m = _f64(o), _f64(m)
    return float(lib().sho_rmse(_d(o), _d(m), C.c_int64(o.size)))


def kling_gupta(o, m, s_r=1.0, s_a=1.0, s_b=1.0):
    o, m = _f64(o), _f64(m)
    return float(lib().sho_kling_gupta(_d(o), _d(m), C.c_int64(o.size), C.c_double(s_r), C.c_double(s_a), C.c_double(s_b)))


def abs_diff_sum(o, m):
    o, m = _f64(o), _f64(m)
    return float(lib().sho_abs_diff_sum(_d(o), _d(m), C.c_int64(o.size)))


def max_abs_average_to_axis(values, dt_us, first, k, n_out):
    """max_abs_average_accessor (core/time_series.h:2198-2267) over aligned coarser periods: the larger of the true averages of
    max(0, v) and max(0, -v) (finite values only are mapped, source_max_abs :2208-2213)"""
    v = _f64(values)
    fin = np.isfinite(v)
    pos = np.where(fin, np.maximum(0.0, v), v)
    neg = np.where(fin, np.maximum(0.0, -v), v)
    a, b = average_to_axis(pos, dt_us, first, k, n_out), average_to_axis(neg, dt_us, first, k, n_out)
    return np.where(a < b, b, a)   # std::max(pos_val, neg_val)


def abs_diff_sum_scaled(o, m, s):
    """abs_diff_sum_goal_function_scaled (core/time_series.h:2435-2448)"""
    r = 0.0
    for tv, dv, sv in zip(_f64(o), _f64(m), _f64(s)):
        if np.isfinite(tv) and np.isfinite(dv) and np.isfinite(sv) and abs(sv) > 1e-20:
            r += abs(tv - dv) / sv
    return r


def catchment_index(cids):
    cid = np.ascontiguousarray(cids, dtype=np.int64)
    cix = np.zeros_like(cid)
    m = np.zeros_like(cid)
    n = lib().sho_catchment_index(C.c_int64(cid.size), cid.ctypes.data_as(c_i64p), cix.ctypes.data_as(c_i64p), m.ctypes.data_as(c_i64p))
    return cix, m[:n].copy()


# ---- statistics readers: cell_statistics, core/cell_model.h:194-406 (plain restatement, cell after cell) ---------------------
CATCHMENT_IX, CELL_IX = 0, 1  # stat_scope (:178-181)


def _stat_matches(catchment_ids, indexes, scope):
    """verify_cids_exist (:197-213) + the cells selected by is_match (:215-218), in cell order"""
    cids = np.asarray(catchment_ids, dtype=np.int64)
    n = cids.size
    if n == 0:
        raise RuntimeError("no cells to make statistics on")
    if len(indexes) == 0:
        return list(range(n))
    if scope == CELL_IX:
        for cid in indexes:
            if cid < 0 or cid > n:
                raise RuntimeError(f"Supplied cell index reference {cid} is ouside valid range 0 ..{n}")
        return [i for i in range(n) if any(cid == i for cid in indexes)]
    known = set(cids.tolist())
    for cid in indexes:
        if cid not in known:
            raise RuntimeError(f"one or more supplied catchment_indexes does not exist:{cid}")
    return [i for i in range(n) if any(cids[i] == cid for cid in indexes)]


def sum_catchment_feature(series, catchment_ids, indexes=(), scope=CATCHMENT_IX):
    """:313-333; series [T][n_cells] -> [T]"""
    r = np.zeros(series.shape[0])
    for i in _stat_matches(catchment_ids, indexes, scope):
        r = r + series[:, i]                     # pts_t::add
    return r


def average_catchment_feature(series, area, catchment_ids, indexes=(), scope=CATCHMENT_IX):
    """:230-268: r += ts * area per cell, then r *= 1 / sum_area"""
    r = np.zeros(series.shape[0])
    sum_area = 0.0
    for i in _stat_matches(catchment_ids, indexes, scope):
        r = r + series[:, i] * area[i]           # pts_t::add_scale (time_series.h:399-401)
        sum_area += area[i]
    return r * (1 / sum_area) if sum_area != 0.0 else r * np.inf


def sum_catchment_feature_value(series, catchment_ids, indexes, j, scope=CATCHMENT_IX):
    """:347-369"""
    r = 0.0
    for i in _stat_matches(catchment_ids, indexes, scope):
        r += series[j, i]
    return r


def average_catchment_feature_value(series, area, catchment_ids, indexes, j, scope=CATCHMENT_IX):
    """:282-308: r / sum_area (a division, unlike the series form)"""
    r, sum_area = 0.0, 0.0
    for i in _stat_matches(catchment_ids, indexes, scope):
        r += series[j, i] * area[i]
        sum_area += area[i]
    return r / sum_area


def catchment_feature(series, catchment_ids, indexes, j, scope=CATCHMENT_IX):
    """:381-402: the selected cells' values at step j, in cell order"""
    return np.array([series[j, i] for i in _stat_matches(catchment_ids, indexes, scope)])


def ae_pot_ratio(kirchner_discharge_m3s, area, ae_scale_factor):
    """actual_evapotranspiration_cell_response_statistics::pot_ratio's per-cell series (api/api.h:1527-1541)"""
    kd = np.asarray(kirchner_discharge_m3s, dtype=np.float64)
    out = np.empty_like(kd)
    for c in range(kd.shape[1]):
        q_mmh = kd[:, c] / ((1 / (3600.0 * 1000.0)) * area[c])                     # m3s_to_mmh, unit_conversion.h:11-15
        out[:, c] = 1.0 - dm_eval("exp", -q_mmh * 3.0 / ae_scale_factor)           # calc_pot_ratio, actual_evapotranspiration.h:40-43
    return out


# ---- projection of point sources onto a time axis: core/time_series.h:144-310, 2033-2072 (oracle/sho_ts.hpp) -----------------
def average_value(t_us, v, p_start_us, p_end_us, ix=-1, linear=False):
    """average_value(source, period, last_idx, linear) -> (value, last_idx); ix = -1 is "no hint" (npos)"""
    t = np.ascontiguousarray(t_us, dtype=np.int64)
    vv = _f64(v)
    hint = C.c_int64(ix)
    r = lib().sho_average_value(t.ctypes.data_as(c_i64p), _d(vv), C.c_int64(t.size), C.c_int64(p_start_us), C.c_int64(p_end_us), C.byref(hint),
                                C.c_int(1 if linear else 0))
    return r, hint.value


def average_accessor(t_us, values, t_end_us, linear, ta_t0_us, ta_dt_us, ta_n):
    """average_accessor<point_ts, fixed_dt>::value(i), i = 0..ta_n-1, for sources sharing one point axis: values [n_points][n_src]"""
    t = np.ascontiguousarray(t_us, dtype=np.int64)
    vv = _f64(values).reshape(t.size, -1)
    out = np.zeros((ta_n, vv.shape[1]))
    lib().sho_average_accessor(t.ctypes.data_as(c_i64p), _d(vv), C.c_int64(t.size), C.c_int64(vv.shape[1]), C.c_int64(t_end_us), C.c_int(1 if linear else 0),
                               C.c_int64(ta_t0_us), C.c_int64(ta_dt_us), C.c_int64(ta_n), _d(out))
    return out


# ---- state tuning: core/model_state_tuning.h:38-118 over region_model::adjust_state_to_target_flow (core/region_model.h:626-637) ----
# dlib 19.16 (pinned in build_support/build_dependencies.sh) is not under /root/reference: find_min_single_variable is restated from
# its published algorithm (dlib/optimization/optimization_line_search.h: bracket by radius doubling, then three-point parabola
# steps with the 0.1 keep-away rule).  PARITY UNPINNED at the dlib boundary: the reference's tests only assert the reached flow to
# 2 decimals (shyft/tests/api/test_region_model_stacks.py:318-333; test/cell_builder_test.cpp:281-296) -- those asserts pass here.
class MinimiserFailure(RuntimeError):
    pass


def _poly_min_3(p1, p2, p3, f1, f2, f3):
    t1 = f1 * (p3 * p3 - p2 * p2) + f2 * (p1 * p1 - p3 * p3) + f3 * (p2 * p2 - p1 * p1)
    t2 = 2 * (f1 * (p3 - p2) + f2 * (p1 - p3) + f3 * (p2 - p1))
    if t2 == 0:
        return p2
    r = t1 / t2
    return r if p1 <= r <= p3 else min(max(p1, r), p3)


def find_min_single_variable(f, start, begin=-1e200, end=1e200, eps=1e-3, max_iter=100, initial_search_radius=1.0):
    """-> (argmin, f(argmin)); raises MinimiserFailure like dlib's argument check / iteration limit"""
    if not (eps > 0 and max_iter > 1 and begin <= start <= end and initial_search_radius > 0):
        raise MinimiserFailure("find_min_single_variable: eps > 0, max_iter > 1 and begin <= starting_point <= end are required")
    radius = initial_search_radius
    evals = 1
    if begin == end:
        return start, f(start)
    p1, p3 = max(start - radius, begin), min(start + radius, end)
    f1, f3 = f(p1), f(p3)
    if start == p1 or start == p3:
        p2 = (p1 + p3) / 2
    else:
        p2 = start
    f2 = f(p2)
    evals += 2
    too_many = "The max number of iterations of single variable optimization have been reached without converging."
    while not (f1 > f2 and f2 < f3):
        if evals >= max_iter:
            raise MinimiserFailure(too_many)
        if p3 - p1 < eps:
            if f1 < min(f2, f3):
                return p1, f1
            if f2 < min(f1, f3):
                return p2, f2
            return p3, f3
        if f1 <= f3:
            if p1 == begin or (f1 == f2 and (end - begin) < radius):
                p3, f3 = p2, f2
                p2 = (p1 + p2) / 2.0
                f2 = f(p2)
            else:
                p3, f3, p2, f2 = p2, f2, p1, f1
                p1 = max(p1 - radius, begin)
                f1 = f(p1)
                radius *= 2
        else:
            if p3 == end or (f2 == f3 and (end - begin) < radius):
                p1, f1 = p2, f2
                p2 = (p3 + p2) / 2.0
                f2 = f(p2)
            else:
                p1, f1, p2, f2 = p2, f2, p3, f3
                p3 = min(p3 + radius, end)
                f3 = f(p3)
                radius *= 2
        evals += 1
    tau = 0.1
    while evals < max_iter and p3 - p1 > eps:
        pm = _poly_min_3(p1, p2, p3, f1, f2, f3)
        if pm < p2:
            d = (p2 - p1) * tau
            if abs(p1 - pm) < d:
                pm = p1 + d
            elif abs(p2 - pm) < d:
                pm = p2 - d
        else:
            d = (p3 - p2) * tau
            if abs(p2 - pm) < d:
                pm = p2 + d
            elif abs(p3 - pm) < d:
                pm = p3 - d
        ratio = abs(p1 - p2) / abs(p2 - p3)
        if not (0.01 < ratio < 100):
            if ratio > 1 and pm > p2:
                pm = (p1 + p2) / 2
            elif pm < p2:
                pm = (p2 + p3) / 2
        fm = f(pm)
        if pm < p2:
            if f1 > fm and fm < f2:
                p3, f3, p2, f2 = p2, f2, pm, fm
            else:
                p1, f1 = pm, fm
        else:
            if f2 > fm and fm < f3:
                p1, f1, p2, f2 = p2, f2, pm, fm
            else:
                p3, f3 = pm, fm
        evals += 1
    if evals >= max_iter:
        raise MinimiserFailure(too_many)
    return p2, f2


def adjust_state_to_target_flow(run_cells, state, q_columns, catchment_ids, wanted_flow_m3s, cids=(), start_step=0, scale_range=3.0,
                                scale_eps=1e-3, max_iter=300, n_steps=1):
    """tune_flow over the oracle's run_cells.

    run_cells(state, start_step, n_steps, cell_mask) -> dict with "avg_discharge" [T][n] (any of the *_run_cells above, curried);
    state [n][k] = s0; q_columns = the ground-storage columns state.adjust_q scales (kirchner.q; sm, uz, lz for hbv_stack).
    -> dict(q_0, q_r, diagnostics, state = the adjusted state, scale, evaluations)."""
    s0 = np.array(state, dtype=np.float64)
    cat = np.asarray(catchment_ids, dtype=np.int64)
    in_scope = np.ones(cat.size, dtype=bool) if len(cids) == 0 else np.isin(cat, np.asarray(cids, dtype=np.int64))
    if len(cids):
        for cid in cids:  # set_catchment_calculation_filter in the adjust_state_model ctor (region_model.h:715-729)
            if cid not in set(cat.tolist()):
                raise RuntimeError("set_catchment_calculation_filter: no cells have supplied cid")
    evaluations = []

    def adjusted(q_scale):
        s = s0.copy()
        for c in q_columns:
            s[in_scope, c] = s[in_scope, c] * q_scale
        return s

    def discharge(q_scale):  # model_state_tuning.h:59-73
        out = run_cells(adjusted(q_scale), start_step, n_steps, in_scope.astype(np.uint8))
        q_sum = 0.0
        for i in range(start_step, start_step + n_steps):
            q_sum += sum_catchment_feature_value(out["avg_discharge"], cat, list(cids), i)
        evaluations.append((q_scale, q_sum / float(n_steps)))
        return q_sum / float(n_steps)

    r = dict(q_0=0.0, q_r=0.0, diagnostics="", state=s0, scale=float("nan"), evaluations=evaluations)
    try:
        q_0 = discharge(1.0)
        scale = wanted_flow_m3s / q_0
        diag = ""
        try:
            if not np.isfinite(q_0):
                raise RuntimeError("the initial simulated discharge is nan")
            scale, _ = find_min_single_variable(lambda x: (discharge(x) - wanted_flow_m3s) ** 2, scale, scale / scale_range,
                                                scale * scale_range, scale * scale_eps, max_iter)
        except Exception as e:  # noqa: BLE001  (the reference catches std::exception)
            diag = f"failed to find solution within {max_iter}, exception was:{e}"
        q_r = discharge(scale)
        r.update(q_0=q_0, q_r=q_r, diagnostics=diag, state=adjusted(scale), scale=scale)
    except Exception as e:  # noqa: BLE001
        r["diagnostics"] = "Failed to tune_flow" + str(e)
    return r


# ---- cell-identified state: api/api_state.h:34-142 (plain restatement; integer work, must be exact) ---------------------------
def cell_state_id_of(geo_row):
    """:58-60 -- (catchment_id, (int)x, (int)y, (int)area); geo_row in GEO_COLS order"""
    return (int(geo_row[4]), int(geo_row[0]), int(geo_row[1]), int(geo_row[3]))  # Python int() truncates toward zero like the C cast


def extract_state(geo, states, cids=()):
    """state_io_handler::extract_state :105-114 -> list of (id, state row) in cell order"""
    out = []
    for i in range(len(geo)):
        if len(cids) == 0 or int(geo[i][4]) in cids:
            out.append((cell_state_id_of(geo[i]), np.array(states[i], dtype=np.float64)))
    return out


def apply_state(geo, states, id_states, cids=()):
    """state_io_handler::apply_state :119-140 -> (new states, missing positions)"""
    st = np.array(states, dtype=np.float64)
    cmap = {}
    for i in range(len(geo)):
        if len(cids) == 0 or int(geo[i][4]) in cids:
            cmap[cell_state_id_of(geo[i])] = i
    missing = []
    for k, (sid, row) in enumerate(id_states):
        if len(cids) == 0 or sid[0] in cids:
            if sid in cmap:
                st[cmap[sid]] = row
            else:
                missing.append(k)
    return st, missing
