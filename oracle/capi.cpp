// ============================================================================
// oracle/capi.cpp -- CPU ORACLE (test infrastructure, NOT product code)
// Flat C entry points over the oracle headers so that tests/, smoke() and
// bench.py's cpu_baseline leg can drive it through ctypes.
// ============================================================================
#include <chrono>
#include <cstring>

#include "sho_detmath.hpp"
#include "sho_core.hpp"
#include "sho_hbv.hpp"
#include "sho_skaugen.hpp"
#include "sho_hps.hpp"
#include "sho_pt_gs_k.hpp"
#include "sho_region.hpp"
#include "sho_ts.hpp"

using namespace sho;

static thread_local std::string g_err;
static int fail(const std::exception& e) { g_err = e.what(); return 1; }

#define SHO_TRY try {
#define SHO_END } catch (const std::exception& e) { return fail(e); } return 0;

static std::vector<geo_cell> make_cells(int64_t n, const double* geo /*[n][12]*/) {
    // columns: x y z area catchment_id radiation_slope_factor glacier lake reservoir forest routing_id routing_distance
    std::vector<geo_cell> c(n);
    for (int64_t i = 0; i < n; ++i) {
        const double* g = geo + i * 12;
        c[i].x = g[0]; c[i].y = g[1]; c[i].z = g[2]; c[i].area = g[3]; c[i].catchment_id = int64_t(g[4]);
        c[i].radiation_slope_factor = g[5]; c[i].glacier = g[6]; c[i].lake = g[7]; c[i].reservoir = g[8]; c[i].forest = g[9];
        c[i].routing_id = int64_t(g[10]); c[i].routing_distance = g[11];
    }
    return c;
}
static std::vector<geo_point> make_points(int64_t n, const double* xyz) {
    std::vector<geo_point> p(n);
    for (int64_t i = 0; i < n; ++i) p[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    return p;
}

extern "C" {

const char* sho_last_error() { return g_err.c_str(); }
int sho_hardware_concurrency() { return int(std::thread::hardware_concurrency()); }
void sho_set_gamma_policy(int mode) { special::g_gamma_policy = mode; }  // tools/gamma_policy_sensitivity.py; 0 = full double (the oracle proper)
#ifdef SHO_COUNT
void sho_cost_buffer(unsigned char* buf) { dm::g_cost_cursor = buf; }
void sho_counters(long long* out, int reset) {  // tools/cost_model.py
    for (int i = 0; i < dm::C_N; ++i) { out[i] = dm::g_cnt.v[i]; if (reset) dm::g_cnt.v[i] = 0; }
}
#endif

// ---- calendar -------------------------------------------------------------
int64_t sho_day_of_year(int64_t t_us) { return int64_t(calendar::day_of_year(t_us)); }
int64_t sho_trim_year(int64_t t_us) { return calendar::trim_year(t_us); }
int64_t sho_calendar_time(int y, int m, int d) { return calendar::time(y, m, d); }

// ---- deterministic elementary functions (sho_detmath.hpp), vectorised for the bit-exactness tests -------------
void sho_dm_eval(int fn, int64_t n, const double* a, const double* b, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = dm::exp(a[i]); break;
            case 1: out[i] = dm::log(a[i]); break;
            case 2: out[i] = dm::pow(a[i], b[i]); break;
            case 3: out[i] = dm::lgamma(a[i]); break;
            case 4: out[i] = special::gamma_p(a[i], b[i]); break;
            default: out[i] = nan_v;
        }
    }
}

// ---- special functions ----------------------------------------------------
double sho_gamma_p(double a, double x) { return special::gamma_p(a, x); }
double sho_gamma_quantile(double alpha, double p) { return routing::gamma_quantile(alpha, p); }

// ---- method units (known-answer tests) ------------------------------------
double sho_pt_potential_evapotranspiration(double albedo, double alpha, double t, double rad, double rh) {
    return priestley_taylor::calculator(albedo, alpha).potential_evapotranspiration(t, rad, rh);
}
double sho_ae_calculate_step(double water_level, double pot, double scale, double snow_fraction) {
    return actual_evapotranspiration::calculate_step(water_level, pot, scale, snow_fraction, HOUR);
}
double sho_glacier_melt_step(double dtf, double t, double sca_m2, double glacier_m2) { return glacier_melt::step(dtf, t, sca_m2, glacier_m2); }

int sho_kirchner_step(double c1, double c2, double c3, double abs_err, double rel_err, int64_t dt_us, double* q, double* q_avg, double p,
                      double e, int* n_accepted, int* n_rejected) {
    SHO_TRY
    kirchner::parameter kp; kp.c1 = c1; kp.c2 = c2; kp.c3 = c3;
    kirchner::calculator k(abs_err, rel_err, kp);
    kirchner::step_stats st;
    k.step(0, dt_us, *q, *q_avg, p, e, &st);
    if (n_accepted) *n_accepted = st.accepted;
    if (n_rejected) *n_rejected = st.rejected;
    SHO_END
}

static gamma_snow::parameter gs_param_from(const double* p /*18*/) {
    gamma_snow::parameter g;
    g.winter_end_day_of_year = size_t(p[0]); g.initial_bare_ground_fraction = p[1]; g.snow_cv = p[2]; g.tx = p[3]; g.wind_scale = p[4];
    g.wind_const = p[5]; g.max_water = p[6]; g.surface_magnitude = p[7]; g.max_albedo = p[8]; g.min_albedo = p[9];
    g.fast_albedo_decay_rate = p[10]; g.slow_albedo_decay_rate = p[11]; g.snowfall_reset_depth = p[12]; g.glacier_albedo = p[13];
    g.calculate_iso_pot_energy = p[14] != 0.0; g.snow_cv_forest_factor = p[15]; g.snow_cv_altitude_factor = p[16]; g.n_winter_days = size_t(p[17]);
    return g;
}
void sho_gs_calc_snow_state(double shape, double scale, double y0, double lambda, double lwd, double max_water_frac, double temp_swe, double* swe,
                            double* sca) {
    gamma_snow::calculator().calc_snow_state(shape, scale, y0, lambda, lwd, max_water_frac, temp_swe, *swe, *sca);
}
double sho_gs_corr_lwc(double z1, double a1, double b1, double z2, double a2, double b2, int* n_eval) {
    return gamma_snow::calculator().corr_lwc(z1, a1, b1, z2, a2, b2, n_eval);
}
double sho_gs_calc_q(double a, double b, double z) { return gamma_snow::calculator().calc_q(a, b, z); }
void sho_gs_reset_snow_pack(const double* gsp, double storage, double* out6 /*sca lwc alpha sdc_melt_mean acc_melt temp_swe*/) {
    auto p = gs_param_from(gsp);
    gamma_snow::calculator().reset_snow_pack(out6[0], out6[1], out6[2], out6[3], out6[4], out6[5], storage, p);
}
// one gamma_snow step; state8 in/out (albedo lwc surface_heat alpha sdc_melt_mean acc_melt iso_pot_energy temp_swe); resp3 out (sca storage outflow)
int sho_gs_step(const double* gsp, double* state8, double* resp3, int64_t t_us, int64_t dt_us, double T, double rad, double prec_mm_h,
                double wind_speed, double rel_hum, double forest_fraction, double altitude) {
    SHO_TRY
    auto p = gs_param_from(gsp);
    gamma_snow::state s; std::memcpy((void*)&s, state8, sizeof(double) * 8);
    gamma_snow::response r;
    gamma_snow::calculator().step(s, r, t_us, dt_us, p, T, rad, prec_mm_h, wind_speed, rel_hum, forest_fraction, altitude);
    std::memcpy(state8, &s, sizeof(double) * 8);
    resp3[0] = r.sca; resp3[1] = r.storage; resp3[2] = r.outflow;
    SHO_END
}

// ---- pt_gs_k region run ---------------------------------------------------
// geo [n_cells][12]; params [n_sets][31]; pset_of_cell [n_cells] (index into params);
// forcing: 5 arrays (temperature, precipitation, radiation, wind_speed, rel_hum), element (step i, cell c) at
//   f[i*f_tstride + c*f_cstride]; state [n_cells][9] in/out;
// resp: 8 nullable series, state series: 9 nullable, element (i,c) at p[i*o_tstride + c*o_cstride] (state series have T+1 steps);
// cell_mask nullable [n_cells] (catchment calculation filter expanded per cell)
int sho_ptgsk_run_cells(int64_t n_cells, const double* geo, int64_t n_sets, const double* params, const int32_t* pset_of_cell, int64_t t0_us,
                        int64_t dt_us, int64_t n_axis, int start_step, int n_steps, const double* f_temp, const double* f_prec,
                        const double* f_rad, const double* f_wind, const double* f_rh, int64_t f_tstride, int64_t f_cstride, double* state,
                        double** resp, double** st_series, double* kirchner_substeps, int64_t o_tstride, int64_t o_cstride, const uint8_t* cell_mask, int ncore) {
    SHO_TRY
    auto cells = make_cells(n_cells, geo);
    std::vector<pt_gs_k::parameter> ps(n_sets);
    for (int64_t k = 0; k < n_sets; ++k) ps[k].set(params + k * 31);
    fixed_dt ta{t0_us, dt_us, size_t(n_axis)};
    pt_gs_k::collectors col;
    for (int r = 0; r < pt_gs_k::N_RESPONSE; ++r) col.resp[r] = resp ? resp[r] : nullptr;
    for (int r = 0; r < pt_gs_k::N_STATE_SERIES; ++r) col.st[r] = st_series ? st_series[r] : nullptr;
    col.kirchner_substeps = kirchner_substeps;
    col.cell_stride = o_cstride; col.time_stride = o_tstride;
    parallel_run(size_t(n_cells), ncore, [&](size_t ci) {
        if (cell_mask && !cell_mask[ci]) return;
        pt_gs_k::state s;
        std::memcpy((void*)&s, state + ci * 9, sizeof(double) * 9);
        pt_gs_k::cell_forcing f{f_temp + ci * f_cstride, f_prec + ci * f_cstride, f_wind + ci * f_cstride, f_rh + ci * f_cstride,
                                f_rad + ci * f_cstride, f_tstride};
        pt_gs_k::run_pt_gs_k(cells[ci], ps[pset_of_cell ? pset_of_cell[ci] : 0], ta, start_step, n_steps, f, s, col, ci);
        std::memcpy(state + ci * 9, &s, sizeof(double) * 9);
    });
    SHO_END
}

// ---- pt_hs_k / hbv_stack region runs -------------------------------------
// state layouts: pt_hs_k [n_cells][3 + 2*n_bins] = (swe, sca, sp[n_bins], sw[n_bins], kirchner.q)
//                hbv_stack [n_cells][5 + 2*n_bins] = (swe, sca, sp[], sw[], soil.sm, tank.uz, tank.lz)
int sho_pthsk_run_cells(int64_t n_cells, const double* geo, int64_t n_sets, const double* params, int n_param, const int32_t* pset_of_cell,
                        int64_t t0_us, int64_t dt_us, int64_t n_axis, int start_step, int n_steps, const double* f_temp, const double* f_prec,
                        const double* f_rad, const double* f_wind, const double* f_rh, int64_t f_tstride, int64_t f_cstride, double* state,
                        int n_state, double** resp, int64_t o_tstride, int64_t o_cstride, int ncore) {
    SHO_TRY
    auto cells = make_cells(n_cells, geo);
    std::vector<pt_hs_k::parameter> ps(n_sets);
    for (int64_t k = 0; k < n_sets; ++k) ps[k].set(params + k * n_param, size_t(n_param));
    fixed_dt ta{t0_us, dt_us, size_t(n_axis)};
    parallel_run(size_t(n_cells), ncore, [&](size_t ci) {
        const auto& p = ps[pset_of_cell ? pset_of_cell[ci] : 0];
        pt_hs_k::state s;
        s.unpack(state + ci * n_state, size_t(n_state));
        pt_hs_k::cell_forcing f{f_temp + ci * f_cstride, f_prec + ci * f_cstride, f_wind + ci * f_cstride, f_rh + ci * f_cstride,
                                f_rad + ci * f_cstride, f_tstride};
        pt_hs_k::run(cells[ci], p, ta, start_step, n_steps, f, s, resp, o_tstride, o_cstride, ci);
        s.pack(state + ci * n_state, size_t(n_state));
    });
    SHO_END
}
int sho_hbv_stack_run_cells(int64_t n_cells, const double* geo, int64_t n_sets, const double* params, int n_param, const int32_t* pset_of_cell,
                            int64_t t0_us, int64_t dt_us, int64_t n_axis, int start_step, int n_steps, const double* f_temp, const double* f_prec,
                            const double* f_rad, const double* f_wind, const double* f_rh, int64_t f_tstride, int64_t f_cstride, double* state,
                            int n_state, double** resp, int64_t o_tstride, int64_t o_cstride, int ncore) {
    SHO_TRY
    auto cells = make_cells(n_cells, geo);
    std::vector<hbv_stack::parameter> ps(n_sets);
    for (int64_t k = 0; k < n_sets; ++k) ps[k].set(params + k * n_param, size_t(n_param));
    fixed_dt ta{t0_us, dt_us, size_t(n_axis)};
    parallel_run(size_t(n_cells), ncore, [&](size_t ci) {
        const auto& p = ps[pset_of_cell ? pset_of_cell[ci] : 0];
        hbv_stack::state s;
        s.unpack(state + ci * n_state, size_t(n_state));
        hbv_stack::cell_forcing f{f_temp + ci * f_cstride, f_prec + ci * f_cstride, f_wind + ci * f_cstride, f_rh + ci * f_cstride,
                                  f_rad + ci * f_cstride, f_tstride};
        hbv_stack::run_hbv_stack(cells[ci], p, ta, start_step, n_steps, f, s, resp, o_tstride, o_cstride, ci);
        s.pack(state + ci * n_state, size_t(n_state));
    });
    SHO_END
}
// pt_ss_k region run: state [n_cells][8] = snow.nu, alpha, sca, swe, free_water, residual, num_units, kirchner.q; sts = SS_N nullable
// state series ([n_axis + 1][cells]), resp = HR_N nullable response series
int sho_ptssk_run_cells(int64_t n_cells, const double* geo, int64_t n_sets, const double* params, int n_param, const int32_t* pset_of_cell,
                        int64_t t0_us, int64_t dt_us, int64_t n_axis, int start_step, int n_steps, const double* f_temp, const double* f_prec,
                        const double* f_rad, const double* f_wind, const double* f_rh, int64_t f_tstride, int64_t f_cstride, double* state,
                        double** resp, double** sts, int64_t o_tstride, int64_t o_cstride, int ncore) {
    SHO_TRY
    auto cells = make_cells(n_cells, geo);
    std::vector<pt_ss_k::parameter> ps(n_sets);
    for (int64_t k = 0; k < n_sets; ++k) ps[k].set(params + k * n_param, size_t(n_param));
    fixed_dt ta{t0_us, dt_us, size_t(n_axis)};
    parallel_run(size_t(n_cells), ncore, [&](size_t ci) {
        const auto& p = ps[pset_of_cell ? pset_of_cell[ci] : 0];
        pt_ss_k::state s;
        s.unpack(state + ci * 8);
        pt_ss_k::cell_forcing f{f_temp + ci * f_cstride, f_prec + ci * f_cstride, f_wind + ci * f_cstride, f_rh + ci * f_cstride,
                                f_rad + ci * f_cstride, f_tstride};
        pt_ss_k::run(cells[ci], p, ta, start_step, n_steps, f, s, resp, sts, o_tstride, o_cstride, ci);
        s.pack(state + ci * 8);
    });
    SHO_END
}
// skaugen::calculator::step (core/skaugen.h:150-339): par8 = alpha_0 d_range unit_size max_water_fraction tx cx ts cfr;
// state7 = nu alpha sca swe free_water residual num_units (in place); out3 = outflow sca swe
int sho_skaugen_step(const double* par8, double* state7, int64_t dt_us, double temp, double prec, double* out3) {
    SHO_TRY
    skaugen::parameter p{par8[0], par8[1], par8[2], par8[3], par8[4], par8[5], par8[6], par8[7]};
    skaugen::state s{state7[0], state7[1], state7[2], state7[3], state7[4], state7[5], (unsigned long)state7[6]};
    skaugen::response r;
    skaugen::step(dt_us, p, temp, prec, s, r);
    state7[0] = s.nu; state7[1] = s.alpha; state7[2] = s.sca; state7[3] = s.swe; state7[4] = s.free_water; state7[5] = s.residual; state7[6] = double(s.num_units);
    out3[0] = r.outflow; out3[1] = r.sca; out3[2] = r.swe;
    SHO_END
}
double sho_skaugen_sca_rel_red(double u, double n, double unit_size, double nu_a, double alpha) {
    try { return skaugen::statistics::sca_rel_red((unsigned long)u, (unsigned long)n, unit_size, nu_a, alpha); } catch (...) { return std::nan(""); }
}
// pt_hps_k region run: state [n_cells][4*nb + 4] = sp[nb], sw[nb], albedo[nb], iso_pot_energy[nb], surface_heat, swe, sca, kirchner.q
int sho_pthpsk_run_cells(int64_t n_cells, const double* geo, int64_t n_sets, const double* params, int n_param, const int32_t* pset_of_cell,
                         int64_t t0_us, int64_t dt_us, int64_t n_axis, int start_step, int n_steps, const double* f_temp, const double* f_prec,
                         const double* f_rad, const double* f_wind, const double* f_rh, int64_t f_tstride, int64_t f_cstride, double* state,
                         int n_state, double** resp, int64_t o_tstride, int64_t o_cstride, int ncore) {
    SHO_TRY
    auto cells = make_cells(n_cells, geo);
    std::vector<pt_hps_k::parameter> ps(n_sets);
    for (int64_t k = 0; k < n_sets; ++k) ps[k].set(params + k * n_param, size_t(n_param));
    fixed_dt ta{t0_us, dt_us, size_t(n_axis)};
    parallel_run(size_t(n_cells), ncore, [&](size_t ci) {
        const auto& p = ps[pset_of_cell ? pset_of_cell[ci] : 0];
        pt_hps_k::state s;
        s.unpack(state + ci * n_state, size_t(n_state));
        pt_hps_k::cell_forcing f{f_temp + ci * f_cstride, f_prec + ci * f_cstride, f_wind + ci * f_cstride, f_rh + ci * f_cstride,
                                 f_rad + ci * f_cstride, f_tstride};
        pt_hps_k::run(cells[ci], p, ta, start_step, n_steps, f, s, resp, o_tstride, o_cstride, ci);
        s.pack(state + ci * n_state, size_t(n_state));
    });
    SHO_END
}
// hbv_physical_snow::calculator::step (core/hbv_physical_snow.h:266-529) with the explicit parameter(s, intervals) constructor (normalised);
// par11 = tx lw cfr wind_scale wind_const surface_magnitude max_albedo min_albedo fast_albedo_decay_rate slow_albedo_decay_rate snowfall_reset_depth;
// state = sp[n], sw[n], albedo[n], iso_pot_energy[n], surface_heat, swe, sca in place (distribute = 1: state.distribute(p) first);
// forcing5 = T rad prec wind_speed rel_hum; out3 = outflow sca storage
int sho_hps_step(const double* s_q, const double* intervals, int n, const double* par11, int iso, int distribute, double* state, int64_t dt_us,
                 const double* forcing5, double* out3) {
    SHO_TRY
    hbv_physical_snow::parameter p(std::vector<double>(s_q, s_q + n), std::vector<double>(intervals, intervals + n));
    p.tx = par11[0]; p.lw = par11[1]; p.cfr = par11[2]; p.wind_scale = par11[3]; p.wind_const = par11[4]; p.surface_magnitude = par11[5];
    p.max_albedo = par11[6]; p.min_albedo = par11[7]; p.fast_albedo_decay_rate = par11[8]; p.slow_albedo_decay_rate = par11[9];
    p.snowfall_reset_depth = par11[10]; p.calculate_iso_pot_energy = iso != 0;
    hbv_physical_snow::state s;
    s.sp.assign(state, state + n); s.sw.assign(state + n, state + 2 * n); s.albedo.assign(state + 2 * n, state + 3 * n);
    s.iso_pot_energy.assign(state + 3 * n, state + 4 * n);
    s.surface_heat = state[4 * n]; s.swe = state[4 * n + 1]; s.sca = state[4 * n + 2];
    if (distribute) s.distribute(p);
    hbv_physical_snow::response r;
    hbv_physical_snow::calculator c(p);
    c.step(s, r, 0, dt_us, forcing5[0], forcing5[1], forcing5[2], forcing5[3], forcing5[4]);
    for (int i = 0; i < n; ++i) { state[i] = s.sp[i]; state[n + i] = s.sw[i]; state[2 * n + i] = s.albedo[i]; state[3 * n + i] = s.iso_pot_energy[i]; }
    state[4 * n] = s.surface_heat; state[4 * n + 1] = s.swe; state[4 * n + 2] = s.sca;
    out3[0] = r.outflow; out3[1] = r.sca; out3[2] = r.storage;
    SHO_END
}
// hbv unit steps for known-answer tests
int sho_hbv_snow_step(const double* s_q /*n*/, const double* intervals /*n*/, int n, const double* par5 /*tx cx ts lw cfr*/, double* sp, double* sw,
                      double* swe, double* sca, int64_t dt_us, double prec, double temp, double* outflow) {
    SHO_TRY
    hbv_snow::parameter p;
    p.s.assign(s_q, s_q + n); p.intervals.assign(intervals, intervals + n);
    p.tx = par5[0]; p.cx = par5[1]; p.ts = par5[2]; p.lw = par5[3]; p.cfr = par5[4];
    hbv_snow::state s; s.sp.assign(sp, sp + n); s.sw.assign(sw, sw + n); s.swe = *swe; s.sca = *sca;
    hbv_snow::response r;
    hbv_snow::calculator c(p);
    c.step(s, r, 0, dt_us, prec, temp);
    std::copy(s.sp.begin(), s.sp.end(), sp); std::copy(s.sw.begin(), s.sw.end(), sw);
    *swe = s.swe; *sca = s.sca; *outflow = r.outflow;
    SHO_END
}
// hbv_snow::state::distribute(p) (core/hbv_snow.h:113-116 over hbv_snow_common.h:45-67)
int sho_hbv_snow_distribute(const double* s_q /*n*/, const double* intervals /*n*/, int n, double lw, double* sp, double* sw, double* swe, double* sca) {
    SHO_TRY
    hbv_snow::parameter p;
    p.s.assign(s_q, s_q + n); p.intervals.assign(intervals, intervals + n); p.lw = lw;
    hbv_snow::state s; s.swe = *swe; s.sca = *sca;
    s.distribute(p);
    std::copy(s.sp.begin(), s.sp.end(), sp); std::copy(s.sw.begin(), s.sw.end(), sw);
    *swe = s.swe; *sca = s.sca;
    SHO_END
}
double sho_hbv_soil_step(double fc, double beta, double* sm, double insoil, double act_evap) {
    hbv_soil::parameter p; p.fc = fc; p.beta = beta;
    hbv_soil::state s; s.sm = *sm; hbv_soil::response r;
    hbv_soil::calculator(p).step(s, r, 0, HOUR, insoil, act_evap);
    *sm = s.sm;
    return r.outflow;
}
double sho_hbv_tank_step(const double* par5 /*uz1 kuz2 kuz1 perc klz*/, double* uz, double* lz, double soil_outflow) {
    hbv_tank::parameter p; p.uz1 = par5[0]; p.kuz2 = par5[1]; p.kuz1 = par5[2]; p.perc = par5[3]; p.klz = par5[4];
    hbv_tank::state s; s.uz = *uz; s.lz = *lz; hbv_tank::response r;
    hbv_tank::calculator(p).step(s, r, 0, HOUR, soil_outflow);
    *uz = s.uz; *lz = s.lz;
    return r.outflow;
}
double sho_hbv_ae_step(double soil_moisture, double pot_evap, double lp, double snow_fraction) {
    return hbv_actual_evapotranspiration::calculate_step(soil_moisture, pot_evap, lp, snow_fraction, HOUR);
}

// ---- interpolation ---------------------------------------------------------
// idw_par: max_members max_distance distance_measure_factor zscale default_temp_gradient gradient_by_equation scale_factor
int sho_idw_run(int kind, int64_t n_src, const double* src_xyz, const double* src_values /*[T][n_src]*/, int64_t n_steps, int64_t n_dst,
                const double* dst_xyz, const double* dst_slope, const double* idw_par, double* out, int64_t o_tstride, int64_t o_cstride, int ncore) {
    SHO_TRY
    idw::parameter p;
    p.max_members = size_t(idw_par[0]); p.max_distance = idw_par[1]; p.distance_measure_factor = idw_par[2]; p.zscale = idw_par[3];
    p.default_temp_gradient = idw_par[4]; p.gradient_by_equation = idw_par[5] != 0.0; p.scale_factor = idw_par[6];
    auto src = make_points(n_src, src_xyz), dst = make_points(n_dst, dst_xyz);
    std::vector<double> slope(dst_slope ? dst_slope : nullptr, dst_slope ? dst_slope + n_dst : nullptr);
    if (!dst_slope) slope.assign(n_dst, 1.0);
    // same partition as inverse_distance.h:524-541: 1 + n_cells/ncore cells per thread
    if (ncore < 2) idw::run_interpolation(idw::model_kind(kind), src, src_values, size_t(n_steps), dst, slope, p, out, o_tstride, o_cstride);
    else {
        size_t per = 1 + size_t(n_dst) / size_t(ncore);
        std::vector<std::thread> th;
        for (size_t b = 0; b < size_t(n_dst); b += per)
            th.emplace_back([&, b]() { idw::run_interpolation(idw::model_kind(kind), src, src_values, size_t(n_steps), dst, slope, p, out, o_tstride, o_cstride, b, b + per); });
        for (auto& t : th) t.join();
    }
    SHO_END
}
// neighbour lists only: out_idx/out_w [n_dst][max_members], out_n [n_dst]
int sho_idw_neighbours(int64_t n_src, const double* src_xyz, int64_t n_dst, const double* dst_xyz, const double* idw_par, int32_t* out_idx,
                       double* out_w, int32_t* out_n) {
    SHO_TRY
    idw::parameter p;
    p.max_members = size_t(idw_par[0]); p.max_distance = idw_par[1]; p.distance_measure_factor = idw_par[2]; p.zscale = idw_par[3];
    auto nb = idw::build_neighbours(make_points(n_src, src_xyz), make_points(n_dst, dst_xyz), p);
    for (size_t j = 0; j < nb.size(); ++j) {
        out_n[j] = int32_t(nb[j].size());
        for (size_t k = 0; k < nb[j].size(); ++k) { out_idx[j * p.max_members + k] = nb[j][k].source; out_w[j * p.max_members + k] = nb[j][k].weight; }
    }
    SHO_END
}
// btk_par: gradient_sd sill nug range zscale fixed_prior_gradient(NaN = the day-of-year sinusoid)
int sho_btk_run(int64_t n_src, const double* src_xyz, const double* src_values, int64_t t0_us, int64_t dt_us, int64_t n_steps, int64_t n_dst,
                const double* dst_xyz, const double* btk_par, double* out, int64_t o_tstride, int64_t o_cstride) {
    SHO_TRY
    btk::parameter p;
    p.gradient_sd = btk_par[0]; p.sill_value = btk_par[1]; p.nug_value = btk_par[2]; p.range_value = btk_par[3]; p.zscale_value = btk_par[4];
    p.fixed_gradient = btk_par[5];
    fixed_dt ta{t0_us, dt_us, size_t(n_steps)};
    btk::btk_interpolation(make_points(n_src, src_xyz), src_values, ta, make_points(n_dst, dst_xyz), p, out, o_tstride, o_cstride);
    SHO_END
}
int sho_btk_covariance(int64_t n_src, const double* src_xyz, int64_t n_dst, const double* dst_xyz, const double* btk_par, double* K, double* k) {
    SHO_TRY
    btk::parameter p;
    p.gradient_sd = btk_par[0]; p.sill_value = btk_par[1]; p.nug_value = btk_par[2]; p.range_value = btk_par[3]; p.zscale_value = btk_par[4];
    la::mat Km, km;
    btk::build_covariance_matrices(make_points(n_src, src_xyz), make_points(n_dst, dst_xyz), p, Km, km);
    std::copy(Km.a.begin(), Km.a.end(), K);
    std::copy(km.a.begin(), km.a.end(), k);
    SHO_END
}
double sho_btk_prior_gradient(int64_t t_us, int64_t dt_us) { return btk::parameter().temperature_gradient(t_us, dt_us); }
// average_value with the caller's hint (replays test/time_series_test.cpp); returns the value, updates *ix
double sho_average_value(const int64_t* t, const double* v, int64_t n, int64_t p_start, int64_t p_end, int64_t* ix, int linear) {
    ts::point_source src{t, v, size_t(n), 1, n ? t[n - 1] : 0};
    size_t last = *ix < 0 ? ts::npos : size_t(*ix);
    const double r = ts::average_value(src, p_start, p_end, last, linear != 0);
    *ix = last == ts::npos ? -1 : int64_t(last);
    return r;
}
// average_accessor of n_src sources sharing one point axis: values [n_points][n_src] -> out [ta_n][n_src]
void sho_average_accessor(const int64_t* t, const double* values, int64_t n_points, int64_t n_src, int64_t t_end, int linear, int64_t ta_t0,
                          int64_t ta_dt, int64_t ta_n, double* out) {
    for (int64_t s = 0; s < n_src; ++s) {
        ts::point_source src{t, values + s, size_t(n_points), size_t(n_src), t_end};
        ts::average_accessor(src, linear != 0, ta_t0, ta_dt, size_t(ta_n), out + s, size_t(n_src));
    }
}
void sho_average_accessor_same_axis(double* v, int64_t n, int64_t dt_us) { for (int64_t i = 0; i < n; ++i) v[i] = average_accessor_same_axis(v[i], dt_us); }

// ---- routing ----------------------------------------------------------------
int sho_make_uhg(int n_steps, double alpha, double beta, double* out, int* n_out) {
    SHO_TRY
    auto w = routing::make_uhg_from_gamma(n_steps, alpha, beta);
    *n_out = int(w.size());
    std::copy(w.begin(), w.end(), out);
    SHO_END
}
int sho_uhg_steps(double distance, double velocity, int64_t dt_us) { return routing::uhg_steps(distance, velocity, dt_us); }
// River network evaluation (routing.h:326-383): cell_discharge element (step i, cell c) at q[i*tstride + c*cstride];
// cells route to cell_rid[c] over cell_distance[c] with cell uhg (velocity/alpha/beta per cell);
// rivers [n_riv][6] = id downstream_id distance velocity alpha beta.  out [3][T] = local_inflow, upstream_inflow, output for river `rid`.
static void river_eval(const routing::network& net, int64_t rid, int64_t n_cells, const double* q, int64_t tstride, int64_t cstride,
                       const int64_t* cell_rid, const double* cell_distance, const double* cell_uhg_par /*[n][3]*/, size_t T, int64_t dt_us,
                       std::vector<double>& local, std::vector<double>& upstream, std::vector<double>& output) {
    local.assign(T, 0.0); upstream.assign(T, 0.0); output.assign(T, 0.0);
    for (int64_t c = 0; c < n_cells; ++c)
        if (cell_rid[c] == rid) {
            auto w = routing::make_uhg_from_gamma(routing::uhg_steps(cell_distance[c], cell_uhg_par[3 * c], dt_us), cell_uhg_par[3 * c + 1], cell_uhg_par[3 * c + 2]);
            for (size_t t = 0; t < T; ++t) local[t] += routing::convolve_value(q + c * cstride, tstride, w, t);
        }
    for (auto up : net.upstreams_by_id(rid)) {
        std::vector<double> l, u, o;
        river_eval(net, up, n_cells, q, tstride, cstride, cell_rid, cell_distance, cell_uhg_par, T, dt_us, l, u, o);
        for (size_t t = 0; t < T; ++t) upstream[t] += o[t];
    }
    const auto& r = net.rid_map.at(rid);
    auto w = routing::make_uhg_from_gamma(routing::uhg_steps(r.distance, r.velocity, dt_us), r.alpha, r.beta);
    std::vector<double> sum(T);
    for (size_t t = 0; t < T; ++t) sum[t] = local[t] + upstream[t];
    for (size_t t = 0; t < T; ++t) output[t] = routing::convolve_value(sum.data(), 1, w, t);
}
int sho_river_flows(int64_t n_riv, const double* rivers, int64_t rid, int64_t n_cells, const double* q, int64_t tstride, int64_t cstride,
                    const int64_t* cell_rid, const double* cell_distance, const double* cell_uhg_par, int64_t T, int64_t dt_us, double* out) {
    SHO_TRY
    routing::network net;
    for (int64_t i = 0; i < n_riv; ++i) {
        routing::river r;
        r.id = int64_t(rivers[6 * i]); r.downstream_id = int64_t(rivers[6 * i + 1]); r.distance = rivers[6 * i + 2];
        r.velocity = rivers[6 * i + 3]; r.alpha = rivers[6 * i + 4]; r.beta = rivers[6 * i + 5];
        net.rid_map[r.id] = r;
    }
    std::vector<double> l, u, o;
    river_eval(net, rid, n_cells, q, tstride, cstride, cell_rid, cell_distance, cell_uhg_par, size_t(T), dt_us, l, u, o);
    std::copy(l.begin(), l.end(), out); std::copy(u.begin(), u.end(), out + T); std::copy(o.begin(), o.end(), out + 2 * T);
    SHO_END
}

// average_accessor of a stair-case series on an aligned coarser axis (core/time_series.h:202-310,2033-2072): target period i covers
// source steps [first + i*k, first + (i+1)*k); area accumulates v*dt over the finite steps, tsum their length
void sho_average_to_axis(const double* v, int64_t n_src, int64_t dt_us, int64_t first, int64_t k, int64_t n_out, double* out) {
    for (int64_t i = 0; i < n_out; ++i) {
        double area = 0.0;
        utctimespan tsum = 0;
        for (int64_t j = 0; j < k; ++j) {
            const int64_t s = first + i * k + j;
            if (s >= n_src) break;
            if (std::isfinite(v[s])) { area += v[s] * to_seconds(dt_us); tsum += dt_us; }
        }
        out[i] = tsum > 0 ? area / to_seconds(tsum) : nan_v;
    }
}

// ---- goal functions -----------------------------------------------------------
double sho_nash_sutcliffe(const double* o, const double* m, int64_t n) { return goal::nash_sutcliffe(o, m, size_t(n)); }
double sho_rmse(const double* o, const double* m, int64_t n) { return goal::rmse(o, m, size_t(n)); }
double sho_kling_gupta(const double* o, const double* m, int64_t n, double s_r, double s_a, double s_b) { return goal::kling_gupta(o, m, size_t(n), s_r, s_a, s_b); }
double sho_abs_diff_sum(const double* o, const double* m, int64_t n) { return goal::abs_diff_sum(o, m, size_t(n)); }

// ---- catchment index (region_model.h:233-249) ----------------------------------
int64_t sho_catchment_index(int64_t n_cells, const int64_t* cid, int64_t* cix_of_cell, int64_t* cix_to_cid) {
    std::vector<geo_cell> c(n_cells);
    for (int64_t i = 0; i < n_cells; ++i) c[i].catchment_id = cid[i];
    auto m = update_ix_to_id_mapping(c);
    for (int64_t i = 0; i < n_cells; ++i) cix_of_cell[i] = int64_t(c[i].catchment_ix);
    std::copy(m.begin(), m.end(), cix_to_cid);
    return int64_t(m.size());
}

}  // extern "C"
