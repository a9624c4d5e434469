// ============================================================================
// oracle/sho_skaugen.hpp -- CPU ORACLE (test infrastructure, NOT product code)
//
// The Skaugen snow routine and the pt_ss_k method stack, restated from
//   core/skaugen.h:24-386        statistics::sca_rel_red, calculator::step, compute_shape_vars
//   core/pt_ss_k.h:28-295        parameter / state / response, run()
//   core/pt_ss_k_cell_model.h    collectors (which series, scale_snow)
// Third-party arithmetic absent from the reference tree (boost 1.68, build_support/build_dependencies.sh:5-8), restated from
// its published algorithms:
//   boost::math::gamma_distribution pdf / cdf / mean   (policy digits10<16>, i.e. full double: skaugen.h:41-42)
//       pdf(x) = gamma_p_derivative(k, x / theta) / theta = (x/theta)^(k-1) e^(-x/theta) / Gamma(k) / theta;  cdf = gamma_p(k, x / theta)
//   boost::math::tools::brent_find_minima(f, a, b, bits = 2)      (special::brent_find_minima of sho_core.hpp)
//   boost::math::tools::bisect(f, lo, hi, eps_tolerance<double>(10), max_iter = 100)
// PARITY STATUS: pinned by the reference's own asserts for the routine (test/skaugen_test.cpp:9-281: accumulation, melt and mass
// balance, liquid water, the melt-down regression) and for the stack (test/pt_ss_k_test.cpp); "parity unpinned" at the boost
// boundary -- the 2-bit Brent and 10-bit bisection of sca_rel_red are discrete searches whose iteration counts depend on last-bit
// values of the gamma pdf, exactly like gamma_snow's corr_lwc.
// ============================================================================
#pragma once
#include "sho_hbv.hpp"

namespace sho {
namespace skaugen {

// boost gamma_distribution<double>(shape, scale)
struct gamma_dist {
    double shape, scale;
    double mean() const { return shape * scale; }
    double pdf(double x) const {  // gamma_p_derivative(shape, x / scale) / scale
        const double a = shape, z = x / scale;
        if (z < 0.0) throw std::runtime_error("gamma pdf: negative argument");
        if (z == 0.0) {
            if (a > 1.0) return 0.0;
            if (a == 1.0) return 1.0 / scale;
            throw std::overflow_error("gamma pdf: overflow at zero");
        }
        return dm::exp(a * dm::log(z) - z - special::lgamma_(a)) / z / scale;
    }
    double cdf(double x) const { return special::gamma_p(shape, x / scale); }
};

// boost::math::tools::eps_tolerance<double>(bits) and bisect(f, min, max, tol, max_iter)
struct eps_tolerance {
    double eps;
    explicit eps_tolerance(unsigned bits) { eps = std::max(double(std::ldexp(1.0f, 1 - int(bits))), 4.0 * std::numeric_limits<double>::epsilon()); }
    bool operator()(double a, double b) const { return std::fabs(a - b) <= eps * std::min(std::fabs(a), std::fabs(b)); }
};
template <class F>
inline std::pair<double, double> bisect(F f, double min, double max, const eps_tolerance& tol, unsigned long max_iter) {
    double fmin = f(min), fmax = f(max);
    if (fmin == 0.0) return {min, min};
    if (fmax == 0.0) return {max, max};
    if (min >= max) throw std::runtime_error("bisect: arguments in wrong order");
    if (fmin * fmax >= 0.0) throw std::runtime_error("bisect: no change of sign, either there is no root to find, or there are multiple roots in the interval");
    unsigned long count = max_iter < 3 ? 0 : max_iter - 3;  // three function invocations so far
    auto sign = [](double v) { return v == 0.0 ? 0 : (v < 0.0 ? -1 : 1); };
    while (count && !tol(min, max)) {
        const double mid = (min + max) / 2;
        const double fmid = f(mid);
        if (mid == max || mid == min) break;
        if (fmid == 0.0) { min = max = mid; break; }
        else if (sign(fmid) * sign(fmin) < 0) { max = mid; fmax = fmid; }
        else { min = mid; fmin = fmid; }
        --count;
    }
    return {min, max};
}

struct parameter {  // skaugen.h:90-110
    double alpha_0 = 40.77, d_range = 113.0, unit_size = 0.1, max_water_fraction = 0.1, tx = 0.16, cx = 2.5, ts = 0.14, cfr = 0.01;
};
struct state {  // :113-139
    double nu = 4.077, alpha = 40.77, sca = 0.0, swe = 0.0, free_water = 0.0, residual = 0.0;
    unsigned long num_units = 0;
};
struct response { double outflow = 0.0, sca = 0.0, swe = 0.0; };

struct statistics {  // :44-88
    double alpha_0, d_range, unit_size;
    static double c(unsigned long n, double d_range) { return dm::exp(-double(n) / d_range); }
    double c(unsigned long n) const { return c(n, d_range); }
    static double sca_rel_red(unsigned long u, unsigned long n, double /*unit_size*/, double nu_a, double alpha) {
        const double nu_m = (double(u) / n) * nu_a;
        const gamma_dist g_m{nu_m, 1.0 / alpha};
        const gamma_dist g_a{nu_a, 1.0 / alpha};
        const double g_a_mean = g_a.mean();
        auto zero_func = [&](double x) { return g_m.pdf(x) - g_a.pdf(x); };
        double lower = g_m.mean();
        const double upper = special::brent_find_minima(zero_func, 0.0, g_a_mean, 2, std::numeric_limits<int>::max());
        while (g_m.pdf(lower) < g_a.pdf(lower)) lower *= 0.9;
        const auto res = bisect(zero_func, lower, upper, eps_tolerance(10), 100);
        const double x = (res.first + res.second) * 0.5;
        const double m = g_m.cdf(x);
        const double a = g_a.cdf(x);
        return a + 1.0 - m;
    }
    double sca_rel_red(unsigned long u, unsigned long n, double nu_a, double alpha) const { return sca_rel_red(u, n, unit_size, nu_a, alpha); }
};

// calculator::compute_shape_vars, :341-383
inline void compute_shape_vars(const statistics& stat, unsigned long nnn, unsigned long n, unsigned long u, double sca, double rel_red_sca, double& alpha,
                               double& nu) {
    const double alpha_0 = stat.alpha_0;
    const double nu_0 = stat.alpha_0 * stat.unit_size;
    const double dyn_var = nu / (alpha * alpha);
    const double init_var = nu_0 / (alpha_0 * alpha_0);
    double tot_var = 0.0;
    double tot_mean = 0.0;
    if (n > 0) {  // accumulation
        if (nnn == 0) {
            tot_var = n * init_var * (1 + (n - 1) * stat.c(n));
            tot_mean = n * nu_0 / alpha_0;
        } else {
            const double old_var_cov = (nnn + n) * init_var * (1 + ((nnn + n) - 1) * stat.c(nnn + n));
            const double new_var_cov = n * init_var * (1 + (n - 1) * stat.c(n));
            tot_var = old_var_cov * sca * sca + new_var_cov * (1.0 - sca) * (1.0 - sca);
            tot_mean = (sca * (nnn + n) + (1.0 - sca) * n) * stat.unit_size;
        }
    }
    if (u > 0) {  // ablation
        const double factor = (dyn_var / (nnn * init_var) + 1.0 + (nnn - 1) * stat.c(nnn)) / (2 * nnn);
        const double non_cond_mean = (nnn - u) * stat.unit_size;
        tot_mean = non_cond_mean / (1.0 - rel_red_sca);
        const unsigned long cond_u = (unsigned long)std::lrint((1.0 - rel_red_sca) * nnn - (nnn - u));
        const double auto_var = cond_u > 0 ? init_var * cond_u * (1.0 + (cond_u - 1.0) * stat.c(cond_u)) : 0.0;
        const double cross_var = cond_u > 0 ? init_var * cond_u * 2.0 * factor * cond_u : 0.0;
        tot_var = dyn_var + auto_var - cross_var;
    }
    if (std::fabs(tot_mean) < 1.0e-7) {
        nu = nu_0;
        alpha = alpha_0;
        return;
    }
    nu = tot_mean * tot_mean / tot_var;
    alpha = nu / (stat.unit_size * std::lrint(tot_mean / stat.unit_size));
}

// calculator::step, :150-339 (rad and wind speed are unused there)
inline void step(utctimespan dt, const parameter& p, double T, double prec_mm_h, state& s, response& r) {
    const double snow_tol = 1.0e-10;
    const double unit_size = p.unit_size;
    const double step_in_days = to_seconds(dt) / 86400.0;
    const double dt_hours = to_seconds(dt) / 3600.0;
    const double prec = prec_mm_h * dt_hours;
    const double corr_prec = std::max(0.0, prec + s.residual);
    s.residual = std::min(0.0, prec + s.residual);
    const double snow = T < p.tx ? corr_prec : 0.0;
    const double rain = T < p.tx ? 0.0 : corr_prec;
    if (s.sca * s.swe < unit_size && snow < snow_tol) {
        r.outflow = (rain + s.sca * (s.swe + s.free_water) + s.residual) / dt_hours;
        r.swe = 0.0;
        s.residual = 0.0;
        if (r.outflow < 0.0) {
            s.residual = r.outflow;
            r.outflow = 0.0;
        }
        s.nu = p.alpha_0 * unit_size;
        s.alpha = p.alpha_0;
        s.sca = 0.0;
        s.swe = 0.0;
        s.free_water = 0.0;
        s.num_units = 0;
        r.sca = s.sca;
        r.swe = s.swe;
        return;
    }
    const double alpha_0 = p.alpha_0;
    double swe = s.swe;
    unsigned long nnn = s.num_units;
    double sca = s.sca;
    double nu = s.nu;
    double alpha = s.alpha;
    if (nnn > 0) nu *= nnn;
    else {
        nu = alpha_0 * p.unit_size;
        alpha = alpha_0;
    }
    double total_new_snow = snow;
    double lwc = s.free_water;
    const double total_storage = swe + lwc;
    double pot_melt = p.cx * step_in_days * (T - p.ts);
    const double refreeze = std::min(std::max(0.0, -pot_melt * p.cfr), lwc);
    total_new_snow += sca * refreeze;
    lwc -= refreeze;
    pot_melt = std::max(0.0, pot_melt);
    const double new_snow_reduction = std::min(pot_melt, total_new_snow);
    pot_melt -= new_snow_reduction;
    total_new_snow -= new_snow_reduction;
    const statistics stat{alpha_0, p.d_range, unit_size};
    unsigned long n = 0;
    if (total_new_snow > unit_size) {  // 1. accumulation
        n = (unsigned long)std::lrint(total_new_snow / unit_size);
        compute_shape_vars(stat, nnn, n, 0, sca, 0.0, alpha, nu);
        nnn = (unsigned long)std::lrint(nnn * sca) + n;
        sca = 1.0;
        swe = nnn * unit_size;
    }
    if (pot_melt > unit_size) {  // 2. melting
        unsigned long u = (unsigned long)std::lrint(pot_melt / unit_size);
        if (nnn < u + 2) {
            nnn = 0;
            alpha = alpha_0;
            nu = alpha_0 * unit_size;
            swe = 0.0;
            lwc = 0.0;
            sca = 0.0;
        } else {
            const double rel_red_sca = stat.sca_rel_red(u, nnn, nu, alpha);
            const double sca_scale_factor = 1.0 - rel_red_sca;
            sca = s.sca * sca_scale_factor;
            swe = (nnn - u) / sca_scale_factor * unit_size;
            if (swe >= nnn * unit_size) {
                u = (unsigned long)(long(nnn * rel_red_sca) + 1);
                swe = (nnn - u) / sca_scale_factor * unit_size;
                if (nnn == u) sca = 0.0;
            }
            if (sca < 0.005) {
                nnn = 0;
                alpha = alpha_0;
                nu = alpha_0 * unit_size;
                swe = 0.0;
                lwc = 0.0;
                sca = 0.0;
            } else {
                compute_shape_vars(stat, nnn, n, u, sca, rel_red_sca, alpha, nu);
                nnn = (unsigned long)std::lrint(swe / unit_size);
                swe = nnn * unit_size;
            }
        }
    }
    if (s.sca * s.swe > sca * swe) lwc += std::max(0.0, s.swe - swe);  // 3. liquid water
    lwc *= std::min(1.0, s.sca / sca);
    lwc = std::min(lwc, swe * p.max_water_fraction);
    double discharge = s.sca * total_storage + snow - sca * (swe + lwc);
    if (discharge < 0.0) {
        s.residual += discharge;
        discharge = 0.0;
    }
    if (rain > swe * p.max_water_fraction - lwc) {  // 4. rain
        discharge += sca * (rain - (swe * p.max_water_fraction - lwc)) + rain * (1.0 - sca);
        lwc = swe * p.max_water_fraction;
    } else {
        lwc += rain;
        discharge += rain * (1.0 - sca);
    }
    if (discharge >= -s.residual) {
        discharge += s.residual;
        s.residual = 0.0;
    }
    if (nnn > 0) nu /= nnn;  // 5.
    r.outflow = discharge / dt_hours;
    r.swe = sca * (swe + lwc);
    r.sca = sca;
    s.nu = nu;
    s.alpha = alpha;
    s.sca = sca;
    s.swe = swe;
    s.free_water = lwc;
    s.num_units = nnn;
}
}  // namespace skaugen

namespace pt_ss_k {
struct parameter {  // core/pt_ss_k.h:28-150, vector order of set() :63-89
    priestley_taylor::parameter pt;
    skaugen::parameter ss;
    actual_evapotranspiration::parameter ae;
    kirchner::parameter kirchner;
    precipitation_correction::parameter p_corr;
    glacier_melt::parameter gm;
    struct { double velocity = 1.0, alpha = 7.0, beta = 0.0; } routing;
    struct { double reservoir_direct_response_fraction = 1.0; } msp;
    static constexpr size_t n_params = 21;
    void set(const double* p, size_t n) {
        if (n != n_params) throw std::runtime_error("pt_ss_k parameter accessor: .set size mismatch");
        int i = 0;
        kirchner.c1 = p[i++]; kirchner.c2 = p[i++]; kirchner.c3 = p[i++];
        ae.ae_scale_factor = p[i++];
        ss.alpha_0 = p[i++]; ss.d_range = p[i++]; ss.unit_size = p[i++]; ss.max_water_fraction = p[i++];
        ss.tx = p[i++]; ss.cx = p[i++]; ss.ts = p[i++]; ss.cfr = p[i++];
        p_corr.scale_factor = p[i++];
        pt.albedo = p[i++]; pt.alpha = p[i++];
        gm.dtf = p[i++];
        routing.velocity = p[i++]; routing.alpha = p[i++]; routing.beta = p[i++];
        gm.direct_response = p[i++];
        msp.reservoir_direct_response_fraction = p[i++];
    }
};
struct state {  // :152-174; flat layout: snow.nu, snow.alpha, snow.sca, snow.swe, snow.free_water, snow.residual, snow.num_units, kirchner.q
    skaugen::state snow;
    double kirchner_q = 0.1;
    static constexpr size_t n_state = 8;
    void unpack(const double* v) {
        snow.nu = v[0]; snow.alpha = v[1]; snow.sca = v[2]; snow.swe = v[3]; snow.free_water = v[4]; snow.residual = v[5];
        snow.num_units = (unsigned long)v[6];
        kirchner_q = v[7];
    }
    void pack(double* v) const {
        v[0] = snow.nu; v[1] = snow.alpha; v[2] = snow.sca; v[3] = snow.swe; v[4] = snow.free_water; v[5] = snow.residual; v[6] = double(snow.num_units);
        v[7] = kirchner_q;
    }
};
using cell_forcing = pt_hs_k::cell_forcing;
// state series of the state collector (pt_ss_k_cell_model.h:150-199) on state.scale_snow (pt_ss_k.h:165-171):
// kirchner_discharge, snow_swe (= (free_water + swe) sca), snow_sca, snow_alpha, snow_nu, snow_lwc (= free_water sca), snow_residual
enum ss_state_id { SS_KIRCHNER_DISCHARGE = 0, SS_SNOW_SWE, SS_SNOW_SCA, SS_SNOW_ALPHA, SS_SNOW_NU, SS_SNOW_LWC, SS_SNOW_RESIDUAL, SS_N };

// core/pt_ss_k.h:195-293; resp = HR_N nullable response series, sts = SS_N nullable state series (n_axis + 1 points)
inline void run(const geo_cell& geo, const parameter& parameter, const fixed_dt& time_axis, int start_step, int n_steps, const cell_forcing& f, state& st,
                double** resp, double** sts, int64_t tstride, int64_t cstride, size_t ci) {
    priestley_taylor::calculator pt(parameter.pt.albedo, parameter.pt.alpha);
    kirchner::calculator kirchner(parameter.kirchner);
    skaugen::response rsnow;
    const double glacier_fraction = geo.glacier;
    const double gm_direct = parameter.gm.direct_response;
    const double gm_routed = 1 - gm_direct;
    const double snow_storage_fraction = geo.snow_storage();
    const double kirchner_routed_prec = geo.reservoir * (1.0 - parameter.msp.reservoir_direct_response_fraction) + geo.lake;
    const double direct_response_fraction = glacier_fraction * gm_direct + geo.reservoir * parameter.msp.reservoir_direct_response_fraction;
    const double kirchner_fraction = 1 - direct_response_fraction;
    const double cell_area_m2 = geo.area;
    const double glacier_area_m2 = geo.area * glacier_fraction;
    auto collect_state = [&](size_t i) {
        if (!sts) return;
        const int64_t o = int64_t(i) * tstride + int64_t(ci) * cstride;
        // scale_snow: swe, free_water scaled; num_units truncated (not collected)
        const double swe = st.snow.swe * snow_storage_fraction, fw = st.snow.free_water * snow_storage_fraction;
        if (sts[SS_KIRCHNER_DISCHARGE]) sts[SS_KIRCHNER_DISCHARGE][o] = mmh_to_m3s(st.kirchner_q, cell_area_m2);
        if (sts[SS_SNOW_SCA]) sts[SS_SNOW_SCA][o] = st.snow.sca;
        if (sts[SS_SNOW_SWE]) sts[SS_SNOW_SWE][o] = (fw + swe) * st.snow.sca;
        if (sts[SS_SNOW_ALPHA]) sts[SS_SNOW_ALPHA][o] = st.snow.alpha;
        if (sts[SS_SNOW_NU]) sts[SS_SNOW_NU][o] = st.snow.nu;
        if (sts[SS_SNOW_LWC]) sts[SS_SNOW_LWC][o] = fw * st.snow.sca;
        if (sts[SS_SNOW_RESIDUAL]) sts[SS_SNOW_RESIDUAL][o] = st.snow.residual;
    };
    size_t i_begin = n_steps > 0 ? start_step : 0;
    size_t i_end = n_steps > 0 ? start_step + n_steps : time_axis.size();
    for (size_t i = i_begin; i < i_end; ++i) {
        const utctime p_start = time_axis.time(i), p_end = p_start + time_axis.dt;
        double temp = f.temp[int64_t(i) * f.stride];
        double rad = f.rad[int64_t(i) * f.stride];
        double rel_hum = f.rel_hum[int64_t(i) * f.stride];
        double prec = f.prec[int64_t(i) * f.stride] * parameter.p_corr.scale_factor;
        collect_state(i);
        skaugen::step(p_end - p_start, parameter.ss, temp, prec, st.snow, rsnow);
        double gm_melt_m3s = glacier_melt::step(parameter.gm.dtf, temp, geo.area * st.snow.sca, glacier_area_m2);
        double pot = pt.potential_evapotranspiration(temp, rad, rel_hum) * to_seconds(HOUR);
        double ae = actual_evapotranspiration::calculate_step(st.kirchner_q, pot, parameter.ae.ae_scale_factor, std::max(st.snow.sca, glacier_fraction),
                                                              p_end - p_start);
        double gm_mmh = m3s_to_mmh(gm_melt_m3s, cell_area_m2);
        double q_avg = 0.0;
        kirchner.step(p_start, p_end, st.kirchner_q, q_avg, rsnow.outflow * snow_storage_fraction + prec * kirchner_routed_prec + gm_routed * gm_mmh, ae);
        double total_discharge = std::max(0.0, prec - ae) * direct_response_fraction + gm_direct * gm_mmh + q_avg * kirchner_fraction;
        double charge_m3s = +mmh_to_m3s(prec, cell_area_m2) - mmh_to_m3s(ae, cell_area_m2) + gm_melt_m3s - mmh_to_m3s(total_discharge, cell_area_m2);
        if (resp) {  // response.scale_snow (:186-191) + all_response_collector / discharge_collector (pt_ss_k_cell_model.h:78-88,119-126)
            const int64_t o = int64_t(i) * tstride + int64_t(ci) * cstride;
            if (resp[HR_AVG_DISCHARGE]) resp[HR_AVG_DISCHARGE][o] = mmh_to_m3s(total_discharge, cell_area_m2);
            if (resp[HR_CHARGE_M3S]) resp[HR_CHARGE_M3S][o] = charge_m3s;
            if (resp[HR_SNOW_OUTFLOW]) resp[HR_SNOW_OUTFLOW][o] = mmh_to_m3s(rsnow.outflow * snow_storage_fraction, cell_area_m2);
            if (resp[HR_SNOW_SCA]) resp[HR_SNOW_SCA][o] = rsnow.sca;
            if (resp[HR_SNOW_SWE]) resp[HR_SNOW_SWE][o] = rsnow.swe * snow_storage_fraction;  // snow_total_stored_water / snow_swe
            if (resp[HR_GLACIER_MELT]) resp[HR_GLACIER_MELT][o] = gm_melt_m3s;
            if (resp[HR_AE_OUTPUT]) resp[HR_AE_OUTPUT][o] = ae;
            if (resp[HR_PE_OUTPUT]) resp[HR_PE_OUTPUT][o] = pot;
        }
        if (i + 1 == i_end) collect_state(i + 1);
    }
}
}  // namespace pt_ss_k
}  // namespace sho
