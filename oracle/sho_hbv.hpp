// ============================================================================
// oracle/sho_hbv.hpp -- CPU ORACLE (test infrastructure, NOT product code)
// hbv_snow / hbv_soil / hbv_tank / hbv_actual_evapotranspiration and the
// pt_hs_k and hbv_stack method stacks.  Follows core/hbv_snow_common.h:14-67,
// core/hbv_snow.h:37-275, core/hbv_soil.h:17-64, core/hbv_tank.h:17-78,
// core/hbv_actual_evapotranspiration.h:11-38, core/pt_hs_k.h:29-283,
// core/hbv_stack.h:30-361 and the *_cell_model.h collectors.
// ============================================================================
#pragma once
#include <sstream>

#include "sho_core.hpp"

namespace sho {

namespace hbv_snow_common {
// core/hbv_snow_common.h:14-43
inline double integrate(const std::vector<double>& f, const std::vector<double>& x, size_t n, double a, double b, bool f_b_is_zero = false) {
    size_t left = 0;
    double area = 0.0;
    double f_l = 0.0;
    double x_l = a;
    while (a > x[left]) ++left;
    if (std::fabs(a - x[left]) > 1.0e-8 && left > 0) {
        --left;
        f_l = (f[left + 1] - f[left]) / (x[left + 1] - x[left]) * (a - x[left]) + f[left];
    } else
        f_l = f[left];
    while (left < n - 1) {
        if (b >= x[left + 1]) {
            area += 0.5 * (f_l + f[left + 1]) * (x[left + 1] - x_l);
            x_l = x[left + 1];
            f_l = f[left + 1];
            ++left;
        } else {
            if (!f_b_is_zero) area += (f_l + 0.5 * (f[left + 1] - f_l) / (x[left + 1] - x_l) * (b - x_l)) * (b - x_l);
            else area += 0.5 * f_l * (b - x_l);
            break;
        }
    }
    return area;
}
// :45-67
template <class parameter>
void distribute_snow(const parameter& p, std::vector<double>& sp, std::vector<double>& sw, double& swe, double& sca) {
    sp = std::vector<double>(p.intervals.size(), 0.0);
    sw = std::vector<double>(p.intervals.size(), 0.0);
    if (swe <= 1.0e-3 || sca <= 1.0e-3) {
        swe = sca = 0.0;
    } else {
        for (size_t i = 0; i < p.intervals.size(); ++i) sp[i] = sca < p.intervals[i] ? 0.0 : p.s[i] * swe;
        auto temp_swe = integrate(sp, p.intervals, p.intervals.size(), 0.0, sca, true);
        if (temp_swe < swe) {
            const double corr1 = swe / temp_swe * p.lw;
            const double corr2 = swe / temp_swe * (1.0 - p.lw);
            for (size_t i = 0; i < p.intervals.size(); ++i) {
                sw[i] = corr1 * sp[i];
                sp[i] *= corr2;
            }
        } else
            sw = std::vector<double>(p.intervals.size(), 0.0);
    }
}
}  // namespace hbv_snow_common

namespace hbv_snow {
using hbv_snow_common::integrate;
struct parameter {  // core/hbv_snow.h:37-92
    std::vector<double> s, intervals;
    double tx = 0.0, cx = 1.0, ts = 0.0, lw = 0.1, cfr = 0.5;
    parameter() { set_std_distribution_and_quantiles(); }
    void set_std_distribution_and_quantiles() {
        s = {1.0, 1.0, 1.0, 1.0, 1.0};
        intervals = {0, 0.25, 0.5, 0.75, 1.0};
        normalize_snow_distribution();
    }
    void normalize_snow_distribution() {
        const double mean = integrate(s, intervals, intervals.size(), intervals[0], intervals.back());
        for (auto& s_ : s) s_ /= mean;
    }
};
struct state {  // :97-121
    std::vector<double> sp, sw;
    double swe = 0.0, sca = 0.0;
    void distribute(const parameter& p, bool force = true) {
        if (force || sp.size() != p.s.size() || sw.size() != p.s.size()) hbv_snow_common::distribute_snow(p, sp, sw, swe, sca);
    }
};
struct response { double outflow = 0.0; double snow_state_swe = 0.0, snow_state_sca = 0.0; };

struct calculator {  // :143-275
    const parameter p;
    explicit calculator(const parameter& p) : p(p) {}
    static void refreeze(double& sp, double& sw, const double rain, const double potmelt, const double lw) {
        if (sp > 0.0) {
            if (sw + rain > -potmelt) {
                sp -= potmelt;
                sw += potmelt + rain;
                if (sw > sp * lw) sw = sp * lw;
            } else {
                sp += sw + rain;
                sw = 0.0;
            }
        }
    }
    static void update_state(double& sp, double& sw, const double rain, const double potmelt, const double lw) {
        if (sp > potmelt) {
            sw += potmelt + rain;
            sp -= potmelt;
            sw = std::min(sw, sp * lw);
        } else if (sp > 0.0)
            sp = sw = 0.0;
    }
    size_t sca_index(double sca) const {
        for (size_t i = 0; i < p.intervals.size() - 1; ++i)
            if (sca >= p.intervals[i] && sca < p.intervals[i + 1]) return i;
        return p.intervals.size() - 1;
    }
    size_t melt_index(double potmelt, const state& s) const {
        for (size_t i = 0; i < p.intervals.size(); ++i)
            if (s.sp[i] < potmelt) return i;
        return p.intervals.size();
    }
    void step(state& s, response& r, utctime t0, utctime t1, double prec_mm_h, double temp) const {  // :195-275
        double swe = s.swe;
        double sca = s.sca;
        const auto& I = p.intervals;
        double step_in_days = to_seconds(t1 - t0) / 86400.0;
        const double dt_hours = to_seconds(t1 - t0) / 3600.0;
        const double prec = prec_mm_h * dt_hours;
        const double total_water = prec + swe;
        double snow, rain;
        if (temp < p.tx) { snow = prec; rain = 0.0; }
        else { snow = 0.0; rain = prec; }
        swe += snow + sca * rain;
        if (swe < 0.1) {
            r.outflow = total_water / dt_hours;
            std::fill(s.sp.begin(), s.sp.end(), 0.0);
            std::fill(s.sw.begin(), s.sw.end(), 0.0);
            s.swe = 0.0;
            s.sca = 0.0;
            return;
        }
        if (snow > 0.0) {
            auto idx = sca_index(sca);
            if (sca > 1.0e-5 && sca < 1.0 - 1.0e-5) {
                if (idx == 0) {
                    s.sp[0] *= sca / (I[1] - I[0]);
                    s.sw[0] *= sca / (I[1] - I[0]);
                } else {
                    s.sp[idx] *= (1.0 + (sca - I[idx]) / (I[idx] - I[idx - 1])) / (1.0 + (I[idx + 1] - I[idx]) / (I[idx] - I[idx - 1]));
                    s.sw[idx] *= (1.0 + (sca - I[idx]) / (I[idx] - I[idx - 1])) / (1.0 + (I[idx + 1] - I[idx]) / (I[idx] - I[idx - 1]));
                }
            }
            for (size_t i = 0; i < p.s.size(); ++i) s.sp[i] += snow * p.s[i];
            sca = I[1];
            for (size_t i = I.size() - 2; i > 0; --i)
                if (p.s[i] > 0.0) { sca = I[i + 1]; break; }
        }
        double potmelt = p.cx * step_in_days * (temp - p.ts);
        const double lw = p.lw;
        if (potmelt < 0.0) {
            potmelt *= p.cfr;
            for (size_t i = 0; i < I.size(); ++i) refreeze(s.sp[i], s.sw[i], rain, potmelt, lw);
        } else {
            size_t idx = melt_index(potmelt, s);
            if (idx == 0) sca = 0.0;
            else if (idx == I.size()) sca = 1.0;
            else {
                if (s.sp[idx] > 0.0) sca = I[idx] - (I[idx] - I[idx - 1]) * (potmelt - s.sp[idx]) / (s.sp[idx - 1] - s.sp[idx]);
                else sca = (1.0 - potmelt / s.sp[idx - 1]) * (sca - I[idx - 1]) + I[idx - 1];
            }
            for (size_t i = 0; i < I.size(); ++i) update_state(s.sp[i], s.sw[i], rain, potmelt, lw);
        }
        if (sca < 1.0e-6) swe = 0.0;
        else {
            bool f_is_zero = sca >= 1.0 ? false : true;
            swe = integrate(s.sp, I, I.size(), 0, sca, f_is_zero);
            swe += integrate(s.sw, I, I.size(), 0, sca, f_is_zero);
        }
        if (total_water < swe) {
            if (total_water - swe < -1.0e-6) {
                std::ostringstream buff;
                buff << "Negative outflow: total_water (" << total_water << ") - swe (" << swe << ") = " << total_water - swe;
                throw std::runtime_error(buff.str());
            } else
                swe = total_water;
        }
        r.outflow = (total_water - swe) / dt_hours;
        s.swe = swe;
        s.sca = sca;
    }
};
}  // namespace hbv_snow

namespace hbv_soil {  // core/hbv_soil.h:17-64
struct parameter { double fc = 300.0, beta = 2.0; };
struct state { double sm = 0.0; };  // explicit state(double sm = 0.0): a default-constructed state has sm = 0 (:28-29)
struct response { double outflow = 0.0; };
struct calculator {
    parameter param;
    explicit calculator(const parameter& p) : param(p) {}
    void step(state& s, response& r, utctime, utctime, double insoil, double act_evap) const {
        double temp = s.sm + insoil;
        double outflow = insoil * dm::pow(temp / param.fc, param.beta);
        r.outflow = outflow > temp ? temp : outflow;
        s.sm = std::max(0.0, s.sm + insoil - r.outflow - act_evap);
    }
};
}  // namespace hbv_soil

namespace hbv_tank {  // core/hbv_tank.h:17-78
struct parameter { double uz1 = 25.0, kuz2 = 0.5, kuz1 = 0.3, perc = 0.8, klz = 0.02; };
struct state { double uz = 20.0, lz = 10.0; };
struct response { double outflow = 0.0; };
struct calculator {
    parameter param;
    explicit calculator(const parameter& p) : param(p) {}
    void step(state& s, response& r, utctime, utctime, double soil_outflow) const {
        double temp = s.uz + soil_outflow;
        double q12 = std::max(0.0, (temp - param.uz1) * param.kuz2);
        double q11 = std::min(temp, param.uz1) * param.kuz1;
        s.uz = s.uz + soil_outflow - param.perc - (q12 + q11);
        double q2 = (s.lz + param.perc) * param.klz;
        s.lz = s.lz + param.perc - q2;
        r.outflow = q12 + q11 + q2;
    }
};
}  // namespace hbv_tank

namespace hbv_actual_evapotranspiration {  // core/hbv_actual_evapotranspiration.h:11-38
struct parameter { double lp = 150.0; };
inline double calculate_step(double soil_moisture, double pot_evapo, double lp, double snow_fraction, utctimespan) {
    return (1.0 - snow_fraction) * (soil_moisture < lp ? pot_evapo * (soil_moisture / lp) : pot_evapo);
}
}  // namespace hbv_actual_evapotranspiration

// response series ids shared by the two HBV-snow stacks (pt_hs_k_cell_model.h:41-93, hbv_stack_cell_model.h:38-93)
enum hs_response_id { HR_AVG_DISCHARGE = 0, HR_CHARGE_M3S, HR_SNOW_SCA, HR_SNOW_SWE, HR_SNOW_OUTFLOW, HR_GLACIER_MELT, HR_AE_OUTPUT, HR_PE_OUTPUT, HR_SOIL_OUTFLOW, HR_N };

namespace pt_hs_k {
struct parameter {  // core/pt_hs_k.h:29-147
    priestley_taylor::parameter pt;
    hbv_snow::parameter hs;
    actual_evapotranspiration::parameter ae;
    kirchner::parameter kirchner;
    precipitation_correction::parameter p_corr;
    glacier_melt::parameter gm;
    struct { double velocity = 1.0, alpha = 7.0, beta = 0.0; } routing;  // routing::uhg_parameter defaults, core/routing.h:30-35
    struct { double reservoir_direct_response_fraction = 1.0; } msp;
    static constexpr size_t n_params = 18;
    void set(const double* p, size_t n) {  // :66-88
        if (n != n_params) throw std::runtime_error("pt_ss_k parameter accessor: .set size missmatch");
        int i = 0;
        kirchner.c1 = p[i++]; kirchner.c2 = p[i++]; kirchner.c3 = p[i++];
        ae.ae_scale_factor = p[i++];
        hs.lw = p[i++]; hs.tx = p[i++]; hs.cx = p[i++]; hs.ts = p[i++]; hs.cfr = p[i++];
        gm.dtf = p[i++];
        p_corr.scale_factor = p[i++];
        pt.albedo = p[i++]; pt.alpha = p[i++];
        routing.velocity = p[i++]; routing.alpha = p[i++]; routing.beta = p[i++];
        gm.direct_response = p[i++];
        msp.reservoir_direct_response_fraction = p[i++];
    }
};
struct state {  // :150-170
    hbv_snow::state snow;
    double kirchner_q = 0.1;
    // flat layout: swe, sca, sp[nb], sw[nb], kirchner.q   (n = 3 + 2*nb; nb = 0 allowed -> bins distributed on run)
    void unpack(const double* v, size_t n) {
        size_t nb = (n - 3) / 2;
        snow.swe = v[0]; snow.sca = v[1];
        snow.sp.assign(v + 2, v + 2 + nb); snow.sw.assign(v + 2 + nb, v + 2 + 2 * nb);
        kirchner_q = v[2 + 2 * nb];
    }
    void pack(double* v, size_t n) const {
        size_t nb = (n - 3) / 2;
        v[0] = snow.swe; v[1] = snow.sca;
        for (size_t i = 0; i < nb; ++i) { v[2 + i] = snow.sp[i]; v[2 + nb + i] = snow.sw[i]; }
        v[2 + 2 * nb] = kirchner_q;
    }
};
struct cell_forcing { const double* temp; const double* prec; const double* wind_speed; const double* rel_hum; const double* rad; int64_t stride; };

// core/pt_hs_k.h:201-283; resp = HR_N nullable series, element (step i, cell c) at p[i*tstride + c*cstride]
inline void run(const geo_cell& geo, const parameter& parameter, const fixed_dt& time_axis, int start_step, int n_steps, const cell_forcing& f,
                state& st, double** resp, int64_t tstride, int64_t cstride, size_t ci) {
    priestley_taylor::calculator pt(parameter.pt.albedo, parameter.pt.alpha);
    hbv_snow::calculator hbv_snow(parameter.hs);
    kirchner::calculator kirchner(parameter.kirchner);
    st.snow.distribute(parameter.hs, false);
    hbv_snow::response rsnow;
    const double glacier_fraction = geo.glacier;
    const double gm_direct = parameter.gm.direct_response;
    const double gm_routed = 1 - gm_direct;
    const double snow_storage_fraction = geo.snow_storage();
    const double kirchner_routed_prec = geo.reservoir * (1.0 - parameter.msp.reservoir_direct_response_fraction) + geo.lake;
    const double direct_response_fraction = glacier_fraction * gm_direct + geo.reservoir * parameter.msp.reservoir_direct_response_fraction;
    const double kirchner_fraction = 1 - direct_response_fraction;
    const double cell_area_m2 = geo.area;
    const double glacier_area_m2 = geo.area * glacier_fraction;
    size_t i_begin = n_steps > 0 ? start_step : 0;
    size_t i_end = n_steps > 0 ? start_step + n_steps : time_axis.size();
    for (size_t i = i_begin; i < i_end; ++i) {
        const utctime p_start = time_axis.time(i), p_end = p_start + time_axis.dt;
        double temp = f.temp[int64_t(i) * f.stride];
        double rad = f.rad[int64_t(i) * f.stride];
        double rel_hum = f.rel_hum[int64_t(i) * f.stride];
        double prec = f.prec[int64_t(i) * f.stride] * parameter.p_corr.scale_factor;
        hbv_snow.step(st.snow, rsnow, p_start, p_end, prec, temp);
        double gm_melt_m3s = glacier_melt::step(parameter.gm.dtf, temp, cell_area_m2 * st.snow.sca, glacier_area_m2);
        double pot = pt.potential_evapotranspiration(temp, rad, rel_hum) * to_seconds(HOUR);
        double ae = actual_evapotranspiration::calculate_step(st.kirchner_q, pot, parameter.ae.ae_scale_factor, std::max(st.snow.sca, glacier_fraction),
                                                              p_end - p_start);
        double gm_mmh = m3s_to_mmh(gm_melt_m3s, cell_area_m2);
        double q_avg = 0.0;
        kirchner.step(p_start, p_end, st.kirchner_q, q_avg, rsnow.outflow * snow_storage_fraction + prec * kirchner_routed_prec + gm_routed * gm_mmh, ae);
        double total_discharge = std::max(0.0, prec - ae) * direct_response_fraction + gm_direct * gm_mmh + q_avg * kirchner_fraction;
        double charge_m3s = +mmh_to_m3s(prec, cell_area_m2) - mmh_to_m3s(ae, cell_area_m2) + gm_melt_m3s - mmh_to_m3s(total_discharge, cell_area_m2);
        if (resp) {  // response.scale_snow (:192-198) + all_response_collector::collect (pt_hs_k_cell_model.h:82-91)
            const int64_t o = int64_t(i) * tstride + int64_t(ci) * cstride;
            if (resp[HR_AVG_DISCHARGE]) resp[HR_AVG_DISCHARGE][o] = mmh_to_m3s(total_discharge, cell_area_m2);
            if (resp[HR_CHARGE_M3S]) resp[HR_CHARGE_M3S][o] = charge_m3s;
            if (resp[HR_SNOW_OUTFLOW]) resp[HR_SNOW_OUTFLOW][o] = mmh_to_m3s(rsnow.outflow * snow_storage_fraction, cell_area_m2);
            if (resp[HR_SNOW_SCA]) resp[HR_SNOW_SCA][o] = st.snow.sca;
            if (resp[HR_SNOW_SWE]) resp[HR_SNOW_SWE][o] = st.snow.swe * snow_storage_fraction;
            if (resp[HR_GLACIER_MELT]) resp[HR_GLACIER_MELT][o] = gm_melt_m3s;
            if (resp[HR_AE_OUTPUT]) resp[HR_AE_OUTPUT][o] = ae;
            if (resp[HR_PE_OUTPUT]) resp[HR_PE_OUTPUT][o] = pot;
        }
    }
}
}  // namespace pt_hs_k

namespace hbv_stack {
struct parameter {  // core/hbv_stack.h:30-163
    priestley_taylor::parameter pt;
    hbv_snow::parameter snow;
    hbv_actual_evapotranspiration::parameter ae;
    hbv_soil::parameter soil;
    hbv_tank::parameter tank;
    precipitation_correction::parameter p_corr;
    glacier_melt::parameter gm;
    struct { double velocity = 1.0, alpha = 7.0, beta = 0.0; } routing;  // routing::uhg_parameter defaults, core/routing.h:30-35
    struct { double reservoir_direct_response_fraction = 1.0; } msp;
    static constexpr size_t n_params = 22;
    void set(const double* p, size_t n) {  // :73-99
        if (n != n_params) throw std::runtime_error("HBV_Stack Parameter Accessor: .set size missmatch");
        int i = 0;
        soil.fc = p[i++]; soil.beta = p[i++];
        ae.lp = p[i++];
        tank.uz1 = p[i++]; tank.kuz2 = p[i++]; tank.kuz1 = p[i++]; tank.perc = p[i++]; tank.klz = p[i++];
        snow.lw = p[i++]; snow.tx = p[i++]; snow.cx = p[i++]; snow.ts = p[i++]; snow.cfr = p[i++];
        p_corr.scale_factor = p[i++];
        pt.albedo = p[i++]; pt.alpha = p[i++];
        gm.dtf = p[i++];
        routing.velocity = p[i++]; routing.alpha = p[i++]; routing.beta = p[i++];
        gm.direct_response = p[i++];
        msp.reservoir_direct_response_fraction = p[i++];
    }
};
struct state {  // :165-180
    hbv_snow::state snow;
    hbv_soil::state soil;
    hbv_tank::state tank;
    // flat layout: swe, sca, sp[nb], sw[nb], soil.sm, tank.uz, tank.lz   (n = 5 + 2*nb)
    void unpack(const double* v, size_t n) {
        size_t nb = (n - 5) / 2;
        snow.swe = v[0]; snow.sca = v[1];
        snow.sp.assign(v + 2, v + 2 + nb); snow.sw.assign(v + 2 + nb, v + 2 + 2 * nb);
        soil.sm = v[2 + 2 * nb]; tank.uz = v[3 + 2 * nb]; tank.lz = v[4 + 2 * nb];
    }
    void pack(double* v, size_t n) const {
        size_t nb = (n - 5) / 2;
        v[0] = snow.swe; v[1] = snow.sca;
        for (size_t i = 0; i < nb; ++i) { v[2 + i] = snow.sp[i]; v[2 + nb + i] = snow.sw[i]; }
        v[2 + 2 * nb] = soil.sm; v[3 + 2 * nb] = tank.uz; v[4 + 2 * nb] = tank.lz;
    }
};
using cell_forcing = pt_hs_k::cell_forcing;

// core/hbv_stack.h:278-361
inline void run_hbv_stack(const geo_cell& geo, const parameter& parameter, const fixed_dt& time_axis, int start_step, int n_steps,
                          const cell_forcing& f, state& st, double** resp, int64_t tstride, int64_t cstride, size_t ci) {
    priestley_taylor::calculator pt(parameter.pt.albedo, parameter.pt.alpha);
    hbv_snow::calculator snow(parameter.snow);
    hbv_soil::calculator soil(parameter.soil);
    hbv_tank::calculator tank(parameter.tank);
    st.snow.distribute(parameter.snow, false);
    hbv_snow::response rsnow;
    hbv_soil::response rsoil;
    hbv_tank::response rtank;
    const double glacier_fraction = geo.glacier;
    const double gm_direct = parameter.gm.direct_response;
    const double gm_routed = 1 - gm_direct;
    const double direct_response_fraction = glacier_fraction * gm_direct + geo.reservoir * parameter.msp.reservoir_direct_response_fraction;
    const double land_fraction = 1 - direct_response_fraction;
    const double cell_area_m2 = geo.area;
    const double glacier_area_m2 = geo.area * glacier_fraction;
    size_t i_begin = n_steps > 0 ? start_step : 0;
    size_t i_end = n_steps > 0 ? start_step + n_steps : time_axis.size();
    for (size_t i = i_begin; i < i_end; ++i) {
        const utctime p_start = time_axis.time(i), p_end = p_start + time_axis.dt;
        double temp = f.temp[int64_t(i) * f.stride];
        double rad = f.rad[int64_t(i) * f.stride];
        double rel_hum = f.rel_hum[int64_t(i) * f.stride];
        double prec = f.prec[int64_t(i) * f.stride] * parameter.p_corr.scale_factor;
        snow.step(st.snow, rsnow, p_start, p_end, prec, temp);
        double gm_melt_m3s = glacier_melt::step(parameter.gm.dtf, temp, geo.area * st.snow.sca, glacier_area_m2);
        double pot = pt.potential_evapotranspiration(temp, rad, rel_hum) * to_seconds(HOUR);
        double ae = hbv_actual_evapotranspiration::calculate_step(st.soil.sm, pot, parameter.ae.lp, std::max(st.snow.sca, glacier_fraction), p_end - p_start);
        double gm_mmh = m3s_to_mmh(gm_melt_m3s, cell_area_m2);
        soil.step(st.soil, rsoil, p_start, p_end, rsnow.outflow, ae);
        tank.step(st.tank, rtank, p_start, p_end, rsoil.outflow + gm_routed * gm_mmh);
        double total_discharge = std::max(0.0, prec - ae) * direct_response_fraction + gm_direct * gm_mmh + rtank.outflow * land_fraction;
        double charge_m3s = +mmh_to_m3s(prec, cell_area_m2) - mmh_to_m3s(ae, cell_area_m2) + gm_melt_m3s - mmh_to_m3s(total_discharge, cell_area_m2);
        if (resp) {  // hbv_stack_cell_model.h:82-92.  response.snow.snow_state is never assigned by run_hbv_stack -> sca = swe = 0
            const int64_t o = int64_t(i) * tstride + int64_t(ci) * cstride;
            if (resp[HR_PE_OUTPUT]) resp[HR_PE_OUTPUT][o] = pot;
            if (resp[HR_SNOW_OUTFLOW]) resp[HR_SNOW_OUTFLOW][o] = mmh_to_m3s(rsnow.outflow, cell_area_m2);
            if (resp[HR_GLACIER_MELT]) resp[HR_GLACIER_MELT][o] = gm_melt_m3s;
            if (resp[HR_SNOW_SCA]) resp[HR_SNOW_SCA][o] = 0.0;
            if (resp[HR_SNOW_SWE]) resp[HR_SNOW_SWE][o] = 0.0;
            if (resp[HR_AE_OUTPUT]) resp[HR_AE_OUTPUT][o] = ae;
            if (resp[HR_SOIL_OUTFLOW]) resp[HR_SOIL_OUTFLOW][o] = rsoil.outflow;
            if (resp[HR_AVG_DISCHARGE]) resp[HR_AVG_DISCHARGE][o] = mmh_to_m3s(total_discharge, cell_area_m2);
            if (resp[HR_CHARGE_M3S]) resp[HR_CHARGE_M3S][o] = charge_m3s;
        }
    }
}
}  // namespace hbv_stack
}  // namespace sho
