// ============================================================================
// oracle/sho_hps.hpp -- CPU ORACLE (test infrastructure, NOT product code)
//
// hbv_physical_snow (the hbv_snow quantile bins driven by gamma_snow's energy balance, one balance per bin) and the pt_hps_k method stack
// (SURVEY.md 8f item 4), restated from
//   core/hbv_physical_snow.h:40-559    parameter / state / calculator::step (refreeze :211-228, update_state :231-238, sca_index :240-245)
//   core/pt_hps_k.h:29-308             parameter (24 values, set / get :63-120) / state / response, run() :201-303
//   core/pt_hps_k_cell_model.h         collectors
// libm calls go through sho_detmath.hpp as everywhere in the oracle (pow(x, 8) and pow(x, 4) as repeated squaring, as in gamma_snow).
// Two oddities of the reference are kept as they are, because results are defined by them:
//   * after a snowfall sca is set from the redistribution FACTORS p.s[i + 1] (hbv_physical_snow.h:357-362), not from the quantiles;
//   * the melt branch's sca formula assigns inside its denominator: (s.sp[idx - 1] = s.sp[idx])  (:473-476).
// PARITY STATUS: pinned by the reference's mass-balance asserts (test/hbv_physical_snow_test.cpp:26-176: snow pack reset, build-up incl. T =
// tx, rain without snow, melt without precipitation) and the stack test (test/pt_hps_k_test.cpp).
// ============================================================================
#pragma once
#include "sho_hbv.hpp"

namespace sho {
namespace hbv_physical_snow {
using hbv_snow_common::integrate;
const double tol = 1.0e-10;  // hbv_physical_snow.h:36

struct parameter {  // :40-129
    std::vector<double> s, intervals;
    double tx = 0.0, lw = 0.1, cfr = 0.5, wind_scale = 2.0, wind_const = 1.0, surface_magnitude = 30.0, max_albedo = 0.9, min_albedo = 0.6,
           fast_albedo_decay_rate = 5.0, slow_albedo_decay_rate = 5.0, snowfall_reset_depth = 5.0;
    bool calculate_iso_pot_energy = false;
    parameter() { set_std_distribution_and_quantiles(); }
    parameter(const std::vector<double>& s_, const std::vector<double>& i_) : s(s_), intervals(i_) { normalize_snow_distribution(); }  // :95-118
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
struct state {  // :132-189
    std::vector<double> sp, sw, albedo, iso_pot_energy;
    double surface_heat = 30000.0, swe = 0.0, sca = 0.0;
    void distribute(const parameter& p, bool force = true) {
        if (force || sp.size() != p.s.size() || sw.size() != p.s.size()) hbv_snow_common::distribute_snow(p, sp, sw, swe, sca);
        if (sp.size() != albedo.size()) {
            albedo = std::vector<double>(sp.size(), 0.4);
            iso_pot_energy = std::vector<double>(sp.size(), 0.0);
        }
    }
};
struct response { double sca = 0.0, storage = 0.0, outflow = 0.0; };

struct calculator {  // :200-531
    const parameter p;
    const double melt_heat = 333660.0, water_heat = 4180.0, ice_heat = 2050.0, sigma = 5.670373e-8;
    const double BB0{0.98 * sigma * dm::pow4(273.15)};
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
    // :266-529
    void step(state& s, response& r, utctime /*t*/, utctimespan dt, const double T, const double rad, const double prec_mm_h, const double wind_speed,
              const double rel_hum) const {
        const auto& I = p.intervals;
        const double prec = prec_mm_h * dt / HOUR;
        const double total_water = prec + s.swe;
        double snow, rain;
        if (T < p.tx) { snow = prec; rain = 0.0; }
        else { snow = 0.0; rain = prec; }
        if (std::fabs(snow + rain - prec) > 1.0e-8) throw std::runtime_error("Mass balance violation!!!!");
        s.swe += snow + s.sca * rain;
        if (s.swe < tol) {
            r.outflow = total_water;
            std::fill(s.sp.begin(), s.sp.end(), 0.0);
            std::fill(s.sw.begin(), s.sw.end(), 0.0);
            s.swe = 0.0;
            s.sca = 0.0;
            r.sca = 0.0;
            r.storage = 0.0;
            std::fill(s.albedo.begin(), s.albedo.end(), p.max_albedo);
            s.surface_heat = 0.0;
            std::fill(s.iso_pot_energy.begin(), s.iso_pot_energy.end(), 0.0);
            return;
        }
        std::vector<double> albedo = s.albedo;
        double surface_heat = s.surface_heat;
        const double min_albedo = p.min_albedo;
        const double max_albedo = p.max_albedo;
        const double albedo_range = max_albedo - min_albedo;
        const double dt_in_days = to_seconds(dt) / to_seconds(DAY);
        const double slow_albedo_decay_rate = (0.5 * albedo_range * dt_in_days / p.slow_albedo_decay_rate);
        const double fast_albedo_decay_rate = dm::pow(2.0, -dt_in_days / p.fast_albedo_decay_rate);
        const double T_k = T + 273.15;
        const double turb = p.wind_scale * wind_speed + p.wind_const;
        double vapour_pressure = (33.864 * (dm::pow8(7.38e-3 * T + 0.8072) - 1.9e-5 * std::fabs(1.8 * T + 48.0) + 1.316e-3) * rel_hum);
        if (T < 0.0) vapour_pressure *= 1.0 + 9.72e-3 * T + 4.2e-5 * T * T;
        if (snow > tol) {
            auto idx = sca_index(s.sca);
            if (s.sca > 1.0e-5 && s.sca < 1.0 - 1.0e-5) {
                if (idx == 0) {
                    s.sp[0] *= s.sca / (I[1] - I[0]);
                    s.sw[0] *= s.sca / (I[1] - I[0]);
                } else {
                    s.sp[idx] *= (1.0 + (s.sca - I[idx]) / (I[idx] - I[idx - 1])) / (1.0 + (I[idx + 1] - I[idx]) / (I[idx] - I[idx - 1]));
                    s.sw[idx] *= (1.0 + (s.sca - I[idx]) / (I[idx] - I[idx - 1])) / (1.0 + (I[idx + 1] - I[idx]) / (I[idx] - I[idx - 1]));
                }
            }
            for (size_t i = 0; i < I.size(); ++i) {
                double currsnow = snow * p.s[i];
                s.sp[i] += currsnow;
                albedo[i] += (currsnow * albedo_range / p.snowfall_reset_depth);
            }
            for (size_t i = I.size() - 2; i > 0; --i)
                if (p.s[i] > 0.0) {
                    s.sca = p.s[i + 1];
                    break;
                } else
                    s.sca = p.s[1];
        } else {
            if (T < 0.0) {
                for (auto& alb : albedo) alb -= slow_albedo_decay_rate;
            } else {
                for (auto& alb : albedo) alb = (min_albedo + fast_albedo_decay_rate * (alb - min_albedo));
            }
        }
        for (auto& alb : albedo) alb = std::max(std::min(alb, max_albedo), min_albedo);
        std::vector<double> effect;
        for (auto alb : albedo) effect.push_back(rad * (1.0 - alb));
        for (auto& eff : effect) eff += (0.98 * sigma * dm::pow(vapour_pressure / T_k, 6.87e-2) * dm::pow4(T_k));
        if (T > 0.0 && snow < tol)
            for (auto& eff : effect) eff += rain * T * water_heat / to_seconds(dt);
        if (T <= 0.0 && rain < tol)
            for (size_t i = 0; i < I.size(); ++i) effect[i] += snow * p.s[i] * T * ice_heat / to_seconds(dt);
        if (p.calculate_iso_pot_energy)
            for (size_t i = 0; i < I.size(); ++i) {
                double iso_effect = (effect[i] - BB0 + turb * (T + 1.7 * (vapour_pressure - 6.12)));
                s.iso_pot_energy[i] += (iso_effect * to_seconds(dt) / melt_heat);
            }
        double sst = std::min(0.0, 1.16 * T - 2.09);
        if (sst > -tol) {
            for (auto& eff : effect) eff += turb * (T + 1.7 * (vapour_pressure - 6.12)) - BB0;
        } else {
            for (auto& eff : effect)
                eff += (turb * (T - sst + 1.7 * (vapour_pressure - 6.132 * dm::exp(0.103 * T - 0.186))) - 0.98 * sigma * dm::pow4(sst + 273.15));
        }
        double delta_sh = -surface_heat;
        surface_heat = p.surface_magnitude * ice_heat * sst * 0.5;
        delta_sh += surface_heat;
        std::vector<double> energy;
        for (auto eff : effect) energy.push_back(eff * to_seconds(dt));
        if (delta_sh > 0.0)
            for (auto& en : energy) en -= delta_sh;
        std::vector<double> potential_melt;
        for (auto en : energy) potential_melt.push_back(en / melt_heat);
        const double lw = p.lw;
        size_t idx = I.size();
        bool any_melt = false;
        for (size_t i = 0; i < I.size(); ++i) {
            if (potential_melt[i] >= tol) {
                any_melt = true;
                if (s.sp[i] < potential_melt[i]) {
                    idx = i;
                    break;
                }
            }
        }
        if (any_melt) {
            if (idx == 0) s.sca = 0.0;
            else if (idx == I.size()) s.sca = 1.0;
            else {
                if (s.sp[idx] > 0.0) {
                    s.sca = (I[idx] - (I[idx] - I[idx - 1]) * (potential_melt[idx] - s.sp[idx]) / (s.sp[idx - 1] = s.sp[idx]));  // sic
                } else {
                    s.sca = (1.0 - potential_melt[idx] / s.sp[idx - 1]) * (s.sca - I[idx - 1]) + I[idx - 1];
                }
            }
        }
        for (size_t i = 0; i < I.size(); ++i) {
            if (potential_melt[i] < tol) refreeze(s.sp[i], s.sw[i], rain, p.cfr * potential_melt[i], lw);
            else update_state(s.sp[i], s.sw[i], rain, potential_melt[i], lw);
        }
        if (s.sca < tol) s.swe = 0.0;
        else {
            bool f_is_zero = s.sca >= 1.0 ? false : true;
            s.swe = integrate(s.sp, I, I.size(), 0, s.sca, f_is_zero);
            s.swe += integrate(s.sw, I, I.size(), 0, s.sca, f_is_zero);
        }
        if (total_water < s.swe) {
            if (total_water - s.swe < -tol) throw std::runtime_error("Negative outflow: total_water - s.swe < 0");
            else s.swe = total_water;
        }
        r.outflow = total_water - s.swe;
        r.sca = s.sca;
        r.storage = s.swe;
        // NOTE: albedo and surface_heat are locals of the reference's step as well: the state keeps its values unless the pack is reset above
        (void)albedo;
        (void)surface_heat;
    }
};
}  // namespace hbv_physical_snow

namespace pt_hps_k {
struct parameter {  // core/pt_hps_k.h:29-158, vector order of set() :63-92
    priestley_taylor::parameter pt;
    hbv_physical_snow::parameter hps;
    actual_evapotranspiration::parameter ae;
    kirchner::parameter kirchner;
    precipitation_correction::parameter p_corr;
    glacier_melt::parameter gm;
    struct { double velocity = 1.0, alpha = 7.0, beta = 0.0; } routing;
    struct { double reservoir_direct_response_fraction = 1.0; } msp;
    static constexpr size_t n_params = 24;
    void set(const double* p, size_t n) {
        if (n != n_params) throw std::runtime_error("pt_ss_k parameter accessor: .set size missmatch");  // the reference's text
        int i = 0;
        kirchner.c1 = p[i++]; kirchner.c2 = p[i++]; kirchner.c3 = p[i++];
        ae.ae_scale_factor = p[i++];
        hps.lw = p[i++]; hps.tx = p[i++]; hps.cfr = p[i++]; hps.wind_scale = p[i++]; hps.wind_const = p[i++]; hps.surface_magnitude = p[i++];
        hps.max_albedo = p[i++]; hps.min_albedo = p[i++]; hps.fast_albedo_decay_rate = p[i++]; hps.slow_albedo_decay_rate = p[i++];
        hps.snowfall_reset_depth = p[i++];
        hps.calculate_iso_pot_energy = std::fabs(p[i++]) < 0.0001 ? false : true;
        gm.dtf = p[i++];
        p_corr.scale_factor = p[i++];
        pt.albedo = p[i++]; pt.alpha = p[i++];
        routing.velocity = p[i++]; routing.alpha = p[i++]; routing.beta = p[i++];
        msp.reservoir_direct_response_fraction = p[i++];
    }
};
// flat layout (the member order of hbv_physical_snow::state, then kirchner): sp[nb], sw[nb], albedo[nb], iso_pot_energy[nb], surface_heat, swe, sca,
// kirchner.q   (n = 4*nb + 4)
struct state {
    hbv_physical_snow::state hps;
    double kirchner_q = 0.1;
    void unpack(const double* v, size_t n) {
        const size_t nb = (n - 4) / 4;
        hps.sp.assign(v, v + nb); hps.sw.assign(v + nb, v + 2 * nb); hps.albedo.assign(v + 2 * nb, v + 3 * nb);
        hps.iso_pot_energy.assign(v + 3 * nb, v + 4 * nb);
        hps.surface_heat = v[4 * nb]; hps.swe = v[4 * nb + 1]; hps.sca = v[4 * nb + 2];
        kirchner_q = v[4 * nb + 3];
    }
    void pack(double* v, size_t n) const {
        const size_t nb = (n - 4) / 4;
        for (size_t i = 0; i < nb; ++i) { v[i] = hps.sp[i]; v[nb + i] = hps.sw[i]; v[2 * nb + i] = hps.albedo[i]; v[3 * nb + i] = hps.iso_pot_energy[i]; }
        v[4 * nb] = hps.surface_heat; v[4 * nb + 1] = hps.swe; v[4 * nb + 2] = hps.sca; v[4 * nb + 3] = kirchner_q;
    }
};
using cell_forcing = pt_hs_k::cell_forcing;

// core/pt_hps_k.h:201-303; resp = HR_N nullable response series
inline void run(const geo_cell& geo, const parameter& parameter, const fixed_dt& time_axis, int start_step, int n_steps, const cell_forcing& f, state& st,
                double** resp, int64_t tstride, int64_t cstride, size_t ci) {
    priestley_taylor::calculator pt(parameter.pt.albedo, parameter.pt.alpha);
    const hbv_physical_snow::calculator hps(parameter.hps);
    kirchner::calculator kirchner(parameter.kirchner);
    st.hps.distribute(parameter.hps, false);
    hbv_physical_snow::response rsnow;
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
        double wind_speed = f.wind_speed[int64_t(i) * f.stride];
        hps.step(st.hps, rsnow, p_start, p_end - p_start, temp, rad, prec, wind_speed, rel_hum);
        double gm_melt_m3s = glacier_melt::step(parameter.gm.dtf, temp, geo.area * st.hps.sca, glacier_area_m2);
        double pot = pt.potential_evapotranspiration(temp, rad, rel_hum) * to_seconds(HOUR);
        double ae = actual_evapotranspiration::calculate_step(st.kirchner_q, pot, parameter.ae.ae_scale_factor, std::max(st.hps.sca, glacier_fraction),
                                                              p_end - p_start);
        double gm_mmh = m3s_to_mmh(gm_melt_m3s, cell_area_m2);
        double q_avg = 0.0;
        kirchner.step(p_start, p_end, st.kirchner_q, q_avg, rsnow.outflow * snow_storage_fraction + prec * kirchner_routed_prec + gm_routed * gm_mmh, ae);
        double total_discharge = std::max(0.0, prec - ae) * direct_response_fraction + gm_direct * gm_mmh + q_avg * kirchner_fraction;
        double charge_m3s = +mmh_to_m3s(prec, cell_area_m2) - mmh_to_m3s(ae, cell_area_m2) + gm_melt_m3s - mmh_to_m3s(total_discharge, cell_area_m2);
        if (resp) {  // response.scale_snow (:193-198) + all_response_collector::collect (pt_hps_k_cell_model.h)
            const int64_t o = int64_t(i) * tstride + int64_t(ci) * cstride;
            if (resp[HR_AVG_DISCHARGE]) resp[HR_AVG_DISCHARGE][o] = mmh_to_m3s(total_discharge, cell_area_m2);
            if (resp[HR_CHARGE_M3S]) resp[HR_CHARGE_M3S][o] = charge_m3s;
            if (resp[HR_SNOW_OUTFLOW]) resp[HR_SNOW_OUTFLOW][o] = rsnow.outflow * snow_storage_fraction;  // hps_outflow stays in mm/h (pt_hps_k_cell_model.h:85)
            if (resp[HR_SNOW_SCA]) resp[HR_SNOW_SCA][o] = rsnow.sca;
            if (resp[HR_SNOW_SWE]) resp[HR_SNOW_SWE][o] = rsnow.storage * snow_storage_fraction;
            if (resp[HR_GLACIER_MELT]) resp[HR_GLACIER_MELT][o] = gm_melt_m3s;
            if (resp[HR_AE_OUTPUT]) resp[HR_AE_OUTPUT][o] = ae;
            if (resp[HR_PE_OUTPUT]) resp[HR_PE_OUTPUT][o] = pot;
        }
    }
}
}  // namespace pt_hps_k
}  // namespace sho
