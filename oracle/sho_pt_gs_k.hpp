// ============================================================================
// oracle/sho_pt_gs_k.hpp -- CPU ORACLE (test infrastructure, NOT product code)
// pt_gs_k stack: parameter / state / response, run_pt_gs_k, collectors.
// Follows core/pt_gs_k.h:39-398 and core/pt_gs_k_cell_model.h:41-262.
// ============================================================================
#pragma once
#include "sho_core.hpp"

namespace sho {
namespace pt_gs_k {

struct parameter {  // core/pt_gs_k.h:39-194
    priestley_taylor::parameter pt;
    gamma_snow::parameter gs;
    actual_evapotranspiration::parameter ae;
    kirchner::parameter kirchner;
    precipitation_correction::parameter p_corr;
    glacier_melt::parameter gm;
    struct { double velocity = 1.0, alpha = 7.0, beta = 0.0; } routing;  // routing::uhg_parameter defaults, core/routing.h:30-35   // routing::uhg_parameter, core/routing.h
    struct { double reservoir_direct_response_fraction = 1.0; } msp;      // core/mstack_param.h
    static constexpr size_t n_params = 31;
    size_t size() const { return n_params; }
    void set(const double* p) {  // :77-112
        int i = 0;
        kirchner.c1 = p[i++]; kirchner.c2 = p[i++]; kirchner.c3 = p[i++];
        ae.ae_scale_factor = p[i++];
        gs.tx = p[i++]; gs.wind_scale = p[i++]; gs.max_water = p[i++]; gs.wind_const = p[i++];
        gs.fast_albedo_decay_rate = p[i++]; gs.slow_albedo_decay_rate = p[i++]; gs.surface_magnitude = p[i++];
        gs.max_albedo = p[i++]; gs.min_albedo = p[i++]; gs.snowfall_reset_depth = p[i++]; gs.snow_cv = p[i++];
        gs.glacier_albedo = p[i++];
        p_corr.scale_factor = p[i++];
        gs.snow_cv_forest_factor = p[i++]; gs.snow_cv_altitude_factor = p[i++];
        pt.albedo = p[i++]; pt.alpha = p[i++];
        gs.initial_bare_ground_fraction = p[i++];
        gs.winter_end_day_of_year = size_t(p[i++]);
        gs.calculate_iso_pot_energy = p[i++] != 0.0 ? true : false;
        gm.dtf = p[i++];
        routing.velocity = p[i++]; routing.alpha = p[i++]; routing.beta = p[i++];
        gs.n_winter_days = size_t(p[i++]);
        gm.direct_response = p[i++];
        msp.reservoir_direct_response_fraction = p[i++];
    }
    double get(size_t i) const {  // :115-153
        switch (i) {
            case 0: return kirchner.c1; case 1: return kirchner.c2; case 2: return kirchner.c3;
            case 3: return ae.ae_scale_factor; case 4: return gs.tx; case 5: return gs.wind_scale;
            case 6: return gs.max_water; case 7: return gs.wind_const; case 8: return gs.fast_albedo_decay_rate;
            case 9: return gs.slow_albedo_decay_rate; case 10: return gs.surface_magnitude; case 11: return gs.max_albedo;
            case 12: return gs.min_albedo; case 13: return gs.snowfall_reset_depth; case 14: return gs.snow_cv;
            case 15: return gs.glacier_albedo; case 16: return p_corr.scale_factor; case 17: return gs.snow_cv_forest_factor;
            case 18: return gs.snow_cv_altitude_factor; case 19: return pt.albedo; case 20: return pt.alpha;
            case 21: return gs.initial_bare_ground_fraction; case 22: return double(gs.winter_end_day_of_year);
            case 23: return gs.calculate_iso_pot_energy ? 1.0 : 0.0; case 24: return gm.dtf;
            case 25: return routing.velocity; case 26: return routing.alpha; case 27: return routing.beta;
            case 28: return double(gs.n_winter_days); case 29: return gm.direct_response;
            case 30: return msp.reservoir_direct_response_fraction;
            default: throw std::runtime_error("PTGSK Parameter Accessor:.get(i) Out of range.");
        }
    }
};

struct state {  // :204-226 ; memory order = 8 gamma_snow doubles then kirchner.q
    gamma_snow::state gs;
    double kirchner_q = 0.1;  // kirchner::state, core/kirchner.h:124-126
    static constexpr size_t n_state = 9;
};

struct response {  // :233-255
    double pt_pot_evapotranspiration = 0.0;
    gamma_snow::response gs;
    double ae = 0.0;
    double kirchner_q_avg = 0.0;
    double gm_melt_m3s = 0.0;
    double total_discharge = 0.0;
    double charge_m3s = 0.0;
};

// Collector output block.  Any pointer may be null (series not collected).  Element (cell c, step i) of
// a response series lives at p[c*cell_stride + i*time_stride]; state series have n+1 points per run.
// Response ids follow all_response_collector (pt_gs_k_cell_model.h:41-94), state ids follow
// state_collector (:146-208).
enum response_id { R_AVG_DISCHARGE = 0, R_CHARGE_M3S, R_SNOW_SCA, R_SNOW_SWE, R_SNOW_OUTFLOW, R_GLACIER_MELT, R_AE_OUTPUT, R_PE_OUTPUT, N_RESPONSE };
enum state_id { S_KIRCHNER_DISCHARGE = 0, S_GS_ALBEDO, S_GS_LWC, S_GS_SURFACE_HEAT, S_GS_ALPHA, S_GS_SDC_MELT_MEAN, S_GS_ACC_MELT, S_GS_ISO_POT_ENERGY, S_GS_TEMP_SWE, N_STATE_SERIES };

struct collectors {
    double* resp[N_RESPONSE] = {nullptr};
    double* st[N_STATE_SERIES] = {nullptr};
    int64_t cell_stride = 0, time_stride = 0;
    double* kirchner_substeps = nullptr;  // optional diagnostics: accepted+rejected try_steps per (cell, step)
};

// core/pt_gs_k.h:312-398.  forcing element (step i) of variable v: f[v][i*fstride]
struct cell_forcing {
    const double* temp; const double* prec; const double* wind_speed; const double* rel_hum; const double* rad;
    int64_t stride;
};

inline void run_pt_gs_k(const geo_cell& geo, const parameter& parameter, const fixed_dt& time_axis, int start_step, int n_steps,
                        const cell_forcing& f, state& st, const collectors& col, size_t cell_index, response* end_response = nullptr) {
    priestley_taylor::calculator pt(parameter.pt.albedo, parameter.pt.alpha);
    gamma_snow::calculator gs;
    kirchner::calculator kirchner(parameter.kirchner);
    response response;
    const double forest_fraction = geo.forest;
    const double glacier_fraction = geo.glacier;
    const double gm_direct = parameter.gm.direct_response;
    const double gm_routed = 1 - gm_direct;
    const double snow_storage_fraction = geo.snow_storage();
    const double kirchner_routed_prec = geo.reservoir * (1.0 - parameter.msp.reservoir_direct_response_fraction) + geo.lake;
    const double direct_response_fraction = glacier_fraction * gm_direct + geo.reservoir * parameter.msp.reservoir_direct_response_fraction;
    const double kirchner_fraction = 1 - direct_response_fraction;
    const double cell_area_m2 = geo.area;
    const double glacier_area_m2 = geo.area * glacier_fraction;
    const double altitude = geo.z;
    const int64_t cs = col.cell_stride * int64_t(cell_index), ts = col.time_stride;

    auto collect_state = [&](size_t i) {  // state.scale_snow (:213-218) + state_collector::collect (pt_gs_k_cell_model.h:193-205)
        if (!col.st[S_KIRCHNER_DISCHARGE]) return;
        const int64_t o = cs + int64_t(i) * ts;
        col.st[S_KIRCHNER_DISCHARGE][o] = mmh_to_m3s(st.kirchner_q, cell_area_m2);
        col.st[S_GS_ALBEDO][o] = st.gs.albedo;
        col.st[S_GS_LWC][o] = st.gs.lwc * snow_storage_fraction;
        col.st[S_GS_SURFACE_HEAT][o] = st.gs.surface_heat;
        col.st[S_GS_ALPHA][o] = st.gs.alpha;
        col.st[S_GS_SDC_MELT_MEAN][o] = st.gs.sdc_melt_mean;
        col.st[S_GS_ACC_MELT][o] = st.gs.acc_melt;
        col.st[S_GS_ISO_POT_ENERGY][o] = st.gs.iso_pot_energy;
        col.st[S_GS_TEMP_SWE][o] = st.gs.temp_swe * snow_storage_fraction;
    };

    size_t i_begin = n_steps > 0 ? start_step : 0;
    size_t i_end = n_steps > 0 ? start_step + n_steps : time_axis.size();
    for (size_t i = i_begin; i < i_end; ++i) {
        const utctime p_start = time_axis.time(i), p_end = p_start + time_axis.dt;
        const utctimespan timespan = p_end - p_start;
        double temp = f.temp[int64_t(i) * f.stride];
        double rad = f.rad[int64_t(i) * f.stride];
        double rel_hum = f.rel_hum[int64_t(i) * f.stride];
        double prec = f.prec[int64_t(i) * f.stride] * parameter.p_corr.scale_factor;  // precipitation_correction.h:39-41
        collect_state(i);

#ifdef SHO_COUNT
        const dm::counters_t cnt0 = dm::g_cnt;
#endif
        gs.step(st.gs, response.gs, p_start, timespan, parameter.gs, temp, rad, prec, f.wind_speed[int64_t(i) * f.stride], rel_hum,
                forest_fraction, altitude);
#ifdef SHO_COUNT
        if (dm::g_cost_cursor) {
            auto d = [&](int k) { const long long v = dm::g_cnt.v[k] - cnt0.v[k]; return (unsigned char)(v > 255 ? 255 : v); };
            dm::g_cost_cursor[0] = d(dm::C_GSER_ITER); dm::g_cost_cursor[1] = d(dm::C_GCF_ITER);
            dm::g_cost_cursor[2] = (unsigned char)(d(dm::C_EXP) + d(dm::C_LOG)); dm::g_cost_cursor[3] = d(dm::C_BRENT_EVAL);
            dm::g_cost_cursor += 4;
        }
#endif
        response.gm_melt_m3s = glacier_melt::step(parameter.gm.dtf, temp, cell_area_m2 * response.gs.sca, glacier_area_m2);
        response.pt_pot_evapotranspiration = pt.potential_evapotranspiration(temp, rad, rel_hum) * to_seconds(HOUR);
        response.ae = actual_evapotranspiration::calculate_step(st.kirchner_q, response.pt_pot_evapotranspiration,
                                                                parameter.ae.ae_scale_factor,
                                                                std::max(response.gs.sca, glacier_fraction), timespan);
        double gm_mmh = m3s_to_mmh(response.gm_melt_m3s, cell_area_m2);
        kirchner::step_stats kst;
        kirchner.step(p_start, p_end, st.kirchner_q, response.kirchner_q_avg,
                      response.gs.outflow * snow_storage_fraction + prec * kirchner_routed_prec + gm_routed * gm_mmh, response.ae,
                      col.kirchner_substeps ? &kst : nullptr);

        response.total_discharge = std::max(0.0, prec - response.ae) * direct_response_fraction + gm_direct * gm_mmh +
                                   response.kirchner_q_avg * kirchner_fraction;
        response.charge_m3s = +mmh_to_m3s(prec, cell_area_m2) - mmh_to_m3s(response.ae, cell_area_m2) + response.gm_melt_m3s -
                              mmh_to_m3s(response.total_discharge, cell_area_m2);
        {  // response.scale_snow (:248-254) + all_response_collector::collect (pt_gs_k_cell_model.h:81-90)
            const int64_t o = cs + int64_t(i) * ts;
            if (col.resp[R_AVG_DISCHARGE]) col.resp[R_AVG_DISCHARGE][o] = mmh_to_m3s(response.total_discharge, cell_area_m2);
            if (col.resp[R_CHARGE_M3S]) col.resp[R_CHARGE_M3S][o] = response.charge_m3s;
            if (col.resp[R_SNOW_SCA]) col.resp[R_SNOW_SCA][o] = response.gs.sca;
            if (col.resp[R_SNOW_OUTFLOW]) col.resp[R_SNOW_OUTFLOW][o] = mmh_to_m3s(response.gs.outflow * snow_storage_fraction, cell_area_m2);
            if (col.resp[R_GLACIER_MELT]) col.resp[R_GLACIER_MELT][o] = response.gm_melt_m3s;
            if (col.resp[R_SNOW_SWE]) col.resp[R_SNOW_SWE][o] = response.gs.storage * snow_storage_fraction;
            if (col.resp[R_AE_OUTPUT]) col.resp[R_AE_OUTPUT][o] = response.ae;
            if (col.resp[R_PE_OUTPUT]) col.resp[R_PE_OUTPUT][o] = response.pt_pot_evapotranspiration;
            if (col.kirchner_substeps) col.kirchner_substeps[o] = double(kst.accepted + kst.rejected);
        }
        if (i + 1 == i_end) collect_state(i + 1);
    }
    if (end_response) *end_response = response;
}

}  // namespace pt_gs_k
}  // namespace sho
