// region_model.hpp -- C++ host shim over the C ABI (include/shyft_b200.h) that keeps the member surface of the reference's
// region_model<cell_t, region_env_t> for the hot path (core/region_model.h:211-1049), so that code written against
//   model.run_interpolation(ip, ta, env); model.revert_to_initial_state(); model.run_cells(); model.catchment_discharges(cr);
// (shyft/orchestration/simulator.py:125-135 through api/boostpython/expose.h:251-290; core/model_calibration.h:830-834) keeps
// its shape.  Errors surface as std::runtime_error carrying the reference's message text.  Header-only, C++17, no CUDA headers.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../shyft_b200.h"

namespace shyft_b200 {

using utctime = int64_t;      // microseconds since epoch (core/utctime_utilities.h:28-34)
using utctimespan = int64_t;

struct fixed_dt {  // core/time_axis.h:74-115
    utctime t = 0;
    utctimespan dt = 0;
    size_t n = 0;
    size_t size() const { return n; }
    utctime time(size_t i) const { return t + utctimespan(i) * dt; }
};

using geo_cell_data = sb2_geo_cell;                       // core/geo_cell_data.h:107-138, flattened
using interpolation_parameter = sb2_interpolation_parameter;  // core/region_model.h:65-95
inline interpolation_parameter default_interpolation_parameter() { interpolation_parameter ip; sb2_interpolation_parameter_default(&ip); return ip; }

// one variable of a_region_environment (api/api.h:137-168): geo-located series already on the model axis
struct geo_point_sources {
    std::vector<double> xyz;     // [n_sources][3]
    std::vector<double> values;  // [n_steps][n_sources]
    size_t size() const { return xyz.size() / 3; }
};
struct region_environment {
    geo_point_sources temperature, precipitation, radiation, wind_speed, rel_hum;
    const geo_point_sources& of(int var) const {
        switch (var) { case SB2_TEMPERATURE: return temperature; case SB2_PRECIPITATION: return precipitation; case SB2_RADIATION: return radiation;
                       case SB2_WIND_SPEED: return wind_speed; default: return rel_hum; }
    }
};

struct q_adjust_result {  // core/model_state_tuning.h:11-16
    double q_0 = 0.0, q_r = 0.0;
    std::string diagnostics;
};

template <int STACK>
class region_model {
    sb2_model* h_ = nullptr;
    void ck(int rc) const { if (rc != 0) throw std::runtime_error(sb2_last_error(h_)); }

  public:
    using parameter_t = std::vector<double>;            // parameter::get(i) order, size = parameter::size()
    using state_t = std::vector<double>;                // one cell's state, flat
    fixed_dt time_axis;
    interpolation_parameter ip_parameter = default_interpolation_parameter();

    // region_model(const std::vector<geo_cell_data>&, const parameter_t&)  (:283-291)
    region_model(const std::vector<geo_cell_data>& geov, const parameter_t& region_param, int device = 0) {
        if (sb2_model_create(STACK, int64_t(geov.size()), geov.data(), device, &h_) != 0) throw std::runtime_error(sb2_last_error(nullptr));
        set_region_parameter(region_param);
    }
    ~region_model() { sb2_model_destroy(h_); }
    region_model(const region_model&) = delete;
    region_model& operator=(const region_model&) = delete;
    sb2_model* handle() const { return h_; }

    size_t size() const { return size_t(sb2_size(h_)); }                                    // :860
    size_t number_of_catchments() const { return size_t(sb2_number_of_catchments(h_)); }    // :318
    std::vector<int64_t> catchment_ids() const { std::vector<int64_t> v(number_of_catchments()); ck(sb2_catchment_ids(h_, v.data())); return v; }

    void set_region_parameter(const parameter_t& p) { ck(sb2_set_region_parameter(h_, p.data(), int(p.size()))); }                       // :646-655
    parameter_t get_region_parameter() const { parameter_t p(size_t(sb2_parameter_size(h_))); ck(sb2_get_region_parameter(h_, p.data(), int(p.size()))); return p; }
    void set_catchment_parameter(int64_t cid, const parameter_t& p) { ck(sb2_set_catchment_parameter(h_, cid, p.data(), int(p.size()))); } // :668-678
    void remove_catchment_parameter(int64_t cid) { ck(sb2_remove_catchment_parameter(h_, cid)); }
    bool has_catchment_parameter(int64_t cid) const { return sb2_has_catchment_parameter(h_, cid) != 0; }
    parameter_t get_catchment_parameter(int64_t cid) const { parameter_t p(size_t(sb2_parameter_size(h_))); ck(sb2_get_catchment_parameter(h_, cid, p.data(), int(p.size()))); return p; }
    void set_catchment_calculation_filter(const std::vector<int64_t>& cids) { ck(sb2_set_catchment_calculation_filter(h_, cids.data(), int(cids.size()))); }  // :715-729

    // states: vector<state_t> flattened [cell][state_size]
    void set_states(const std::vector<double>& states) {                                                                                       // :802-809
        const size_t ss = size_t(sb2_state_size(h_));
        if (states.size() % ss != 0) throw std::runtime_error("set_states: the flattened state vector is not a whole number of cell states");
        ck(sb2_set_states(h_, states.data(), int64_t(states.size() / ss)));  // a wrong cell count is the reference's "Length of the state vector ..." error
    }
    void get_states(std::vector<double>& states) const { states.resize(size() * size_t(sb2_state_size(h_))); ck(sb2_get_states(h_, states.data(), int64_t(size()))); }
    void revert_to_initial_state() { ck(sb2_revert_to_initial_state(h_)); }                                                                   // :814-818
    void adjust_q(double q_scale, const std::vector<int64_t>& cids) { ck(sb2_adjust_q(h_, q_scale, cids.data(), int(cids.size()))); }          // :831-837
    void set_collector_mode(int bits) { ck(sb2_set_collector_mode(h_, bits)); }
    // state_io_handler (api/api_state.h:93-142): ids and flattened states of the cells of `cids`; apply returns the unmatched positions
    void extract_state(const std::vector<int64_t>& cids, std::vector<sb2_cell_state_id>& ids, std::vector<double>& states) const {
        ids.resize(size());
        states.resize(size() * size_t(sb2_state_size(h_)));
        int64_t n = 0;
        ck(sb2_extract_state(h_, cids.data(), int(cids.size()), ids.data(), states.data(), &n));
        ids.resize(size_t(n));
        states.resize(size_t(n) * size_t(sb2_state_size(h_)));
    }
    std::vector<int64_t> apply_state(const std::vector<sb2_cell_state_id>& ids, const std::vector<double>& states, const std::vector<int64_t>& cids) {
        std::vector<int64_t> missing(ids.size() + 1);
        int64_t n = 0;
        if (states.size() != ids.size() * size_t(sb2_state_size(h_))) throw std::runtime_error("apply_state: states must be [ids][state_size]");
        ck(sb2_apply_state(h_, int64_t(ids.size()), ids.data(), states.data(), cids.data(), int(cids.size()), missing.data(), &n));
        missing.resize(size_t(n));
        return missing;
    }
    // :626-637; q_adjust_result of core/model_state_tuning.h:11-16
    q_adjust_result adjust_state_to_target_flow(double wanted_flow_m3s, const std::vector<int64_t>& cids, size_t start_step = 0, double scale_range = 3.0,
                                                double scale_eps = 1e-3, size_t max_iter = 300, size_t n_steps = 1) {
        sb2_q_adjust_result r{};
        ck(sb2_adjust_state_to_target_flow(h_, wanted_flow_m3s, cids.data(), int(cids.size()), int64_t(start_step), scale_range, scale_eps,
                                           int64_t(max_iter), int64_t(n_steps), &r));
        return q_adjust_result{r.q_0, r.q_r, std::string(r.diagnostics)};
    }

    void initialize_cell_environment(const fixed_dt& ta) { ck(sb2_initialize_cell_environment(h_, ta.t, ta.dt, int64_t(ta.n))); time_axis = ta; }  // :359-364
    bool interpolate(const interpolation_parameter& ip, const region_environment& env, bool best_effort = true) {                                   // :397-527
        for (int v = 0; v < SB2_N_FORCING; ++v) {
            const geo_point_sources& s = env.of(v);
            // the C ABI takes no element counts: a mis-shaped vector would be over-read during the upload
            if (s.xyz.size() % 3 != 0) throw std::runtime_error("interpolate: source xyz must be [n_sources][3]");
            if (s.values.size() != time_axis.n * s.size()) throw std::runtime_error("interpolate: source values must be [n_steps][n_sources]");
            ck(sb2_set_sources(h_, v, int64_t(s.size()), s.size() ? s.xyz.data() : nullptr, s.size() ? s.values.data() : nullptr));
        }
        int ok = 0;
        ck(sb2_interpolate(h_, &ip, best_effort ? 1 : 0, &ok));
        ip_parameter = ip;
        return ok != 0;
    }
    // a variable's series on their own point axis (any resolution), projected onto the model axis on the device (:426-438)
    void set_sources_on_axis(int var, const std::vector<double>& xyz, const std::vector<int64_t>& t_us, int64_t t_end_us,
                             const std::vector<double>& values /* [points][sources] */, int point_interpretation) {
        if (xyz.size() % 3 != 0 || values.size() != t_us.size() * (xyz.size() / 3))
            throw std::runtime_error("set_sources_on_axis: xyz must be [n_sources][3] and values [n_points][n_sources]");
        ck(sb2_set_sources_on_axis(h_, var, int64_t(xyz.size() / 3), xyz.data(), int64_t(t_us.size()), t_us.data(), t_end_us, values.data(),
                                   point_interpretation));
    }
    bool run_interpolation(const interpolation_parameter& ip, const fixed_dt& ta, const region_environment& env, bool best_effort = true) {         // :546-549
        initialize_cell_environment(ta);
        return interpolate(ip, env, best_effort);
    }
    bool is_cell_env_ts_ok() { int ok = 0; ck(sb2_is_cell_env_ts_ok(h_, &ok)); return ok != 0; }                                                  // :954-962

    void run_cells(size_t /*use_ncore*/ = 0, int start_step = 0, int n_steps = 0) { ck(sb2_run_cells(h_, start_step, n_steps)); }                   // :578-597

    // catchment_discharges(TSV&) (:873-885): cr[cix] = the catchment's series over the whole axis
    void catchment_discharges(std::vector<std::vector<double>>& cr) const { fetch(cr, sb2_catchment_discharges); }
    void catchment_charges(std::vector<std::vector<double>>& cr) const { fetch(cr, sb2_catchment_charges); }
    // cell.rc.<series> of every cell, [step][cell]
    std::vector<double> response(int series) const {
        std::vector<double> v(time_axis.n * size());
        ck(sb2_get_response(h_, series, 0, int64_t(time_axis.n), v.data(), SB2_TIME_MAJOR));
        return v;
    }
    // statistics readers (api/api.h:178-1600 over core/cell_model.h:194-406): kind / series / op / scope are the SB2_STAT_* / SB2_SCOPE_*
    // constants, e.g. basic_cell_statistics::discharge(cids) = statistics(SB2_STAT_RESPONSE, SB2_R_AVG_DISCHARGE, cids, SB2_STAT_SUM)
    std::vector<double> statistics(int kind, int series, const std::vector<int64_t>& indexes, int op, int scope = SB2_SCOPE_CATCHMENT_IX) const {
        std::vector<double> v(time_axis.n + ((kind == SB2_STAT_STATE || kind == SB2_STAT_AE_POT_RATIO) ? 1 : 0));
        ck(sb2_statistics_series(h_, kind, series, indexes.data(), int(indexes.size()), scope, op, 0, int64_t(v.size()), v.data()));
        return v;
    }
    double statistics_value(int kind, int series, const std::vector<int64_t>& indexes, size_t ith_timestep, int op,
                            int scope = SB2_SCOPE_CATCHMENT_IX) const {
        double v = 0.0;
        ck(sb2_statistics_series(h_, kind, series, indexes.data(), int(indexes.size()), scope, op == SB2_STAT_AREA_AVERAGE ? SB2_STAT_AREA_AVERAGE_VALUE : op,
                                 int64_t(ith_timestep), 1, &v));
        return v;
    }
    std::vector<double> statistics_cells(int kind, int series, const std::vector<int64_t>& indexes, size_t ith_timestep,
                                         int scope = SB2_SCOPE_CATCHMENT_IX) const {
        std::vector<double> v(size());
        int64_t n = 0;
        ck(sb2_statistics_cells(h_, kind, series, indexes.data(), int(indexes.size()), scope, int64_t(ith_timestep), v.data(), &n));
        v.resize(size_t(n));
        return v;
    }
    std::vector<double> river_output_flow_m3s(int64_t rid) { std::vector<double> v(time_axis.n); ck(sb2_river_flows(h_, rid, 0, int64_t(time_axis.n), nullptr, nullptr, v.data())); return v; }  // :926-933
    void set_river_network(const std::vector<double>& rivers6) {
        if (rivers6.size() % 6 != 0) throw std::runtime_error("set_river_network: rivers must be [n][6]");
        ck(sb2_set_river_network(h_, int64_t(rivers6.size() / 6), rivers6.data()));
    }

  private:
    template <class F>
    void fetch(std::vector<std::vector<double>>& cr, F f) const {
        const size_t nc = number_of_catchments(), T = time_axis.n;
        std::vector<double> flat(T * nc);
        ck(f(h_, 0, int64_t(T), flat.data()));
        cr.assign(nc, std::vector<double>(T));
        for (size_t t = 0; t < T; ++t)
            for (size_t k = 0; k < nc; ++k) cr[k][t] = flat[t * nc + k];
    }
};

using pt_gs_k_region_model = region_model<SB2_PT_GS_K>;
using pt_hs_k_region_model = region_model<SB2_PT_HS_K>;
using hbv_stack_region_model = region_model<SB2_HBV_STACK>;

}  // namespace shyft_b200
