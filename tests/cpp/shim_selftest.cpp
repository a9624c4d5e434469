// Self test of the C++ host shim: the reference's own 20 cells x 240 h region (shyft/tests/api/test_region_model_stacks.py:145-304)
// written the way the reference's C++ tests drive a region model (test/region_model_test.cpp:44-177), with the reference's literals.
#include <cmath>
#include <cstdio>
#include <stdexcept>

#include "shyft_b200/region_model.hpp"

using namespace shyft_b200;

static int fails = 0;
#define CHECK_NEAR(a, b, tol) do { double a_ = (a), b_ = (b); if (!(std::fabs(a_ - b_) <= (tol))) { std::printf("FAIL %s:%d %s = %.17g, expected %.17g +- %g\n", __FILE__, __LINE__, #a, a_, b_, double(tol)); ++fails; } } while (0)
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

int main() {
    const int n = 20, T = 240;
    std::vector<geo_cell_data> cells(n);
    for (int i = 0; i < n; ++i) cells[i] = geo_cell_data{500 + 1000.0 * i, 500.0, 500.0 * i / n, 1000.0 * 1000.0, 1, 0.9, 0.01, 0.05, 0.19, 0.3, 1, 0.0};
    pt_gs_k_region_model::parameter_t p{-2.439, 0.966, -0.10, 1.5, -0.5, 2.0, 0.1, 1.0, 5.0, 5.0, 30.0, 0.9, 0.6, 5.0, 0.4, 0.4, 1.0, 0.1, 0.0001, 0.2, 1.26,
                                        0.04,   100.0, 0.0,  6.0, 1.0,  7.0, 0.0, 221.0, 0.0, 1.0};
    try {
        pt_gs_k_region_model model(cells, p);
        CHECK(model.size() == size_t(n) && model.number_of_catchments() == 1);
        fixed_dt ta{1420070400LL * 1000000LL, 3600LL * 1000000LL, size_t(T)};  // 2015-01-01, hourly
        region_environment env;
        auto one = [&](double v) { geo_point_sources s; s.xyz = {cells[n / 2].x, cells[n / 2].y, cells[n / 2].z}; s.values.assign(T, v); return s; };
        env.temperature = one(10.0); env.precipitation = one(5.0); env.radiation = one(300.0); env.wind_speed = one(2.0); env.rel_hum = one(0.7);
        auto ip = default_interpolation_parameter();
        ip.use_idw_for_temperature = 1;
        CHECK(model.run_interpolation(ip, ta, env));
        CHECK(model.is_cell_env_ts_ok());
        std::vector<double> s0(size_t(n) * 9);
        for (int i = 0; i < n; ++i) { double st[9] = {0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, 40.0}; for (int k = 0; k < 9; ++k) s0[size_t(i) * 9 + k] = st[k]; }
        model.set_states(s0);
        model.set_collector_mode(SB2_COLLECT_ALL | SB2_COLLECT_STATE);
        model.run_cells();
        auto charge = model.response(SB2_R_CHARGE_M3S);
        double c0 = 0; for (int i = 0; i < n; ++i) c0 += charge[i];
        CHECK_NEAR(c0, -110.6998, 1e-4);
        CHECK_NEAR(charge[0] + charge[1] + charge[3], -16.7138, 1e-4);
        auto ae = model.response(SB2_R_AE_OUTPUT);
        double ae_max = 0; for (int t = 0; t < T; ++t) { double s = 0; for (int i = 0; i < n; ++i) s += ae[size_t(t) * n + i]; ae_max = std::fmax(ae_max, s / n); }
        CHECK_NEAR(ae_max, 0.189214067680088, 1e-12);
        std::vector<std::vector<double>> cr;
        model.catchment_discharges(cr);
        CHECK(cr.size() == 1 && cr[0].size() == size_t(T) && cr[0][0] >= 130.0);
        model.set_river_network({1, 0, 3000.0, 1 / 3.60, 7.0, 0.0});
        CHECK_NEAR(model.river_output_flow_m3s(1)[8], 28.061248025828114, 1e-10);
        // chunked run equals one-shot (region_model.h:573-575)
        model.revert_to_initial_state();
        for (int k = 0; k < 10; ++k) model.run_cells(0, 24 * k, 24);
        std::vector<std::vector<double>> cr2;
        model.catchment_discharges(cr2);
        CHECK(cr2[0] == cr[0]);
        // state tuning, the reference's Python sequence (test_region_model_stacks.py:317-326) through the shim
        model.revert_to_initial_state();
        model.run_cells(0, 10, 2);
        const std::vector<int64_t> all_cids;
        const double q_avg = (model.statistics_value(SB2_STAT_RESPONSE, SB2_R_AVG_DISCHARGE, all_cids, 10, SB2_STAT_SUM) +
                              model.statistics_value(SB2_STAT_RESPONSE, SB2_R_AVG_DISCHARGE, all_cids, 11, SB2_STAT_SUM)) / 2.0;
        model.revert_to_initial_state();
        const q_adjust_result adj = model.adjust_state_to_target_flow(0.7 * q_avg, all_cids, 10, 3.0, 1e-3, 350, 2);
        CHECK(adj.diagnostics.empty());
        CHECK_NEAR(adj.q_r, 0.7 * q_avg, 0.005);
        CHECK_NEAR(adj.q_0, q_avg, 0.005);
        // cell-identified state (api/api_state.h:93-142): extract, modify, apply, extract again
        std::vector<sb2_cell_state_id> ids;
        std::vector<double> st;
        model.extract_state(all_cids, ids, st);
        CHECK(ids.size() == size_t(n) && ids[3].cid == 1 && ids[3].x == 3500 && ids[3].y == 500 && ids[3].area == 1000000);
        for (int i = 0; i < n; ++i) st[size_t(i) * 9 + 8] = 100.0 + i;
        ids[5].x += 1;  // no such cell
        const std::vector<int64_t> missing = model.apply_state(ids, st, all_cids);
        CHECK(missing.size() == 1 && missing[0] == 5);
        std::vector<double> back;
        model.get_states(back);
        CHECK(back[4 * 9 + 8] == 104.0 && back[5 * 9 + 8] != 105.0 && back[19 * 9 + 8] == 119.0);
        // argument validation with the reference's messages (:586-592)
        bool threw = false;
        try { model.run_cells(0, T, 1); } catch (const std::runtime_error& e) { threw = std::string(e.what()).find("start_step must in range") != std::string::npos; }
        CHECK(threw);
    } catch (const std::exception& e) {
        std::printf("FAIL exception: %s\n", e.what());
        ++fails;
    }
    std::printf(fails ? "shim selftest: %d failure(s)\n" : "shim selftest ok\n", fails);
    return fails ? 1 : 0;
}
