// ============================================================================
// oracle/sho_region.hpp -- CPU ORACLE (test infrastructure, NOT product code)
// Region level: IDW and BTK interpolation, routing, goal functions, the
// threaded run_cells work queue.  Follows core/inverse_distance.h:142-472,
// core/bayesian_kriging.h:41-402, core/routing.h:119-421,
// core/time_series.h:966-975,2316-2448, core/region_model.h:233-249,397-527,
// 578-597,873-900,991-1021.
// ============================================================================
#pragma once
#include <atomic>
#include <map>
#include <mutex>
#include <thread>

#include "sho_core.hpp"

namespace sho {

struct geo_point { double x = 0, y = 0, z = 0; };

// core/geo_point.h:41-43
inline double distance_measure(const geo_point& a, const geo_point& b, double p, double zscale) {
    return dm::pow((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z) * zscale * zscale, p / 2.0);
}
// core/geo_point.h:49-51
inline double zscaled_distance(const geo_point& a, const geo_point& b, double zscale) {
    return std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z) * zscale * zscale);
}

// average_accessor of a stair-case source that already sits on the model axis
// (core/time_series.h:202-310,2033-2072): area = to_seconds(dt)*v; value = area/to_seconds(tsum).
inline double average_accessor_same_axis(double v, utctimespan dt) {
    if (!std::isfinite(v)) return nan_v;
    return (to_seconds(dt) * v) / to_seconds(dt);
}

// ---------------------------------------------------------------------------
// inverse_distance  (core/inverse_distance.h)
// ---------------------------------------------------------------------------
namespace idw {
enum model_kind { TEMPERATURE = 0, PRECIPITATION = 1, RADIATION = 2, WIND_SPEED = 3, REL_HUM = 4 };

struct parameter {  // :38-74 (the union of the three parameter structs)
    size_t max_members = 10;
    double max_distance = 200000.0;
    double distance_measure_factor = 2.0;
    double zscale = 1.0;
    double default_temp_gradient = -0.006;  // temperature_parameter
    bool gradient_by_equation = false;      // temperature_parameter
    double scale_factor = 1.02;             // precipitation_parameter
};

struct source_weight { int source; double weight; };

// step 1 of run_interpolation (:160-203): per destination the reachable sources, at most max_members
inline std::vector<std::vector<source_weight>> build_neighbours(const std::vector<geo_point>& src, const std::vector<geo_point>& dst,
                                                               const parameter& p) {
    const double max_weight = 1.0;
    geo_point o{0, 0, 0}, m{p.max_distance, 0, 0};
    const double min_weight = 1.0 / distance_measure(o, m, p.distance_measure_factor, p.zscale);
    std::vector<std::vector<source_weight>> out;
    out.reserve(dst.size());
    std::vector<source_weight> swl;
    for (const auto& d : dst) {
        swl.clear();
        for (size_t k = 0; k < src.size(); ++k) {
            double w = std::min(max_weight, 1.0 / distance_measure(d, src[k], p.distance_measure_factor, p.zscale));
            if (w >= min_weight) swl.push_back({int(k), w});
        }
        if (swl.size() > p.max_members) {
            std::partial_sort(swl.begin(), swl.begin() + p.max_members, swl.end(),
                              [](const source_weight& a, const source_weight& b) { return a.weight > b.weight; });
            swl.resize(p.max_members);
        }
        out.push_back(swl);
    }
    return out;
}

// 3x3 solve with partial pivoting; false when (numerically) singular.  Stands in for
// arma::solve(..., solve_opts::no_approx) at :296-300 (LAPACK dgesv; not in the tree).
inline bool solve3(double A[3][3], double b[3], double x[3]) {
    int piv[3] = {0, 1, 2};
    double scale = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(A[i][j]));
    if (!(scale > 0)) return false;
    for (int c = 0; c < 3; ++c) {
        int best = c;
        for (int r = c + 1; r < 3; ++r) if (std::fabs(A[piv[r]][c]) > std::fabs(A[piv[best]][c])) best = r;
        std::swap(piv[c], piv[best]);
        double d = A[piv[c]][c];
        if (std::fabs(d) < 1e-14 * scale) return false;
        for (int r = c + 1; r < 3; ++r) {
            double f = A[piv[r]][c] / d;
            for (int j = c; j < 3; ++j) A[piv[r]][j] -= f * A[piv[c]][j];
            b[piv[r]] -= f * b[piv[c]];
        }
    }
    for (int c = 2; c >= 0; --c) {
        double s = b[piv[c]];
        for (int j = c + 1; j < 3; ++j) s -= A[piv[c]][j] * x[j];
        x[c] = s / A[piv[c]][c];
    }
    return std::isfinite(x[0]) && std::isfinite(x[1]) && std::isfinite(x[2]);
}

// temperature_gradient_scale_computer::compute (:285-316) over the valid neighbours in list order
inline double temperature_gradient(const std::vector<geo_point>& pts, const std::vector<double>& temps, const parameter& p) {
    const double minimum_z_distance = 50.0;
    size_t n = pts.size();
    if (p.gradient_by_equation && n > 3) {
        double A[3][3], b[3], x[3];
        for (int r = 0; r < 3; ++r) {
            A[r][0] = pts[r + 1].x - pts[0].x; A[r][1] = pts[r + 1].y - pts[0].y; A[r][2] = pts[r + 1].z - pts[0].z;
            b[r] = temps[r + 1] - temps[0];
        }
        if (solve3(A, b, x)) return x[2];
    }
    if (n > 1) {
        size_t mx_i = 0, mn_i = 0;
        for (size_t i = 0; i < n; ++i) {
            double h = pts[i].z;
            if (h < pts[mn_i].z) mn_i = i;
            else if (h > pts[mx_i].z) mx_i = i;
        }
        double mi_mx_dz = pts[mx_i].z - pts[mn_i].z;
        return mi_mx_dz > minimum_z_distance ? (temps[mx_i] - temps[mn_i]) / mi_mx_dz : p.default_temp_gradient;
    }
    return p.default_temp_gradient;
}

// step 2 of run_interpolation (:214-249).  src_values element (step i, source k): v[i*n_src + k];
// output element (step i, destination j): out[i*out_tstride + j*out_cstride]
inline void run_interpolation(model_kind kind, const std::vector<geo_point>& src, const double* src_values, size_t n_steps,
                              const std::vector<geo_point>& dst, const std::vector<double>& dst_slope_factor, const parameter& p,
                              double* out, int64_t out_tstride, int64_t out_cstride, size_t j_begin = 0, size_t j_end = size_t(-1)) {
    auto nb = build_neighbours(src, dst, p);
    const size_t n_src = src.size();
    if (j_end > dst.size()) j_end = dst.size();
    std::vector<geo_point> gpts;
    std::vector<double> gt;
    for (size_t i = 0; i < n_steps; ++i) {
        const double* v = src_values + i * n_src;
        for (size_t j = j_begin; j < j_end; ++j) {
            double scale;
            if (kind == TEMPERATURE) {
                gpts.clear(); gt.clear();
                for (const auto& sw : nb[j]) if (std::isfinite(v[sw.source])) { gpts.push_back(src[sw.source]); gt.push_back(v[sw.source]); }
                scale = temperature_gradient(gpts, gt, p);
            } else if (kind == PRECIPITATION) scale = p.scale_factor;
            else scale = 1.0;
            double sum_weights = 0, sum_weight_value = 0;
            for (const auto& sw : nb[j]) {
                double sv = v[sw.source];
                if (std::isfinite(sv)) {
                    double tv;
                    switch (kind) {
                        case TEMPERATURE: tv = sv + scale * (dst[j].z - src[sw.source].z); break;                       // :367-369
                        case PRECIPITATION: tv = sv * dm::pow(scale, (dst[j].z - src[sw.source].z) / 100.0); break;   // :422-426
                        case RADIATION: tv = sv * dst_slope_factor[j]; break;                                          // :392-394
                        default: tv = sv;
                    }
                    sum_weight_value += sw.weight * tv;
                    sum_weights += sw.weight;
                }
            }
            out[int64_t(i) * out_tstride + int64_t(j) * out_cstride] = sum_weight_value / sum_weights;
        }
    }
}
}  // namespace idw

// ---------------------------------------------------------------------------
// small dense linear algebra standing in for armadillo 9.200.6 (absent): row-major matrices
// ---------------------------------------------------------------------------
namespace la {
struct mat {
    size_t r = 0, c = 0;
    std::vector<double> a;
    mat() {}
    mat(size_t r, size_t c, double v = 0.0) : r(r), c(c), a(r * c, v) {}
    double& operator()(size_t i, size_t j) { return a[i * c + j]; }
    double operator()(size_t i, size_t j) const { return a[i * c + j]; }
};
inline mat mul(const mat& A, const mat& B) {
    mat C(A.r, B.c);
    for (size_t i = 0; i < A.r; ++i)
        for (size_t k = 0; k < A.c; ++k) {
            const double aik = A(i, k);
            for (size_t j = 0; j < B.c; ++j) C(i, j) += aik * B(k, j);
        }
    return C;
}
inline mat tr(const mat& A) {
    mat T(A.c, A.r);
    for (size_t i = 0; i < A.r; ++i) for (size_t j = 0; j < A.c; ++j) T(j, i) = A(i, j);
    return T;
}
inline mat sub(const mat& A, const mat& B) { mat C = A; for (size_t i = 0; i < C.a.size(); ++i) C.a[i] -= B.a[i]; return C; }
inline mat add(const mat& A, const mat& B) { mat C = A; for (size_t i = 0; i < C.a.size(); ++i) C.a[i] += B.a[i]; return C; }
inline mat eye(size_t n) { mat I(n, n); for (size_t i = 0; i < n; ++i) I(i, i) = 1.0; return I; }
// inverse by LU with partial pivoting (what LAPACK dgetrf/dgetri does for arma::mat::i())
inline mat inv(const mat& A_) {
    const size_t n = A_.r;
    mat A = A_, I = eye(n);
    for (size_t c = 0; c < n; ++c) {
        size_t best = c;
        for (size_t r = c + 1; r < n; ++r) if (std::fabs(A(r, c)) > std::fabs(A(best, c))) best = r;
        if (A(best, c) == 0.0) throw std::runtime_error("inv(): matrix seems singular");
        if (best != c) for (size_t j = 0; j < n; ++j) { std::swap(A(c, j), A(best, j)); std::swap(I(c, j), I(best, j)); }
        const double d = A(c, c);
        for (size_t r = c + 1; r < n; ++r) {
            const double f = A(r, c) / d;
            if (f == 0.0) continue;
            for (size_t j = c; j < n; ++j) A(r, j) -= f * A(c, j);
            for (size_t j = 0; j < n; ++j) I(r, j) -= f * I(c, j);
        }
    }
    for (size_t cc = n; cc-- > 0;) {
        const double d = A(cc, cc);
        for (size_t j = 0; j < n; ++j) I(cc, j) /= d;
        for (size_t r = 0; r < cc; ++r) {
            const double f = A(r, cc);
            if (f == 0.0) continue;
            for (size_t j = 0; j < n; ++j) I(r, j) -= f * I(cc, j);
        }
    }
    return I;
}
// numerical rank of a symmetric 2x2 (arma::rank: singular values above max(dim)*s_max*eps)
inline int rank_sym22(const mat& H) {
    const double a = H(0, 0), b = 0.5 * (H(0, 1) + H(1, 0)), d = H(1, 1);
    const double tr_ = a + d, det = a * d - b * b;
    const double disc = std::sqrt(std::max(0.0, 0.25 * tr_ * tr_ - det));
    const double l1 = std::fabs(0.5 * tr_ + disc), l2 = std::fabs(0.5 * tr_ - disc);
    const double smax = std::max(l1, l2), smin = std::min(l1, l2);
    const double tolr = 2.0 * smax * std::numeric_limits<double>::epsilon();
    return int(smax > tolr) + int(smin > tolr);
}
}  // namespace la

// ---------------------------------------------------------------------------
// bayesian_kriging  (core/bayesian_kriging.h)
// ---------------------------------------------------------------------------
namespace btk {
struct parameter {  // :204-230
    double gradient_sd = 0.0025;
    double sill_value = 25.0;
    double nug_value = 0.5;
    double range_value = 200000.0;
    double zscale_value = 20.0;
    double fixed_gradient = nan_v;  // test hook: the reference's test Parameter returns a constant prior gradient (test/bayesian_kriging_test.cpp:62-84)
    double temperature_gradient(utctime p_start, utctimespan p_span) const {  // :220-223
        if (std::isfinite(fixed_gradient)) return fixed_gradient;
        const double doy = double(calendar::day_of_year(p_start + p_span / 2));
        return 1.18e-3 * std::sin(6.2831 / 365 * (doy + 79.0)) - 5.48e-3;
    }
};

// :93-125  K (n x n, source-source) and k (n x m, source-destination)
inline void build_covariance_matrices(const std::vector<geo_point>& src, const std::vector<geo_point>& dst, const parameter& p, la::mat& K,
                                      la::mat& k) {
    const size_t n = src.size(), m = dst.size();
    K = la::mat(n, n);
    for (size_t i = 0; i < n; ++i) {
        K(i, i) = 1.0 * (p.sill_value - p.nug_value);  // K.eye; K.diag() *= zero_dist_cov
        for (size_t j = i + 1; j < n; ++j) {
            const double d = zscaled_distance(src[i], src[j], p.zscale_value);
            K(i, j) = K(j, i) = (p.sill_value - p.nug_value) * dm::exp(-d / p.range_value);
        }
    }
    k = la::mat(n, m);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < m; ++j)
            k(i, j) = (p.sill_value - p.nug_value) * dm::exp(-zscaled_distance(src[i], dst[j], p.zscale_value) / p.range_value);
}

struct operators { la::mat F, E_beta_w, omega, GH_inv, BM; };

// the operator build of :300-316 (full) and :362-374 (reduced to the rows `valid`)
inline operators build_operators(const la::mat& F_full, const la::mat& f, const la::mat& K_full, const la::mat& k_full,
                                 const std::vector<size_t>& valid, const parameter& p, bool check_rank) {
    const size_t n = valid.size(), m = f.c;
    la::mat F(n, 2), K(n, n), k(n, m);
    for (size_t i = 0; i < n; ++i) {
        F(i, 0) = F_full(valid[i], 0); F(i, 1) = F_full(valid[i], 1);
        for (size_t j = 0; j < n; ++j) K(i, j) = K_full(valid[i], valid[j]);
        for (size_t j = 0; j < m; ++j) k(i, j) = k_full(valid[i], j);
    }
    la::mat K_inv = la::inv(K);
    la::mat Ft = la::tr(F);
    la::mat H_inv = la::mul(la::mul(Ft, K_inv), F);
    if (check_rank && la::rank_sym22(H_inv) == 1)
        throw std::runtime_error("The bayestian temperature kriging algorithm needs at least two sources at different heights.");
    la::mat H = la::inv(H_inv);
    la::mat G_inv = H_inv;
    G_inv(1, 1) += 1 / (p.gradient_sd * p.gradient_sd);
    la::mat G = la::inv(G_inv);
    operators o;
    o.F = F;
    o.GH_inv = la::mul(G, H_inv);
    o.BM = la::mul(la::tr(la::sub(f, la::mul(la::mul(Ft, K_inv), k))), la::sub(la::eye(2), o.GH_inv));
    o.E_beta_w = la::mul(la::mul(H, Ft), K_inv);
    o.omega = la::mul(la::tr(k), K_inv);
    return o;
}

// :280-402.  src_values element (step i, source k): v[i*n_src + k]; out as idw::run_interpolation
inline void btk_interpolation(const std::vector<geo_point>& src, const double* src_values, const fixed_dt& ta,
                              const std::vector<geo_point>& dst, const parameter& p, double* out, int64_t out_tstride, int64_t out_cstride) {
    const size_t n = src.size(), m = dst.size();
    la::mat F(n, 2), f(2, m), K, k;
    for (size_t i = 0; i < n; ++i) { F(i, 0) = 1.0; F(i, 1) = src[i].z; }  // :146-162
    for (size_t j = 0; j < m; ++j) { f(0, j) = 1.0; f(1, j) = dst[j].z; }
    build_covariance_matrices(src, dst, p, K, k);
    std::vector<size_t> all(n);
    for (size_t i = 0; i < n; ++i) all[i] = i;
    operators full = build_operators(F, f, K, k, all, p, true), red;
    const operators* op = nullptr;
    std::vector<size_t> valid, prev_valid;
    std::vector<double> temperatures;
    la::mat ft = la::tr(f);
    for (size_t t_step = 0; t_step < ta.size(); ++t_step) {
        temperatures.clear();
        prev_valid = valid;
        valid.clear();
        for (size_t s = 0; s < n; ++s) {
            double v = src_values[t_step * n + s];
            if (std::isfinite(v)) { valid.push_back(s); temperatures.push_back(v); }
        }
        if (valid != prev_valid || valid.size() == 0) {
            if (valid.size() == 0) throw std::runtime_error("bayesian kriging temperature: No valid sources for time period, giving up.");
            if (valid.size() == n) op = &full;
            else { red = build_operators(F, f, K, k, valid, p, false); op = &red; }
        }
        la::mat E_beta_pri(2, 1);
        E_beta_pri(0, 0) = 0.0;
        E_beta_pri(1, 0) = p.temperature_gradient(ta.time(t_step), ta.dt);
        la::mat T_obs(valid.size(), 1);
        for (size_t i = 0; i < valid.size(); ++i) T_obs(i, 0) = temperatures[i];
        la::mat beta_hat = la::mul(op->E_beta_w, T_obs);
        la::mat T_hat = la::add(la::mul(ft, beta_hat), la::mul(op->omega, la::sub(T_obs, la::mul(op->F, beta_hat))));
        la::mat E_temp_post = la::sub(T_hat, la::mul(op->BM, la::sub(beta_hat, E_beta_pri)));
        for (size_t j = 0; j < m; ++j) out[int64_t(t_step) * out_tstride + int64_t(j) * out_cstride] = E_temp_post(j, 0);
    }
}
}  // namespace btk

// ---------------------------------------------------------------------------
// routing  (core/routing.h) -- gamma pdf / quantile stand in for boost::math::gamma_distribution
// ---------------------------------------------------------------------------
namespace routing {
inline double gamma_pdf(double alpha, double x) {  // pdf of Gamma(shape alpha, scale 1)
    if (x < 0) return 0.0;
    if (x == 0) return 0.0;  // boost 1.68 gamma.hpp pdf(): "if(x == 0) return 0;" for every shape
    return dm::exp((alpha - 1.0) * dm::log(x) - x - dm::lgamma(alpha));
}
inline double gamma_quantile(double alpha, double pq) {  // inverse of P(alpha, x) by bracketed Newton on full-double P
    double lo = 0.0, hi = std::max(1.0, alpha);
    while (special::gamma_p(alpha, hi) < pq) hi *= 2.0;
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 200; ++it) {
        double fx = special::gamma_p(alpha, x) - pq;
        if (fx > 0) hi = x; else lo = x;
        double d = gamma_pdf(alpha, x);
        double xn = d > 0 ? x - fx / d : 0.5 * (lo + hi);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (std::fabs(xn - x) <= 1e-15 * std::fabs(x)) { x = xn; break; }
        x = xn;
    }
    return x;
}
// :399-421
inline std::vector<double> make_uhg_from_gamma(int n_steps, double alpha, double base) {
    std::vector<double> r;
    if (n_steps > 1) {
        double s = 0.0;
        double x_max = gamma_quantile(alpha, 0.99);
        double d = x_max / double(n_steps);
        for (int i = 0; i < n_steps; ++i) {
            double x = d * i;
            double y = std::max(0.0, gamma_pdf(alpha, x) + base);
            s += y;
            r.push_back(y);
        }
        if (s > 0.0) for (auto& y : r) y /= s;
        else for (auto& y : r) y = 1 / double(n_steps);
    }
    if (r.size() == 0) r.push_back(1.0);
    return r;
}
inline int uhg_steps(double distance, double velocity, utctimespan dt) {  // :119-123, :326-330
    double steps = (distance / velocity) / to_seconds(dt);
    return int(steps + 0.5);
}
// convolve_w_ts::value with convolve_policy::USE_ZERO  (core/time_series.h:966-975)
inline double convolve_value(const double* ts, int64_t stride, const std::vector<double>& w, size_t i) {
    double v = 0.0;
    for (size_t j = 0; j < w.size(); ++j) v += j <= i ? w[j] * ts[int64_t(i - j) * stride] : 0.0;
    return v;
}
struct river { int64_t id = 0; int64_t downstream_id = 0; double distance = 0.0; double velocity = 1.0, alpha = 7.0, beta = 0.0; };

struct network {
    std::map<int64_t, river> rid_map;
    std::vector<int64_t> upstreams_by_id(int64_t rid) const {  // :205-213
        std::vector<int64_t> r;
        for (const auto& kv : rid_map) if (kv.second.downstream_id == rid) r.push_back(kv.first);
        return r;
    }
};
}  // namespace routing

// ---------------------------------------------------------------------------
// goal functions  (core/time_series.h:2316-2448); KGE uses dlib::running_scalar_covariance semantics
// ---------------------------------------------------------------------------
namespace goal {
inline double nash_sutcliffe(const double* o, const double* m, size_t n) {  // :2316-2343
    if (n == 0) throw std::runtime_error("nash_sutcliffe needs equal sized ts accessors with elements >1");
    double sd2 = 0, obs_avg = 0;
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::isfinite(o[i]) && std::isfinite(m[i])) { double d = o[i] - m[i]; sd2 += d * d; obs_avg += o[i]; ++cnt; }
    obs_avg /= double(cnt);
    double so2 = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::isfinite(o[i]) && std::isfinite(m[i])) { double d = o[i] - obs_avg; so2 += d * d; }
    return sd2 / so2;
}
inline double rmse(const double* o, const double* m, size_t n) {  // :2358-2376
    if (n == 0) throw std::runtime_error("rmse needs equal sized ts accessors with elements >1");
    double sd2 = 0, obs_avg = 0;
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::isfinite(o[i]) && std::isfinite(m[i])) { double d = o[i] - m[i]; sd2 += d * d; obs_avg += o[i]; ++cnt; }
    obs_avg /= double(cnt);
    return cnt ? std::sqrt(sd2 / cnt) / obs_avg : nan_v;
}
inline double kling_gupta(const double* o, const double* m, size_t n, double s_r, double s_a, double s_b) {  // :2396-2418
    // dlib::running_scalar_covariance<double>: sums of x, y, xx, yy, xy; covariance/variance with n-1
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0, cnt = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::isfinite(o[i]) && std::isfinite(m[i])) { sx += o[i]; sy += m[i]; sxx += o[i] * o[i]; syy += m[i] * m[i]; sxy += o[i] * m[i]; cnt += 1; }
    double qo = sx / cnt, qs = sy / cnt;
    double var_x = 1 / (cnt - 1) * (sxx - sx * sx / cnt), var_y = 1 / (cnt - 1) * (syy - sy * sy / cnt);
    double cov = 1 / (cnt - 1) * (sxy - sx * sy / cnt);
    double uo = std::sqrt(var_x), us = std::sqrt(var_y);
    double r = cov / std::sqrt(var_x * var_y);
    double a = qs / qo, b = us / uo;
    if (!std::isfinite(a)) a = 1.0;
    if (!std::isfinite(b)) b = 1.0;
    double eds2 = (s_r != 0.0 ? dm::pow(s_r * (r - 1), 2) : 0.0) + (s_a != 0.0 ? dm::pow(s_a * (a - 1), 2) : 0.0) +
                  (s_b != 0.0 ? dm::pow(s_b * (b - 1), 2) : 0.0);
    return std::sqrt(eds2);
}
inline double abs_diff_sum(const double* o, const double* m, size_t n) {  // :2421-2431
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) if (std::isfinite(o[i]) && std::isfinite(m[i])) s += std::fabs(o[i] - m[i]);
    return s;
}
}  // namespace goal

// ---------------------------------------------------------------------------
// region model pieces
// ---------------------------------------------------------------------------
// core/region_model.h:233-249: cix in order of first appearance of the catchment id
inline std::vector<int64_t> update_ix_to_id_mapping(std::vector<geo_cell>& cells) {
    std::map<int64_t, int64_t> cid_to_cix;
    std::vector<int64_t> cix_to_cid;
    for (auto& c : cells) {
        auto f = cid_to_cix.find(c.catchment_id);
        if (f == cid_to_cix.end()) {
            cid_to_cix[c.catchment_id] = int64_t(cix_to_cid.size());
            c.catchment_ix = cix_to_cid.size();
            cix_to_cid.push_back(c.catchment_id);
        } else c.catchment_ix = size_t(f->second);
    }
    return cix_to_cid;
}

// core/region_model.h:991-1021: mutex-protected counter, one whole cell per grab, use_ncore workers
template <class F>
inline void parallel_run(size_t n_cells, int use_ncore, F&& run_one) {
    if (n_cells == 0) return;
    if (use_ncore <= 0) throw std::runtime_error("parallel_run: use_ncore is zero ");
    std::mutex pos_mx, err_mx;
    size_t pos = 0;
    std::string err;
    std::vector<std::thread> th;
    for (int i = 0; i < use_ncore; ++i)
        th.emplace_back([&]() {
            try {
                while (true) {
                    size_t ci;
                    {
                        std::lock_guard<std::mutex> lock(pos_mx);
                        if (pos < n_cells) ci = pos++;
                        else break;
                    }
                    run_one(ci);
                }
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lock(err_mx);
                if (err.empty()) err = e.what();
            }
        });
    for (auto& t : th) t.join();
    if (!err.empty()) throw std::runtime_error(err);
}

}  // namespace sho
