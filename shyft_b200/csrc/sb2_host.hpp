// sb2_host.hpp -- host-side support of the C ABI: UTC calendar slice, small dense algebra for the
// station x station part of Bayesian temperature kriging, gamma unit hydrographs.
//
// Follows (paths relative to the reference root):
//   calendar day arithmetic        core/utctime_utilities.h:342-365, core/utctime_utilities.cpp:230-255
//   BTK operator algebra           core/bayesian_kriging.h:300-316,362-374 (armadillo in the reference)
//   make_uhg_from_gamma            core/routing.h:399-421 (boost::math::gamma_distribution in the reference)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "sb2_math.cuh"

namespace sb2 {
namespace host {

// ---- time: utctime = int64 microseconds (core/utctime_utilities.h:28-34) ---------------------------------
constexpr int64_t kUsec = 1000000LL;
constexpr int64_t kUnixDay = 2440588LL;

inline int64_t julian_day(int y, int mo, int d) {
    const int a = (14 - mo) / 12;
    const int yy = y + 4800 - a;
    const int mm = mo + 12 * a - 3;
    return d + ((153 * mm + 2) / 5) + 365LL * yy + (yy / 4) - (yy / 100) + (yy / 400) - 32045;
}
inline int64_t julian_day_of(int64_t t_us) { return (86400LL * kUnixDay + t_us / kUsec) / 86400; }
inline int year_of_julian_day(int64_t jdn) {
    const int64_t a = jdn + 32044;
    const int64_t b = (4 * a + 3) / 146097;
    const int64_t c = a - ((146097 * b) / 4);
    const int64_t d = (4 * c + 3) / 1461;
    const int64_t e = c - (1461 * d) / 4;
    const int64_t m = (5 * e + 2) / 153;
    return int(100 * b + d - 4800 + (m / 10));
}
// calendar::day_of_year, UTC (core/utctime_utilities.cpp:230-235)
inline int day_of_year(int64_t t_us) {
    const int64_t jdn = julian_day_of(t_us);
    return int(1 + jdn - julian_day(year_of_julian_day(jdn), 1, 1));
}
// seconds since calendar::trim(t, YEAR), UTC (core/utctime_utilities.cpp:248-255)
inline int64_t seconds_of_year(int64_t t_us) {
    const int64_t jdn = julian_day_of(t_us);
    const int64_t year_start_us = (julian_day(year_of_julian_day(jdn), 1, 1) - kUnixDay) * 86400LL * kUsec;
    return (t_us - year_start_us) / kUsec;
}

// ---- dense row-major matrices, only what BTK's S x S algebra needs -----------------------------------------
struct Mat {
    size_t r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(size_t r_, size_t c_) : r(r_), c(c_), a(r_ * c_, 0.0) {}
    double& operator()(size_t i, size_t j) { return a[i * c + j]; }
    double operator()(size_t i, size_t j) const { return a[i * c + j]; }
};
inline Mat matmul(const Mat& A, const Mat& B) {  // C(i,j) = sum_k A(i,k)*B(k,j), k ascending
    Mat C(A.r, B.c);
    for (size_t i = 0; i < A.r; ++i)
        for (size_t k = 0; k < A.c; ++k) {
            const double aik = A(i, k);
            for (size_t j = 0; j < B.c; ++j) C(i, j) += aik * B(k, j);
        }
    return C;
}
inline Mat transpose(const Mat& A) {
    Mat T(A.c, A.r);
    for (size_t i = 0; i < A.r; ++i)
        for (size_t j = 0; j < A.c; ++j) T(j, i) = A(i, j);
    return T;
}
// Gauss-Jordan inverse with partial pivoting (stands in for arma::mat::i() = LAPACK getrf/getri)
inline Mat inverse(const Mat& A_) {
    const size_t n = A_.r;
    Mat A = A_, I(n, n);
    for (size_t i = 0; i < n; ++i) I(i, i) = 1.0;
    for (size_t col = 0; col < n; ++col) {
        size_t piv = col;
        for (size_t r = col + 1; r < n; ++r)
            if (std::fabs(A(r, col)) > std::fabs(A(piv, col))) piv = r;
        if (A(piv, col) == 0.0) throw std::runtime_error("inv(): matrix seems singular");
        if (piv != col)
            for (size_t j = 0; j < n; ++j) { std::swap(A(col, j), A(piv, j)); std::swap(I(col, j), I(piv, j)); }
        const double d = A(col, col);
        for (size_t r = col + 1; r < n; ++r) {
            const double f = A(r, col) / d;
            if (f == 0.0) continue;
            for (size_t j = col; j < n; ++j) A(r, j) -= f * A(col, j);
            for (size_t j = 0; j < n; ++j) I(r, j) -= f * I(col, j);
        }
    }
    for (size_t col = n; col-- > 0;) {
        const double d = A(col, col);
        for (size_t j = 0; j < n; ++j) I(col, j) /= d;
        for (size_t r = 0; r < col; ++r) {
            const double f = A(r, col);
            if (f == 0.0) continue;
            for (size_t j = 0; j < n; ++j) I(r, j) -= f * I(col, j);
        }
    }
    return I;
}
// numerical rank of a symmetric 2x2 (arma::rank: singular values above max(dim)*s_max*eps)
inline int rank_sym22(const Mat& H) {
    const double a = H(0, 0), b = 0.5 * (H(0, 1) + H(1, 0)), d = H(1, 1);
    const double tr = a + d, det = a * d - b * b;
    const double disc = std::sqrt(std::max(0.0, 0.25 * tr * tr - det));
    const double l1 = std::fabs(0.5 * tr + disc), l2 = std::fabs(0.5 * tr - disc);
    const double smax = std::max(l1, l2), smin = std::min(l1, l2);
    const double tol = 2.0 * smax * 2.220446049250313e-16;
    return int(smax > tol) + int(smin > tol);
}

// The station-side BTK operators for the station subset `valid` (core/bayesian_kriging.h:300-316 full, :362-374 reduced):
//   K_inv (n x n), FtKinv = F.t()*K_inv (2 x n), M22 = I - G*H_inv (2 x 2), E_beta_w = H*F.t()*K_inv (2 x n)
struct BtkStationOps {
    Mat K_inv, FtKinv, M22, E_beta_w;
    std::vector<double> z;  // station heights of the subset (column 1 of F)
};
inline BtkStationOps btk_station_ops(const std::vector<double>& sxyz, const std::vector<int>& valid, double sill, double nug, double range,
                                     double zscale, double gradient_sd, bool check_rank) {
    const size_t n = valid.size();
    Mat K(n, n), F(n, 2);
    const double c0 = sill - nug;
    for (size_t i = 0; i < n; ++i) {
        const double* a = &sxyz[3 * valid[i]];
        F(i, 0) = 1.0;
        F(i, 1) = a[2];
        K(i, i) = 1.0 * c0;
        for (size_t j = i + 1; j < n; ++j) {
            const double* b = &sxyz[3 * valid[j]];
            const double d = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]) * zscale * zscale);
            K(i, j) = K(j, i) = c0 * sb_exp(-d / range);
        }
    }
    BtkStationOps o;
    o.K_inv = inverse(K);
    const Mat Ft = transpose(F);
    o.FtKinv = matmul(Ft, o.K_inv);
    const Mat H_inv = matmul(o.FtKinv, F);
    if (check_rank && rank_sym22(H_inv) == 1)
        throw std::runtime_error("The bayestian temperature kriging algorithm needs at least two sources at different heights.");
    const Mat H = inverse(H_inv);
    Mat G_inv = H_inv;
    G_inv(1, 1) += 1 / (gradient_sd * gradient_sd);
    const Mat G = inverse(G_inv);
    const Mat GH_inv = matmul(G, H_inv);
    o.M22 = Mat(2, 2);
    o.M22(0, 0) = 1.0 - GH_inv(0, 0); o.M22(0, 1) = 0.0 - GH_inv(0, 1);
    o.M22(1, 0) = 0.0 - GH_inv(1, 0); o.M22(1, 1) = 1.0 - GH_inv(1, 1);
    o.E_beta_w = matmul(matmul(H, Ft), o.K_inv);
    o.z.resize(n);
    for (size_t i = 0; i < n; ++i) o.z[i] = F(i, 1);
    return o;
}
// bayesian_kriging::parameter::temperature_gradient(period) (core/bayesian_kriging.h:220-223): DOY sinusoid at the period midpoint
inline double btk_prior_gradient(int64_t p_start_us, int64_t dt_us) {
    const double doy = double(day_of_year(p_start_us + dt_us / 2));
    return 1.18e-3 * std::sin(6.2831 / 365 * (doy + 79.0)) - 5.48e-3;
}

// ---- gamma unit hydrograph (core/routing.h:399-421) -----------------------------------------------------------
inline double gamma_p_full(double a, double x) {  // regularised lower incomplete gamma, full double
    if (!(x > 0.0)) return 0.0;
    if (std::isinf(x)) return 1.0;
    const double pre = sb_exp(a * sb_log(x) - x - sb_lgamma(a));
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 2000; ++n) { ap += 1.0; del *= x / ap; sum += del; if (del < sum * 1.0e-16) break; }
        return sum * pre;
    }
    const double tiny = 1.0e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const double an = -double(i) * (double(i) - a);
        b += 2.0;
        d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
        c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < 1.0e-16) break;
    }
    return 1.0 - pre * h;
}
inline double gamma_pdf(double alpha, double x) {
    if (x < 0) return 0.0;
    if (x == 0) return 0.0;  // boost 1.68 gamma_distribution pdf returns 0 at x == 0 for every shape
    return sb_exp((alpha - 1.0) * sb_log(x) - x - sb_lgamma(alpha));
}
inline double gamma_quantile(double alpha, double pq) {  // bracketed Newton on P(alpha, x)
    double lo = 0.0, hi = std::max(1.0, alpha);
    while (gamma_p_full(alpha, hi) < pq) hi *= 2.0;
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 200; ++it) {
        const double fx = gamma_p_full(alpha, x) - pq;
        if (fx > 0) hi = x; else lo = x;
        const double d = gamma_pdf(alpha, x);
        double xn = d > 0 ? x - fx / d : 0.5 * (lo + hi);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (std::fabs(xn - x) <= 1e-15 * std::fabs(x)) { x = xn; break; }
        x = xn;
    }
    return x;
}
inline int uhg_steps(double distance, double velocity, int64_t dt_us) {  // routing.h:119-123,326-330
    const double steps = (distance / velocity) / (double(dt_us) / double(kUsec));
    return int(steps + 0.5);
}
inline std::vector<double> make_uhg_from_gamma(int n_steps, double alpha, double beta) {
    std::vector<double> r;
    if (n_steps > 1) {
        double s = 0.0;
        const double x_max = gamma_quantile(alpha, 0.99);
        const double d = x_max / double(n_steps);
        for (int i = 0; i < n_steps; ++i) {
            const double y = std::max(0.0, gamma_pdf(alpha, d * i) + beta);
            s += y;
            r.push_back(y);
        }
        if (s > 0.0) for (auto& y : r) y /= s;
        else for (auto& y : r) y = 1 / double(n_steps);
    }
    if (r.empty()) r.push_back(1.0);
    return r;
}

// ---- one-dimensional minimiser of the state tuning (core/model_state_tuning.h:98-108) ---------------------------------
// The reference calls dlib 19.16 find_min_single_variable(f, x, begin, end, eps, max_iter) (dlib/optimization/optimization_line_search.h;
// the library is not under /root/reference).  Its published algorithm, written out here: (1) three points p1 < p2 < p3 around the start
// (search radius 1, clipped to [begin, end]); (2) walk / shrink until f1 > f2 < f3, doubling the radius on every outward step;
// (3) shrink the bracket with the minimum of the parabola through the three points (Fletcher eq. 4.2.1), kept at least a tenth of the
// sub-interval away from the points it would split and pushed to the wider side when one side is > 100 x the other; stop when
// p3 - p1 <= eps.  Throws like dlib does: argument check first, "exceeded the allowable number of iterations" on max_iter.
struct MinimiserFailure : std::runtime_error { using std::runtime_error::runtime_error; };

inline double parabola_min_3pt(double p1, double p2, double p3, double f1, double f2, double f3) {
    const double num = f1 * (p3 * p3 - p2 * p2) + f2 * (p1 * p1 - p3 * p3) + f3 * (p2 * p2 - p1 * p1);
    const double den = 2 * (f1 * (p3 - p2) + f2 * (p1 - p3) + f3 * (p2 - p1));
    if (den == 0) return p2;
    const double r = num / den;
    if (p1 <= r && r <= p3) return r;
    return std::min(std::max(p1, r), p3);
}
template <class F>
double find_min_single_variable(F&& f, double& x, double begin, double end, double eps, long max_iter, double radius = 1.0) {
    if (!(eps > 0 && max_iter > 1 && begin <= x && x <= end && radius > 0))
        throw MinimiserFailure("find_min_single_variable: eps > 0, max_iter > 1 and begin <= starting_point <= end are required");
    long evals = 1;
    if (begin == end) return f(x);
    double p1 = std::max(x - radius, begin), p3 = std::min(x + radius, end), p2;
    double f1 = f(p1), f3 = f(p3), f2;
    if (x == p1 || x == p3) { p2 = (p1 + p3) / 2; f2 = f(p2); }
    else { p2 = x; f2 = f(x); }
    evals += 2;
    while (!(f1 > f2 && f2 < f3)) {
        if (evals >= max_iter) throw MinimiserFailure("The max number of iterations of single variable optimization have been reached without converging.");
        if (p3 - p1 < eps) {
            if (f1 < std::min(f2, f3)) { x = p1; return f1; }
            if (f2 < std::min(f1, f3)) { x = p2; return f2; }
            x = p3; return f3;
        }
        if (f1 <= f3) {  // the low side is the left one
            if (p1 == begin || (f1 == f2 && (end - begin) < radius)) { p3 = p2; f3 = f2; p2 = (p1 + p2) / 2.0; f2 = f(p2); }
            else { p3 = p2; f3 = f2; p2 = p1; f2 = f1; p1 = std::max(p1 - radius, begin); f1 = f(p1); radius *= 2; }
        } else {
            if (p3 == end || (f2 == f3 && (end - begin) < radius)) { p1 = p2; f1 = f2; p2 = (p3 + p2) / 2.0; f2 = f(p2); }
            else { p1 = p2; f1 = f2; p2 = p3; f2 = f3; p3 = std::min(p3 + radius, end); f3 = f(p3); radius *= 2; }
        }
        ++evals;
    }
    const double tau = 0.1;
    while (evals < max_iter && p3 - p1 > eps) {
        double pm = parabola_min_3pt(p1, p2, p3, f1, f2, f3);
        if (pm < p2) {
            const double d = (p2 - p1) * tau;
            if (std::fabs(p1 - pm) < d) pm = p1 + d;
            else if (std::fabs(p2 - pm) < d) pm = p2 - d;
        } else {
            const double d = (p3 - p2) * tau;
            if (std::fabs(p2 - pm) < d) pm = p2 + d;
            else if (std::fabs(p3 - pm) < d) pm = p3 - d;
        }
        const double ratio = std::fabs(p1 - p2) / std::fabs(p2 - p3);
        if (!(ratio < 100 && ratio > 0.01)) {
            if (ratio > 1 && pm > p2) pm = (p1 + p2) / 2;
            else if (pm < p2) pm = (p2 + p3) / 2;
        }
        const double fm = f(pm);
        if (pm < p2) {
            if (f1 > fm && fm < f2) { p3 = p2; f3 = f2; p2 = pm; f2 = fm; }
            else { p1 = pm; f1 = fm; }
        } else {
            if (f2 > fm && fm < f3) { p1 = p2; f1 = f2; p2 = pm; f2 = fm; }
            else { p3 = pm; f3 = fm; }
        }
        ++evals;
    }
    if (evals >= max_iter) throw MinimiserFailure("The max number of iterations of single variable optimization have been reached without converging.");
    x = p2;
    return f2;
}

}  // namespace host
}  // namespace sb2
