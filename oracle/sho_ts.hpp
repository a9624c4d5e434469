// ============================================================================
// oracle/sho_ts.hpp -- CPU ORACLE (test infrastructure, NOT product code)
//
// Projection of a point source onto a time axis: hint_based_search, accumulate_value, average_value and
// average_accessor::value, restated line by line from core/time_series.h:144-310 and :2033-2072.
// A source is a sequence of points (t[i] in microseconds, v[i]) with a total period that ends at t_end
// (point_ts over a fixed_dt axis: t0 + n*dt; over a point_dt axis: its explicit end).
// Pinned by the reference's own known answers (test/time_series_test.cpp:480-640), tests/test_oracle_resampling.py.
// ============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>

#include "sho_core.hpp"

namespace sho {
namespace ts {

constexpr size_t npos = std::string::npos;

struct point_source {
    const int64_t* t;   // [n] point times, strictly increasing
    const double* v;    // value of point i at v[i * stride]
    size_t n;
    size_t stride;
    int64_t t_end;      // total_period().end
    int64_t time(size_t i) const { return t[i]; }
    double value(size_t i) const { return v[i * stride]; }
    // index of the last point with t <= tx; npos before the first point (test_timeseries / point_ts::index_of within the period)
    size_t index_of(int64_t tx) const {
        if (n == 0 || tx < t[0]) return npos;
        return size_t(std::upper_bound(t, t + n, tx) - t) - 1;
    }
};

// core/time_series.h:144-178
inline size_t hint_based_search(const point_source& source, int64_t p_start, size_t i) {
    const size_t n = source.n;
    if (n == 0) return npos;
    if (i != npos && i < n) {
        const size_t max_directional_search = 5;
        int64_t ti = source.time(i);
        if (ti == p_start) {
            return i;
        } else if (ti < p_start) {
            if (i == n - 1) return i;
            const size_t i_max = std::min(i + max_directional_search, n);
            while (++i < i_max) {
                ti = source.time(i);
                if (ti < p_start) continue;
                return ti > p_start ? i - 1 : i;
            }
            return (i < n) ? source.index_of(p_start) : n - 1;
        } else {
            if (i == 0) return 0;  // :165-166 -- NOT npos: the source's first point becomes the left anchor
            const size_t i_min = i - std::min(i, max_directional_search);
            do {
                ti = source.time(--i);
                if (ti > p_start) continue;
                return i;
            } while (i > i_min);
            return i > 0 ? source.index_of(p_start) : npos;
        }
    }
    return source.index_of(p_start);
}

// core/time_series.h:202-291: area under the non-NaN parts of f(t) over [p_start, p_end), tsum = their length [us]
inline double accumulate_value(const point_source& source, int64_t p_start, int64_t p_end, size_t& last_idx, int64_t& tsum, bool linear = true,
                               bool strict_linear_between = true) {
    const size_t n = source.n;
    const bool extrapolate_flat = !linear || (linear && !strict_linear_between);
    if (n == 0) return nan_v;
    size_t i = hint_based_search(source, p_start, last_idx);
    int64_t l_t = 0;
    double l_v = 0.0;
    bool l_finite = false;
    if (i == npos) {
        i = 0;
        last_idx = 0;
        if (strict_linear_between) {
            l_t = source.time(i); l_v = source.value(i); ++i;
            l_finite = std::isfinite(l_v);
            if (!(l_t >= p_start && l_t < p_end)) return nan_v;  // !p.contains(l.t)
        }
    }
    double area = 0.0;
    tsum = 0;
    while (true) {
        if (!l_finite) {
            l_t = source.time(i); l_v = source.value(i); ++i;
            l_finite = std::isfinite(l_v);
            if (i == n) {
                if (l_finite && l_t < p_end) {
                    if (extrapolate_flat) {
                        const int64_t dt = p_end - std::max(p_start, l_t);
                        tsum += dt;
                        area += to_seconds(dt) * l_v;
                    }
                }
                break;
            }
            if (l_t >= p_end) break;
        } else {
            const int64_t r_t = source.time(i);
            const double r_v = source.value(i);
            ++i;
            const bool r_finite = std::isfinite(r_v);
            const int64_t px_start = std::max(l_t, p_start), px_end = std::min(r_t, p_end);
            int64_t dt = px_end - px_start;  // utcperiod::timespan()
            if (linear && r_finite) {
                const double a = (r_v - l_v) / to_seconds(r_t - l_t);
                const double b = r_v - a * to_seconds(r_t);
                area += to_seconds(dt) * (0.5 * a * to_seconds(px_start + px_end) + b);
                tsum += dt;
            } else {
                if (extrapolate_flat) {
                    area += l_v * to_seconds(dt);
                    tsum += dt;
                }
            }
            if (i == n) {
                if (r_finite && r_t < p_end) {
                    if (extrapolate_flat) {
                        dt = p_end - r_t;
                        tsum += dt;
                        area += to_seconds(dt) * r_v;
                    }
                }
                break;
            }
            if (r_t >= p_end) break;
            l_finite = r_finite;
            l_t = r_t;
            l_v = r_v;
        }
    }
    last_idx = i - 1;
    return tsum ? area : nan_v;
}

// core/time_series.h:306-310
inline double average_value(const point_source& source, int64_t p_start, int64_t p_end, size_t& last_idx, bool linear = true) {
    int64_t tsum = 0;
    const double area = accumulate_value(source, p_start, p_end, last_idx, tsum, linear);
    return tsum > 0 ? area / to_seconds(tsum) : nan_v;
}

// average_accessor<S, fixed_dt>::value(i) for i = 0..n-1 in sequence (core/time_series.h:2033-2072), extension policy USE_NAN,
// linear = (source.point_interpretation() == POINT_INSTANT_VALUE); last_idx starts at 0 as in the constructor (:2045)
inline void average_accessor(const point_source& source, bool linear, int64_t ta_t0, int64_t ta_dt, size_t ta_n, double* out, size_t out_stride) {
    size_t last_idx = 0;
    for (size_t i = 0; i < ta_n; ++i) {
        const int64_t t = ta_t0 + int64_t(i) * ta_dt;
        out[i * out_stride] = t >= source.t_end ? nan_v : average_value(source, t, t + ta_dt, last_idx, linear);
    }
}

}  // namespace ts
}  // namespace sho
