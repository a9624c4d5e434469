// sb2_goal.cuh -- the calibration goal function on the device.
//
// Follows model_calibration::optimizer::run (core/model_calibration.h:830-899): per target the simulated property series
// (sum of catchment discharge / charge over the target's catchments :742-761, or the area-weighted catchment mean of
// snow covered area / snow water equivalent :765-790) is resampled onto the target's axis with average_accessor semantics
// (true average of the stair-case over each target period, core/time_series.h:202-310,2033-2072) and compared with the
// observed series by NASH_SUTCLIFFE / KLING_GUPTA / ABS_DIFF / RMSE (core/time_series.h:2316-2448); the goal is the
// scale_factor-weighted mean over the targets with a finite value (:881-887).
#pragma once
#include <stdint.h>

#include "sb2_math.cuh"

namespace sb2 {

enum { GOAL_NASH_SUTCLIFFE = 0, GOAL_KLING_GUPTA = 1, GOAL_ABS_DIFF = 2, GOAL_RMSE = 3,
       GOAL_ABS_DIFF_SCALED = 4 };  // ABS_DIFF of a CELL_CHARGE target: abs_diff_sum_goal_function_scaled (core/time_series.h:2435-2448)

struct GoalTarget {
    const double* series;     // simulated series [n_ens][T][n_col] (device): catchment sums, or a per-target snow series
    int64_t ens_stride;       // doubles between ensemble members of `series`
    int32_t n_col;
    int32_t pad_;
    const double* obs;        // [n] observed values on the target axis (device)
    int64_t first_step;       // model step where target period 0 starts
    int32_t steps_per_period; // model steps per target period
    int32_t n;                // target periods
    const int32_t* cix;       // [n_cix] catchment indices summed into the property (device)
    int32_t n_cix;
    int32_t calc_mode;
    double s_r, s_a, s_b;
    double dt_seconds;        // model step length
    int64_t rows;             // rows of `series` (the model axis; or, for a target whose axis is not aligned with it, the target periods
                              // themselves: series = the property already projected by average_accessor_kernel, one row per period)
    const double* scale;      // [n] max_abs_average_accessor values for such a pre-projected target (null: evaluated here)
};

// property[t] for one (ensemble member, target): sum over the target's catchments of series[t][cix]; then the period average.
// One block per (target, member); out[e * n_targets + k] = partial goal value.
__global__ void __launch_bounds__(256) goal_kernel(const GoalTarget* __restrict__ targets, int n_targets, double* __restrict__ out) {
    __shared__ double red[6][256];
    __shared__ double obs_avg_s;
    const GoalTarget t = targets[blockIdx.x];
    const int e = blockIdx.y;
    const double* s = t.series + (int64_t)e * t.ens_stride;
    const int n_catch = t.n_col;
    const int64_t T = t.rows;
    // average_accessor of the stair-case sum over target period i; `scale` = max_abs_average_accessor of the same period
    // (core/time_series.h:2198-2267): the larger of the period averages of max(0, v) and max(0, -v)
    auto sim_value2 = [&](int i, double& scale) {
        double area = 0.0, tsum = 0.0, apos = 0.0, aneg = 0.0;
        for (int j = 0; j < t.steps_per_period; ++j) {
            const int64_t step = t.first_step + (int64_t)i * t.steps_per_period + j;
            if (step >= T) break;
            double v = 0.0;
            for (int c = 0; c < t.n_cix; ++c) v += s[step * n_catch + t.cix[c]];
            if (isfinite(v)) {
                area += v * t.dt_seconds; tsum += t.dt_seconds;
                apos += dmax(0.0, v) * t.dt_seconds; aneg += dmax(0.0, -v) * t.dt_seconds;
            }
        }
        scale = t.scale != nullptr ? t.scale[i] : (tsum > 0.0 ? dmax(apos / tsum, aneg / tsum) : nan_());
        return tsum > 0.0 ? area / tsum : nan_();
    };
    auto sim_value = [&](int i) { double unused; return sim_value2(i, unused); };
    // pass 1: sums over the periods where both are finite
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;  // meaning depends on calc_mode
    for (int i = threadIdx.x; i < t.n; i += blockDim.x) {
        double sv;
        const double o = t.obs[i], m = sim_value2(i, sv);
        if (isfinite(o) && isfinite(m)) {
            const double d = o - m;
            if (t.calc_mode == GOAL_ABS_DIFF_SCALED) { if (isfinite(sv) && fabs(sv) > 1e-20) a0 += fabs(d) / sv; }
            else if (t.calc_mode == GOAL_KLING_GUPTA) { a0 += o; a1 += m; a2 += o * o; a3 += m * m; a4 += o * m; a5 += 1.0; }
            else if (t.calc_mode == GOAL_ABS_DIFF) { a0 += fabs(d); }
            else { a0 += d * d; a1 += o; a5 += 1.0; }
        }
    }
    red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2; red[3][threadIdx.x] = a3; red[4][threadIdx.x] = a4; red[5][threadIdx.x] = a5;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w)
            for (int q = 0; q < 6; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + w];
        __syncthreads();
    }
    const double s0 = red[0][0], s1 = red[1][0], s2 = red[2][0], s3 = red[3][0], s4 = red[4][0], cnt = red[5][0];
    double result = 0.0;
    if (t.calc_mode == GOAL_NASH_SUTCLIFFE) {
        if (threadIdx.x == 0) obs_avg_s = s1 / cnt;
        __syncthreads();
        const double obs_avg = obs_avg_s;
        double b = 0.0;  // pass 2: sum (o - mean(o))^2 over the same periods
        for (int i = threadIdx.x; i < t.n; i += blockDim.x) {
            const double o = t.obs[i], m = sim_value(i);
            if (isfinite(o) && isfinite(m)) { const double d = o - obs_avg; b += d * d; }
        }
        __syncthreads();
        red[0][threadIdx.x] = b;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if (threadIdx.x < w) red[0][threadIdx.x] += red[0][threadIdx.x + w];
            __syncthreads();
        }
        result = s0 / red[0][0];
    } else if (t.calc_mode == GOAL_RMSE) {
        const double obs_avg = s1 / cnt;
        result = cnt > 0.0 ? sqrt(s0 / cnt) / obs_avg : nan_();
    } else if (t.calc_mode == GOAL_ABS_DIFF || t.calc_mode == GOAL_ABS_DIFF_SCALED) {
        result = s0;
    } else {  // Kling-Gupta over dlib::running_scalar_covariance semantics (n-1 denominators)
        const double qo = s0 / cnt, qs = s1 / cnt;
        const double var_x = 1 / (cnt - 1) * (s2 - s0 * s0 / cnt), var_y = 1 / (cnt - 1) * (s3 - s1 * s1 / cnt);
        const double cov = 1 / (cnt - 1) * (s4 - s0 * s1 / cnt);
        const double uo = sqrt(var_x), us = sqrt(var_y);
        const double r = cov / sqrt(var_x * var_y);
        double a = qs / qo, b = us / uo;
        if (!isfinite(a)) a = 1.0;
        if (!isfinite(b)) b = 1.0;
        const double er = t.s_r != 0.0 ? (t.s_r * (r - 1)) * (t.s_r * (r - 1)) : 0.0;
        const double ea = t.s_a != 0.0 ? (t.s_a * (a - 1)) * (t.s_a * (a - 1)) : 0.0;
        const double eb = t.s_b != 0.0 ? (t.s_b * (b - 1)) * (t.s_b * (b - 1)) : 0.0;
        result = sqrt(er + ea + eb);
    }
    if (threadIdx.x == 0) out[(int64_t)e * n_targets + blockIdx.x] = result;
}

// The property series of one target on the model axis, for targets whose own axis is not aligned with it:
// out[t] = sum over the target's catchments of series[t][cix]; mode 1 / 2 = max(0, v) / max(0, -v) of it (source_max_abs, finite values only)
__global__ void goal_property_kernel(const double* __restrict__ series, int64_t rows, int n_col, const int32_t* __restrict__ cix, int n_cix, int mode,
                                     double* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    double v = 0.0;
    for (int c = 0; c < n_cix; ++c) v += series[t * n_col + cix[c]];
    if (mode != 0 && isfinite(v)) v = mode == 1 ? dmax(0.0, v) : dmax(0.0, -v);
    out[t] = v;
}
__global__ void goal_max_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dmax(a[i], b[i]);  // std::max(pos_val, neg_val), core/time_series.h:2261
}

// area-weighted catchment means of a per-cell series: out[t][k] = sum_{cells of k} v[t][c]*area[c] / sum area  (model_calibration.h:765-790)
__global__ void catchment_area_mean_kernel(const double* __restrict__ v /* [rows][n_cells] */, const double* __restrict__ area,
                                           const int32_t* __restrict__ cell_ptr, const int32_t* __restrict__ cell_of_catch, int n_catch,
                                           int64_t rows, int64_t n_cells, double* __restrict__ out /* [rows][n_catch] */) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * n_catch) return;
    const int64_t t = idx / n_catch;
    const int k = int(idx % n_catch);
    double s = 0.0, a = 0.0;
    for (int j = cell_ptr[k]; j < cell_ptr[k + 1]; ++j) {
        const int c = cell_of_catch[j];
        s += v[t * n_cells + c] * area[c];
        a += area[c];
    }
    out[idx] = s * (1 / a);
}

}  // namespace sb2
