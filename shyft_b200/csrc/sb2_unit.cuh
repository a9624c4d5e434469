// sb2_unit.cuh -- per-function evaluation kernel behind sb2_unit_eval: lets the tests compare single device functions
// (deterministic math, incomplete gamma, Brent's lwc correction, snow state, one Kirchner step) with the oracle, the way
// the reference unit-tests its methods (test/gamma_snow_test.cpp, test/kirchner_test.cpp).
#pragma once
#include "sb2_ptgsk.cuh"

namespace sb2 {

enum { UNIT_EXP = 0, UNIT_LOG, UNIT_POW, UNIT_LGAMMA, UNIT_GAMMA_P, UNIT_CORR_LWC, UNIT_CALC_SNOW_STATE, UNIT_KIRCHNER_STEP, UNIT_N };

__global__ void unit_eval_kernel(int fn, int64_t n, const double* __restrict__ in, int n_in, double* __restrict__ out, int n_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* a = in + i * n_in;
    double* o = out + i * n_out;
    switch (fn) {
        case UNIT_EXP: o[0] = sb_exp(a[0]); break;
        case UNIT_LOG: o[0] = sb_log(a[0]); break;
        case UNIT_POW: o[0] = sb_pow(a[0], a[1]); break;
        case UNIT_LGAMMA: o[0] = sb_lgamma(a[0]); break;
        case UNIT_GAMMA_P: o[0] = gamma_p(a[0], a[1]); break;
        case UNIT_CORR_LWC: o[0] = gs_corr_lwc(a[0], a[1], a[2], a[3], a[4]); break;
        case UNIT_CALC_SNOW_STATE: {
            double lg_key = nan_(), lg_val = 0.0;
            gs_calc_snow_state(a[0], a[1], a[2], a[3], a[4], a[5], a[6], o[0], o[1], lg_key, lg_val);
            break;
        }
        case UNIT_KIRCHNER_STEP: {
            double q = a[4], q_avg = 0.0;
            const bool ok = kirchner_step(a[0], a[1], a[2], a[3], q, q_avg, a[5], a[6]);
            o[0] = q; o[1] = q_avg; o[2] = ok ? 1.0 : 0.0;
            break;
        }
        default: break;
    }
}

}  // namespace sb2
