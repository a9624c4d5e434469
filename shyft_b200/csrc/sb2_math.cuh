// sb2_math.cuh -- fp64 device math used by the cell-stack kernels (sm_100a).
//
// The reference evaluates these through boost 1.68 (absent from its tree):
//   gamma_p / lgamma      core/gamma_snow.h:195-201   (boost::math, reduced-precision policies)
//   brent_find_minima     core/gamma_snow.h:216-226   (bits = 12, 60 iterations)
// The kernels use full-double evaluations with a fixed operation order; tests/ checks them against
// the CPU oracle, which states the same algorithms independently.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace sb2 {

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }  // std::max(a,b): (a < b) ? b : a
__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }  // std::min(a,b): (b < a) ? b : a

// x^a e^-x / Gamma(a), the prefix shared by P(a,x) and its a-derivative term (gamma_snow.h:245,254)
__device__ __forceinline__ double gamma_prefix(double a, double x, double lgamma_a) { return exp(a * log(x) - x - lgamma_a); }

// Regularised lower incomplete gamma P(a,x) given the prefix: series for x < a+1, Lentz continued fraction otherwise.
__device__ __noinline__ double gamma_p_with_prefix(double a, double x, double pre) {
    const double eps = 1.0e-16;
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 2000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (del < sum * eps) break;
        }
        return sum * pre;
    }
    const double tiny = 1.0e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        const double an = -double(i) * (double(i) - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return 1.0 - pre * h;
}

__device__ __forceinline__ double gamma_p(double a, double x, double lgamma_a) {
    if (!(x > 0.0)) return 0.0;
    if (isinf(x)) return 1.0;
    return gamma_p_with_prefix(a, x, gamma_prefix(a, x, lgamma_a));
}
__device__ __forceinline__ double gamma_p(double a, double x) { return gamma_p(a, x, lgamma(a)); }

}  // namespace sb2
