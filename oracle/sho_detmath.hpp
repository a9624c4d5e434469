// ============================================================================
// oracle/sho_detmath.hpp -- CPU ORACLE (test infrastructure, NOT product code)
//
// Deterministic elementary functions.  The reference calls libm (exp, log, pow) and boost (lgamma); their
// results differ by an ulp from platform to platform, and gamma_snow's Brent search (core/gamma_snow.h:214-227)
// amplifies such last-bit noise to 1e-4-level differences in liquid water content whenever its objective is flat
// (measured on the B200: CUDA libm vs glibc gave 2.7 % of cell-steps off by more than 1e-9).  A 1e-9 parity
// statement between two machines is therefore only meaningful if both evaluate the SAME operation sequence.
// This header is that sequence, written with IEEE-754 +,-,*,/, sqrt and explicit fma() only (no implicit contraction,
// no libm), so that any
// conforming machine -- the host CPU here, an sm_100a SM in shyft_b200/csrc/sb2_math.cuh -- produces identical
// bits.  Accuracy (checked in tests/test_oracle_detmath.py against libm / scipy): exp, log < 1.5 ulp;
// lgamma abs error < 1e-14 on (0.05, 200); pow(x,y) = exp(y*log(x)) except the exact cases y = 0, 0.5, 1, 2.
// The oracle's pinned known answers (tests/test_oracle_*) are re-checked with these functions in place.
// ============================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

// Optional operation counters (tools/cost_model.py builds a second copy of the library with -DSHO_COUNT, single-threaded runs only).
#ifdef SHO_COUNT
namespace sho { namespace dm {
struct counters_t { long long v[16]; };
enum { C_EXP, C_LOG, C_LGAMMA, C_GSER_CALLS, C_GSER_ITER, C_GCF_CALLS, C_GCF_ITER, C_BRENT_CALLS, C_BRENT_EVAL, C_KIR_TRY, C_KIR_REJECT, C_SNOW_STATE,
       C_GS_ACTIVE, C_CELL_STEPS, C_N };
inline counters_t g_cnt{};
inline unsigned char* g_cost_cursor = nullptr;  // tools/cost_model.py: 4 bytes per cell-step {series terms, fraction convergents, exp+log, brent evals} of gamma_snow
} }
#define SHO_CNT(i, n) (::sho::dm::g_cnt.v[::sho::dm::i] += (n))
#else
#define SHO_CNT(i, n) ((void)0)
#endif

namespace sho {
namespace dm {

inline uint64_t bits_of(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
inline double from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
inline double pow2i(int k) { return from_bits(uint64_t(k + 1023) << 52); }  // 2^k, -1022 <= k <= 1023

// exp(x): k = rint(x/ln2) by the magic-number add (t = x/ln2 + 1.5*2^52: low word of t = k), r = x - k*ln2 (two fused steps),
// exp(r) = 1 + (r + r^2 Q(r)) with the degree-11 Taylor Q split into even and odd Horner chains in r^2, scaled by 2^k
inline double exp(double x) {
    SHO_CNT(C_EXP, 1);
    if (x != x) return x;
    if (x > 709.782712893384) return std::numeric_limits<double>::infinity();
    if (x < -745.1332191019412) return 0.0;
    const double t = std::fma(x, 1.44269504088896338700e+00, 6755399441055744.0);
    const double kf = t - 6755399441055744.0;
    double r = std::fma(kf, -6.93147180369123816490e-01, x);
    r = std::fma(kf, -1.90821492927058770002e-10, r);
    const double z = r * r;
    double qe = 1.0 / 479001600.0, qo = 1.0 / 6227020800.0;
    qe = std::fma(qe, z, 1.0 / 3628800.0);
    qo = std::fma(qo, z, 1.0 / 39916800.0);
    qe = std::fma(qe, z, 1.0 / 40320.0);
    qo = std::fma(qo, z, 1.0 / 362880.0);
    qe = std::fma(qe, z, 1.0 / 720.0);
    qo = std::fma(qo, z, 1.0 / 5040.0);
    qe = std::fma(qe, z, 1.0 / 24.0);
    qo = std::fma(qo, z, 1.0 / 120.0);
    qe = std::fma(qe, z, 0.5);
    qo = std::fma(qo, z, 1.0 / 6.0);
    const double Q = std::fma(r, qo, qe);
    double p = 1.0 + std::fma(z, Q, r);
    int k = int(uint32_t(bits_of(t) & 0xffffffffULL));
    if (k > 1023) { p *= pow2i(1023); k -= 1023; }
    if (k < -1022) { p *= pow2i(k + 1000); return p * pow2i(-1000); }      // one rounding into the subnormal range
    return p * pow2i(k);
}

// log(x): x = 2^e * m, m in (sqrt(1/2), sqrt(2)], f = m-1, s = f/(2+f), log(1+f) = f - f^2/2 + s*(f^2/2 + R(s^2)),
// R(z) = z * sum_{k=0..9} 2/(2k+3) z^k  (the atanh series; even and odd Horner chains in z^2)
inline double log(double x) {
    SHO_CNT(C_LOG, 1);
    if (x != x || x < 0.0) return std::numeric_limits<double>::quiet_NaN();
    if (x == 0.0) return -std::numeric_limits<double>::infinity();
    if (x == std::numeric_limits<double>::infinity()) return x;
    int e = 0;
    if (x < 2.2250738585072014e-308) { x *= 18014398509481984.0; e = -54; }  // subnormal: scale by 2^54
    const uint64_t u = bits_of(x);
    e += int((u >> 52) & 0x7ff) - 1023;
    double m = from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s, w = z * z;
    double re = 2.0 / 19.0, ro = 2.0 / 21.0;
    re = std::fma(re, w, 2.0 / 15.0);
    ro = std::fma(ro, w, 2.0 / 17.0);
    re = std::fma(re, w, 2.0 / 11.0);
    ro = std::fma(ro, w, 2.0 / 13.0);
    re = std::fma(re, w, 2.0 / 7.0);
    ro = std::fma(ro, w, 2.0 / 9.0);
    re = std::fma(re, w, 2.0 / 3.0);
    ro = std::fma(ro, w, 2.0 / 5.0);
    const double R = z * std::fma(z, ro, re);
    const double hfsq = 0.5 * f * f;
    const double dk = double(e);
    const double t = std::fma(s, hfsq + R, dk * 1.90821492927058770002e-10);
    return std::fma(dk, 6.93147180369123816490e-01, f - (hfsq - t));
}

// pow(x, y) for x >= 0: exact for y = 0, 1, 2, 0.5; exp(y*log(x)) otherwise
inline double pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (y == 1.0) return x;
    if (y == 2.0) return x * x;
    if (y == 0.5) return std::sqrt(x);
    if (x == 0.0) return y > 0.0 ? 0.0 : std::numeric_limits<double>::infinity();
    return dm::exp(y * dm::log(x));
}
inline double pow4(double x) { const double x2 = x * x; return x2 * x2; }
inline double pow8(double x) { double y = x * x; y = y * y; return y * y; }

// lgamma(a), a > 0: shift a up to >= 12 by the recurrence, then the Stirling series
inline double lgamma(double a) {
    SHO_CNT(C_LGAMMA, 1);
    double prod = 1.0;
    while (a < 12.0) { prod *= a; a += 1.0; }
    const double ai = 1.0 / a, ai2 = ai * ai;
    double s = 1.0 / 156.0;
    s = std::fma(-ai2, s, 691.0 / 360360.0);
    s = std::fma(-ai2, s, 1.0 / 1188.0);
    s = std::fma(-ai2, s, 1.0 / 1680.0);
    s = std::fma(-ai2, s, 1.0 / 1260.0);
    s = std::fma(-ai2, s, 1.0 / 360.0);
    s = std::fma(-ai2, s, 1.0 / 12.0);
    s = ai * s;
    return (((a - 0.5) * dm::log(a) - a) + 0.91893853320467274178) + s - dm::log(prod);
}

}  // namespace dm
}  // namespace sho
