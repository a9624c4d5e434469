// sb2_math.cuh -- fp64 math used by the cell-stack kernels (sm_100a) and by the host-side operator builds.
//
// Deterministic elementary functions: exp, log, pow, lgamma written with IEEE-754 +,-,*,/, sqrt and explicit fma() only (no
// implicit contraction: the library is built with -fmad=false; no CUDA libm), so that the device, the host part of this library
// and any other conforming machine produce identical bits.  Why: gamma_snow's Brent search (core/gamma_snow.h:214-227)
// amplifies last-bit differences between math libraries to 1e-4-level differences in liquid water content; a 1e-9 parity
// statement is only meaningful over one fixed operation sequence (DESIGN.md "Deterministic math").  The sequence:
//   exp(x)    k = rint(x/ln2) (magic-number add), r = fma(k,-ln2_lo, fma(k,-ln2_hi,x)), degree-13 Taylor polynomial (even/odd Horner, fma), scaled by 2^k
//   log(x)    x = 2^e*m, m in (sqrt(1/2), sqrt(2)], f = m-1, s = f/(2+f), f - f^2/2 + s*(f^2/2 + R(s^2)), R = atanh series to s^20 (even/odd Horner)
//   pow(x,y)  exact for y = 0, 1, 2, 0.5; exp(y*log(x)) otherwise (x >= 0)
//   lgamma(a) recurrence up to a >= 12, then the Stirling series to 1/a^13
// The reference evaluates through libm and boost 1.68 (absent from its tree):
//   gamma_p / lgamma      core/gamma_snow.h:195-201   (boost::math, reduced-precision policies)
//   brent_find_minima     core/gamma_snow.h:216-226   (bits = 12, 60 iterations)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define SB2_HD __host__ __device__ __forceinline__

namespace sb2 {

SB2_HD double from_bits(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
SB2_HD unsigned long long bits_of(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long u; memcpy(&u, &x, 8); return u;
#endif
}
SB2_HD double pow2i(int k) { return from_bits((unsigned long long)(k + 1023) << 52); }  // 2^k, -1022 <= k <= 1023
SB2_HD double inf_() { return from_bits(0x7ff0000000000000ULL); }
SB2_HD double nan_() { return from_bits(0x7ff8000000000000ULL); }

// Polynomial coefficients.  On the device they live in constant memory so that every fma takes its coefficient as a
// constant-bank operand (a 64-bit immediate costs two extra moves per use); the host reads the same initialisers.
#define SB2_EXP_COEFFS {0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, \
                        1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0}
#define SB2_LOG_COEFFS {2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0, 2.0 / 17.0, 2.0 / 19.0, 2.0 / 21.0}
#define SB2_MISC_COEFFS {1.44269504088896338700e+00, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 6.93147180369123816490e-01, \
                         1.90821492927058770002e-10}
__constant__ double kExpC_dev[12] = SB2_EXP_COEFFS;
__constant__ double kLogC_dev[10] = SB2_LOG_COEFFS;
__constant__ double kMiscC_dev[5] = SB2_MISC_COEFFS;
static const double kExpC_host[12] = SB2_EXP_COEFFS;
static const double kLogC_host[10] = SB2_LOG_COEFFS;
static const double kMiscC_host[5] = SB2_MISC_COEFFS;
#ifdef __CUDA_ARCH__
#define SB2_EXPC(i) kExpC_dev[i]
#define SB2_LOGC(i) kLogC_dev[i]
#define SB2_MISC(i) kMiscC_dev[i]
#else
#define SB2_EXPC(i) kExpC_host[i]
#define SB2_LOGC(i) kLogC_host[i]
#define SB2_MISC(i) kMiscC_host[i]
#endif

// reduced argument and polynomial of exp: returns p = exp(r) for x = k*ln2 + r, |x| < 746
//   k = round-to-nearest(x/ln2) by the magic-number add (t = x/ln2 + 1.5*2^52: the low word of t is k, t - 1.5*2^52 is k as a double)
//   exp(r) = 1 + (r + r^2 Q(r)),  Q = Qe(r^2) + r Qo(r^2): two independent Horner chains, every fma takes ONE coefficient -- on sm_100a
//   a constant operand of DFMA has to come from a uniform register, and one per instruction is what the ISA allows
SB2_HD double sb_exp_core(double x, int& k) {
    const double t = fma(x, SB2_MISC(0), 6755399441055744.0);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, SB2_MISC(1), x);
    r = fma(kf, SB2_MISC(2), r);
    const double z = r * r;
    double qe = SB2_EXPC(10), qo = SB2_EXPC(11);
    qe = fma(qe, z, SB2_EXPC(8));
    qo = fma(qo, z, SB2_EXPC(9));
    qe = fma(qe, z, SB2_EXPC(6));
    qo = fma(qo, z, SB2_EXPC(7));
    qe = fma(qe, z, SB2_EXPC(4));
    qo = fma(qo, z, SB2_EXPC(5));
    qe = fma(qe, z, SB2_EXPC(2));
    qo = fma(qo, z, SB2_EXPC(3));
    qe = fma(qe, z, SB2_EXPC(0));
    qo = fma(qo, z, SB2_EXPC(1));
    const double Q = fma(r, qo, qe);
    k = int(unsigned(bits_of(t) & 0xffffffffULL));
    return 1.0 + fma(z, Q, r);
}
// the general path: NaN, overflow, underflow into the subnormal range
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
double sb_exp_slow(double x) {
    if (x != x) return x;
    if (x > 709.782712893384) return inf_();
    if (x < -745.1332191019412) return 0.0;
    int k;
    double p = sb_exp_core(x, k);
    if (k > 1023) { p *= pow2i(1023); k -= 1023; }
    if (k < -1022) { p *= pow2i(k + 1000); return p * pow2i(-1000); }
    return p * pow2i(k);
}
SB2_HD double sb_exp_inl(double x) {
    if (!(fabs(x) < 690.0)) return sb_exp_slow(x);
    int k;
    const double p = sb_exp_core(x, k);
    // |k| <= 996 and p in (0.7, 1.5): p * 2^k is exact and normal, i.e. an add on the exponent field
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    return p * pow2i(k);
#endif
}
// The same value with the polynomial outside any branch: warp-synchronous callers (ptgsk_response_kernel) keep their control flow
// uniform, so the compiler hoists the coefficients out of the loops; the general path patches the result of the rare lane.
__device__ __forceinline__ double sb_exp_flat(double x) {
#ifdef __CUDA_ARCH__
    int k;
    const double p = sb_exp_core(x, k);
    double r = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    if (!(fabs(x) < 690.0)) r = sb_exp_slow(x);
    return r;
#else
    return sb_exp_inl(x);
#endif
}

// Out-of-line copies for the device: the cell-step kernels call exp/log from some thirty places; one shared copy of each
// keeps the step loop inside the instruction cache (inlined everywhere the pt_gs_k kernel was 136 KB of SASS and 22 % of
// its issue slots stalled on instruction fetch).  SB2_MATH_INLINE=1 restores full inlining (tuning only).
#ifndef SB2_MATH_INLINE
#define SB2_MATH_INLINE 0
#endif
#if defined(__CUDA_ARCH__) && !SB2_MATH_INLINE
#define SB2_MATH_FN __device__ __noinline__
#else
#define SB2_MATH_FN SB2_HD
#endif
SB2_MATH_FN double sb_exp(double x) { return sb_exp_inl(x); }

SB2_HD double sb_log_inl(double x) {
    if (x != x || x < 0.0) return nan_();
    if (x == 0.0) return -inf_();
    if (x == inf_()) return x;
    int e = 0;
    if (x < 2.2250738585072014e-308) { x *= 18014398509481984.0; e = -54; }
    const unsigned long long u = bits_of(x);
    e += int((u >> 52) & 0x7ff) - 1023;
    double m = from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s, w = z * z;
    double re = SB2_LOGC(8), ro = SB2_LOGC(9);  // even / odd Horner chains in w = z^2: one coefficient per fma (see sb_exp_core)
    re = fma(re, w, SB2_LOGC(6));
    ro = fma(ro, w, SB2_LOGC(7));
    re = fma(re, w, SB2_LOGC(4));
    ro = fma(ro, w, SB2_LOGC(5));
    re = fma(re, w, SB2_LOGC(2));
    ro = fma(ro, w, SB2_LOGC(3));
    re = fma(re, w, SB2_LOGC(0));
    ro = fma(ro, w, SB2_LOGC(1));
    const double R = z * fma(z, ro, re);
    const double hfsq = 0.5 * f * f;
    const double dk = double(e);
    const double t = fma(s, hfsq + R, dk * SB2_MISC(4));
    return fma(dk, SB2_MISC(3), f - (hfsq - t));
}
SB2_MATH_FN double sb_log(double x) { return sb_log_inl(x); }
// the same value without early returns (special operands patched by selects at the end), for code that must stay branch-free
__device__ __forceinline__ double sb_log_flat(double x0) {
    const bool sub = x0 < 2.2250738585072014e-308;
    const double x = sub ? x0 * 18014398509481984.0 : x0;
    const unsigned long long u = bits_of(x);
    int e = int((u >> 52) & 0x7ff) - 1023 + (sub ? -54 : 0);
    double m = from_bits((u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    const bool hi = m > 1.4142135623730951;
    m = hi ? m * 0.5 : m;
    e += hi ? 1 : 0;
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s, w = z * z;
    double re = SB2_LOGC(8), ro = SB2_LOGC(9);  // even / odd Horner chains in w = z^2: one coefficient per fma (see sb_exp_core)
    re = fma(re, w, SB2_LOGC(6));
    ro = fma(ro, w, SB2_LOGC(7));
    re = fma(re, w, SB2_LOGC(4));
    ro = fma(ro, w, SB2_LOGC(5));
    re = fma(re, w, SB2_LOGC(2));
    ro = fma(ro, w, SB2_LOGC(3));
    re = fma(re, w, SB2_LOGC(0));
    ro = fma(ro, w, SB2_LOGC(1));
    const double R = z * fma(z, ro, re);
    const double hfsq = 0.5 * f * f;
    const double dk = double(e);
    const double t = fma(s, hfsq + R, dk * SB2_MISC(4));
    double r = fma(dk, SB2_MISC(3), f - (hfsq - t));
    if (x0 == inf_()) r = x0;
    if (x0 == 0.0) r = -inf_();
    if (x0 != x0 || x0 < 0.0) r = nan_();
    return r;
}
__device__ __forceinline__ double sb_pow_flat(double x, double y) {  // y is a literal at every call site: the exact cases fold away
    if (y == 0.0) return 1.0;
    if (y == 1.0) return x;
    if (y == 2.0) return x * x;
    if (y == 0.5) return sqrt(x);
    const double r = sb_exp_flat(y * sb_log_flat(x));
    return x == 0.0 ? (y > 0.0 ? 0.0 : inf_()) : r;
}

SB2_HD double sb_pow(double x, double y) {
    if (y == 0.0) return 1.0;
    if (y == 1.0) return x;
    if (y == 2.0) return x * x;
    if (y == 0.5) return sqrt(x);
    if (x == 0.0) return y > 0.0 ? 0.0 : inf_();
    return sb_exp(y * sb_log(x));
}
SB2_HD double sb_pow4(double x) { const double x2 = x * x; return x2 * x2; }
SB2_HD double sb_pow8(double x) { double y = x * x; y = y * y; return y * y; }

SB2_MATH_FN double sb_lgamma(double a) {
    double prod = 1.0;
    while (a < 12.0) { prod *= a; a += 1.0; }
    const double ai = 1.0 / a, ai2 = ai * ai;
    double s = 1.0 / 156.0;
    s = fma(-ai2, s, 691.0 / 360360.0);
    s = fma(-ai2, s, 1.0 / 1188.0);
    s = fma(-ai2, s, 1.0 / 1680.0);
    s = fma(-ai2, s, 1.0 / 1260.0);
    s = fma(-ai2, s, 1.0 / 360.0);
    s = fma(-ai2, s, 1.0 / 12.0);
    s = ai * s;
    return (((a - 0.5) * sb_log(a) - a) + 0.91893853320467274178) + s - sb_log(prod);
}


// num / den for den > 0 where num is often exactly zero (no precipitation, no outflow, ...): +-0 / den = +-0 exactly, and a zero
// numerator sends CUDA's division through its ~90-instruction general path (ncu: three such calls per step of the snow kernel)
// (the compiler turns `num == 0 ? num : num / den` back into an unconditional division of num, so the zero lanes divide 1 / den
// behind an optimisation barrier)
__device__ __forceinline__ double div_pos(double num, double den) {
    const bool z = num == 0.0;
    double n1 = z ? 1.0 : num;
    asm("" : "+d"(n1));  // opaque to the optimiser, or the select folds away again
    const double q = n1 / den;
    return z ? num : q;
}

SB2_HD double dmax(double a, double b) { return (a < b) ? b : a; }  // std::max(a,b): (a < b) ? b : a
SB2_HD double dmin(double a, double b) { return b < a ? b : a; }  // std::min(a,b): (b < a) ? b : a

// x^a e^-x / Gamma(a), the prefix shared by P(a,x) and its a-derivative term (gamma_snow.h:245,254)
__device__ __forceinline__ double gamma_prefix(double a, double x, double lgamma_a) { return sb_exp(a * sb_log(x) - x - lgamma_a); }

// Regularised lower incomplete gamma P(a,x) given the prefix: series for x < a+1, continued fraction for Q otherwise, both
// without a division per term (an fp64 division is ~30 instructions; the series runs 10-40 terms per call and was the single
// hottest loop of the pt_gs_k kernel).  The series is carried as the fraction P/Q (Q *= a+n, P = P (a+n) + x^n), the continued
// fraction by the forward recurrence of its convergents A_i/B_i; exact 2^-500 rescaling keeps them in range.
// Terms are taken four (convergents two) at a time with the tests after each group: short dependent chains, one branch per group.
__device__ __forceinline__ double gamma_p_with_prefix_inl(double a, double x, double pre) {
    const double eps = 1.0e-16;
    const double small = 3.0549363634996047e-151;  // 2^-500
    if (x < a + 1.0) {
        double ap = a, P = 1.0, Q = a, xn = 1.0;
        for (int n = 0; n < 500; ++n) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ap += 1.0;
                xn *= x;
                Q *= ap;
                P = fma(P, ap, xn);
            }
            if (xn < P * eps) break;
            if (__double2hiint(Q) > 0x5f300000) { Q *= small; P *= small; xn *= small; }  // Q > 2^500 (Q > 0: high word decides)
        }
        return (P / Q) * pre;
    }
    double b = x + 1.0 - a, di = 0.0;
    double A1 = 1.0, B1 = 0.0, A = b, B = 1.0;
    for (int i = 0; i < 1000; ++i) {
        di += 1.0;
        double an = -di * (di - a);
        b += 2.0;
        A1 = fma(b, A, an * A1);
        B1 = fma(b, B, an * B1);
        di += 1.0;
        an = -di * (di - a);
        b += 2.0;
        A = fma(b, A1, an * A);
        B = fma(b, B1, an * B);
        const double m1 = A * B1, m0 = A1 * B;
        if (fabs(m1 - m0) < eps * fabs(m1)) break;
        if ((__double2hiint(A) & 0x7fffffff) > 0x5f300000) { A *= small; B *= small; A1 *= small; B1 *= small; }  // |A| > 2^500
    }
    return 1.0 - pre * (B / A);
}
__device__ __noinline__ double gamma_p_with_prefix(double a, double x, double pre) { return gamma_p_with_prefix_inl(a, x, pre); }

// Two incomplete gamma values at once: P(a1, x1) and P(a2, x2), each evaluated exactly as by gamma_p_with_prefix_inl (same terms,
// same tests, same rescaling), but in ONE series loop and ONE continued-fraction loop that advance both problems together: two
// independent dependency chains per lane and max(n1, n2) instead of n1 + n2 passes.  A problem that has converged has its result
// latched and is carried along idle.  Bit-identical to two single evaluations (tests/test_gpu_units.py) but measured SLOWER where it
// was tried (calc_snow_state: the second value is rarely needed; the Brent objective: -7 %), so no production kernel calls it.
__device__ __forceinline__ void gamma_p_pair_inl(double a1, double x1, bool need1, double pre1, double a2, double x2, bool need2, double pre2,
                                                 double& P1, double& P2) {
    const double eps = 1.0e-16;
    const double small = 3.0549363634996047e-151;  // 2^-500
    const bool s1 = need1 && x1 < a1 + 1.0, s2 = need2 && x2 < a2 + 1.0;
    const bool c1 = need1 && !s1, c2 = need2 && !s2;
    if (s1 || s2) {
        double apa = a1, Qa = a1, Pa = 1.0, xa = 1.0;
        double apb = a2, Qb = a2, Pb = 1.0, xb = 1.0;
        double Pa_f = 1.0, Qa_f = a1, Pb_f = 1.0, Qb_f = a2;
        bool ra = s1, rb = s2;
        for (int n = 0; n < 500; ++n) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                apa += 1.0;
                apb += 1.0;
                xa *= x1;
                xb *= x2;
                Qa *= apa;
                Qb *= apb;
                Pa = fma(Pa, apa, xa);
                Pb = fma(Pb, apb, xb);
            }
            if (ra) { Pa_f = Pa; Qa_f = Qa; ra = !(xa < Pa * eps); }
            if (rb) { Pb_f = Pb; Qb_f = Qb; rb = !(xb < Pb * eps); }
            if (!(ra || rb)) break;
            if (__double2hiint(Qa) > 0x5f300000) { Qa *= small; Pa *= small; xa *= small; }
            if (__double2hiint(Qb) > 0x5f300000) { Qb *= small; Pb *= small; xb *= small; }
        }
        if (s1) P1 = (Pa_f / Qa_f) * pre1;
        if (s2) P2 = (Pb_f / Qb_f) * pre2;
    }
    if (c1 || c2) {
        double ba = x1 + 1.0 - a1, bb = x2 + 1.0 - a2, di = 0.0;
        double A1a = 1.0, B1a = 0.0, Aa = ba, Ba = 1.0;
        double A1b = 1.0, B1b = 0.0, Ab = bb, Bb = 1.0;
        double Aa_f = Aa, Ba_f = Ba, Ab_f = Ab, Bb_f = Bb;
        bool ra = c1, rb = c2;
        for (int i = 0; i < 1000; ++i) {
            di += 1.0;
            double ana = -di * (di - a1), anb = -di * (di - a2);
            ba += 2.0;
            bb += 2.0;
            A1a = fma(ba, Aa, ana * A1a);
            B1a = fma(ba, Ba, ana * B1a);
            A1b = fma(bb, Ab, anb * A1b);
            B1b = fma(bb, Bb, anb * B1b);
            di += 1.0;
            ana = -di * (di - a1);
            anb = -di * (di - a2);
            ba += 2.0;
            bb += 2.0;
            Aa = fma(ba, A1a, ana * Aa);
            Ba = fma(ba, B1a, ana * Ba);
            Ab = fma(bb, A1b, anb * Ab);
            Bb = fma(bb, B1b, anb * Bb);
            if (ra) {
                const double m1 = Aa * B1a, m0 = A1a * Ba;
                Aa_f = Aa; Ba_f = Ba;
                ra = !(fabs(m1 - m0) < eps * fabs(m1));
            }
            if (rb) {
                const double m1 = Ab * B1b, m0 = A1b * Bb;
                Ab_f = Ab; Bb_f = Bb;
                rb = !(fabs(m1 - m0) < eps * fabs(m1));
            }
            if (!(ra || rb)) break;
            if ((__double2hiint(Aa) & 0x7fffffff) > 0x5f300000) { Aa *= small; Ba *= small; A1a *= small; B1a *= small; }
            if ((__double2hiint(Ab) & 0x7fffffff) > 0x5f300000) { Ab *= small; Bb *= small; A1b *= small; B1b *= small; }
        }
        if (c1) P1 = 1.0 - pre1 * (Ba_f / Aa_f);
        if (c2) P2 = 1.0 - pre2 * (Bb_f / Ab_f);
    }
}

__device__ __noinline__ void gamma_p_pair(double a1, double x1, double pre1, double a2, double x2, double pre2, double& P1, double& P2) {
    gamma_p_pair_inl(a1, x1, true, pre1, a2, x2, true, pre2, P1, P2);
}

__device__ __forceinline__ double gamma_p(double a, double x, double lgamma_a) {
    if (!(x > 0.0)) return 0.0;
    if (x == inf_()) return 1.0;
    return gamma_p_with_prefix(a, x, gamma_prefix(a, x, lgamma_a));
}
__device__ __forceinline__ double gamma_p(double a, double x) { return gamma_p(a, x, sb_lgamma(a)); }

}  // namespace sb2
