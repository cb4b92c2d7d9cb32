/*
 * pht_math.h -- bit-reproducible FP64 exp/log shared by the CUDA kernels and
 * the host-side checkers.
 *
 * Why this exists: tier-1 parity (BASELINE.json north_star) asks for bit-exact
 * integer sufficient statistics against the reference C samplers.  Every
 * categorical scan, ARMS accept/squeeze test and Brent branch in the reference
 * (e.g. src/Simulate_AbsCTMC_gt_Bladt_MHRS.c:61,101; src/arms.c:464;
 * src/utility.c:273-331) is a floating-point comparison, so a 1-ulp difference
 * between glibc's and CUDA's libm flips decisions.  Both sides therefore use
 * THIS implementation: the device code calls pht_exp/pht_log directly and the
 * reference build links `exp`/`log` to them (oracle/shim/rshim.c).  Only
 * IEEE-754 add/mul/div and explicitly written fma() are used, evaluated in a
 * fixed order; host code must be built with -ffp-contract=off and device code
 * with -fmad=false so neither compiler introduces contractions of its own.
 *
 * Accuracy: < 1 ulp for both functions (checked against libm in
 * tests/test_math.py).  Algorithms are the classical ones: Cody-Waite range
 * reduction + Taylor/Horner for exp, the fdlibm-style atanh series for log.
 */
#ifndef PHT_MATH_H
#define PHT_MATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PHT_HD __host__ __device__ __forceinline__
#else
#define PHT_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define PHT_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define PHT_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

PHT_HD uint64_t pht_d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PHT_HD double pht_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}

/* exp(x).  k = round(x/ln2); r = x - k ln2 (two-step, fma); Taylor degree 13 on
 * |r| <= ln2/2 (remainder < 5e-18 relative); scaling by 2^k in two exact
 * halves so that k = 1024 and subnormal results are handled without branches
 * on the common path. */
PHT_HD double pht_exp(double x) {
    if (!(x == x)) return x + x;                          /* NaN */
    if (x > 709.782712893384) return pht_u2d(0x7ff0000000000000ULL);
    if (x < -745.1332191019412) return 0.0;
    const double LOG2E  = 1.4426950408889634074;          /* 1/ln2 */
    const double LN2_HI = 6.93147180559945286227e-01;     /* nearest double to ln2 */
    const double LN2_LO = 2.31904681384629955842e-17;     /* ln2 - LN2_HI */
    const double SHIFT  = 6755399441055744.0;             /* 1.5 * 2^52 */
    double kd = PHT_FMA(x, LOG2E, SHIFT) - SHIFT;          /* round-to-nearest integer */
    double r = PHT_FMA(-kd, LN2_HI, x);
    r = PHT_FMA(-kd, LN2_LO, r);
    double p = 1.6059043836821613e-10;                    /* 1/13! */
    p = PHT_FMA(p, r, 2.08767569878681e-09);              /* 1/12! */
    p = PHT_FMA(p, r, 2.505210838544172e-08);             /* 1/11! */
    p = PHT_FMA(p, r, 2.755731922398589e-07);             /* 1/10! */
    p = PHT_FMA(p, r, 2.7557319223985893e-06);            /* 1/9!  */
    p = PHT_FMA(p, r, 2.48015873015873e-05);              /* 1/8!  */
    p = PHT_FMA(p, r, 1.984126984126984e-04);             /* 1/7!  */
    p = PHT_FMA(p, r, 1.388888888888889e-03);             /* 1/6!  */
    p = PHT_FMA(p, r, 8.333333333333333e-03);             /* 1/5!  */
    p = PHT_FMA(p, r, 4.1666666666666664e-02);            /* 1/4!  */
    p = PHT_FMA(p, r, 1.6666666666666666e-01);            /* 1/3!  */
    p = PHT_FMA(p, r, 0.5);
    p = PHT_FMA(p, r, 1.0);
    p = PHT_FMA(p, r, 1.0);
    int k = (int)kd;
    int k1 = k >> 1, k2 = k - k1;                         /* both within [-538, 512] */
    double s1 = pht_u2d((uint64_t)(int64_t)(k1 + 1023) << 52);
    double s2 = pht_u2d((uint64_t)(int64_t)(k2 + 1023) << 52);
    return (p * s1) * s2;
}

/* log(x).  x = 2^e * m, m in [sqrt(1/2), sqrt(2)); f = m - 1; s = f/(2+f);
 * log(1+f) = f - f^2/2 + s*(f^2/2 + R(s^2)) with the degree-7 even polynomial
 * R from fdlibm (e_log.c; Sun Microsystems 1993, freely distributable). */
PHT_HD double pht_log(double x) {
    const double LN2_HI = 6.93147180369123816490e-01;     /* 0x3fe62e42fee00000 */
    const double LN2_LO = 1.90821492927058770002e-10;     /* 0x3dea39ef35793c76 */
    const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
                 L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
                 L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
                 L7 = 1.479819860511658591e-01;
    uint64_t ux = pht_d2u(x);
    int e = 0;
    if (ux >= 0x7ff0000000000000ULL) {                    /* inf, NaN, negative */
        if (ux == 0x7ff0000000000000ULL) return x;        /* +inf */
        if (!(x == x)) return x + x;                      /* NaN */
        if (x == 0.0) return -pht_u2d(0x7ff0000000000000ULL);   /* -0 */
        return pht_u2d(0x7ff8000000000000ULL);            /* x < 0 */
    }
    if (ux < 0x0010000000000000ULL) {                     /* zero or subnormal */
        if (ux == 0) return -pht_u2d(0x7ff0000000000000ULL);
        x *= 18014398509481984.0;                         /* 2^54 */
        ux = pht_d2u(x);
        e = -54;
    }
    /* bring the mantissa into [sqrt(1/2), sqrt(2)) */
    uint64_t hx = ux >> 32;
    hx += 0x3ff00000 - 0x3fe6a09e;
    e += (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffff) + 0x3fe6a09e;
    double m = pht_u2d((hx << 32) | (ux & 0xffffffffULL));
    double f = m - 1.0;
    double hfsq = 0.5 * f * f;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * PHT_FMA(w, PHT_FMA(w, L6, L4), L2);
    double t2 = z * PHT_FMA(w, PHT_FMA(w, PHT_FMA(w, L7, L5), L3), L1);
    double R = t2 + t1;
    double dk = (double)e;
    return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f);
}

#endif /* PHT_MATH_H */
