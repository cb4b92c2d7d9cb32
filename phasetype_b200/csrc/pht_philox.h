/*
 * pht_philox.h -- the engine's random-number contract, shared by device and host.
 *
 * The reference draws every variate from R's global generator
 * (unif_rand/runif/rexp; call sites listed in SURVEY.md section 8(c)), which is
 * sequential and therefore cannot be reproduced by a parallel sampler.  The
 * engine instead defines a counter-based stream (Philox4x32-10, Salmon et al.,
 * SC'11) addressed by
 *
 *      key     = 64-bit seed (one per chain)
 *      counter = { block, substream, observation, iteration }
 *
 * and uniform number `d` of a substream is half (d & 1) of block (d >> 1).
 * Substreams: for MHRS the substream is the rejection-attempt index of the
 * observation (the reference calls LJMA_GUI(), i.e. R_FlushConsole(), exactly
 * once per attempt: src/Simulate_AbsCTMC_gt_Bladt_MHRS.c:120, which is the hook
 * the checker's R shim uses to advance the substream); ECS and DCS use
 * substream 0 for the whole path.  The parameter update uses observation
 * 0xFFFFFFFF and substream = parameter index.
 *
 * Derived variates (these DEFINE the R nmath boundary for parity purposes, as
 * R itself is not part of the reference tree):
 *      unif_rand()  = (x + 0.5) * 2^-52,  x = top 52 bits of a 64-bit word
 *      exp_rand()   = -log(unif_rand())            (one uniform per variate)
 *      runif(a,b)   = a + (b - a) * unif_rand()
 *      rexp(scale)  = scale * exp_rand()
 */
#ifndef PHT_PHILOX_H
#define PHT_PHILOX_H

#include <stdint.h>
#include "pht_math.h"

#define PHT_PHILOX_M0 0xD2511F53u
#define PHT_PHILOX_M1 0xCD9E8D57u
#define PHT_PHILOX_W0 0x9E3779B9u
#define PHT_PHILOX_W1 0xBB67AE85u

#define PHT_OBS_PARAM 0xFFFFFFFFu   /* "observation" slot used by the parameter update */

typedef struct { uint32_t v[4]; } pht_u32x4;

PHT_HD pht_u32x4 pht_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                   uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)PHT_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHT_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHT_PHILOX_W0; k1 += PHT_PHILOX_W1;
    }
    pht_u32x4 out; out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

/* Philox4x32-10 with the ten round keys precomputed (rk[2r] = k0 + r W0, rk[2r+1] = k1 + r W1): the kernels
 * read them from the kernel-parameter constant bank as immediate operands of the XORs. */
#define PHT_PHILOX_ROUND(RK0, RK1) { \
        const uint64_t p0_ = (uint64_t)PHT_PHILOX_M0 * c0, p1_ = (uint64_t)PHT_PHILOX_M1 * c2; \
        const uint32_t n0_ = (uint32_t)(p1_ >> 32) ^ c1 ^ (RK0), n2_ = (uint32_t)(p0_ >> 32) ^ c3 ^ (RK1); \
        c1 = (uint32_t)p1_; c3 = (uint32_t)p0_; c0 = n0_; c2 = n2_; }
typedef struct { uint32_t k[20]; } pht_roundkeys;
PHT_HD void pht_roundkeys_init(pht_roundkeys *rk, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) { rk->k[2 * r] = k0 + (uint32_t)r * PHT_PHILOX_W0; rk->k[2 * r + 1] = k1 + (uint32_t)r * PHT_PHILOX_W1; }
}

/* map 64 random bits to the open interval (0,1): 52 bits, centred */
PHT_HD double pht_u01(uint32_t lo, uint32_t hi) {
    /* (x + 0.5) * 2^-52 with x the top 52 bits, built without an int->double conversion:
     * 1.x (a double in [1,2)) minus 1 is x * 2^-52 exactly, and adding 2^-53 is exact too */
    const uint64_t x = (((uint64_t)hi << 32) | lo) >> 12;
    /* 1.x - (1 - 2^-53) = (2x + 1) * 2^-53 is exactly representable, so this single subtraction is exact
     * (and equal to the two-step form (1.x - 1) + 2^-53) */
    return pht_u2d(0x3ff0000000000000ULL | x) - 0.99999999999999988898;
}

/* A positioned stream: draws are numbered 0,1,2,... inside (iter, obs, sub). */
typedef struct {
    uint32_t k0, k1;        /* seed */
    uint32_t iter, obs, sub;
    uint32_t d;             /* index of the next uniform */
    double spare;           /* second half of the last block when d is odd */
} pht_stream;

PHT_HD void pht_stream_seek(pht_stream *s, uint32_t iter, uint32_t obs, uint32_t sub, uint32_t d) {
    s->iter = iter; s->obs = obs; s->sub = sub; s->d = d; s->spare = 0.0;
    if (d & 1u) {
        pht_u32x4 b = pht_philox4x32_10(d >> 1, sub, obs, iter, s->k0, s->k1);
        s->spare = pht_u01(b.v[2], b.v[3]);
    }
}

PHT_HD double pht_stream_unif(pht_stream *s) {
    double u;
    if (s->d & 1u) {
        u = s->spare;
    } else {
        pht_u32x4 b = pht_philox4x32_10(s->d >> 1, s->sub, s->obs, s->iter, s->k0, s->k1);
        u = pht_u01(b.v[0], b.v[1]);
        s->spare = pht_u01(b.v[2], b.v[3]);
    }
    s->d++;
    return u;
}

/* direct access to uniform number d of a substream (stateless) */
PHT_HD double pht_unif_at(uint32_t k0, uint32_t k1, uint32_t iter, uint32_t obs, uint32_t sub, uint32_t d) {
    pht_u32x4 b = pht_philox4x32_10(d >> 1, sub, obs, iter, k0, k1);
    return (d & 1u) ? pht_u01(b.v[2], b.v[3]) : pht_u01(b.v[0], b.v[1]);
}

/* ---- Gamma(shape a, scale) for the conjugate update and the prior start draw.
 * R's rgamma (Ahrens-Dieter GD/GS, nmath/rgamma.c) is outside the reference
 * tree; the engine defines its own: Marsaglia-Tsang (2000) squeeze method with
 * Marsaglia polar normals, boosted by u^(1/a) when a < 1.  All arithmetic goes
 * through pht_exp/pht_log/sqrt so host and device agree bit for bit. */
#if defined(__CUDA_ARCH__)
#define PHT_SQRT(x) __dsqrt_rn(x)
#else
#define PHT_SQRT(x) __builtin_sqrt(x)
#endif

PHT_HD double pht_norm_polar(pht_stream *s) {
    for (;;) {
        double x = 2.0 * pht_stream_unif(s) - 1.0;
        double y = 2.0 * pht_stream_unif(s) - 1.0;
        double r2 = x * x + y * y;
        if (r2 < 1.0 && r2 > 0.0) {
            return x * PHT_SQRT(-2.0 * pht_log(r2) / r2);
        }
    }
}

PHT_HD double pht_rgamma(pht_stream *s, double a, double scale) {
    double boost = 1.0;
    if (a < 1.0) {
        double u = pht_stream_unif(s);
        boost = pht_exp(pht_log(u) / a);
        a += 1.0;
    }
    double dd = a - 1.0 / 3.0;
    double c = 1.0 / PHT_SQRT(9.0 * dd);
    for (;;) {
        double x, v;
        do {
            x = pht_norm_polar(s);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        double u = pht_stream_unif(s);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * dd * v * scale;
        if (pht_log(u) < 0.5 * x2 + dd * (1.0 - v + pht_log(v))) return boost * dd * v * scale;
    }
}

#endif /* PHT_PHILOX_H */
