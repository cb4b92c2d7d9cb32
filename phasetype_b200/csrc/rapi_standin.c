/*
 * rapi_standin.c -- default definitions of the handful of R API symbols the host routine
 * LJMA_Gibbs uses (Rprintf, R_FlushConsole, GetRNGstate, PutRNGstate, unif_rand), so that
 * libpht_b200.so can be loaded outside R (ctypes, C tests).  They are weak: inside an R
 * session the interpreter's own definitions are found first and these are never used.
 * An R package build simply leaves this file out.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <stdint.h>

static uint64_t standin_state = 0x5048ULL;
static int standin_seeded = 0;

__attribute__((weak)) void Rprintf(const char *fmt, ...) {
    if (getenv("PHT_B200_QUIET")) return;
    va_list ap; va_start(ap, fmt); vfprintf(stdout, fmt, ap); va_end(ap);
}
__attribute__((weak)) void R_FlushConsole(void) { fflush(stdout); }
__attribute__((weak)) void GetRNGstate(void) {
    if (!standin_seeded) {
        const char *s = getenv("PHT_B200_RSEED");
        if (s) standin_state = strtoull(s, NULL, 0);
        standin_seeded = 1;
    }
}
__attribute__((weak)) void PutRNGstate(void) {}
__attribute__((weak)) double unif_rand(void) {          /* splitmix64 -> (0,1) */
    uint64_t z = (standin_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return ((double)(z >> 12) + 0.5) * 2.220446049250313e-16;
}
