/* Stand-in for R's <R.h>, only for building the checker in oracle/ (R is not
 * installed in this image).  It declares the part of the R API the reference's
 * src/*.c use (list: SURVEY.md section 8(c)) and routes exp/log to the engine's
 * bit-reproducible implementations so reference and GPU kernels round alike.
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library. */
#ifndef PHT_SHIM_R_H
#define PHT_SHIM_R_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <stdarg.h>
#include <stddef.h>

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define R_INLINE inline

#ifndef PHT_SHIM_LIBM          /* `make ref-libm` keeps the platform's exp/log (mismatch-rate report only) */
double phtshim_exp(double);
double phtshim_log(double);
#define exp phtshim_exp
#define log phtshim_log
#endif

void Rprintf(const char *, ...);
void REprintf(const char *, ...);
char *R_alloc(size_t, int);
void GetRNGstate(void);
void PutRNGstate(void);
double unif_rand(void);
double exp_rand(void);
double norm_rand(void);
void R_FlushConsole(void);
void R_CheckUserInterrupt(void);
#define R_Calloc(n, t) ((t *)calloc((size_t)(n), sizeof(t)))
#define R_Free(p) free(p)
#endif
