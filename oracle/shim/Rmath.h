/* Stand-in for <Rmath.h>: the four nmath entry points the reference calls. */
#ifndef PHT_SHIM_RMATH_H
#define PHT_SHIM_RMATH_H
double runif(double a, double b);
double rexp(double scale);
double rgamma(double shape, double scale);
double dexp(double x, double scale, int give_log);
#endif
