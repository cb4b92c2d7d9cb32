/* Stand-in for <Rinternals.h>: only the two SEXP type codes src/Registrations.c names (values as in R's own header).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef PHT_SHIM_RINTERNALS_H
#define PHT_SHIM_RINTERNALS_H
#define INTSXP 13
#define REALSXP 14
#endif
