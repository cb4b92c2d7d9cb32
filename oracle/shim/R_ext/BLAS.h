/* Stand-in for <R_ext/BLAS.h>.  dgemv/dgemm resolve to plain-loop versions in
 * rshim.c that follow the association order of the Fortran reference BLAS R
 * ships by default (libRblas), so results do not depend on a vendor kernel. */
#ifndef PHT_SHIM_BLAS_H
#define PHT_SHIM_BLAS_H
#define F77_CALL(x) phtshim_##x
#define FCONE
void phtshim_dgemv(const char *trans, const int *m, const int *n, const double *alpha,
                   const double *a, const int *lda, const double *x, const int *incx,
                   const double *beta, double *y, const int *incy);
void phtshim_dgemm(const char *transa, const char *transb, const int *m, const int *n, const int *k,
                   const double *alpha, const double *a, const int *lda, const double *b, const int *ldb,
                   const double *beta, double *c, const int *ldc);
#endif
