/* Stand-in for <R_ext/Lapack.h>: forwards to the LAPACK inside the OpenBLAS
 * that ships with the scipy wheel (symbol prefix scipy_, LP64). */
#ifndef PHT_SHIM_LAPACK_H
#define PHT_SHIM_LAPACK_H
#include <R_ext/BLAS.h>
void phtshim_dgeevx(const char *balanc, const char *jobvl, const char *jobvr, const char *sense,
                    const int *n, double *a, const int *lda, double *wr, double *wi,
                    double *vl, const int *ldvl, double *vr, const int *ldvr,
                    int *ilo, int *ihi, double *scale, double *abnrm,
                    double *rconde, double *rcondv, double *work, const int *lwork,
                    int *iwork, int *info);
void phtshim_dgetrf(const int *m, const int *n, double *a, const int *lda, int *ipiv, int *info);
void phtshim_dgetri(const int *n, double *a, const int *lda, int *ipiv, double *work, const int *lwork, int *info);
#endif
