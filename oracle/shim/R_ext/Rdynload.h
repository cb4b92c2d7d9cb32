/* Stand-in for <R_ext/Rdynload.h>: the registration types and entry points src/Registrations.c uses, with the layout
 * of R's own header.  The stand-in R_registerRoutines (regstub.c) records the table so that a test can inspect what
 * useDynLib(PhaseType, .registration = TRUE) would bind.  TEST INFRASTRUCTURE ONLY. */
#ifndef PHT_SHIM_RDYNLOAD_H
#define PHT_SHIM_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef unsigned int R_NativePrimitiveArgType;
typedef struct { const char *name; DL_FUNC fun; int numArgs; R_NativePrimitiveArgType *types; } R_CMethodDef;
typedef R_CMethodDef R_FortranMethodDef;
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef R_CallMethodDef R_ExternalMethodDef;
typedef struct _DllInfo { const R_CMethodDef *c_methods; int dynamic_symbols; int force_symbols; } DllInfo;
typedef int Rboolean;
int R_registerRoutines(DllInfo *info, const R_CMethodDef *const croutines, const R_CallMethodDef *const callRoutines,
                       const R_FortranMethodDef *const fortranRoutines, const R_ExternalMethodDef *const externalRoutines);
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value);
Rboolean R_forceSymbols(DllInfo *info, Rboolean value);
#endif
