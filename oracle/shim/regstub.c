/* R's side of the registration handshake, recorded instead of acted upon (see R_ext/Rdynload.h in this directory).
 * TEST INFRASTRUCTURE ONLY. */
#include <stddef.h>
#include <string.h>
#include <R_ext/Rdynload.h>
int R_registerRoutines(DllInfo *info, const R_CMethodDef *const croutines, const R_CallMethodDef *const callRoutines,
                       const R_FortranMethodDef *const fortranRoutines, const R_ExternalMethodDef *const externalRoutines) {
    (void)callRoutines; (void)fortranRoutines; (void)externalRoutines;
    info->c_methods = croutines; return 1;
}
Rboolean R_useDynamicSymbols(DllInfo *info, Rboolean value) { int old = info->dynamic_symbols; info->dynamic_symbols = value; return old; }
Rboolean R_forceSymbols(DllInfo *info, Rboolean value) { int old = info->force_symbols; info->force_symbols = value; return old; }

/* what a test asks: run the package's init routine and report the .C entry `name` */
void R_init_PhaseType(DllInfo *info);
int phtreg_lookup(const char *name, void **fun, int *num_args, unsigned int *types, int max_types, int *dynamic, int *force) {
    DllInfo info; memset(&info, 0, sizeof(info)); info.dynamic_symbols = 1;
    R_init_PhaseType(&info);
    *dynamic = info.dynamic_symbols; *force = info.force_symbols;
    for (const R_CMethodDef *m = info.c_methods; m && m->name; m++)
        if (strcmp(m->name, name) == 0) {
            *fun = (void *)m->fun; *num_args = m->numArgs;
            for (int i = 0; i < m->numArgs && i < max_types; i++) types[i] = m->types[i];
            return 0;
        }
    return -1;
}
