/* Stand-in for <R_ext/Utils.h> (nothing needed beyond R.h). */
