/* stub of R.h for the syntax check of r_shim.c (tests/test_r_package.py): there is no R in this image */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
char *R_alloc(size_t, int);            /* R_ext/Memory.h: transient storage, freed at the end of the .Call */
#endif
