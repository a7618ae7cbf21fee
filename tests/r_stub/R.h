/* stub of R.h for the syntax check of r_shim.c (tests/test_r_package.py): there is no R in this image */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
#endif
