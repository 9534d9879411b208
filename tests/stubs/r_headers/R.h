/* Minimal declaration-only stand-in for R's <R.h>: just enough of the public C API for
 * `gcc -fsyntax-only` to type-check integration/r-package/src/speigen_glue.c in an image without R.
 * Signatures follow "Writing R Extensions" (R >= 4.0).  Never linked, never shipped. */
#ifndef STUB_R_H
#define STUB_R_H
#include <stddef.h>
#include <stdint.h>
void Rf_error(const char*, ...) __attribute__((noreturn));
void R_CheckUserInterrupt(void);
#endif
