/* .Call glue between the R closure spEigen() (R/spEigen.R:46) and the C ABI of libspeigen_b200.so
 * (include/speigen_b200.h).  Thin by design: pointers and sizes in, list(vectors, standard_vectors, values) out
 * (R/spEigen.R:162).  Errors are raised with Rf_error only after the library has returned.
 * Cannot be compiled in the build image (no R headers); kept in sync with INTEGRATION.md. */
#include <stdint.h>
#include <string.h>
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include "speigen_b200.h"

SEXP speigen_call(SEXP X, SEXP q, SEXP rho, SEXP data, SEXP d, SEXP V, SEXP thres, SEXP gpus) {
  SEXP dim = getAttrib(X, R_DimSymbol);
  const int64_t nrow = INTEGER(dim)[0], ncol = INTEGER(dim)[1];      /* m*m may exceed 2^31: never use length() as int */
  const int cplx = isComplex(X);
  const int nq = (int)asReal(q);
  speigen_cfg cfg;
  speigen_cfg_default(&cfg);
  cfg.n_gpus = asInteger(gpus);
  speigen_out out;
  memset(&out, 0, sizeof(out));
  SEXP vec = PROTECT(allocMatrix(cplx ? CPLXSXP : REALSXP, (int)ncol, nq));
  SEXP std = PROTECT(allocMatrix(cplx ? CPLXSXP : REALSXP, (int)ncol, nq));
  SEXP val = PROTECT(allocVector(REALSXP, nq));
  out.vectors = cplx ? (double*)COMPLEX(vec) : REAL(vec);             /* Rcomplex = {double r, i} = interleaved */
  out.standard_vectors = cplx ? (double*)COMPLEX(std) : REAL(std);
  out.values = REAL(val);
  const double* dp = isNull(d) ? NULL : REAL(d);
  const double* Vp = isNull(V) ? NULL : (cplx ? (const double*)COMPLEX(V) : REAL(V));
  const double* Xp = cplx ? (const double*)COMPLEX(X) : REAL(X);
  const int rc = cplx ? speigen_run_c64(Xp, nrow, ncol, asLogical(data), asReal(q), asReal(rho), dp, Vp, asReal(thres), &cfg, &out)
                      : speigen_run_f64(Xp, nrow, ncol, asLogical(data), asReal(q), asReal(rho), dp, Vp, asReal(thres), &cfg, &out);
  if (rc != SPEIGEN_OK) {
    UNPROTECT(3);
    Rf_error("%s", speigen_last_error());
  }
  SEXP res = PROTECT(allocVector(VECSXP, 3));
  SEXP nm = PROTECT(allocVector(STRSXP, 3));
  SET_VECTOR_ELT(res, 0, vec);
  SET_VECTOR_ELT(res, 1, std);
  SET_VECTOR_ELT(res, 2, val);
  SET_STRING_ELT(nm, 0, mkChar("vectors"));
  SET_STRING_ELT(nm, 1, mkChar("standard_vectors"));
  SET_STRING_ELT(nm, 2, mkChar("values"));
  setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(5);
  return res;
}

/* spEigenCov(S, q, rho, thres) -> list(vectors, values, cov)                      R/spEigenCov.R:48,135 */
SEXP speigencov_call(SEXP S, SEXP q, SEXP rho, SEXP thres, SEXP device) {
  SEXP dim = getAttrib(S, R_DimSymbol);
  const int64_t nrow = INTEGER(dim)[0], ncol = INTEGER(dim)[1];
  const int cplx = isComplex(S);
  speigencov_cfg cfg;
  speigencov_cfg_default(&cfg);
  cfg.device = asInteger(device);
  speigencov_out out;
  memset(&out, 0, sizeof(out));
  SEXP vec = PROTECT(allocMatrix(cplx ? CPLXSXP : REALSXP, (int)ncol, (int)ncol));
  SEXP val = PROTECT(allocVector(REALSXP, (R_xlen_t)ncol));
  SEXP cov = PROTECT(allocMatrix(cplx ? CPLXSXP : REALSXP, (int)ncol, (int)ncol));
  out.vectors = cplx ? (double*)COMPLEX(vec) : REAL(vec);            /* Rcomplex = {double r, i} = interleaved */
  out.values = REAL(val);
  out.cov = cplx ? (double*)COMPLEX(cov) : REAL(cov);
  const int rc = cplx ? speigencov_run_c64((const double*)COMPLEX(S), nrow, ncol, asReal(q), asReal(rho), asReal(thres), &cfg, &out)
                      : speigencov_run_f64(REAL(S), nrow, ncol, asReal(q), asReal(rho), asReal(thres), &cfg, &out);
  if (rc != SPEIGEN_OK) {
    UNPROTECT(3);
    Rf_error("%s", speigen_last_error());
  }
  SEXP res = PROTECT(allocVector(VECSXP, 3));
  SEXP nm = PROTECT(allocVector(STRSXP, 3));
  SET_VECTOR_ELT(res, 0, vec);
  SET_VECTOR_ELT(res, 1, val);
  SET_VECTOR_ELT(res, 2, cov);
  SET_STRING_ELT(nm, 0, mkChar("vectors"));
  SET_STRING_ELT(nm, 1, mkChar("values"));
  SET_STRING_ELT(nm, 2, mkChar("cov"));
  setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(5);
  return res;
}

/* spEigen(cov(X), q, ...) with cov(X) formed on the device: X is the n x m data matrix */
SEXP speigen_cov_of_data_call(SEXP X, SEXP q, SEXP rho, SEXP d, SEXP V, SEXP thres, SEXP gpus) {
  SEXP dim = getAttrib(X, R_DimSymbol);
  const int64_t n = INTEGER(dim)[0], m = INTEGER(dim)[1];
  const int nq = (int)asReal(q);
  speigen_cfg cfg;
  speigen_cfg_default(&cfg);
  cfg.n_gpus = asInteger(gpus);
  speigen_out out;
  memset(&out, 0, sizeof(out));
  SEXP vec = PROTECT(allocMatrix(REALSXP, (int)m, nq));
  SEXP std = PROTECT(allocMatrix(REALSXP, (int)m, nq));
  SEXP val = PROTECT(allocVector(REALSXP, nq));
  out.vectors = REAL(vec);
  out.standard_vectors = REAL(std);
  out.values = REAL(val);
  const int rc = speigen_run_cov_of_data_f64(REAL(X), n, m, asReal(q), asReal(rho), isNull(d) ? NULL : REAL(d),
                                             isNull(V) ? NULL : REAL(V), asReal(thres), &cfg, &out);
  if (rc != SPEIGEN_OK) {
    UNPROTECT(3);
    Rf_error("%s", speigen_last_error());
  }
  SEXP res = PROTECT(allocVector(VECSXP, 3));
  SEXP nm = PROTECT(allocVector(STRSXP, 3));
  SET_VECTOR_ELT(res, 0, vec);
  SET_VECTOR_ELT(res, 1, std);
  SET_VECTOR_ELT(res, 2, val);
  SET_STRING_ELT(nm, 0, mkChar("vectors"));
  SET_STRING_ELT(nm, 1, mkChar("standard_vectors"));
  SET_STRING_ELT(nm, 2, mkChar("values"));
  setAttrib(res, R_NamesSymbol, nm);
  UNPROTECT(5);
  return res;
}

SEXP speigen_shutdown_call(void) {
  speigen_shutdown();
  return R_NilValue;
}

static const R_CallMethodDef calls[] = {
    {"C_speigen_call", (DL_FUNC)&speigen_call, 8},
    {"C_speigencov_call", (DL_FUNC)&speigencov_call, 5},
    {"C_speigen_cov_of_data_call", (DL_FUNC)&speigen_cov_of_data_call, 7},
    {"C_speigen_shutdown_call", (DL_FUNC)&speigen_shutdown_call, 0},
    {NULL, NULL, 0}};

void R_init_sparseEigen(DllInfo* dll) {
  R_registerRoutines(dll, NULL, calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}

/* library.dynam.unload(): stop the reaper thread of the kept-alive solver and give the device / pinned memory back
 * (also reachable from R through .onUnload -> C_speigen_shutdown_call) */
void R_unload_sparseEigen(DllInfo* dll) {
  (void)dll;
  speigen_shutdown();
}
