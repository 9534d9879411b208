# Drop-in replacement body for sparseEigen::spEigenCov() (reference: R/spEigenCov.R:48-136).
# Same signature, defaults, stop() conditions (in the reference's order) and return list(vectors, values, cov);
# the numerics (eigen(S), the m x m Procrustes updates, the m x m x m products) run in libspeigen_b200.so.
# NOT testable in the build image (no R); the same C ABI is exercised from Python in tests/test_gpu_cov.py.
spEigenCov <- function(S, q = 1, rho = 0.5, thres = 1e-9) {
  S <- as.matrix(S)                                                                          # R/spEigenCov.R:52
  m <- ncol(S)
  if (m == 1) stop("Data is univariate!")                                                    # :54
  if (q > m) stop("The number of sparse eigenvectors q should not be larger than m.")        # :55
  if (anyNA(S) || anyNA(q) || anyNA(rho)) stop("This function cannot handle NAs.")           # :56
  if (!isSymmetric.matrix(S)) stop("The covariance matrix is not symmetric")                 # :57
  if ((q %% 1) != 0 || q <= 0) stop("The input argument q should be a positive integer.")    # :58
  if (rho <= 0) stop("The input argument rho should be positive.")                           # :59
  if (!is.complex(S)) storage.mode(S) <- "double"
  # "not PSD" (:63) and "low-rank" (:65) are raised by the library from the device eigen-decomposition
  .Call(C_speigencov_call, S, as.double(q), as.double(rho), as.double(thres),
        as.integer(getOption("sparseEigen.device", 0L)))
}

# spEigen(cov(X), q, ...) for a real n x m data matrix X without ever forming cov(X) on the host
# (README.Rmd:108 calls spEigen(cov(X), q); forming cov(X) in R costs 2 n m^2 flops on the CPU and 8 m^2 bytes over PCIe).
spEigenOfCov <- function(X, q = 1, rho = 0.5, d = NA, V = NA, thres = 1e-9) {
  X <- as.matrix(X)
  if (ncol(X) == 1) stop("Data is univariate!")
  if (anyNA(X) || anyNA(q) || anyNA(rho)) stop("This function cannot handle NAs.")
  if ((q %% 1) != 0 || q <= 0) stop("The input argument q should be a positive integer.")
  if (rho <= 0) stop("The input argument rho should be positive.")
  d <- if (length(d) == 1 && is.na(d)) NULL else as.double(d)
  V <- if (length(V) == 1 && is.na(V)) NULL else V
  storage.mode(X) <- "double"
  .Call(C_speigen_cov_of_data_call, X, as.double(q), as.double(rho), d, V, as.double(thres),
        as.integer(getOption("sparseEigen.gpus", 1L)))
}
