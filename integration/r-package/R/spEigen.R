# Drop-in replacement body for sparseEigen::spEigen() (reference: R/spEigen.R:46-163).
# Same signature, defaults, stop() conditions and return list; the numerics run in libspeigen_b200.so.
# NOT testable in the build image (no R); the same C ABI is exercised from Python in tests/.
spEigen <- function(X, q = 1, rho = 0.5, data = FALSE, d = NA, V = NA, thres = 1e-9) {
  X <- as.matrix(X)                                                                          # R/spEigen.R:50
  m <- ncol(X)
  if (m == 1) stop("Data is univariate!")                                                    # :52
  if (anyNA(X) || anyNA(q) || anyNA(rho)) stop("This function cannot handle NAs.")           # :53
  if ((q %% 1) != 0 || q <= 0) stop("The input argument q should be a positive integer.")    # :54
  if (rho <= 0) stop("The input argument rho should be positive.")                           # :55
  d <- if (length(d) == 1 && is.na(d)) NULL else as.double(d)      # NA -> seq(1, 0.5, length.out = q)   (:63-64)
  V <- if (length(V) == 1 && is.na(V)) NULL else V                 # NA -> Vx[, 1:q]                     (:86-87)
  if (!is.null(V)) storage.mode(V) <- if (is.complex(X)) "complex" else "double"
  if (!is.complex(X)) storage.mode(X) <- "double"
  .Call(C_speigen_call, X, as.double(q), as.double(rho), as.logical(data), d, V, as.double(thres),
        as.integer(getOption("sparseEigen.gpus", 1L)))
}

.onUnload <- function(libpath) {
  .Call(C_speigen_shutdown_call)
  library.dynam.unload("sparseEigen", libpath)
}
