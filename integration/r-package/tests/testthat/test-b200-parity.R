# Extra testthat file for a machine that has R, a B200 and the library built (not runnable in the build image).
# The reference's own tests/testthat/test-sparseEigen.R runs unchanged against the replaced closures (it pins error
# raising and output lengths); this file adds what it does not pin: agreement with the pure-R implementation.
# `sparseEigenRef` is the CRAN build of the reference installed under another name (R CMD INSTALL --library=ref).
context("B200 drop-in vs the pure-R reference")

spiked <- function(m, n, q, card, seed = 42) {
  set.seed(seed)
  sp <- card * m
  V <- matrix(0, m, q)
  V[cbind(seq(1, q * sp), rep(1:q, each = sp))] <- 1 / sqrt(sp)
  Z <- matrix(rnorm(n * m), n, m)
  Z + (Z %*% V) %*% diag(sqrt(100 * seq(q, 1)) - 1, q) %*% t(V)
}

test_that("spEigen agrees with the reference on covariance and data input", {
  skip_if_not_installed("sparseEigenRef")
  X <- spiked(500, 100, 3, 0.1)
  for (data in c(FALSE, TRUE)) {
    A <- if (data) X else cov(X)
    ref <- sparseEigenRef::spEigen(A, 3, data = data)
    new <- spEigen(A, 3, data = data)
    expect_equal(new$values, ref$values, tolerance = 1e-11)
    expect_identical(new$vectors != 0, ref$vectors != 0)
    expect_true(all(abs(colSums(new$vectors * ref$vectors)) > 1 - 1e-5))
  }
})

test_that("spEigenCov agrees with the reference", {
  skip_if_not_installed("sparseEigenRef")
  S <- cov(spiked(200, 600, 3, 0.1))
  ref <- sparseEigenRef::spEigenCov(S, 3)
  new <- spEigenCov(S, 3)
  expect_equal(new$values, ref$values, tolerance = 1e-9)
  expect_identical(new$vectors[, 1:3] != 0, ref$vectors[, 1:3] != 0)
  expect_true(all(abs(colSums(new$vectors[, 1:3] * ref$vectors[, 1:3])) > 1 - 1e-8))
  expect_equal(new$cov, ref$cov, tolerance = 1e-6)
})

test_that("spEigenOfCov(X) equals spEigen(cov(X))", {
  X <- spiked(500, 100, 3, 0.1)
  a <- spEigen(cov(X), 3)
  b <- spEigenOfCov(X, 3)
  expect_equal(a$values, b$values, tolerance = 1e-11)
  expect_identical(a$vectors != 0, b$vectors != 0)
})
