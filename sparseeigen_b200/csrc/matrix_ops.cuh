// One-time streaming passes over the resident matrix: validation (R/spEigen.R:53,75), centring (:59-60)
// and the rho scale inputs (:72,:82).
#pragma once
#include "common.cuh"

namespace speigen {

constexpr int kScanBlocks = 148 * 4;

// Covariance shard scan.  A: rows x (m * (1+cplx)) doubles, pitch ld; local row jl is global column lo+jl of S.
// full_check != 0: out[0] = sum|S - S^H| over this shard's rows, out[1] = sum|S| (the transposed entries are read from
// the shard that owns them through shard_ptrs[] - peer-mapped pointers when the matrix is row-sharded over several GPUs
// of one process; nullptr for a single shard).  Always: out[2] = max Re diag, out[3] = number of NaN entries (as double).
int launch_scan_cov(const double* A, int64_t rows, int64_t m, int64_t ld, int64_t lo, int cplx, int full_check,
                    const double* const* shard_ptrs, int nshards, int64_t rows_per_shard, double* part, int* ticket,
                    double* out4, cudaStream_t st);
// Hermitian-half upload, device side: rebuild every element whose super-tile (edge `block`, host_pack.h) was not uploaded as the
// conjugate of its transposed partner (shard_ptrs as in launch_scan_cov).
int launch_mirror_cov(double* A, int64_t rows, int64_t m, int64_t ld, int64_t lo, int cplx, int64_t block,
                      const double* const* shard_ptrs, int nshards, int64_t rows_per_shard, cudaStream_t st);
// Data shard (rows = variables, contiguous = local samples).  sums[j*(1+cplx)+e] = sum_i A[j][i]
int launch_row_sums(const double* A, int64_t rows, int64_t nloc, int64_t ld, int cplx, double* sums, cudaStream_t st);
// A[j][i] -= mean[j] ; sq[j] = sum_i |A[j][i]|^2   (mean = sums / n_total)
int launch_center_sqnorm(double* A, int64_t rows, int64_t nloc, int64_t ld, int cplx, const double* sums,
                         double inv_n, double* sq, cudaStream_t st);
// out[0] = max_j v[j], out[1] = count of NaN in v
int launch_vec_max(const double* v, int64_t n, double* out2, cudaStream_t st);

}  // namespace speigen
