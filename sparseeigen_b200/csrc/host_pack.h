// Host side of the Hermitian-half upload (solver.cu, Solver::load_host_half).
//
// A covariance matrix is Hermitian (R/spEigen.R:75 stops otherwise), so only one tile of every transposed pair has to cross
// PCIe: the host workers copy the chosen tile into the pinned staging ring and - while the two tiles are in cache anyway -
// evaluate isSymmetric.matrix's sums  sum|S - S^H| , sum|S|  against the partner tile, which never leaves the host.  The device
// rebuilds the other tile by a transposed copy (k_mirror_cov, matrix_ops.cu).  Nothing here touches CUDA: the planner and the
// tile compare are exercised by the CPU tests through speigen_host_half_upload_emulate().
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace speigen {

// Square super-tiles of kHalfBlock x kHalfBlock elements decide which side of a transposed pair is uploaded.
constexpr int64_t kHalfBlock = 2048;

// Is the tile holding element (row, col) one that crosses PCIe?  Diagonal super-tiles always do; of an off-diagonal pair
// (I, J), (J, I) exactly one does, alternating with the parity of I + J so that every row block (= every GPU of a row-sharded
// matrix) uploads about half of its tiles.
#if defined(__CUDACC__)
__host__ __device__
#endif
inline bool half_tile_uploaded(int64_t row, int64_t col, int64_t block) {
  const int64_t I = row / block, J = col / block;
  if (I == J) return true;
  return (((I + J) & 1) == 0) ? (I < J) : (I > J);
}

struct HalfUnit {
  int dev;             // index of the device whose row range holds the rows
  int64_t r0, nr;      // global rows of A (= columns of S), elements
  int64_t c0, nc;      // columns, elements
  bool diag;           // the unit lies in a diagonal super-tile (its partner is inside the same tile)
};

// Cut the uploaded tiles of every device's row range [lo, hi) into units of whole rows that fit `slot_bytes`.
// `e` = doubles per element (1 real, 2 complex).
std::vector<HalfUnit> plan_half_upload(int64_t m, int e, const std::vector<std::pair<int64_t, int64_t>>& dev_ranges,
                                       int64_t block, size_t slot_bytes);

// Sums of isSymmetric.matrix over one unit.  a: nr x nc elements, row pitch lda doubles (the unit, e.g. the staged copy);
// b: the partner, nc x nr elements, row pitch ldb doubles.  Adds  sum |a_ij - conj(b_ji)|  to acc[0],  sum |a_ij|  to acc[1]
// and  sum |b_ji|  to acc[2].  `scratch` holds at least 2 * 128 * 128 doubles.
void hermitian_unit_sums(const double* a, size_t lda, const double* b, size_t ldb, int64_t nr, int64_t nc, int e,
                         double* scratch, double acc[3]);
// The worker's inner step: walk the unit in 128 x 128 sub-tiles; `stage` != nullptr: copy every sub-tile of `a` into the dense
// staging slot (row pitch nc * e doubles, streaming stores) while it is in cache; `compare`: accumulate the sums as above.
// `stage` without `compare` (the one-shot call defers the compare to background workers) copies whole rows instead.
void half_unit_pack_and_sums(const double* a, size_t lda, const double* b, size_t ldb, int64_t nr, int64_t nc, int e,
                             double* stage, bool compare, double* scratch, double acc[3]);
// same arithmetic without SIMD (reference for the tests; also the fallback on CPUs without AVX2)
void hermitian_unit_sums_scalar(const double* a, size_t lda, const double* b, size_t ldb, int64_t nr, int64_t nc, int e,
                                double acc[3]);
bool host_has_avx2();
// What a unit's sums contribute to  diff = sum_ij |S_ij - conj(S_ji)|  and  sabs = sum_ij |S_ij|  over the whole matrix: a unit of
// an off-diagonal tile stands for its partner too; the units of a diagonal tile visit every element of the tile exactly once.
inline void half_unit_contribution(const HalfUnit& u, const double acc[3], double* diff, double* sabs) {
  *diff = u.diag ? acc[0] : 2.0 * acc[0];
  *sabs = u.diag ? acc[1] : acc[1] + acc[2];
}
// The whole procedure on the host (no GPU): pack every unit through a staging buffer into A_out (m x m elements, the layout of
// S itself), rebuild the tiles that were not "uploaded" by the transposed conjugate copy, and return isSymmetric's sums.
// A_out must not alias S.  Used by the CPU tests; the product path is Solver::load_host_half + k_mirror_cov.
void emulate_half_upload(const double* S, int64_t m, int e, int world, int64_t block, size_t slot_bytes, int use_simd,
                         double* A_out, double* diff, double* sabs);

}  // namespace speigen
