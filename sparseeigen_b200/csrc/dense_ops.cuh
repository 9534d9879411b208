// Dense m x m FP64 kernels of the spEigenCov() path (R/spEigenCov.R:48-160) and of the on-device cov(X):
//   * a TMA-staged FP64 tensor-core (DMMA) GEMM  C = alpha * op(A) * op(B)  (replaces the m x m x m `%*%` of
//     R/spEigenCov.R:106,114,135,154 and the `cov(X)` / `t(X) %*% Conj(X)` callers run before spEigen());
//   * a block one-sided Jacobi SVD (batched panel Gram -> batched 64 x 64 Jacobi -> batched panel rotation) from which
//     both the full `eigen(S)` of :62 and the Procrustes/polar update `svd()` + `u %*% h(v)` of :157-159 are built;
//   * the column-major elementwise / column-reduction kernels of the iteration (:147-153, :110-114, :131-133).
// All matrices here are column-major (R's layout) with a padded leading dimension.
#pragma once
#include "common.cuh"

namespace speigen {

// ---- GEMM -------------------------------------------------------------------------------------------------------
// Operand storage: `kmajor` != 0 -> the contracted index is contiguous (A given as a K x M column-major matrix, i.e.
// op(A) = A^T); `kmajor` == 0 -> the free index is contiguous (A given as M x K column-major, op(A) = A).
// For B: kmajor -> K x N column-major (op(B) = B); otherwise N x K column-major (op(B) = B^T).
struct GemmOperand {
  const double* ptr = nullptr;
  int64_t ld = 0;      // leading dimension in doubles (even; base 16-byte aligned)
  int kmajor = 0;
};
struct GemmEpilogue {
  double alpha = 1.0;
  double beta = 0.0;          // C = alpha*acc + beta*C
  int symmetric = 0;          // 1: op(A)*op(B) is known symmetric (A == B): only tiles on/above the diagonal are computed, mirrored on store
};
constexpr int kGemmTM = 128, kGemmTN = 128, kGemmTK = 32;
int launch_gemm(int64_t M, int64_t N, int64_t K, const GemmOperand& A, const GemmOperand& B, double* C, int64_t ldc,
                const GemmEpilogue& epi, int num_sms, cudaStream_t st);

// ---- block one-sided Jacobi SVD ------------------------------------------------------------------------------------
constexpr int kJB = 32;             // columns per block; a panel is two adjacent blocks
constexpr int kJP = 2 * kJB;        // panel width (64)
constexpr int kJRows = 128;         // rows per tile in the panel kernels
struct JacobiWork {
  double* gram_part = nullptr;      // [max_panels][nsplit][kJP*kJP]
  double* wp = nullptr;             // [max_panels][kJP*kJP] rotation of each panel (column-major, block swap folded in)
  double* offmax = nullptr;         // device [2]: max relative off-diagonal seen since the last reset; hand-off timeout flag
  int nsplit = 1;
  int inner_sweeps = 0;             // cap on the 64 x 64 Jacobi sweeps per panel visit (0 = run to convergence)
  // The rotation of the accumulated right factor W is off the critical path (the next round's Gram needs only X): with a
  // side stream it runs concurrently with the next round's Gram / 64 x 64 Jacobi (which occupies 32 of 148 SMs).
  // split 64 x 64 Jacobi (k_jeig_split): per-panel rotation list [63][32]{c, s} and handshake flags (zero-initialised)
  double* rot = nullptr;
  int* flags = nullptr;
  int split_threads = 1024;        // CTA size of k_jeig_split (>= 544: the solver updates 528 blocks; 16 or 32 builder warps)
  int publish_every = 8;           // rounds per publish of the rotation list (power of two; 64 = once per sweep)
  cudaStream_t side = nullptr;
  static constexpr int kWpSlots = 4;             // ring of rotation buffers so jeig(r+1..r+3) never overwrites an unconsumed W_p
  size_t wp_stride = 0;                          // doubles per slot
  cudaEvent_t ev_eig[kWpSlots] = {};
  cudaEvent_t ev_applied[kWpSlots] = {};
};
// One round of the odd-even block ordering on X (mp x mp, ld) and the accumulated right factor W (same shape):
// parity 0 pairs blocks (0,1),(2,3),..; parity 1 pairs (1,2),(3,4),...
int launch_jacobi_round(double* X, double* W, int64_t mp, int64_t ld, int parity, const JacobiWork& jw, cudaStream_t st,
                        int round_index = 0);

// ---- column-major helpers --------------------------------------------------------------------------------------------
// out[j] = sum_i A[i,j] * B[i,j]   (one block per column, fixed order; B == A gives squared column norms)
int launch_cm_coldots(const double* A, const double* B, int64_t rows, int64_t cols, int64_t ld, double* out, cudaStream_t st);
// out[j] = min_i |A[i,j]|, i < rows
int launch_cm_colabsmin(const double* A, int64_t rows, int64_t cols, int64_t ld, double* out, cudaStream_t st);
// out[j] = sum_i g(|A[i,j]|), the penalty of R/spEigenCov.R:110-112
int launch_cm_penalty(const double* A, int64_t rows, int64_t cols, int64_t ld, StageConsts sc, double* out, cudaStream_t st);
// A[:, j] *= s[j]   (inv != 0: A[:, j] /= s[j])
int launch_cm_scale_cols(double* A, int64_t rows, int64_t cols, int64_t ld, const double* s, int inv, cudaStream_t st);
// Out[:, j] = In[:, perm[j]]
int launch_cm_gather_cols(const double* In, int64_t rows, int64_t cols, int64_t ld, const int* perm, double* Out, cudaStream_t st);
// A = diag-padded copy: Dst (mp x mp, ld) <- Src (m x m, lds) in the top-left block, `pad_diag` on the padded diagonal, 0 elsewhere;
// `shift` is subtracted from the first m diagonal entries (S_hat = S - shift*I, R/spEigenCov.R:69)
int launch_cm_pad_copy(const double* Src, int64_t m, int64_t lds, double* Dst, int64_t mp, int64_t ld, double shift,
                       double pad_diag, cudaStream_t st);
int launch_cm_identity(double* A, int64_t mp, int64_t ld, cudaStream_t st);
// Procrustes target of R/spEigenCov.R:147-157:  B = -(H1 + T*diag(1/Xi)),  H1[:, j<q] = (w(|V|) - w(absmin_j)) * V * rho_j;
// padded diagonal = 1
int launch_cm_build_target(const double* V, const double* T, const double* inv_xi, const double* rho, const double* absmin,
                           int q, StageConsts sc, int64_t m, int64_t mp, int64_t ld, double* B, cudaStream_t st);
// M = 1.5*I - 0.5*G   (Newton-Schulz polish of the polar factor)
int launch_cm_ns_matrix(const double* G, int64_t mp, int64_t ld, double* M, cudaStream_t st);
// threshold (strict <) and squared column norms of the thresholded matrix: nrm2[j]   (R/spEigenCov.R:131-132)
int launch_cm_thres_norms(const double* V, int64_t m, int64_t ld, double thres, double* nrm2, cudaStream_t st);
// Out[i,j] = (|V[i,j]| < thres ? 0 : V[i,j]) * rs[i] * cs[j]   (rs/cs nullable); Out is m x m with ldo
int launch_cm_thres_scale(const double* V, int64_t m, int64_t ld, double thres, const double* rs, const double* cs,
                          double* Out, int64_t ldo, cudaStream_t st);
// symmetry / NaN scan of a column-major m x m matrix: out[0] = sum|S - S^T|, out[1] = sum|S|, out[2] = #NaN, out[3] = sum S^2
int launch_cm_scan_sym(const double* S, int64_t m, int64_t ld, double* part, double* out4, cudaStream_t st);
// column sums of an n x m column-major matrix (one block per column)
int launch_cm_colsums(const double* X, int64_t n, int64_t m, int64_t ld, double* out, cudaStream_t st);
// X[:, j] -= mean[j]
int launch_cm_center(double* X, int64_t n, int64_t m, int64_t ld, const double* sums, double inv_n, cudaStream_t st);


// ---- complex matrices in real embedded form ---------------------------------------------------------------------------
// A complex m x m matrix C = Cr + i Ci is carried as the real 2h x 2h matrix [[Cr, -Ci], [Ci, Cr]] (h = m rounded up to 64;
// rows/columns [0, h) hold real parts, [h, 2h) imaginary parts).  The embedding is a *-homomorphism: products, Hermitian
// transposes, unitary / Hermitian-positive-definite factors and therefore polar factors map onto their real counterparts,
// so the real GEMM and Jacobi kernels serve complex input unchanged.
// Dst (2h x 2h, ld) <- embed(Src - shift*I), Src = m x m column-major interleaved {re, im}; padded diagonal = pad_diag
int launch_ce_embed(const double* Src_interleaved, int64_t m, double* Dst, int64_t h, int64_t ld, double shift,
                    double pad_diag, cudaStream_t st);
// V (embedded) <- complex eigenvectors x + i y taken from the real eigenvectors [x; y] = W[:, perm[j]], j < m; unit vectors beyond
int launch_ce_select_eigvecs(const double* W, const int* perm, int64_t m, int64_t h, int64_t ld, double* V, cudaStream_t st);
// out[j] = min_i |V[i,j]| (complex modulus), j < cols
int launch_ce_colabsmin(const double* V, int64_t m, int64_t h, int64_t cols, int64_t ld, double* out, cudaStream_t st);
// out[j] = sum_i g(|V[i,j]|)
int launch_ce_penalty(const double* V, int64_t m, int64_t h, int64_t cols, int64_t ld, StageConsts sc, double* out, cudaStream_t st);
// embedded Procrustes target (R/spEigenCov.R:147-157 for complex V)
int launch_ce_build_target(const double* V, const double* T, const double* inv_xi, const double* rho, const double* absmin,
                           int q, StageConsts sc, int64_t m, int64_t h, int64_t ld, double* B, cudaStream_t st);
// nrm2[j] = sum_i |v_ij|^2 over entries with |v_ij| >= thres
int launch_ce_thres_norms(const double* V, int64_t m, int64_t h, int64_t ld, double thres, double* nrm2, cudaStream_t st);
// Out (embedded) <- embed( (|v| < thres ? 0 : v) * rs[i] * cs[j] ), zero padding
int launch_ce_thres_scale(const double* V, int64_t m, int64_t h, int64_t ld, double thres, const double* rs, const double* cs,
                          double* Out, cudaStream_t st);
// Dst (m x m interleaved complex, column-major) <- the complex matrix carried by the embedded Src
int launch_ce_extract(const double* Src, int64_t m, int64_t h, int64_t ld, double* Dst_interleaved, cudaStream_t st);

}  // namespace speigen
