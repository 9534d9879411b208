// m x q panel kernels of the spEigen MM iteration (everything except the streamed contraction):
// reweighting (R/spEigen.R:170-176), Gram/polar building blocks (:130-131,:180-182), SQUAREM
// extrapolation (:121-127), penalty/objective (:133-138) and the final threshold (:158-160).
// All reductions are two-stage with a fixed summation order -> bit-reproducible, also across GPUs.
#pragma once
#include "common.cuh"

namespace speigen {

constexpr int kPanelBlocks = 148;   // stage-1 reduction blocks (one per SM)
constexpr int kTileRows = 32;       // rows staged per tile in the panel kernels

struct PanelDims {
  int64_t m = 0;       // valid rows
  int q = 0;           // logical columns
  int cplx = 0;
  int ncr = 0;         // real columns = q * (1 + cplx)
  int pitch = 0;       // doubles
  int nblk = kPanelBlocks;   // stage-1 blocks actually launched for this panel (small m: fewer partials to sum)
};
inline PanelDims make_dims(int64_t m, int q, int cplx) {
  PanelDims d;
  d.m = m; d.q = q; d.cplx = cplx; d.ncr = q * (1 + cplx); d.pitch = panel_pitch(d.ncr);
  const int64_t nb = (m + 127) / 128;
  d.nblk = static_cast<int>(nb < 1 ? 1 : (nb > kPanelBlocks ? kPanelBlocks : nb));
  return d;
}
inline int64_t padded_rows(int64_t m) { return (m + kRowPad - 1) / kRowPad * kRowPad; }

// Small-matrix solver modes (K3).
enum SmallMode {
  SMALL_INVSQRT = 0,   // M = W diag(lambda^-1/2) W^H         (polar / orthonormalisation)
  SMALL_EIGVECS = 1,   // M = W sorted by decreasing lambda, evals out (Rayleigh-Ritz)
};

// Scratch shared by the panel kernels of one device.
struct PanelScratch {
  double* gram_part = nullptr;   // [kPanelBlocks][gram_len]
  double* col_part = nullptr;    // [kPanelBlocks][2*kMaxCols]
  int* ticket = nullptr;         // last-block-done counter (self-resetting)
};

int launch_gram(const PanelDims& pd, const double* A, const double* B, double* gram_part, cudaStream_t st);
// warm (nullable): {q, count, W0[q*q*(1+cplx)]} per call site - Jacobi warm start from the previous eigenvectors
int launch_small(const PanelDims& pd, const double* gram_part, int nparts, int mode, double* Msmall, double* evals,
                 int* status, double* warm, cudaStream_t st);
// Out = A * Msmall ; optional fused stage-1 reductions of Out: Gram partials and column min|.| partials.
int launch_apply(const PanelDims& pd, const double* A, const double* Msmall, double* Out, double* gram_part_out,
                 double* absmin_part_out, int* ticket, cudaStream_t st);
// B = d*Y - (w(|V|) - w(amin)) * V * rho  (+ Gram partials of B); `wmax` carries the column minima of |V|
int launch_reweight(const PanelDims& pd, const double* Y, const double* V, const double* d, const double* rho,
                    const double* wmax, StageConsts sc, double* B, double* gram_part_out, cudaStream_t st);
// Out = V - 2 a R + a^2 U,  R = V1 - V, U = V2 - V1 - R   (+ Gram partials of Out)   R/spEigen.R:121-127
int launch_extrapolate(const PanelDims& pd, const double* V, const double* V1, const double* V2, double a, double* Out,
                       double* gram_part_out, cudaStream_t st);
// column min|.| partials of a panel, and wmax[c] = w(min_j |V[j,c]|) from them (R/spEigen.R:176 `apply(w,2,max)`).
// absmin_part: [kPanelBlocks + 1][kMaxCols]; the last row receives the final column minima (last-block reduction)
int launch_absmin(const PanelDims& pd, const double* V, double* absmin_part, int* ticket, cudaStream_t st);
inline const double* absmin_final(const double* absmin_part) { return absmin_part + static_cast<size_t>(kPanelBlocks) * kMaxCols; }
// norms[0..1] = {||V1 - V||_F, ||V2 - 2 V1 + V||_F}
int launch_squarem_norms(const PanelDims& pd, const double* V, const double* V1, const double* V2,
                         const PanelScratch& sc, double* norms, cudaStream_t st);
// out[0] = F = sum_c d_c quad_c - sum_c rho_c gsum_c, out[1..q] = quad, out[1+q..2q] = gsum   (R/spEigen.R:133-138)
// quad_c = fixed-order sum of the contraction's per-row-tile dot partials.
int launch_objective(const PanelDims& pd, const double* V0, StageConsts stc, const double* dots_part, int n_dot_parts,
                     const double* d, const double* rho, const PanelScratch& sc, double* out, cudaStream_t st);
// threshold (strict <) + column normalisation (R/spEigen.R:158-160), written column-major
int launch_finalize(const PanelDims& pd, const double* V, double thres, const PanelScratch& sc, double* norms_tmp,
                    double* out_colmajor, cudaStream_t st);
int launch_panel_to_colmajor(const PanelDims& pd, const double* P, double* out_colmajor, cudaStream_t st);
int launch_colmajor_to_panel(const PanelDims& pd, const double* in_colmajor, double* P, cudaStream_t st);
// complex operand expansion for the real DMMA kernel: Vhat[2i] = V[i], Vhat[2i+1] = -+ i*V[i]
int launch_hat(const PanelDims& pd, const double* V, double* Vhat, int conj_form, cudaStream_t st);
// complex transposed-form fix-up: Pm (2n x 2q real products) -> That (hat-expanded complex T, conj form); in place ok
int launch_tr_combine(int64_t n, int q, const double* Pm, double* That, cudaStream_t st);
// deterministic pseudo-random start block for the subspace iteration
int launch_random_panel(const PanelDims& pd, uint64_t seed, double* P, cudaStream_t st);
// out[c] = || ZW[:,c] - evals[c] * QW[:,c] ||_2
int launch_residual_norms(const PanelDims& pd, const double* ZW, const double* QW, const double* evals,
                          const PanelScratch& sc, double* out, cudaStream_t st);
// out[c] = max_j Re(diag) helper etc. live in the driver.
int launch_axpby(const PanelDims& pd, const double* X, const double* Y, double alpha, double beta, double* Out, cudaStream_t st);
int launch_axpbypcz(const PanelDims& pd, const double* X, const double* Y, const double* Z, double a, double b, double c,
                    double* Out, cudaStream_t st);
int launch_scale_columns(const PanelDims& pd, double* P, const double* scale, cudaStream_t st);
// copy the first q_out logical columns of a (wider) panel into a q_out-column panel
int launch_take_columns(const PanelDims& pd_in, const double* In, const PanelDims& pd_out, double* Out, cudaStream_t st);
// deflated initialiser: Out = first `keep` logical columns of In (zeros beyond); Y[:, :keep] = X[:, :keep];
// Mh = hat(A^H B) from the cross-Gram partials of launch_gram(A, B)
int launch_mask_columns(const PanelDims& pd, const double* In, int keep_logical, double* Out, cudaStream_t st);
int launch_blend_columns(const PanelDims& pd, double* Y, const double* X, int keep_logical, cudaStream_t st);
int launch_gram_sum_hat(const PanelDims& pd, const double* gram_part, int nparts, double* Mh, cudaStream_t st);
// fixed-order sum of nparts partial panels (data mode, multi-GPU): Out = sum_r In[r]
int launch_sum_panels(const PanelDims& pd, const double* In, int nparts, size_t part_stride, double* Out, cudaStream_t st);

// per-block partial column dots Re(conj(V).*Y): dots_part[kPanelBlocks][q]
int launch_coldots(const PanelDims& pd, const double* V, const double* Y, double* dots_part, cudaStream_t st);
// out[i] = sum_k parts[k][i], fixed order (len <= 128)
int launch_sum_parts(const double* parts, int nparts, int len, double* out, cudaStream_t st);

}  // namespace speigen
