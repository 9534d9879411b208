// spEigenCov() on one B200 and the on-device cov(X) (§8(f)-2).
//
// Reference loop (R/spEigenCov.R:92-128), per MM iteration: A = V diag(1/Xi) (:100); Procrustes update
// V = polar(-(H1 + S_hat A)) through a full m x m svd() (:146-160); eigenvalue update from Re diag(-A^H S_hat A) of the
// OLD A through eigvalAlgo (:106-107, :164-338); objective (:110-114).  Here:
//   * T = S_hat V is ONE m x m x m DMMA GEMM per iteration (dense_ops.cu); it serves the objective of this iteration
//     (diag(V^T S V) = coldots(V, T) + shift), the next iteration's S_hat A = T diag(1/Xi) and its diag(A^T S_hat A);
//   * the m x m SVD / the initial eigen(S) are a block one-sided Jacobi on the device (batched DMMA panel Gram ->
//     batched 64 x 64 Jacobi -> batched DMMA panel rotation), the polar factor is (X Sigma^-1) W^T (one more GEMM) plus
//     one Newton-Schulz polish step (two GEMMs);
//   * eigvalAlgo is a short sequential pooling algorithm over m scalars: it stays on the host.
// No CPU fallback: every m x m operation runs on the GPU; the host sees O(m) scalars per iteration.
#include "dense_ops.cuh"
#include "contract.cuh"
#include "solver.cuh"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

namespace speigen {

#define SPE_TRY2(x)                      \
  do {                                   \
    int rc__ = (x);                      \
    if (rc__ != SPEIGEN_OK) return rc__; \
  } while (0)

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------------------
// eigvalAlgo (R/spEigenCov.R:164-338): ordered eigenvalue update.  g is 1-based inside (index 0 unused).
// ---------------------------------------------------------------------------------------------------------------
namespace {

// `1/length(ind) * sum(g[ind])` as R evaluates it: sum() accumulates in long double in index order, rounds to double,
// then multiplies by the reciprocal of the count.
double r_mean(const std::vector<double>& g, const std::vector<int>& idx) {
  long double acc = 0.0L;
  for (int t : idx) acc += static_cast<long double>(g[t]);
  return (1.0 / static_cast<double>(idx.size())) * static_cast<double>(acc);
}
double r_mean_range(const std::vector<double>& g, int i, int j) {
  long double acc = 0.0L;
  for (int t = i; t <= j; ++t) acc += static_cast<long double>(g[t]);
  return (1.0 / static_cast<double>(j - i + 1)) * static_cast<double>(acc);
}

// The inner `while (1)` shared by R/spEigenCov.R:221-239 and :257-274: absorb the smallest remaining violators into the
// pooled set until its mean exceeds none of the remaining ones, then assign the mean to the pooled entries.
void pool_violators(std::vector<double>& g, std::vector<double>& g_new, std::vector<int> ind, std::vector<int> red) {
  auto mean_of = [&](const std::vector<int>& v) { return r_mean(g, v); };
  double tmp = 0.0;
  for (;;) {
    if (!red.empty()) {
      double mn = g[red[0]];
      for (int t : red) mn = std::min(mn, g[t]);
      std::vector<int> rest;
      for (int t : red) {
        if (g[t] == mn) {
          if (std::find(ind.begin(), ind.end(), t) == ind.end()) ind.push_back(t);
        } else {
          rest.push_back(t);
        }
      }
      red.swap(rest);
      tmp = mean_of(ind);
    }
    bool done = true;
    if (!red.empty())
      for (int t : red)
        if (tmp > g[t]) { done = false; break; }
    if (done) {
      const double mu = mean_of(ind);
      for (int t : ind) g_new[t] = mu;
      return;
    }
  }
}

bool first_q_in_order(const std::vector<double>& g, int64_t m, int q) {
  // sum(order(g)[1:q] != seq(1:q)) == 0 with R's stable order(): entry i (1-based, i <= q) must be the i-th smallest,
  // ties resolved by index
  for (int i = 1; i <= q; ++i) {
    int64_t smaller = 0;
    for (int64_t t = 1; t <= m; ++t) {
      if (t == i) continue;
      if (g[t] < g[i] || (g[t] == g[i] && t < i)) ++smaller;
    }
    if (smaller != i - 1) return false;
  }
  return true;
}

}  // namespace

int eigval_algo_host(const double* g_in, int64_t m, double x, int q, int strict_r, double* xi_out) {
  if (q < 1 || q >= m) { set_error("eigvalAlgo needs 1 <= q < m"); return SPEIGEN_ERR_ARG; }
  std::vector<double> g(m + 1), g_new;
  g[0] = 0.0;
  for (int64_t t = 0; t < m; ++t) g[t + 1] = g_in[t];
  g_new = g;
  int mult = 0;
  auto violators = [&](std::vector<int>& out, bool equal_only) {
    out.clear();
    for (int64_t t = q + 1; t <= m; ++t)
      if (equal_only ? (g[q] == g[t]) : (g[q] >= g[t])) out.push_back(static_cast<int>(t));
  };
  std::vector<int> swap_ind, ind, eq;
  constexpr int kMaxPasses = 10000;   // the reference's `while (1)` (:171,:290) has no exit on some inputs
  bool converged = false;
  if (q > 1) {
    for (int pass = 0; pass < kMaxPasses; ++pass) {             // :171
      int flg = 0;
      int i = 1;
      while (i <= q - 1) {                                      // :177
        if (g[i] >= g[i + 1]) {
          if (i == q - 1) flg = 1;
          int j = i + 1;                                        // :185 block of consecutive violations
          while (j <= q - 1) {
            if (g[j] >= g[j + 1]) {
              if (j == q - 1) flg = 1;
              ++j;
            } else {
              break;
            }
          }
          if (flg == 0) {                                       // :198-200
            const double mu = r_mean_range(g, i, j);
            for (int t = i; t <= j; ++t) g_new[t] = mu;
          } else {                                              // :202-242 the block reaches q: look at q+1..m
            violators(swap_ind, false);
            if (swap_ind.empty()) {
              const double mu = r_mean_range(g, i, j);
              for (int t = i; t <= j; ++t) g_new[t] = mu;
              break;
            }
            ind.clear();
            for (int t = 1; t <= q - i + 1; ++t) ind.push_back(t);   // seq(i:q) == 1..(q-i+1)  (reference quirk, :215,:218)
            if (mult == 1) {
              violators(eq, true);
              for (int t : eq)
                if (std::find(ind.begin(), ind.end(), t) == ind.end()) ind.push_back(t);
            }
            pool_violators(g, g_new, ind, swap_ind);
            i = j;                                              // :241
          }
        } else if (i == q - 1 && flg == 0) {                    // :245-275 only q against q+1..m
          violators(swap_ind, false);
          ind.assign(1, q);
          if (mult == 1) {
            violators(eq, true);
            for (int t : eq) ind.push_back(t);
          }
          pool_violators(g, g_new, ind, swap_ind);
        }
        ++i;
      }
      g = g_new;                                                // :279
      if (first_q_in_order(g, m, q)) { converged = true; break; }
      mult = 1;
    }
  } else {                                                      // q == 1, :289-335
    for (int pass = 0; pass < kMaxPasses; ++pass) {
      violators(swap_ind, false);
      ind.assign(1, q);
      if (mult == 1) {
        violators(eq, true);
        for (int t : eq) ind.push_back(t);
      }
      if (swap_ind.empty()) {
        if (strict_r) { set_error("eigvalAlgo: object 'tmp' not found (R/spEigenCov.R:312, q == 1 without violators)"); return SPEIGEN_ERR_ARG; }
      } else {
        double mn = g[swap_ind[0]];
        for (int t : swap_ind) mn = std::min(mn, g[t]);
        for (int t : swap_ind)
          if (g[t] == mn && std::find(ind.begin(), ind.end(), t) == ind.end()) ind.push_back(t);
      }
      const double mu = r_mean(g, ind);
      for (int t : ind) g_new[t] = mu;                                    // both branches of :312-319
      g = g_new;
      if (first_q_in_order(g, m, q)) { converged = true; break; }
    }
  }
  if (!converged) {
    set_error("eigvalAlgo does not terminate on this input (the reference's while(1) would loop forever)");
    return SPEIGEN_ERR_ARG;
  }
  for (int64_t t = 0; t < m; ++t) {
    const double phi = (1.0 + std::sqrt(1.0 + 4.0 * x * g[t + 1])) / (2.0 * x);   // :286,:333
    xi_out[t] = 1.0 / phi;
  }
  return SPEIGEN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Device context for the dense path
// ---------------------------------------------------------------------------------------------------------------
struct DenseCtx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t st = nullptr;
  int64_t m = 0, mp = 0, ld = 0;
  std::vector<double*> mats;         // mp x ld matrices
  JacobiWork jw;
  double* vec[6] = {};               // length-mp device vectors
  int* perm = nullptr;
  double* scan_part = nullptr;
  double* h_pinned = nullptr;        // 4*mp + 64 doubles
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_join = nullptr;
  double gemm_ms = 0.0;
  int gemms = 0;
  int max_sweeps = 40;
  double jtol = 0.0;
  double early_tol = 1e-9;
  int verbose = 0;

  int create(int dev, int64_t m_, int nmats) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
      cudaGetLastError();
      set_error("no CUDA device available: speigen_b200 has no CPU fallback");
      return SPEIGEN_ERR_NO_DEVICE;
    }
    if (dev < 0 || dev >= ndev) { set_error("device %d out of range", dev); return SPEIGEN_ERR_ARG; }
    device = dev;
    SPE_CUDA(cudaSetDevice(dev));
    cudaDeviceProp pr;
    SPE_CUDA(cudaGetDeviceProperties(&pr, dev));
    if (pr.major < 10) { set_error("speigen_b200 needs an sm_100a device (found sm_%d%d)", pr.major, pr.minor); return SPEIGEN_ERR_NO_DEVICE; }
    num_sms = pr.multiProcessorCount;
    m = m_;
    mp = round_up(m, kJRows);
    ld = mp;
    SPE_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int i = 0; i < nmats; ++i) {
      double* p = nullptr;
      if (cudaMalloc(&p, static_cast<size_t>(mp) * ld * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        set_error("out of device memory for %d matrices of %lld x %lld", nmats, (long long)mp, (long long)mp);
        return SPEIGEN_ERR_NOMEM;
      }
      mats.push_back(p);
    }
    const int max_panels = static_cast<int>(mp / kJP);
    const int chunks = static_cast<int>(mp / kJRows);
    jw.nsplit = std::max(1, std::min(chunks, 16));
    if (const char* e = std::getenv("SPEIGEN_JACOBI_NSPLIT")) jw.nsplit = std::max(1, std::min(std::min(chunks, 16), std::atoi(e)));
    jw.inner_sweeps = 1;   // one 64 x 64 Jacobi sweep per panel visit: same outer sweep count as running it to convergence (measured)
    if (const char* e = std::getenv("SPEIGEN_JACOBI_INNER")) jw.inner_sweeps = std::atoi(e);
    if (const char* e = std::getenv("SPEIGEN_JACOBI_EARLY")) early_tol = std::atof(e);
    if (const char* e = std::getenv("SPEIGEN_JACOBI_VERBOSE")) verbose = std::atoi(e);
    SPE_CUDA(cudaMalloc(&jw.gram_part, static_cast<size_t>(max_panels) * jw.nsplit * kJP * kJP * sizeof(double)));
    {
      const char* e = std::getenv("SPEIGEN_JEIG_SPLIT");
      if (!(e && e[0] == '0')) {
        SPE_CUDA(cudaMalloc(&jw.rot, static_cast<size_t>(max_panels) * (kJP - 1) * (kJP / 2) * 2 * sizeof(double)));
        SPE_CUDA(cudaMalloc(&jw.flags, static_cast<size_t>(max_panels) * sizeof(int)));
        SPE_CUDA(cudaMemset(jw.flags, 0, static_cast<size_t>(max_panels) * sizeof(int)));
        if (const char* pe = std::getenv("SPEIGEN_JEIG_PUBLISH")) jw.publish_every = std::max(1, std::atoi(pe));
        if (const char* te = std::getenv("SPEIGEN_JEIG_THREADS")) {
          const int t = std::atoi(te);
          if (t == 576 || t == 640 || t == 768 || t == 1024) jw.split_threads = t;
        }
      }
    }
    jw.wp_stride = static_cast<size_t>(max_panels) * kJP * kJP;
    SPE_CUDA(cudaMalloc(&jw.wp, JacobiWork::kWpSlots * jw.wp_stride * sizeof(double)));
    {
      const char* e = std::getenv("SPEIGEN_JACOBI_SIDE");
      if (!(e && e[0] == '0')) {
        SPE_CUDA(cudaStreamCreateWithFlags(&jw.side, cudaStreamNonBlocking));
        for (int i = 0; i < JacobiWork::kWpSlots; ++i) {
          SPE_CUDA(cudaEventCreateWithFlags(&jw.ev_eig[i], cudaEventDisableTiming));
          SPE_CUDA(cudaEventCreateWithFlags(&jw.ev_applied[i], cudaEventDisableTiming));
        }
        SPE_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      }
    }
    SPE_CUDA(cudaMalloc(&jw.offmax, 2 * sizeof(double)));
    for (auto& v : vec) SPE_CUDA(cudaMalloc(&v, static_cast<size_t>(mp) * sizeof(double)));
    SPE_CUDA(cudaMalloc(&perm, static_cast<size_t>(mp) * sizeof(int)));
    SPE_CUDA(cudaMalloc(&scan_part, static_cast<size_t>(mp) * 4 * sizeof(double)));
    SPE_CUDA(cudaMallocHost(&h_pinned, (static_cast<size_t>(mp) * 4 + 64) * sizeof(double)));
    SPE_CUDA(cudaEventCreate(&ev0));
    SPE_CUDA(cudaEventCreate(&ev1));
    jtol = std::sqrt(static_cast<double>(mp)) * 2.220446049250313e-16;
    return SPEIGEN_OK;
  }
  void destroy() {
    cudaSetDevice(device);
    for (double* p : mats) cudaFree(p);
    mats.clear();
    cudaFree(jw.gram_part); cudaFree(jw.wp); cudaFree(jw.offmax); cudaFree(jw.rot); cudaFree(jw.flags);
    if (jw.side) {
      cudaStreamSynchronize(jw.side);
      for (int i = 0; i < JacobiWork::kWpSlots; ++i) { cudaEventDestroy(jw.ev_eig[i]); cudaEventDestroy(jw.ev_applied[i]); }
      cudaEventDestroy(ev_join);
      cudaStreamDestroy(jw.side);
    }
    for (auto& v : vec) { cudaFree(v); v = nullptr; }
    cudaFree(perm); cudaFree(scan_part);
    if (h_pinned) cudaFreeHost(h_pinned);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (st) cudaStreamDestroy(st);
    jw = JacobiWork();
    perm = nullptr; scan_part = nullptr; h_pinned = nullptr; ev0 = ev1 = nullptr; st = nullptr;
  }
  int sync() { SPE_CUDA(cudaStreamSynchronize(st)); return SPEIGEN_OK; }
  int d2h(double* dst, const double* src, size_t n) {
    SPE_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    return sync();
  }
  int h2d(double* dst, const double* src, size_t n) {
    SPE_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return SPEIGEN_OK;
  }
  // C = alpha op(A) op(B) on mp-padded matrices (M, N, K as given), device-timed
  int gemm(int64_t M, int64_t N, int64_t K, const double* A, int ak, const double* B, int bk, double* C, double alpha = 1.0,
           int symmetric = 0) {
    GemmOperand a, b;
    a.ptr = A; a.ld = ld; a.kmajor = ak;
    b.ptr = B; b.ld = ld; b.kmajor = bk;
    GemmEpilogue e;
    e.alpha = alpha; e.symmetric = symmetric;
    SPE_CUDA(cudaEventRecord(ev0, st));
    SPE_TRY2(launch_gemm(M, N, K, a, b, C, ld, e, num_sms, st));
    SPE_CUDA(cudaEventRecord(ev1, st));
    SPE_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    SPE_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    gemm_ms += ms;
    ++gemms;
    return SPEIGEN_OK;
  }
  // Block one-sided Jacobi: on exit X = X_in * W has mutually orthogonal columns, W orthogonal.
  int jacobi(double* X, double* W, int* sweeps_out) {
    SPE_TRY2(launch_cm_identity(W, mp, ld, st));
    if (jw.side) {   // the side stream rotates W: it starts after the identity fill and joins at the end of every sweep
      SPE_CUDA(cudaEventRecord(ev_join, st));
      SPE_CUDA(cudaStreamWaitEvent(jw.side, ev_join, 0));
    }
    const int nblocks = static_cast<int>(mp / kJB);
    int sweeps = 0;
    int round_index = 0;
    bool converged = false;
    for (; sweeps < max_sweeps;) {
      SPE_CUDA(cudaMemsetAsync(jw.offmax, 0, 2 * sizeof(double), st));
      for (int r = 0; r < nblocks; ++r) SPE_TRY2(launch_jacobi_round(X, W, mp, ld, r & 1, jw, st, round_index++));
      if (jw.side) {
        SPE_CUDA(cudaEventRecord(ev_join, jw.side));
        SPE_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
      }
      ++sweeps;
      double off2[2] = {0.0, 0.0};
      SPE_TRY2(d2h(off2, jw.offmax, 2));
      const double off = off2[0];
      if (off2[1] != 0.0) {
        set_error("block Jacobi: the rotation hand-off between the solver and builder CTAs timed out");
        return SPEIGEN_ERR_CUDA;
      }
      if (verbose) std::fprintf(stderr, "[jacobi] sweep %d max relative off-diagonal %.3e\n", sweeps, off);
      // `off` is measured BEFORE each panel is rotated.  The cyclic Jacobi iteration converges quadratically, so a sweep
      // that started below `early_tol` ends near off^2 << jtol: the confirming sweep is skipped (the Newton-Schulz polish
      // of the callers squares the remaining non-orthogonality once more).
      if (!(off > jtol) || off <= early_tol) { converged = true; break; }
    }
    if (sweeps_out) *sweeps_out = sweeps;
    if (!converged) {
      set_error("block Jacobi did not reach the orthogonality tolerance in %d sweeps (increase jacobi_max_sweeps)", sweeps);
      return SPEIGEN_ERR_UNSUPPORTED;
    }
    return SPEIGEN_OK;
  }
  // U = polar(B): B in `X` (destroyed), result in `U`; W, tmp are scratch matrices
  int polar(double* X, double* W, double* U, double* tmp, int polish, int* sweeps_out) {
    SPE_TRY2(jacobi(X, W, sweeps_out));
    SPE_TRY2(launch_cm_coldots(X, X, mp, mp, ld, vec[0], st));
    double* h = h_pinned;
    SPE_TRY2(d2h(h, vec[0], mp));
    for (int64_t j = 0; j < mp; ++j) h[j] = h[j] > 0.0 ? 1.0 / std::sqrt(h[j]) : 0.0;
    SPE_TRY2(h2d(vec[0], h, mp));
    SPE_TRY2(launch_cm_scale_cols(X, mp, mp, ld, vec[0], 0, st));
    // U = (X Sigma^-1) W^T
    SPE_TRY2(gemm(mp, mp, mp, X, 0, W, 0, U));
    if (polish) {   // U <- U (1.5 I - 0.5 U^T U)
      SPE_TRY2(gemm(mp, mp, mp, U, 1, U, 1, tmp, 1.0, 1));
      SPE_TRY2(launch_cm_ns_matrix(tmp, mp, ld, W, st));
      SPE_TRY2(gemm(mp, mp, mp, U, 0, W, 1, X));
      SPE_CUDA(cudaMemcpyAsync(U, X, static_cast<size_t>(mp) * ld * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return SPEIGEN_OK;
  }
  // eigen(S): S (padded with zeros) in `X` (destroyed); eigenvectors (sorted, decreasing) in `V`, values on the host
  int eigh(double* X, double* W, double* V, std::vector<double>& lam, int* sweeps_out) {
    SPE_TRY2(jacobi(X, W, sweeps_out));
    SPE_TRY2(launch_cm_coldots(W, X, mp, mp, ld, vec[0], st));   // Rayleigh quotients w^T (S w): signed eigenvalues
    double* h = h_pinned;
    SPE_TRY2(d2h(h, vec[0], mp));
    std::vector<int> order(mp);
    std::iota(order.begin(), order.end(), 0);
    // the zero-padded coordinates (eigenvalue exactly 0, eigenvector e_i, i >= m) must sort behind every real one
    SPE_TRY2(launch_cm_coldots(W, W, m, mp, ld, vec[1], st));     // mass of each eigenvector inside the first m rows
    double* hm = h_pinned + mp;
    SPE_TRY2(d2h(hm, vec[1], mp));
    // Rayleigh quotients of the (slightly non-normalised) rotation product: divide by w^T w
    SPE_TRY2(launch_cm_coldots(W, W, mp, mp, ld, vec[1], st));
    double* hn = h_pinned + 2 * mp;
    SPE_TRY2(d2h(hn, vec[1], mp));
    for (int64_t j = 0; j < mp; ++j)
      if (hn[j] > 0.0) h[j] /= hn[j];
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      const bool pa = hm[a] < 0.5, pb = hm[b] < 0.5;
      if (pa != pb) return pb;            // real before padded
      return h[a] > h[b];
    });
    lam.resize(m);
    for (int64_t j = 0; j < m; ++j) lam[j] = h[order[j]];
    SPE_CUDA(cudaMemcpyAsync(perm, order.data(), static_cast<size_t>(mp) * sizeof(int), cudaMemcpyHostToDevice, st));
    SPE_TRY2(launch_cm_gather_cols(W, mp, mp, ld, perm, V, st));
    // W is a product of ~1e4 plane rotations per column: its columns drift from unit norm by ~1e-13.  One Newton-Schulz
    // step V <- V (1.5 I - 0.5 V^T V) restores orthonormality to working precision (X and W are free scratch here).
    SPE_TRY2(gemm(mp, mp, mp, V, 1, V, 1, X, 1.0, 1));
    SPE_TRY2(launch_cm_ns_matrix(X, mp, ld, W, st));
    SPE_TRY2(gemm(mp, mp, mp, V, 0, W, 1, X));
    SPE_CUDA(cudaMemcpyAsync(V, X, static_cast<size_t>(mp) * ld * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return sync();
  }
  // eigen(S) for a Hermitian S carried in embedded form (dimension mp = 2h): every eigenvalue of S appears twice in the
  // embedding, with real eigenvectors [x; y] and [-y; x] that are the same complex vector x + i y up to the phase i.
  // After sorting, one vector of each adjacent pair is kept.  X = embed(S) zero-padded (destroyed); V receives the
  // embedded, sorted, re-orthonormalised eigenvector matrix.
  int eigh_embedded(double* X, double* W, double* V, int64_t mc, int64_t h, std::vector<double>& lam, int* sweeps_out) {
    SPE_TRY2(jacobi(X, W, sweeps_out));
    SPE_TRY2(launch_cm_coldots(W, X, mp, mp, ld, vec[0], st));
    SPE_TRY2(launch_cm_coldots(W, W, mp, mp, ld, vec[1], st));
    double* hq = h_pinned;
    double* hn = h_pinned + mp;
    SPE_CUDA(cudaMemcpyAsync(hq, vec[0], mp * sizeof(double), cudaMemcpyDeviceToHost, st));
    SPE_CUDA(cudaMemcpyAsync(hn, vec[1], mp * sizeof(double), cudaMemcpyDeviceToHost, st));
    // mass of each real eigenvector inside the un-padded coordinates (rows [0, mc) and [h, h + mc))
    SPE_TRY2(launch_cm_coldots(W, W, mc, mp, ld, vec[0], st));
    SPE_TRY2(launch_cm_coldots(W + h, W + h, mc, mp, ld, vec[1], st));
    double* hm0 = h_pinned + 2 * mp;
    double* hm1 = h_pinned + 3 * mp;
    SPE_CUDA(cudaMemcpyAsync(hm0, vec[0], mp * sizeof(double), cudaMemcpyDeviceToHost, st));
    SPE_CUDA(cudaMemcpyAsync(hm1, vec[1], mp * sizeof(double), cudaMemcpyDeviceToHost, st));
    SPE_TRY2(sync());
    for (int64_t j = 0; j < mp; ++j)
      if (hn[j] > 0.0) hq[j] /= hn[j];
    std::vector<int> order(mp);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      const bool pa = (hm0[a] + hm1[a]) < 0.5 * hn[a], pb = (hm0[b] + hm1[b]) < 0.5 * hn[b];
      if (pa != pb) return pb;
      return hq[a] > hq[b];
    });
    lam.resize(mc);
    std::vector<int> pick(mp, 0);
    const double scale = std::fabs(hq[order[0]]);
    for (int64_t k = 0; k < mc; ++k) {
      const double l0 = hq[order[2 * k]], l1 = hq[order[2 * k + 1]];
      if (std::fabs(l0 - l1) > 1e-9 * scale) {
        set_error("complex eigen-decomposition: eigenvalue pair %lld of the embedded matrix does not match (%.17g vs %.17g)",
                  (long long)k, l0, l1);
        return SPEIGEN_ERR_CUDA;
      }
      lam[k] = 0.5 * (l0 + l1);
      pick[k] = order[2 * k];
    }
    SPE_CUDA(cudaMemcpyAsync(perm, pick.data(), static_cast<size_t>(mp) * sizeof(int), cudaMemcpyHostToDevice, st));
    SPE_TRY2(launch_ce_select_eigvecs(W, perm, mc, h, ld, V, st));
    SPE_TRY2(gemm(mp, mp, mp, V, 1, V, 1, X, 1.0, 1));
    SPE_TRY2(launch_cm_ns_matrix(X, mp, ld, W, st));
    SPE_TRY2(gemm(mp, mp, mp, V, 0, W, 1, X));
    SPE_CUDA(cudaMemcpyAsync(V, X, static_cast<size_t>(mp) * ld * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return sync();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// spEigenCov driver
// ---------------------------------------------------------------------------------------------------------------
static DenseCtx* g_cov_ctx = nullptr;
static std::mutex g_cov_mtx;
static void release_cov_context_locked() {
  if (g_cov_ctx) {
    g_cov_ctx->destroy();
    delete g_cov_ctx;
    g_cov_ctx = nullptr;
  }
}

// cplx: S_host is interleaved {re, im}; every m x m matrix is carried in real embedded form of dimension 2h (dense_ops.cuh)
static int run_speigencov(const double* S_host, int64_t m, int q, double rho, double thres, const speigencov_cfg& cfg,
                          speigencov_out* out, bool cplx) {
  const double t_start = now_s();
  const int64_t h = cplx ? round_up(m, 64) : 0;
  const int64_t dim = cplx ? 2 * h : m;
  // One context (7 matrices of the padded dimension, Jacobi work space, streams) is kept alive between calls of the same
  // size on the same device: creating and freeing it costs ~40 ms of a 0.5 s call at m = 2000.  speigen_shutdown() or
  // process exit releases it; contexts above 4 GB are not kept.
  std::lock_guard<std::mutex> lock(g_cov_mtx);
  if (g_cov_ctx && (g_cov_ctx->device != cfg.device || g_cov_ctx->m != dim)) release_cov_context_locked();
  int rc = SPEIGEN_OK;
  if (!g_cov_ctx) {
    g_cov_ctx = new DenseCtx();
    // matrices: 0 S_hat, 1 V, 2 T, 3 X (Jacobi work / target), 4 W, 5 tmp, 6 raw S (m x m in an mp x ld buffer; complex: flat)
    rc = g_cov_ctx->create(cfg.device, dim, 7);
    if (rc != SPEIGEN_OK) { release_cov_context_locked(); return rc; }
  } else {
    SPE_CUDA(cudaSetDevice(cfg.device));
  }
  DenseCtx& cx = *g_cov_ctx;
  cx.gemm_ms = 0.0;
  cx.gemms = 0;
  cx.max_sweeps = cfg.jacobi_max_sweeps > 0 ? cfg.jacobi_max_sweeps : 40;
  cx.jtol = cfg.jacobi_tol > 0.0 ? cfg.jacobi_tol : std::sqrt(static_cast<double>(cx.mp)) * 2.220446049250313e-16;
  auto body = [&]() -> int {
    const int64_t mp = cx.mp, ld = cx.ld;
    double *Sh = cx.mats[0], *V = cx.mats[1], *T = cx.mats[2], *X = cx.mats[3], *W = cx.mats[4], *tmp = cx.mats[5],
           *Sraw = cx.mats[6];
    cudaStream_t st = cx.st;
    const int64_t rows_v = cplx ? mp : m;   // rows that carry a column of V / T (embedded: real and imaginary halves)
    if (cplx) {
      SPE_CUDA(cudaMemcpyAsync(Sraw, S_host, static_cast<size_t>(2) * m * m * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
      SPE_CUDA(cudaMemcpy2DAsync(Sraw, ld * sizeof(double), S_host, m * sizeof(double), m * sizeof(double), m,
                                 cudaMemcpyHostToDevice, st));
    }
    SPE_TRY2(cx.sync());
    out->seconds_upload = now_s() - t_start;
    const double t_init = now_s();
    // anyNA / isSymmetric.matrix / ||S||_F                                                  R/spEigenCov.R:56-57,73
    if (cplx) {   // the embedding of S is symmetric iff S is Hermitian; every sum below counts each entry twice
      SPE_TRY2(launch_ce_embed(Sraw, m, X, h, ld, 0.0, 0.0, st));
      SPE_TRY2(launch_cm_scan_sym(X, mp, ld, cx.scan_part, cx.vec[2], st));
    } else {
      SPE_TRY2(launch_cm_scan_sym(Sraw, m, ld, cx.scan_part, cx.vec[2], st));
    }
    double sc4[4];
    SPE_TRY2(cx.d2h(sc4, cx.vec[2], 4));
    if (cplx) sc4[3] *= 0.5;
    if (cfg.check_na && sc4[2] > 0.0) { set_error("This function cannot handle NAs."); return SPEIGEN_ERR_NA; }
    if (cfg.check_symmetric) {
      const double tol = 100.0 * 2.220446049250313e-16;
      const bool sym = sc4[1] > 0.0 ? (sc4[0] / sc4[1] <= tol) : (sc4[0] == 0.0);
      if (!sym) { set_error("The covariance matrix is not symmetric"); return SPEIGEN_ERR_NOT_SYMMETRIC; }
    }
    const double fro = std::sqrt(sc4[3]);
    // eigen(S)                                                                               :62-68
    std::vector<double> Xi;
    int esw = 0;
    if (cplx) {
      SPE_TRY2(cx.eigh_embedded(X, W, V, m, h, Xi, &esw));
    } else {
      SPE_TRY2(launch_cm_pad_copy(Sraw, m, ld, X, mp, ld, 0.0, 0.0, st));
      SPE_TRY2(cx.eigh(X, W, V, Xi, &esw));
    }
    out->eig_sweeps = esw;
    for (int64_t j = 0; j < m; ++j)
      if (Xi[j] <= cfg.psd_tol) { set_error("The covariance matrix is not PSD."); return SPEIGEN_ERR_NOT_PSD; }
    int64_t rank = 0;
    for (int64_t j = 0; j < m; ++j) {
      if (Xi[j] < 0.0) Xi[j] = 0.0;
      if (Xi[j] > cfg.rank_tol) ++rank;
    }
    if (rank < m) { set_error("The covariance matrix is low-rank."); return SPEIGEN_ERR_LOW_RANK; }
    const double xi_max = Xi[0];
    const double shift = xi_max + cfg.shift;                                                  // :69
    if (cplx) SPE_TRY2(launch_ce_embed(Sraw, m, Sh, h, ld, shift, 0.0, st));
    else SPE_TRY2(launch_cm_pad_copy(Sraw, m, ld, Sh, mp, ld, shift, 0.0, st));
    std::vector<double> rho_vec(q);
    for (int j = 0; j < q; ++j) rho_vec[j] = rho * cfg.rho_scale * fro * Xi[j] / Xi[0];       // :73
    double* d_rho = cx.vec[3];
    SPE_TRY2(cx.h2d(d_rho, rho_vec.data(), q));
    // schedule                                                                               :80-88
    const int K = cfg.n_continuation;
    std::vector<double> pp(K + 1), ee(K + 1), tol(K + 1);
    SPE_TRY2(speigencov_schedule(&cfg, pp.data(), ee.data(), tol.data()));
    // T = S_hat V and its column dots for the start point
    double* d_dots = cx.vec[4];
    double* d_vn = cx.vec[5];
    std::vector<double> dots(m), vn(m), gq(q), inv_xi(mp, 0.0), F_v;
    auto refresh_T = [&]() -> int {
      SPE_TRY2(cx.gemm(mp, mp, mp, Sh, 1, V, 1, T));                                          // S_hat symmetric: S_hat^T V
      SPE_TRY2(launch_cm_coldots(V, T, rows_v, m, ld, d_dots, st));                           // Re(v_j^H t_j)
      SPE_TRY2(launch_cm_coldots(V, V, rows_v, m, ld, d_vn, st));
      double* h = cx.h_pinned;
      SPE_CUDA(cudaMemcpyAsync(h, d_dots, m * sizeof(double), cudaMemcpyDeviceToHost, st));
      SPE_CUDA(cudaMemcpyAsync(h + mp, d_vn, m * sizeof(double), cudaMemcpyDeviceToHost, st));
      SPE_TRY2(cx.sync());
      std::memcpy(dots.data(), h, m * sizeof(double));
      std::memcpy(vn.data(), h + mp, m * sizeof(double));
      return SPEIGEN_OK;
    };
    SPE_TRY2(refresh_T());
    out->seconds_init = now_s() - t_init;
    const double t_loop = now_s();
    int k = 0, svd_sweeps = 0;
    std::vector<double> g(m), Xi_new(m);
    for (int stage = 0; stage <= K; ++stage) {                                                // :92
      const StageConsts sc = make_stage(pp[stage], ee[stage]);
      int flg = 1;
      for (;;) {
        ++k;
        // eigenvector update: B = -(H1 + S_hat V diag(1/Xi)), V <- polar(B)                   :100-103,:146-160
        for (int64_t j = 0; j < m; ++j) inv_xi[j] = 1.0 / Xi[j];
        SPE_TRY2(cx.h2d(cx.vec[2], inv_xi.data(), mp));
        if (cplx) {
          SPE_TRY2(launch_ce_colabsmin(V, m, h, q, ld, cx.vec[1], st));
          SPE_TRY2(launch_ce_build_target(V, T, cx.vec[2], d_rho, cx.vec[1], q, sc, m, h, ld, X, st));
        } else {
          SPE_TRY2(launch_cm_colabsmin(V, m, q, ld, cx.vec[1], st));
          SPE_TRY2(launch_cm_build_target(V, T, cx.vec[2], d_rho, cx.vec[1], q, sc, m, mp, ld, X, st));
        }
        // eigenvalue update from the OLD A = V diag(1/Xi): g_j = -a_j^T S_hat a_j = -dots_j / Xi_j^2   :106-107
        for (int64_t j = 0; j < m; ++j) g[j] = -dots[j] * inv_xi[j] * inv_xi[j];
        SPE_TRY2(eigval_algo_host(g.data(), m, xi_max, q, 0, Xi_new.data()));
        int sw = 0;
        SPE_TRY2(cx.polar(X, W, V, tmp, cfg.polish, &sw));
        svd_sweeps += sw;
        Xi = Xi_new;
        // objective                                                                          :110-114
        SPE_TRY2(refresh_T());
        if (cplx) SPE_TRY2(launch_ce_penalty(V, m, h, q, ld, sc, cx.vec[1], st));
        else SPE_TRY2(launch_cm_penalty(V, m, q, ld, sc, cx.vec[1], st));
        SPE_TRY2(cx.d2h(gq.data(), cx.vec[1], q));
        double F = 0.0;
        for (int64_t j = 0; j < m; ++j) F += std::log(Xi[j]);
        double tr = 0.0;
        for (int64_t j = 0; j < m; ++j) tr += (dots[j] + shift * vn[j]) / Xi[j];             // diag(V^T S V) / Xi
        F += tr;
        for (int j = 0; j < q; ++j) F += rho_vec[j] * gq[j];
        F_v.push_back(F);
        if (out->F && k <= out->trace_capacity) out->F[k - 1] = F;
        if (out->stage_of_k && k <= out->trace_capacity) out->stage_of_k[k - 1] = stage + 1;
        if (cfg.verbose) std::fprintf(stderr, "[speigencov] stage %d k %d F %.15g sweeps %d\n", stage + 1, k, F, sw);
        if (flg == 0) {                                                                       // :117-122
          const double prev = F_v[k - 2];
          const double rel = std::fabs(F - prev) / std::max(1.0, std::fabs(prev));
          if (rel <= tol[stage] || k >= cfg.max_iter) break;
        }
        flg = 0;
      }
    }
    out->iterations = k;
    out->svd_sweeps = svd_sweeps;
    out->seconds_loop = now_s() - t_loop;
    // threshold, the reference's row-wise normalisation, cov                                 :131-135
    if (cplx) SPE_TRY2(launch_ce_thres_norms(V, m, h, ld, thres, cx.vec[0], st));
    else SPE_TRY2(launch_cm_thres_norms(V, m, ld, thres, cx.vec[0], st));
    std::vector<double> nrm(m);
    SPE_TRY2(cx.d2h(nrm.data(), cx.vec[0], m));
    for (int64_t j = 0; j < m; ++j) nrm[j] = 1.0 / std::sqrt(nrm[j]);
    SPE_TRY2(cx.h2d(cx.vec[0], nrm.data(), m));
    std::memcpy(out->values, Xi.data(), m * sizeof(double));
    if (cplx) {
      SPE_TRY2(launch_ce_thres_scale(V, m, h, ld, thres, cx.vec[0], nullptr, X, st));         // row i scaled by nrm[i] (:133)
      SPE_TRY2(launch_ce_extract(X, m, h, ld, Sraw, st));
      SPE_CUDA(cudaMemcpyAsync(out->vectors, Sraw, static_cast<size_t>(2) * m * m * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (out->cov) {
        SPE_TRY2(cx.h2d(cx.vec[1], Xi.data(), m));
        SPE_TRY2(launch_ce_thres_scale(X, m, h, ld, 0.0, nullptr, cx.vec[1], W, st));         // embed(Vout diag(Xi))
        SPE_TRY2(cx.gemm(mp, mp, mp, W, 0, X, 0, tmp));                                       // embed((Vout Xi) Vout^H)
        SPE_TRY2(launch_ce_extract(tmp, m, h, ld, Sraw, st));
        SPE_CUDA(cudaMemcpyAsync(out->cov, Sraw, static_cast<size_t>(2) * m * m * sizeof(double), cudaMemcpyDeviceToHost, st));
      }
    } else {
      SPE_TRY2(launch_cm_thres_scale(V, m, ld, thres, cx.vec[0], nullptr, X, ld, st));        // row i scaled by nrm[i] (:133)
      SPE_CUDA(cudaMemcpy2DAsync(out->vectors, m * sizeof(double), X, ld * sizeof(double), m * sizeof(double), m,
                                 cudaMemcpyDeviceToHost, st));
      if (out->cov) {
        SPE_TRY2(cx.h2d(cx.vec[1], Xi.data(), m));
        SPE_TRY2(launch_cm_thres_scale(X, m, ld, 0.0, nullptr, cx.vec[1], W, ld, st));        // Vout diag(Xi)
        SPE_TRY2(cx.gemm(m, m, m, W, 0, X, 0, tmp));                                          // (Vout Xi) Vout^T
        SPE_CUDA(cudaMemcpy2DAsync(out->cov, m * sizeof(double), tmp, ld * sizeof(double), m * sizeof(double), m,
                                   cudaMemcpyDeviceToHost, st));
      }
    }
    SPE_TRY2(cx.sync());
    out->gemms = cx.gemms;
    out->seconds_gemm = cx.gemm_ms * 1e-3;
    return SPEIGEN_OK;
  };
  rc = body();
  const size_t ctx_bytes = static_cast<size_t>(7) * cx.mp * cx.ld * sizeof(double);
  if (rc != SPEIGEN_OK || ctx_bytes > (static_cast<size_t>(4) << 30)) release_cov_context_locked();
  return rc;
}

void release_cov_context() {
  std::lock_guard<std::mutex> lock(g_cov_mtx);
  release_cov_context_locked();
}

}  // namespace speigen

using namespace speigen;

struct DeviceRestoreCov {
  int prev = -1;
  DeviceRestoreCov() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  ~DeviceRestoreCov() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" {

void speigencov_cfg_default(speigencov_cfg* c) {
  std::memset(c, 0, sizeof(*c));
  c->max_iter = 3000;
  c->n_continuation = 5;
  c->p1 = 1.0;
  c->pk = 5.0;
  c->tol_factor = 1e-1;
  c->eps_factor = 1e-1;
  c->shift = 1e-6;
  c->rank_tol = 1e-9;
  c->psd_tol = -1e-10;
  c->rho_scale = 10.0;
  c->check_symmetric = 1;
  c->check_na = 1;
  c->device = 0;
  c->jacobi_max_sweeps = 40;
  c->jacobi_tol = 0.0;
  c->polish = 1;
  c->verbose = 0;
}

// R/spEigenCov.R:52-60 in the reference's order (the NA / symmetry scans of S itself run on the device)
int speigencov_validate_args(int64_t nrow, int64_t ncol, double q, double rho) {
  if (ncol == 1) { set_error("Data is univariate!"); return SPEIGEN_ERR_UNIVARIATE; }
  if (ncol < 1) { set_error("S must be a matrix"); return SPEIGEN_ERR_ARG; }
  if (!std::isnan(q) && q > static_cast<double>(ncol)) {
    set_error("The number of sparse eigenvectors q should not be larger than m.");
    return SPEIGEN_ERR_Q_GT_M;
  }
  if (std::isnan(q) || std::isnan(rho)) { set_error("This function cannot handle NAs."); return SPEIGEN_ERR_NA; }
  if (nrow != ncol) { set_error("The covariance matrix is not symmetric"); return SPEIGEN_ERR_NOT_SYMMETRIC; }
  if (std::fmod(q, 1.0) != 0.0 || q <= 0) { set_error("The input argument q should be a positive integer."); return SPEIGEN_ERR_Q; }
  if (rho <= 0) { set_error("The input argument rho should be positive."); return SPEIGEN_ERR_RHO; }
  return SPEIGEN_OK;
}

int speigencov_schedule(const speigencov_cfg* cfg_in, double* p, double* eps, double* tol) {
  speigencov_cfg cfg;
  if (cfg_in) cfg = *cfg_in; else speigencov_cfg_default(&cfg);
  const int K = cfg.n_continuation;
  if (K < 0) { set_error("n_continuation must be >= 0"); return SPEIGEN_ERR_ARG; }
  const double gamma = K > 0 ? std::pow(cfg.pk / cfg.p1, 1.0 / K) : 1.0;       // R/spEigenCov.R:83
  for (int j = 0; j <= K; ++j) {
    const double ppj = std::pow(10.0, -(cfg.p1 * std::pow(gamma, j)));         // :84-85
    p[j] = ppj;
    tol[j] = ppj * cfg.tol_factor;                                             // :87
    eps[j] = ppj * cfg.eps_factor;                                             // :88
  }
  return SPEIGEN_OK;
}

int speigencov_struct_sizes(int32_t* cfg_bytes, int32_t* out_bytes) {
  if (cfg_bytes) *cfg_bytes = static_cast<int32_t>(sizeof(speigencov_cfg));
  if (out_bytes) *out_bytes = static_cast<int32_t>(sizeof(speigencov_out));
  return SPEIGEN_OK;
}

int speigencov_eigval_algo(const double* g, int64_t m, double x, int32_t q, int32_t strict_r, double* xi_out) {
  if (!g || !xi_out) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  return eigval_algo_host(g, m, x, q, strict_r, xi_out);
}

int speigencov_run_f64(const double* S, int64_t nrow, int64_t ncol, double q, double rho, double thres,
                       const speigencov_cfg* cfg_in, speigencov_out* out) {
  DeviceRestoreCov restore__;
  if (!S || !out || !out->vectors || !out->values) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  speigencov_cfg cfg;
  if (cfg_in) cfg = *cfg_in; else speigencov_cfg_default(&cfg);
  int rc = speigencov_validate_args(nrow, ncol, q, rho);
  if (rc != SPEIGEN_OK) return rc;
  if (static_cast<int64_t>(q) >= ncol) {
    set_error("spEigenCov with q == m is not supported (the reference's g[(q+1):m] runs backwards there)");
    return SPEIGEN_ERR_UNSUPPORTED;
  }
  out->iterations = out->eig_sweeps = out->svd_sweeps = out->gemms = 0;
  out->seconds_upload = out->seconds_init = out->seconds_loop = out->seconds_gemm = 0.0;
  return run_speigencov(S, ncol, static_cast<int>(q), rho, thres, cfg, out, false);
}

int speigencov_run_c64(const double* S, int64_t nrow, int64_t ncol, double q, double rho, double thres,
                       const speigencov_cfg* cfg_in, speigencov_out* out) {
  DeviceRestoreCov restore__;
  if (!S || !out || !out->vectors || !out->values) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  speigencov_cfg cfg;
  if (cfg_in) cfg = *cfg_in; else speigencov_cfg_default(&cfg);
  int rc = speigencov_validate_args(nrow, ncol, q, rho);
  if (rc != SPEIGEN_OK) return rc;
  if (static_cast<int64_t>(q) >= ncol) {
    set_error("spEigenCov with q == m is not supported (the reference's g[(q+1):m] runs backwards there)");
    return SPEIGEN_ERR_UNSUPPORTED;
  }
  out->iterations = out->eig_sweeps = out->svd_sweeps = out->gemms = 0;
  out->seconds_upload = out->seconds_init = out->seconds_loop = out->seconds_gemm = 0.0;
  return run_speigencov(S, ncol, static_cast<int>(q), rho, thres, cfg, out, true);
}

// ---- dense kernel-level entry points --------------------------------------------------------------------------
int speigen_dense_gemm(int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, int a_kmajor, const double* B,
                       int64_t ldb, int b_kmajor, double* C, int64_t ldc, double alpha, int symmetric, int device,
                       float* ms_out) {
  DeviceRestoreCov restore__;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    set_error("no CUDA device available: speigen_b200 has no CPU fallback");
    return SPEIGEN_ERR_NO_DEVICE;
  }
  SPE_CUDA(cudaSetDevice(device));
  cudaDeviceProp pr;
  SPE_CUDA(cudaGetDeviceProperties(&pr, device));
  const int64_t a_rows = a_kmajor ? K : M, a_cols = a_kmajor ? M : K;
  const int64_t b_rows = b_kmajor ? K : N, b_cols = b_kmajor ? N : K;
  const int64_t dlda = round_up(a_rows, 2), dldb = round_up(b_rows, 2), dldc = round_up(M, 2);
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  cudaStream_t st;
  SPE_CUDA(cudaStreamCreate(&st));
  SPE_CUDA(cudaMalloc(&dA, static_cast<size_t>(dlda) * a_cols * sizeof(double)));
  SPE_CUDA(cudaMalloc(&dB, static_cast<size_t>(dldb) * b_cols * sizeof(double)));
  SPE_CUDA(cudaMalloc(&dC, static_cast<size_t>(dldc) * N * sizeof(double)));
  SPE_CUDA(cudaMemcpy2DAsync(dA, dlda * sizeof(double), A, lda * sizeof(double), a_rows * sizeof(double), a_cols, cudaMemcpyHostToDevice, st));
  SPE_CUDA(cudaMemcpy2DAsync(dB, dldb * sizeof(double), B, ldb * sizeof(double), b_rows * sizeof(double), b_cols, cudaMemcpyHostToDevice, st));
  SPE_CUDA(cudaMemsetAsync(dC, 0, static_cast<size_t>(dldc) * N * sizeof(double), st));
  GemmOperand a, b;
  a.ptr = dA; a.ld = dlda; a.kmajor = a_kmajor;
  b.ptr = dB; b.ld = dldb; b.kmajor = b_kmajor;
  GemmEpilogue e;
  e.alpha = alpha; e.symmetric = symmetric;
  cudaEvent_t e0, e1;
  SPE_CUDA(cudaEventCreate(&e0));
  SPE_CUDA(cudaEventCreate(&e1));
  SPE_CUDA(cudaEventRecord(e0, st));
  int rc = launch_gemm(M, N, K, a, b, dC, dldc, e, pr.multiProcessorCount, st);
  SPE_CUDA(cudaEventRecord(e1, st));
  if (rc == SPEIGEN_OK) {
    SPE_CUDA(cudaMemcpy2DAsync(C, ldc * sizeof(double), dC, dldc * sizeof(double), M * sizeof(double), N, cudaMemcpyDeviceToHost, st));
    SPE_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  cudaStreamDestroy(st);
  return rc;
}

int speigen_dense_gemm_time(int64_t M, int64_t N, int64_t K, int a_kmajor, int b_kmajor, int symmetric, int device,
                            int32_t reps, float* ms_each) {
  DeviceRestoreCov restore__;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    set_error("no CUDA device available: speigen_b200 has no CPU fallback");
    return SPEIGEN_ERR_NO_DEVICE;
  }
  SPE_CUDA(cudaSetDevice(device));
  cudaDeviceProp pr;
  SPE_CUDA(cudaGetDeviceProperties(&pr, device));
  const int64_t a_rows = a_kmajor ? K : M, a_cols = a_kmajor ? M : K;
  const int64_t b_rows = b_kmajor ? K : N, b_cols = b_kmajor ? N : K;
  const int64_t dlda = round_up(a_rows, 2), dldb = round_up(b_rows, 2), dldc = round_up(M, 2);
  double *dA = nullptr, *dB = nullptr, *dC = nullptr;
  cudaStream_t st;
  SPE_CUDA(cudaStreamCreate(&st));
  SPE_CUDA(cudaMalloc(&dA, static_cast<size_t>(dlda) * a_cols * sizeof(double)));
  SPE_CUDA(cudaMalloc(&dB, static_cast<size_t>(dldb) * b_cols * sizeof(double)));
  SPE_CUDA(cudaMalloc(&dC, static_cast<size_t>(dldc) * N * sizeof(double)));
  // deterministic pseudo-random fill (values in (-1, 1))
  {
    std::vector<double> h(static_cast<size_t>(1) << 20);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (double& v : h) { s = s * 6364136223846793005ull + 1442695040888963407ull; v = static_cast<double>(static_cast<int64_t>(s >> 11)) * (1.0 / 9007199254740992.0) * 2.0 - 1.0; }
    auto fill = [&](double* d, size_t n) -> int {
      for (size_t off = 0; off < n; off += h.size()) {
        const size_t c = std::min(h.size(), n - off);
        SPE_CUDA(cudaMemcpyAsync(d + off, h.data(), c * sizeof(double), cudaMemcpyHostToDevice, st));
      }
      return SPEIGEN_OK;
    };
    SPE_TRY2(fill(dA, static_cast<size_t>(dlda) * a_cols));
    SPE_TRY2(fill(dB, static_cast<size_t>(dldb) * b_cols));
    SPE_CUDA(cudaStreamSynchronize(st));
  }
  GemmOperand a, b;
  a.ptr = dA; a.ld = dlda; a.kmajor = a_kmajor;
  b.ptr = dB; b.ld = dldb; b.kmajor = b_kmajor;
  GemmEpilogue e;
  e.symmetric = symmetric;
  cudaEvent_t e0, e1;
  SPE_CUDA(cudaEventCreate(&e0));
  SPE_CUDA(cudaEventCreate(&e1));
  int rc = SPEIGEN_OK;
  for (int r = 0; r < reps && rc == SPEIGEN_OK; ++r) {
    SPE_CUDA(cudaEventRecord(e0, st));
    rc = launch_gemm(M, N, K, a, b, dC, dldc, e, pr.multiProcessorCount, st);
    SPE_CUDA(cudaEventRecord(e1, st));
    SPE_CUDA(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms_each[r], e0, e1);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  cudaStreamDestroy(st);
  return rc;
}

int speigen_dense_polar(const double* B, int64_t m, double* U, int32_t polish, int device, int32_t* sweeps_out) {
  DeviceRestoreCov restore__;
  if (!B || !U || m < 1) { set_error("bad argument"); return SPEIGEN_ERR_ARG; }
  DenseCtx cx;
  int rc = cx.create(device, m, 5);
  if (rc == SPEIGEN_OK) {
    auto body = [&]() -> int {
      double *X = cx.mats[0], *W = cx.mats[1], *Um = cx.mats[2], *tmp = cx.mats[3], *raw = cx.mats[4];
      SPE_CUDA(cudaMemcpy2DAsync(raw, cx.ld * sizeof(double), B, m * sizeof(double), m * sizeof(double), m, cudaMemcpyHostToDevice, cx.st));
      SPE_TRY2(launch_cm_pad_copy(raw, m, cx.ld, X, cx.mp, cx.ld, 0.0, 1.0, cx.st));
      int sw = 0;
      SPE_TRY2(cx.polar(X, W, Um, tmp, polish, &sw));
      if (sweeps_out) *sweeps_out = sw;
      SPE_CUDA(cudaMemcpy2DAsync(U, m * sizeof(double), Um, cx.ld * sizeof(double), m * sizeof(double), m, cudaMemcpyDeviceToHost, cx.st));
      return cx.sync();
    };
    rc = body();
  }
  cx.destroy();
  return rc;
}

int speigen_dense_eigh(const double* S, int64_t m, double* vectors, double* values, int device, int32_t* sweeps_out) {
  DeviceRestoreCov restore__;
  if (!S || !vectors || !values || m < 1) { set_error("bad argument"); return SPEIGEN_ERR_ARG; }
  DenseCtx cx;
  int rc = cx.create(device, m, 4);
  if (rc == SPEIGEN_OK) {
    auto body = [&]() -> int {
      double *X = cx.mats[0], *W = cx.mats[1], *V = cx.mats[2], *raw = cx.mats[3];
      SPE_CUDA(cudaMemcpy2DAsync(raw, cx.ld * sizeof(double), S, m * sizeof(double), m * sizeof(double), m, cudaMemcpyHostToDevice, cx.st));
      SPE_TRY2(launch_cm_pad_copy(raw, m, cx.ld, X, cx.mp, cx.ld, 0.0, 0.0, cx.st));
      std::vector<double> lam;
      int sw = 0;
      SPE_TRY2(cx.eigh(X, W, V, lam, &sw));
      if (sweeps_out) *sweeps_out = sw;
      std::memcpy(values, lam.data(), m * sizeof(double));
      SPE_CUDA(cudaMemcpy2DAsync(vectors, m * sizeof(double), V, cx.ld * sizeof(double), m * sizeof(double), m, cudaMemcpyDeviceToHost, cx.st));
      return cx.sync();
    };
    rc = body();
  }
  cx.destroy();
  return rc;
}

// ---- cov(X) on the device ----------------------------------------------------------------------
// Xc (n x m, ldx) on the device of `d`, already centred: S[:, lo:hi] / (n - 1) into the shard buffer of `d`.
static int cov_slab_from_centred(Solver& S, Dev& d, const double* Xc, int64_t n, int64_t ldx) {
  GemmOperand a, b;
  a.ptr = Xc; a.ld = ldx; a.kmajor = 1;                       // K x M = n x m
  b.ptr = Xc + d.lo * ldx; b.ld = ldx; b.kmajor = 1;          // K x N = n x (hi - lo)
  GemmEpilogue e;
  e.alpha = 1.0 / static_cast<double>(n - 1);
  e.symmetric = (S.prob.world == 1) ? 1 : 0;                  // the whole S on one GPU: upper tiles + mirror
  return launch_gemm(S.m, d.hi - d.lo, n, a, b, d.A, d.a_ld, e, d.num_sms, d.st);
}

static int load_cov_of_data_impl(Solver& S, const double* X_src, int64_t n, int64_t ldx_src, bool src_on_device) {
  if (S.data || S.cplx) { set_error("cov-of-data loading needs a real covariance-mode solver"); return SPEIGEN_ERR_ARG; }
  if (n < 2) { set_error("cov(X) needs at least two samples"); return SPEIGEN_ERR_ARG; }
  const int64_t m = S.m;
  const int64_t ldx = round_up(n, 2);
  for (Dev& d : S.devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    double* Xc = nullptr;
    double* sums = nullptr;
    if (cudaMalloc(&Xc, static_cast<size_t>(ldx) * m * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      set_error("out of device memory for the %lld x %lld data matrix", (long long)n, (long long)m);
      return SPEIGEN_ERR_NOMEM;
    }
    SPE_CUDA(cudaMalloc(&sums, static_cast<size_t>(m) * sizeof(double)));
    int rc = SPEIGEN_OK;
    auto body = [&]() -> int {
      SPE_CUDA(cudaMemcpy2DAsync(Xc, ldx * sizeof(double), X_src, ldx_src * sizeof(double), n * sizeof(double), m,
                                 src_on_device ? cudaMemcpyDefault : cudaMemcpyHostToDevice, d.st));
      SPE_TRY2(launch_cm_colsums(Xc, n, m, ldx, sums, d.st));                    // cov() centres the columns
      SPE_TRY2(launch_cm_center(Xc, n, m, ldx, sums, 1.0 / static_cast<double>(n), d.st));
      SPE_TRY2(cov_slab_from_centred(S, d, Xc, n, ldx));
      SPE_CUDA(cudaStreamSynchronize(d.st));
      return SPEIGEN_OK;
    };
    rc = body();
    cudaFree(Xc);
    cudaFree(sums);
    if (rc != SPEIGEN_OK) return rc;
  }
  S.loaded = true;
  S.host_verified = false;
  return SPEIGEN_OK;
}

int speigen_solver_load_cov_of_data(speigen_solver* s, const double* X_host, int64_t n) {
  DeviceRestoreCov restore__;
  if (!s || !X_host) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  return load_cov_of_data_impl(s->impl, X_host, n, n, false);
}
int speigen_solver_load_cov_of_data_device(speigen_solver* s, const double* X_dev, int64_t n, int64_t ldx) {
  DeviceRestoreCov restore__;
  if (!s || !X_dev) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  return load_cov_of_data_impl(s->impl, X_dev, n, ldx, true);
}

int speigen_run_cov_of_data_f64(const double* X, int64_t n, int64_t m, double q, double rho, const double* d,
                                const double* V0, double thres, const speigen_cfg* cfg_in, speigen_out* out) {
  DeviceRestoreCov restore__;
  if (!X || !out) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  speigen_cfg cfg;
  if (cfg_in) cfg = *cfg_in; else speigen_cfg_default(&cfg);
  int rc = speigen_validate_args(m, m, q, rho);
  if (rc != SPEIGEN_OK) return rc;
  speigen_problem pb;
  std::memset(&pb, 0, sizeof(pb));
  pb.m = m; pb.n = m; pb.q = static_cast<int32_t>(q);
  pb.world = cfg.n_gpus > 0 ? cfg.n_gpus : 1;
  pb.local_gpus = pb.world;
  speigen_solver* s = nullptr;
  rc = speigen_solver_create(&s, &pb, &cfg);
  if (rc != SPEIGEN_OK) return rc;
  rc = s->impl.init_comm_all();
  const double t0 = now_s();
  if (rc == SPEIGEN_OK) rc = load_cov_of_data_impl(s->impl, X, n, n, false);
  out->seconds_upload = now_s() - t0;
  if (rc == SPEIGEN_OK) rc = s->impl.prepare();
  if (rc == SPEIGEN_OK) rc = s->impl.init_vectors(V0, out);
  if (rc == SPEIGEN_OK) rc = s->impl.run(rho, d, thres, out);
  speigen_solver_destroy(s);
  return rc;
}

int speigen_cov_f64(const double* X, int64_t n, int64_t m, double* S_out, int device, float* ms_gemm_out) {
  DeviceRestoreCov restore__;
  if (!X || !S_out || n < 2 || m < 1) { set_error("bad argument"); return SPEIGEN_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    set_error("no CUDA device available: speigen_b200 has no CPU fallback");
    return SPEIGEN_ERR_NO_DEVICE;
  }
  SPE_CUDA(cudaSetDevice(device));
  cudaDeviceProp pr;
  SPE_CUDA(cudaGetDeviceProperties(&pr, device));
  const int64_t ldx = round_up(n, 2), lds = round_up(m, 2);
  double *Xc = nullptr, *sums = nullptr, *dS = nullptr;
  cudaStream_t st;
  SPE_CUDA(cudaStreamCreate(&st));
  SPE_CUDA(cudaMalloc(&Xc, static_cast<size_t>(ldx) * m * sizeof(double)));
  SPE_CUDA(cudaMalloc(&sums, static_cast<size_t>(m) * sizeof(double)));
  SPE_CUDA(cudaMalloc(&dS, static_cast<size_t>(lds) * m * sizeof(double)));
  SPE_CUDA(cudaMemcpy2DAsync(Xc, ldx * sizeof(double), X, n * sizeof(double), n * sizeof(double), m, cudaMemcpyHostToDevice, st));
  int rc = launch_cm_colsums(Xc, n, m, ldx, sums, st);
  if (rc == SPEIGEN_OK) rc = launch_cm_center(Xc, n, m, ldx, sums, 1.0 / static_cast<double>(n), st);
  GemmOperand a;
  a.ptr = Xc; a.ld = ldx; a.kmajor = 1;
  GemmEpilogue e;
  e.alpha = 1.0 / static_cast<double>(n - 1);
  e.symmetric = 1;
  cudaEvent_t e0, e1;
  SPE_CUDA(cudaEventCreate(&e0));
  SPE_CUDA(cudaEventCreate(&e1));
  SPE_CUDA(cudaEventRecord(e0, st));
  if (rc == SPEIGEN_OK) rc = launch_gemm(m, m, n, a, a, dS, lds, e, pr.multiProcessorCount, st);
  SPE_CUDA(cudaEventRecord(e1, st));
  if (rc == SPEIGEN_OK) {
    SPE_CUDA(cudaMemcpy2DAsync(S_out, m * sizeof(double), dS, lds * sizeof(double), m * sizeof(double), m, cudaMemcpyDeviceToHost, st));
    SPE_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_gemm_out) *ms_gemm_out = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(Xc); cudaFree(sums); cudaFree(dS);
  cudaStreamDestroy(st);
  return rc;
}

}  // extern "C"
