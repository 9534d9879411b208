/* speigen_b200 — C ABI of the B200-native spEigen() hot path.
 *
 * The reference package (dppalomar/sparseEigen) is pure R with no native interface (NAMESPACE:3-4 has
 * no useDynLib, there is no src/).  The only interface that exists is the R closure
 *     spEigen(X, q = 1, rho = 0.5, data = FALSE, d = NA, V = NA, thres = 1e-9)      R/spEigen.R:46
 * returning list(vectors, standard_vectors, values)                                  R/spEigen.R:162
 * so these entry points are what a `.Call` glue for that closure binds (INTEGRATION.md shows the glue).
 * Plain C types only; matrices are column-major exactly as R stores them; complex numbers are
 * interleaved {re, im} doubles (Rcomplex == double[2]).  The library never keeps caller pointers after
 * a call returns and never frees caller memory.  Errors are returned as codes, never thrown/longjmp'd.
 */
#ifndef SPEIGEN_B200_H
#define SPEIGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPEIGEN_ABI_VERSION 1

/* ---- status codes: 1..7 mirror the reference's stop() conditions ------------------------------ */
#define SPEIGEN_OK 0
#define SPEIGEN_ERR_UNIVARIATE 1      /* "Data is univariate!"                         R/spEigen.R:52 */
#define SPEIGEN_ERR_NA 2              /* "This function cannot handle NAs."            R/spEigen.R:53 */
#define SPEIGEN_ERR_Q 3               /* "q should be a positive integer."             R/spEigen.R:54 */
#define SPEIGEN_ERR_RHO 4             /* "rho should be positive."                     R/spEigen.R:55 */
#define SPEIGEN_ERR_RANK 5            /* "q should not be larger than rank(X)."        R/spEigen.R:69,79 */
#define SPEIGEN_ERR_NOT_SYMMETRIC 6   /* "The covariance matrix is not symmetric."     R/spEigen.R:75 */
#define SPEIGEN_ERR_NOT_PD 7          /* "The covariance matrix is not PD."            R/spEigen.R:77 */
#define SPEIGEN_ERR_CUDA 20
#define SPEIGEN_ERR_NCCL 21
#define SPEIGEN_ERR_NO_DEVICE 22      /* no CUDA device: there is NO CPU fallback */
#define SPEIGEN_ERR_ARG 23
#define SPEIGEN_ERR_NOMEM 24
#define SPEIGEN_ERR_UNSUPPORTED 25
#define SPEIGEN_ERR_NUMERIC 26        /* a Procrustes target (G - H, R/spEigen.R:180) was rank deficient or non-finite: the reference's
                                         svd() would return an arbitrary completion; this library refuses instead of returning it */
#define SPEIGEN_ERR_NOT_CONVERGED 27  /* the block subspace iteration that replaces eigen()/svd() (R/spEigen.R:68,76) did not reach
                                         init_tol within init_max_passes */
#define SPEIGEN_ERR_TIMEOUT 28        /* a GPU waited longer than exchange_timeout_ms for a peer in the fused exchange; the ranks'
                                         exchange epochs are no longer aligned: destroy the solver (the one-shot calls do) */

/* ---- configuration: the constants the reference hard-codes, plus device options ---------------- */
typedef struct speigen_cfg {
  int32_t max_iter;          /* 1000   R/spEigen.R:47  */
  int32_t n_continuation;    /* K = 10 R/spEigen.R:94  (K+1 stages) */
  double p1;                 /* 1      R/spEigen.R:95  */
  double pk;                 /* 7      R/spEigen.R:96  */
  double tol_factor;         /* 1e-2   R/spEigen.R:101 */
  double backtrack_slack;    /* 1e-9   R/spEigen.R:140 */
  double rank_tol;           /* 1e-9   R/spEigen.R:69,79 */
  double pd_tol;             /* -1e-10 R/spEigen.R:77 */
  int32_t max_backtracks;    /* 200: cap on the reference's unbounded `while(1)` R/spEigen.R:126 */
  /* initialiser (block subspace iteration replacing eigen()/svd(), R/spEigen.R:68,76) */
  int32_t init_max_passes;   /* 300 */
  double init_tol;           /* 2e-14: max_c ||S v_c - theta_c v_c|| <= init_tol * theta_1 */
  int32_t init_block;        /* 0 = auto (q rounded up to the DMMA n-tile) */
  uint64_t init_seed;
  int32_t check_symmetric;   /* 1 */
  int32_t check_na;          /* 1 */
  /* devices */
  int32_t n_gpus;            /* single-process mode: devices 0..n_gpus-1 (0 -> 1) */
  int32_t first_device;      /* single-process mode: first CUDA ordinal */
  int32_t fused_epilogue;    /* 1: reweighting fused into the contraction epilogue where the data flow allows */
  int32_t verbose;
  int32_t pd_check_passes;   /* 4: block power passes on (theta_1 I - S) whose smallest Ritz value of S is tested against pd_tol
                                (R/spEigen.R:77 tests the full spectrum; a Ritz value < pd_tol proves S is not PD) */
  int32_t p2p_exchange;      /* 1: row-sharded covariance exchanges its m x q row blocks with peer-memory stores fused into the
                                contraction epilogue (NVLink P2P / CUDA IPC); 0: NCCL allgather after the kernel */
  int32_t exchange_timeout_ms;   /* 10000: bound on every device-side wait for a peer (fused exchange flags); on expiry the call
                                    returns SPEIGEN_ERR_TIMEOUT instead of hanging the other GPUs */
  int32_t allow_unconverged_init;   /* 0: an initialiser that misses init_tol is an error (SPEIGEN_ERR_NOT_CONVERGED) */
} speigen_cfg;

void speigen_cfg_default(speigen_cfg* cfg);

/* ---- results ------------------------------------------------------------------------------------
 * Caller allocates vectors / standard_vectors (m*q doubles, or 2*m*q for complex) and values (q).
 * The trace arrays are optional (NULL to skip); the reference computes F_v but does not return it
 * (R/spEigen.R:90,162) — it is exposed here as a side channel for parity tests. */
typedef struct speigen_out {
  double* vectors;             /* m x q column-major          R/spEigen.R:158-160 */
  double* standard_vectors;    /* m x q column-major          R/spEigen.R:162 (Vx[,1:q]) */
  double* values;              /* q                           R/spEigen.R:162 */
  double* F;                   /* [trace_capacity] objective per outer iteration k (optional) */
  double* step_a;              /* [trace_capacity] accepted SQUAREM step per k (optional) */
  int32_t* backtracks_per_k;   /* [trace_capacity] (optional) */
  int32_t* stage_of_k;         /* [trace_capacity] (optional) */
  int32_t trace_capacity;
  /* counters written by the library */
  int32_t iterations;          /* k at exit */
  int32_t backtracks;          /* total */
  int32_t contractions;        /* streamed passes over S (or X) in the MM loop */
  int32_t init_passes;         /* subspace-iteration passes */
  int32_t init_converged;
  double init_residual;        /* max_c ||S v_c - theta_c v_c|| / theta_1 at exit of the initialiser */
  double seconds_upload;
  double seconds_init;
  double seconds_loop;
  double seconds_setup;        /* one-shot calls: device allocation, peer mappings (0 when a cached context was reused) */
  double seconds_prepare;      /* validation scan / centring (R/spEigen.R:53,60,75) */
  int32_t upload_staged;       /* bit 0: the host matrix was pageable and went through the pinned staging ring;
                                  bit 1: Hermitian-half upload (one tile of every transposed pair crossed PCIe, isSymmetric
                                  evaluated on the host, the other tile rebuilt on the device) */
  int32_t polar_passes_max;    /* largest number of Gram passes any polar factor of the MM loop needed (2 = the usual case) */
} speigen_out;

/* ---- one-shot drop-in calls (host pointers) — what the .Call glue binds ------------------------
 * X: nrow x ncol column-major.  is_data == 0: covariance (nrow == ncol == m).  is_data != 0: data matrix,
 * n = nrow samples, m = ncol variables (centred by the library, R/spEigen.R:59-60).
 * d: q weights or NULL (default seq(1, 0.5, length.out = q), R/spEigen.R:64).  V0: m x q start or NULL. */
int speigen_run_f64(const double* X, int64_t nrow, int64_t ncol, int is_data, double q, double rho, const double* d,
                    const double* V0, double thres, const speigen_cfg* cfg, speigen_out* out);
int speigen_run_c64(const double* X_interleaved, int64_t nrow, int64_t ncol, int is_data, double q, double rho,
                    const double* d, const double* V0_interleaved, double thres, const speigen_cfg* cfg,
                    speigen_out* out);

/* Releases the device context the one-shot calls keep alive between calls of the same (small) shape. */
int speigen_shutdown(void);
const char* speigen_last_error(void);
int speigen_abi_version(void);
/* sizeof(speigen_cfg / speigen_out / speigen_problem) as compiled into the library: lets a foreign-language binding
 * (ctypes, .Call glue built against an older header) detect a layout mismatch instead of corrupting memory. */
int speigen_struct_sizes(int32_t* cfg_bytes, int32_t* out_bytes, int32_t* problem_bytes);
int speigen_device_count(void);

/* ---- host-only helpers (usable without a GPU) -------------------------------------------------- */
int speigen_validate_args(int64_t nrow, int64_t ncol, double q, double rho);      /* R/spEigen.R:52-55 */
int speigen_default_d(int32_t q, double* d);                                      /* R/spEigen.R:64 */
int speigen_schedule(const speigen_cfg* cfg, double* p, double* eps, double* tol);/* R/spEigen.R:94-102, K+1 each */
int speigen_stage_constants(double p, double eps, double* c1, double* c2, double* w0); /* R/spEigen.R:110-112 */
int speigen_shard_range(int64_t n, int32_t nparts, int32_t part, int64_t* lo, int64_t* hi);
int32_t speigen_panel_pitch(int32_t real_columns);
int64_t speigen_matrix_ld(int64_t contiguous_doubles);

/* ---- resident solver (device-side inputs; multi-process sharding; kernel-level entry points) ----
 * One solver per process drives the local GPU(s).  world/rank describe the row shard (covariance) or
 * sample shard (data matrix) this process owns.  Used by bench.py, the parity tests and by the
 * one-shot calls above. */
typedef struct speigen_solver speigen_solver;

typedef struct speigen_problem {
  int64_t m;            /* variables */
  int64_t n;            /* samples (data mode) ; ignored for covariance */
  int32_t q;
  int32_t is_complex;
  int32_t is_data;
  int32_t world;        /* total shards (GPUs) */
  int32_t rank;         /* first shard owned by this process */
  int32_t local_gpus;   /* shards owned by this process (devices first_device..) */
} speigen_problem;

int speigen_solver_create(speigen_solver** s, const speigen_problem* prob, const speigen_cfg* cfg);
int speigen_solver_destroy(speigen_solver* s);
/* NCCL bootstrap for multi-process runs: rank 0 makes the id, the launcher broadcasts the 128 bytes. */
int speigen_nccl_unique_id(void* id128);
int speigen_solver_init_comm(speigen_solver* s, const void* id128);
/* Load the full host matrix (column-major); each local device keeps only its shard. */
int speigen_solver_load_host(speigen_solver* s, const double* X_host);
/* Device pointer to local shard `local_idx` in library layout, so a caller can fill it on the device:
 * covariance: rows = owned columns of S (lo..hi), each row = one full column of S (m entries), pitch ld doubles;
 * data: rows = variables (m), each row = owned samples (lo..hi), pitch ld doubles. */
int speigen_solver_shard_buffer(speigen_solver* s, int32_t local_idx, double** dev_ptr, int64_t* rows,
                                int64_t* row_len_doubles, int64_t* ld_doubles, int64_t* lo, int64_t* hi);
/* Validation scan + centring + rho scale inputs (max diag / max column energy).  R/spEigen.R:53,59-60,72,75,82 */
int speigen_solver_prepare(speigen_solver* s);
/* Block subspace iteration for the top-q eigenpairs (R/spEigen.R:68-71,76-81 replacement). */
int speigen_solver_init(speigen_solver* s, const double* V0_host_or_null, speigen_out* out);
/* The MM loop (R/spEigen.R:105-160). */
int speigen_solver_run(speigen_solver* s, double rho, const double* d_or_null, double thres, speigen_out* out);

/* kernel-level entry points (host column-major panels in/out) */
int speigen_solver_contract(speigen_solver* s, const double* U_host, int32_t ncols, double* Y_host);
int speigen_solver_polar(speigen_solver* s, const double* A_host, int32_t ncols, double* U_host);
int speigen_solver_mm_update(speigen_solver* s, const double* V_host, const double* d, const double* rho_vec,
                             double p, double eps, int32_t fused, double* V1_host);
int speigen_solver_objective(speigen_solver* s, const double* V_host, const double* d, const double* rho_vec,
                             double p, double eps, double* F_out);
int speigen_solver_eig_small(speigen_solver* s, const double* C_host, int32_t n, double* W_host, double* evals);
/* timing: `reps` back-to-back contractions of a resident random panel, CUDA events on the launching stream,
 * L2 flushed between repetitions when flush_l2 != 0.  ms_each[reps]. */
int speigen_solver_time_contract(speigen_solver* s, int32_t ncols, int32_t reps, int32_t flush_l2, float* ms_each);
/* Device-to-device load of shard `local_idx` from a caller buffer already in HBM (same row meaning as
 * speigen_solver_shard_buffer; src_ld = source row pitch in doubles). */
int speigen_solver_load_device(speigen_solver* s, int32_t local_idx, const double* dev_src, int64_t src_ld);
/* Same, for rows [row0, row0 + nrows) of the shard only (lets a caller stream a matrix larger than its staging buffer). */
int speigen_solver_load_device_rows(speigen_solver* s, int32_t local_idx, int64_t row0, int64_t nrows,
                                    const double* dev_src, int64_t src_ld);
/* `nsteps` back-to-back MM updates V <- spEigenMMupdate(V) (R/spEigen.R:167-183) on the resident iterate at
 * continuation stage `stage` (0-based).  ms_out == NULL: asynchronous.  ms_out != NULL: the nsteps are bracketed by
 * CUDA events on the launching stream (synchronised on both sides) and the elapsed device time is returned. */
int speigen_solver_mm_steps(speigen_solver* s, int32_t nsteps, int32_t stage, double rho, const double* d_or_null,
                            float* ms_out);
int speigen_solver_sync(speigen_solver* s);
/* Diagnostics of the last fused tail kernel (device 0): out48[0] SQUAREM step a, [1] ||R||, [2] ||U||, [3] Gram passes, [4] Jacobi
 * sweeps, [7] number of clock stamps, [8..] phase clock of CTA 0 in ns since kernel entry (produce, then per pass: Gram, exchange,
 * solve, apply).  Measurement side channel only. */
int speigen_solver_tail_diag(speigen_solver* s, double* out48);
/* Hermitian-half upload of a covariance matrix (R/spEigen.R:75 guarantees S = S^H): how the last load moved the data. */
int speigen_solver_upload_info(speigen_solver* s, int32_t* staged, int32_t* half, int32_t* host_verified);
/* Host-only test hooks of that upload (no GPU): the complete procedure - unit plan, staging, isSymmetric's sums
 * sum|S - S^H| and sum|S| (R/spEigen.R:75), transposed rebuild - emulated in host memory.  S, A_out: m x m column-major
 * (interleaved complex), A_out must not alias S.  use_simd: 0 scalar reference, 1 the fused pack-and-compare pass (AVX2 where the
 * CPU has it), 2 the one-shot call's flow (row-wise pack without compare, the sums in a second pass over the caller's matrix). */
int speigen_host_half_upload_emulate(const double* S, int64_t m, int32_t is_complex, int32_t world, int64_t block,
                                     int64_t slot_bytes, int32_t use_simd, double* A_out, double* diff, double* sabs);
int speigen_host_has_avx2(void);
/* Bytes the most recent host-matrix load of this process (resident or one-shot) moved host -> device. */
int64_t speigen_last_upload_bytes(void);
/* Wall-clock side channel of the most recent one-shot call (speigen_run_*) of this process: out6[0] seconds spent waiting for the
 * deferred host evaluation of isSymmetric / anyNA (R/spEigen.R:53,75) after the device work, [1] seconds that evaluation took,
 * [2] seconds spent releasing a solver that is not kept, [3] the whole call, [4] 1 if the call found a kept solver of its shape
 * (matrices <= 256 MB: kept until speigen_shutdown; larger ones: kept SPEIGEN_KEEPALIVE_MS, default 2000, after their last call,
 * then released by a reaper thread), [5] how many solvers that reaper has released so far.  Measurement only. */
int speigen_last_call_times(double* out6);
/* Test hook: a device-side wait for an exchange epoch that never comes, bounded by budget_ms -> SPEIGEN_ERR_TIMEOUT (28). */
int speigen_solver_selftest_timeout(speigen_solver* s, int32_t budget_ms);
/* 0: single GPU (no exchange), 1: NCCL collectives, 2: fused peer-store exchange (valid after speigen_solver_init_comm) */
int speigen_solver_exchange_mode(speigen_solver* s);
/* CUDA-event timing of the contraction kernel launches of local device 0 (for the roofline line of bench.py). */
int speigen_solver_kernel_timing(speigen_solver* s, int32_t enable);
int speigen_solver_kernel_time(speigen_solver* s, double* ms_total, int32_t* launches);
int64_t speigen_solver_launch_count(speigen_solver* s);
int speigen_solver_stream(speigen_solver* s, int32_t local_idx, void** cuda_stream);

/* ==== SURVEY §8(f)-1: spEigenCov(S, q = 1, rho = 0.5, thres = 1e-9)                     R/spEigenCov.R:48 ========
 * returning list(vectors = m x m, values = m, cov = m x m)                                R/spEigenCov.R:135
 * Joint estimation of sparse eigenvectors and eigenvalues, real or complex Hermitian S (complex matrices run through the
 * same real kernels in embedded form [[Sr, -Si], [Si, Sr]]).  Additional status codes mirror its own stop() conditions. */
#define SPEIGEN_ERR_Q_GT_M 8          /* "q should not be larger than m."               R/spEigenCov.R:55 */
#define SPEIGEN_ERR_NOT_PSD 9         /* "The covariance matrix is not PSD."           R/spEigenCov.R:63 */
#define SPEIGEN_ERR_LOW_RANK 10       /* "The covariance matrix is low-rank."          R/spEigenCov.R:65 */

typedef struct speigencov_cfg {
  int32_t max_iter;          /* 3000   R/spEigenCov.R:49 */
  int32_t n_continuation;    /* K = 5  R/spEigenCov.R:80 (K+1 stages) */
  double p1;                 /* 1      :81 */
  double pk;                 /* 5      :82 */
  double tol_factor;         /* 1e-1   :87 */
  double eps_factor;         /* 1e-1   :88 */
  double shift;              /* 1e-6   :69 */
  double rank_tol;           /* 1e-9   :65 */
  double psd_tol;            /* -1e-10 :63 (`<=`) */
  double rho_scale;          /* 10     :73 */
  int32_t check_symmetric;   /* 1 */
  int32_t check_na;          /* 1 */
  int32_t device;            /* CUDA ordinal */
  int32_t jacobi_max_sweeps; /* 40: block one-sided Jacobi sweeps per SVD / EVD */
  double jacobi_tol;         /* 0 -> sqrt(m) * eps: relative column orthogonality at which the sweeps stop */
  int32_t polish;            /* 1: one Newton-Schulz step on the polar factor (orthonormal to working precision) */
  int32_t verbose;
} speigencov_cfg;
void speigencov_cfg_default(speigencov_cfg* cfg);

typedef struct speigencov_out {
  double* vectors;           /* m x m column-major   R/spEigenCov.R:131-133 */
  double* values;            /* m                    R/spEigenCov.R:107 */
  double* cov;               /* m x m column-major   R/spEigenCov.R:135 (nullable: skipped) */
  double* F;                 /* [trace_capacity] objective per iteration (optional; R/spEigenCov.R:114 computes but does not return it) */
  int32_t* stage_of_k;       /* [trace_capacity] (optional) */
  int32_t trace_capacity;
  int32_t iterations;
  int32_t eig_sweeps;        /* Jacobi sweeps of the initial eigen-decomposition */
  int32_t svd_sweeps;        /* total Jacobi sweeps over all Procrustes updates */
  int32_t gemms;             /* m x m x m GEMM launches */
  double seconds_upload;
  double seconds_init;
  double seconds_loop;
  double seconds_gemm;       /* device time inside the m x m x m GEMM launches (CUDA events) */
} speigencov_out;

/* one-shot drop-in call (host pointers): S is m x m column-major */
int speigencov_run_f64(const double* S, int64_t nrow, int64_t ncol, double q, double rho, double thres,
                       const speigencov_cfg* cfg, speigencov_out* out);
/* complex Hermitian S, interleaved {re, im}; vectors / cov are 2*m*m doubles (interleaved), values m doubles */
int speigencov_run_c64(const double* S_interleaved, int64_t nrow, int64_t ncol, double q, double rho, double thres,
                       const speigencov_cfg* cfg, speigencov_out* out);
/* R/spEigenCov.R:52-60 without the matrix scans (host only) */
int speigencov_validate_args(int64_t nrow, int64_t ncol, double q, double rho);
int speigencov_schedule(const speigencov_cfg* cfg, double* p, double* eps, double* tol);   /* :80-88, K+1 each */
int speigencov_struct_sizes(int32_t* cfg_bytes, int32_t* out_bytes);   /* layout self-check for foreign-language bindings */
/* eigvalAlgo(g, x, q), R/spEigenCov.R:164-338: ordered eigenvalue update (host only; xi_out[m]).
 * Returns SPEIGEN_ERR_ARG where R would stop (q == 1 without violators reads an undefined `tmp` - this build pools {q} alone
 * unless strict_r != 0). */
int speigencov_eigval_algo(const double* g, int64_t m, double x, int32_t q, int32_t strict_r, double* xi_out);

/* ---- dense kernel-level entry points (host column-major matrices in/out; used by the parity tests) -------------- */
/* C (M x N) = alpha * op(A) * op(B).  a_kmajor != 0: A is K x M (op(A) = A^T), else M x K.  b_kmajor != 0: B is K x N, else N x K
 * (op(B) = B^T).  symmetric != 0 (M == N, result known symmetric): upper tiles computed, mirrored.  ms_out (nullable): device time. */
int speigen_dense_gemm(int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, int a_kmajor, const double* B,
                       int64_t ldb, int b_kmajor, double* C, int64_t ldc, double alpha, int symmetric, int device,
                       float* ms_out);
/* Polar factor U = u v^T of svd(B) (R/spEigenCov.R:157-159), B is m x m. */
int speigen_dense_polar(const double* B, int64_t m, double* U, int32_t polish, int device, int32_t* sweeps_out);
/* Full symmetric eigen-decomposition, decreasing eigenvalues (eigen(S), R/spEigenCov.R:62). */
int speigen_dense_eigh(const double* S, int64_t m, double* vectors, double* values, int device, int32_t* sweeps_out);
/* `reps` back-to-back C = A^T B launches on resident random operands (M x N x K), CUDA events; ms_each[reps] */
int speigen_dense_gemm_time(int64_t M, int64_t N, int64_t K, int a_kmajor, int b_kmajor, int symmetric, int device,
                            int32_t reps, float* ms_each);

/* ==== SURVEY §8(f)-2: the upstream cov(X) on the device ==========================================================
 * Callers run spEigen(cov(X), q) (README.Rmd:108); forming S is a 2 n m^2 FP64 GEMM that would otherwise run in R
 * and cross PCIe as 8 m^2 bytes.  X is n x m column-major on the host; S = cov(X) (centred, / (n - 1)) is formed by the
 * GEMM kernel straight into each device's shard of the resident solver (covariance mode).  Follow with
 * speigen_solver_prepare / _init / _run. */
int speigen_solver_load_cov_of_data(speigen_solver* s, const double* X_host, int64_t n);
/* the same from a device-resident X (n x m column-major, leading dimension ldx) on local device 0's ordinal */
int speigen_solver_load_cov_of_data_device(speigen_solver* s, const double* X_dev, int64_t n, int64_t ldx);
/* one-shot: spEigen(cov(X), q, rho, d = d, V = V0, thres = thres) */
int speigen_run_cov_of_data_f64(const double* X, int64_t n, int64_t m, double q, double rho, const double* d,
                                const double* V0, double thres, const speigen_cfg* cfg, speigen_out* out);
/* cov(X) alone (host in/out): S_out is m x m column-major */
int speigen_cov_f64(const double* X, int64_t n, int64_t m, double* S_out, int device, float* ms_gemm_out);

#ifdef __cplusplus
}
#endif
#endif /* SPEIGEN_B200_H */
