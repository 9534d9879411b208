// extern "C" surface of libspeigen_b200.so (declared in include/speigen_b200.h).
#include "solver.cuh"
#include "host_pack.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <cstdlib>
#include <vector>

namespace speigen {
const char* last_error();
long long launch_count();
void release_cov_context();   // cov_solver.cu: the dense context speigencov_run_* keeps between calls
void release_stage_ring();    // solver.cu: the pinned staging ring of the pageable-upload path
}  // namespace speigen
using namespace speigen;

#define API_TRY(x)                       \
  do {                                   \
    int rc__ = (x);                      \
    if (rc__ != SPEIGEN_OK) return rc__; \
  } while (0)

// Every entry point leaves the caller's current CUDA device as it found it (a host such as PyTorch caches it).
struct DeviceRestore {
  int prev = -1;
  DeviceRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
};

static double wall_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

void speigen_cfg_default(speigen_cfg* c) {
  std::memset(c, 0, sizeof(*c));
  c->max_iter = 1000;
  c->n_continuation = 10;
  c->p1 = 1.0;
  c->pk = 7.0;
  c->tol_factor = 1e-2;
  c->backtrack_slack = 1e-9;
  c->rank_tol = 1e-9;
  c->pd_tol = -1e-10;
  c->max_backtracks = 200;
  c->init_max_passes = 300;
  c->init_tol = 2e-14;
  c->init_block = 0;
  c->init_seed = 42;
  c->check_symmetric = 1;
  c->check_na = 1;
  c->n_gpus = 1;
  c->first_device = 0;
  c->fused_epilogue = 1;
  c->verbose = 0;
  c->pd_check_passes = 4;
  c->p2p_exchange = 1;
  c->exchange_timeout_ms = 10000;
  c->allow_unconverged_init = 0;
}

const char* speigen_last_error(void) { return last_error(); }
int speigen_abi_version(void) { return SPEIGEN_ABI_VERSION; }
int speigen_struct_sizes(int32_t* cfg_bytes, int32_t* out_bytes, int32_t* problem_bytes) {
  if (cfg_bytes) *cfg_bytes = static_cast<int32_t>(sizeof(speigen_cfg));
  if (out_bytes) *out_bytes = static_cast<int32_t>(sizeof(speigen_out));
  if (problem_bytes) *problem_bytes = static_cast<int32_t>(sizeof(speigen_problem));
  return SPEIGEN_OK;
}
int speigen_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// R/spEigen.R:52-55 (anyNA(X) is checked on the device during solver_prepare)
int speigen_validate_args(int64_t nrow, int64_t ncol, double q, double rho) {
  (void)nrow;
  if (ncol == 1) { set_error("Data is univariate!"); return SPEIGEN_ERR_UNIVARIATE; }
  if (ncol < 1) { set_error("X must be a matrix"); return SPEIGEN_ERR_ARG; }
  if (std::isnan(q) || std::isnan(rho)) { set_error("This function cannot handle NAs."); return SPEIGEN_ERR_NA; }
  if (std::fmod(q, 1.0) != 0.0 || q <= 0) { set_error("The input argument q should be a positive integer."); return SPEIGEN_ERR_Q; }
  if (rho <= 0) { set_error("The input argument rho should be positive."); return SPEIGEN_ERR_RHO; }
  return SPEIGEN_OK;
}

// seq(from = 1, to = 0.5, length.out = q)   R/spEigen.R:64
int speigen_default_d(int32_t q, double* d) {
  if (q < 1 || !d) return SPEIGEN_ERR_ARG;
  if (q == 1) { d[0] = 1.0; return SPEIGEN_OK; }
  for (int i = 0; i < q; ++i) d[i] = 1.0 - 0.5 * static_cast<double>(i) / static_cast<double>(q - 1);
  return SPEIGEN_OK;
}

// pp <- 10^-(p1 * (pk/p1)^((0:K)/K)); tol <- pp * 1e-2; Eps <- pp    R/spEigen.R:94-102
int speigen_schedule(const speigen_cfg* cfg, double* p, double* eps, double* tol) {
  speigen_cfg def;
  if (!cfg) { speigen_cfg_default(&def); cfg = &def; }
  const int K = cfg->n_continuation;
  const double gamma = std::pow(cfg->pk / cfg->p1, 1.0 / K);
  for (int i = 0; i <= K; ++i) {
    const double e = cfg->p1 * std::pow(gamma, static_cast<double>(i));
    const double pp = std::pow(10.0, -e);
    if (p) p[i] = pp;
    if (eps) eps[i] = pp;
    if (tol) tol[i] = pp * cfg->tol_factor;
  }
  return SPEIGEN_OK;
}

int speigen_stage_constants(double p, double eps, double* c1, double* c2, double* w0) {
  const StageConsts s = make_stage(p, eps);
  if (c1) *c1 = s.c1;
  if (c2) *c2 = s.c2;
  if (w0) *w0 = s.w0;
  return SPEIGEN_OK;
}

int speigen_shard_range(int64_t n, int32_t nparts, int32_t part, int64_t* lo, int64_t* hi) {
  if (nparts < 1 || part < 0 || part >= nparts || n < 0) return SPEIGEN_ERR_ARG;
  const int64_t per = (n + nparts - 1) / nparts;
  const int64_t l = std::min<int64_t>(n, per * part);
  if (lo) *lo = l;
  if (hi) *hi = std::min<int64_t>(n, l + per);
  return SPEIGEN_OK;
}
int32_t speigen_panel_pitch(int32_t real_columns) { return panel_pitch(real_columns); }
int64_t speigen_matrix_ld(int64_t contiguous_doubles) { return (contiguous_doubles + 15) / 16 * 16; }

// ---- resident solver --------------------------------------------------------------------------
int speigen_solver_create(speigen_solver** s, const speigen_problem* prob, const speigen_cfg* cfg) {
  DeviceRestore restore_device__;
  if (!s || !prob) return SPEIGEN_ERR_ARG;
  speigen_cfg def;
  if (!cfg) { speigen_cfg_default(&def); cfg = &def; }
  speigen_solver* h = new (std::nothrow) speigen_solver();
  if (!h) return SPEIGEN_ERR_NOMEM;
  const int rc = h->impl.create(prob, cfg);
  if (rc != SPEIGEN_OK) { h->impl.destroy(); delete h; *s = nullptr; return rc; }
  *s = h;
  return SPEIGEN_OK;
}
int speigen_solver_destroy(speigen_solver* s) {
  DeviceRestore restore_device__;
  if (!s) return SPEIGEN_OK;
  s->impl.destroy();
  delete s;
  return SPEIGEN_OK;
}
int speigen_nccl_unique_id(void* id128) {
  const NcclApi* nc = nccl_api();
  if (!nc) return SPEIGEN_ERR_NCCL;
  ncclUniqueId id;
  const int r = nc->GetUniqueId(&id);
  if (r != NCCL_SUCCESS) { set_error("ncclGetUniqueId failed: %s", nc->GetErrorString(r)); return SPEIGEN_ERR_NCCL; }
  std::memcpy(id128, &id, sizeof(id));
  return SPEIGEN_OK;
}
int speigen_solver_init_comm(speigen_solver* s, const void* id128) {
  DeviceRestore restore_device__;
  if (!s) return SPEIGEN_ERR_ARG;
  return id128 ? s->impl.init_comm_rank(id128) : s->impl.init_comm_all();
}
int speigen_solver_load_host(speigen_solver* s, const double* X) {
  DeviceRestore restore_device__; return s->impl.load_host(X); }
int speigen_solver_shard_buffer(speigen_solver* s, int32_t li, double** dev_ptr, int64_t* rows, int64_t* row_len,
                                int64_t* ld, int64_t* lo, int64_t* hi) {
  DeviceRestore restore_device__;
  if (!s || li < 0 || li >= static_cast<int>(s->impl.devs.size())) return SPEIGEN_ERR_ARG;
  Dev& d = s->impl.devs[li];
  if (dev_ptr) *dev_ptr = d.A;
  if (rows) *rows = d.a_rows;
  if (row_len) *row_len = d.a_kd;
  if (ld) *ld = d.a_ld;
  if (lo) *lo = d.lo;
  if (hi) *hi = d.hi;
  s->impl.loaded = true;
  s->impl.host_verified = false;
  return SPEIGEN_OK;
}
int speigen_solver_prepare(speigen_solver* s) {
  DeviceRestore restore_device__; return s->impl.prepare(); }
int speigen_solver_init(speigen_solver* s, const double* V0, speigen_out* out) {
  DeviceRestore restore_device__; return s->impl.init_vectors(V0, out); }
int speigen_solver_run(speigen_solver* s, double rho, const double* d, double thres, speigen_out* out) {
  DeviceRestore restore_device__;
  if (std::isnan(rho)) { set_error("This function cannot handle NAs."); return SPEIGEN_ERR_NA; }
  if (rho <= 0) { set_error("The input argument rho should be positive."); return SPEIGEN_ERR_RHO; }
  return s->impl.run(rho, d, thres, out);
}
int64_t speigen_solver_launch_count(speigen_solver*) {
  DeviceRestore restore_device__; return launch_count(); }
int speigen_solver_stream(speigen_solver* s, int32_t li, void** st) {
  DeviceRestore restore_device__;
  if (!s || li < 0 || li >= static_cast<int>(s->impl.devs.size())) return SPEIGEN_ERR_ARG;
  *st = s->impl.devs[li].st;
  return SPEIGEN_OK;
}

// ---- kernel-level entry points ----------------------------------------------------------------
static int ensure_maps(Solver& S) {
  if (!S.devs[0].Astd.ptr) return S.build_tensor_maps();
  return SPEIGEN_OK;
}

int speigen_solver_contract(speigen_solver* s, const double* U_host, int32_t ncols, double* Y_host) {
  DeviceRestore restore_device__;
  Solver& S = s->impl;
  if (ncols < 1 || ncols > S.bq) { set_error("contract: ncols %d outside 1..%d", ncols, S.bq); return SPEIGEN_ERR_ARG; }
  API_TRY(ensure_maps(S));
  const PanelDims pd = make_dims(S.m, ncols, S.cplx);
  API_TRY(S.op_upload_panel(U_host, BQ, pd));
  API_TRY(S.op_contract(BQ, pd, BZ, NBUF, NBUF, false, make_stage(0.1, 0.1)));
  return S.op_download_panel(BZ, pd, Y_host);
}

// status of the tail kernels after a kernel-level call (sticky device flag -> error code)
static int tail_status(Solver& S) {
  API_TRY(S.sync_all());
  int st = 0;
  Dev& d = S.devs[0];
  cudaSetDevice(d.ordinal);
  SPE_CUDA(cudaMemcpy(&st, d.tstatus, sizeof(int), cudaMemcpyDeviceToHost));
  return S.check_tail_status(static_cast<double>(st));
}

int speigen_solver_polar(speigen_solver* s, const double* A_host, int32_t ncols, double* U_host) {
  DeviceRestore restore_device__;
  Solver& S = s->impl;
  if (ncols < 1 || ncols > S.bq) { set_error("polar: ncols %d outside 1..%d", ncols, S.bq); return SPEIGEN_ERR_ARG; }
  const PanelDims pd = make_dims(S.m, ncols, S.cplx);
  API_TRY(S.op_upload_panel(A_host, BQ, pd));
  TailSpec ts;                       // the fused polar tail (K3): Gram passes + Jacobi/Taylor + applies in one kernel
  ts.producer = TAIL_PANEL; ts.T = BQ; ts.out = BZ; ts.out_all = true;
  API_TRY(S.op_tail(ts, pd, make_stage(0.1, 0.1)));
  API_TRY(S.wait_latest_tail());
  API_TRY(tail_status(S));
  return S.op_download_panel(BZ, pd, U_host);
}

int speigen_solver_mm_update(speigen_solver* s, const double* V_host, const double* d, const double* rho_vec, double p,
                             double eps, int32_t fused, double* V1_host) {
  DeviceRestore restore_device__;
  Solver& S = s->impl;
  API_TRY(ensure_maps(S));
  const PanelDims pd = S.pdq;
  const StageConsts sc = make_stage(p, eps);
  API_TRY(S.op_upload_panel(V_host, BV, pd));
  API_TRY(S.set_qvecs(d, rho_vec));
  {
    TailSpec ts;   // column minima of V (the weights' column max, R/spEigen.R:176)
    ts.T = BV; ts.out = BV; ts.skip_polar = true; ts.absmin = true;
    API_TRY(S.op_tail(ts, pd, sc));
  }
  TailSpec ts;
  ts.T = BB; ts.out = BV1; ts.out_all = true; ts.absmin = true;   // cold Jacobi start: the result does not depend on earlier calls
  if (fused && !(S.data && S.prob.world > 1)) {
    API_TRY(S.op_contract_loop(BV, pd, NBUF, BB, BV, false, sc));
    ts.producer = TAIL_PANEL;
  } else {
    API_TRY(S.op_contract_loop(BV, pd, BY, NBUF, NBUF, false, sc));
    ts.producer = TAIL_REWEIGHT; ts.in0 = BY; ts.in1 = BV; ts.use_prev = true;
  }
  API_TRY(S.op_tail(ts, pd, sc));
  API_TRY(S.wait_latest_tail());
  API_TRY(tail_status(S));
  return S.op_download_panel(BV1, pd, V1_host);
}

int speigen_solver_objective(speigen_solver* s, const double* V_host, const double* d, const double* rho_vec, double p,
                             double eps, double* F_out) {
  DeviceRestore restore_device__;
  Solver& S = s->impl;
  API_TRY(ensure_maps(S));
  const PanelDims pd = S.pdq;
  const StageConsts sc = make_stage(p, eps);
  API_TRY(S.op_upload_panel(V_host, BV0, pd));
  API_TRY(S.set_qvecs(d, rho_vec));
  {
    TailSpec ts;   // penalty column sums of V0 (R/spEigen.R:133-134)
    ts.T = BV0; ts.out = BV0; ts.skip_polar = true; ts.absmin = true; ts.gsum = true;
    API_TRY(S.op_tail(ts, pd, sc));
  }
  API_TRY(S.op_contract_loop(BV0, pd, BY0, NBUF, BV0, true, sc));
  API_TRY(S.op_finish_objective(pd));
  API_TRY(tail_status(S));
  *F_out = S.h_res[0];
  return SPEIGEN_OK;
}

int speigen_solver_eig_small(speigen_solver* s, const double* C_host, int32_t nn, double* W_host, double* evals) {
  DeviceRestore restore_device__;
  Solver& S = s->impl;
  if (nn < 1 || nn > S.bq) { set_error("eig_small: n %d outside 1..%d", nn, S.bq); return SPEIGEN_ERR_ARG; }
  const int e = 1 + S.cplx;
  const PanelDims pd = make_dims(S.m, nn, S.cplx);
  const int ncr = nn * e;
  // one "partial" in the real-Gram layout k_small sums: C[a][b] = (R[2a][2b] + R[2a+1][2b+1]) + i (R[2a][2b+1] - R[2a+1][2b])
  std::vector<double> rm(static_cast<size_t>(ncr) * ncr, 0.0), wm(static_cast<size_t>(ncr) * ncr);
  for (int a = 0; a < nn; ++a)
    for (int b = 0; b < nn; ++b) {
      const double* src = C_host + (static_cast<size_t>(b) * nn + a) * e;   // column-major input
      if (S.cplx) {
        rm[static_cast<size_t>(2 * a) * ncr + 2 * b] = src[0];
        rm[static_cast<size_t>(2 * a) * ncr + 2 * b + 1] = src[1];
      } else {
        rm[static_cast<size_t>(a) * ncr + b] = src[0];
      }
    }
  Dev& d = S.devs[0];
  cudaSetDevice(d.ordinal);
  SPE_CUDA(cudaMemcpyAsync(d.ps.gram_part, rm.data(), rm.size() * sizeof(double), cudaMemcpyHostToDevice, d.st));
  API_TRY(launch_small(pd, d.ps.gram_part, 1, SMALL_EIGVECS, d.Msmall, d.evals, d.status, nullptr, d.st));
  SPE_CUDA(cudaMemcpyAsync(wm.data(), d.Msmall, wm.size() * sizeof(double), cudaMemcpyDeviceToHost, d.st));
  SPE_CUDA(cudaMemcpyAsync(evals, d.evals, nn * sizeof(double), cudaMemcpyDeviceToHost, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  for (int a = 0; a < nn; ++a)          // hat form -> column-major (complex) eigenvectors
    for (int b = 0; b < nn; ++b) {
      double* dst = W_host + (static_cast<size_t>(b) * nn + a) * e;
      if (S.cplx) {
        dst[0] = wm[static_cast<size_t>(2 * a) * ncr + 2 * b];
        dst[1] = wm[static_cast<size_t>(2 * a) * ncr + 2 * b + 1];
      } else {
        dst[0] = wm[static_cast<size_t>(a) * ncr + b];
      }
    }
  return SPEIGEN_OK;
}

int speigen_solver_time_contract(speigen_solver* s, int32_t ncols, int32_t reps, int32_t flush_l2, float* ms_each) {
  DeviceRestore restore_device__;
  Solver& S = s->impl;
  API_TRY(ensure_maps(S));
  return S.time_contract(ncols, reps, flush_l2, ms_each);
}

int speigen_solver_load_device(speigen_solver* s, int32_t li, const double* dev_src, int64_t src_ld) {
  DeviceRestore restore_device__;
  return s->impl.load_device(li, dev_src, src_ld);
}
int speigen_solver_load_device_rows(speigen_solver* s, int32_t li, int64_t row0, int64_t nrows, const double* dev_src, int64_t src_ld) {
  DeviceRestore restore_device__;
  return s->impl.load_device_rows(li, row0, nrows, dev_src, src_ld);
}
int speigen_solver_mm_steps(speigen_solver* s, int32_t nsteps, int32_t stage, double rho, const double* d, float* ms_out) {
  DeviceRestore restore_device__;
  return s->impl.mm_steps(nsteps, stage, rho, d, ms_out);
}
// Test hook for the bounded device-side waits: waits on the own mailbox for an exchange epoch that no peer will ever publish,
// with the given budget, and returns what a solve would return in that situation (SPEIGEN_ERR_TIMEOUT).
int speigen_solver_selftest_timeout(speigen_solver* s, int32_t budget_ms) {
  DeviceRestore restore_device__;
  if (!s) return SPEIGEN_ERR_ARG;
  Solver& S = s->impl;
  Dev& d = S.devs[0];
  cudaSetDevice(d.ordinal);
  API_TRY(launch_wait_flags(d.mail->stat_flag, kFlagStride, 1, S.stat_epoch + 1000000ull, d.tstatus,
                            static_cast<long long>(budget_ms) * 1000000ll, d.st));
  return tail_status(S);
}
// Host-only test hooks of the Hermitian-half upload (host_pack.h): no GPU needed.
int speigen_host_half_upload_emulate(const double* S, int64_t m, int32_t is_complex, int32_t world, int64_t block,
                                     int64_t slot_bytes, int32_t use_simd, double* A_out, double* diff, double* sabs) {
  if (!S || !A_out || !diff || !sabs || m < 1 || world < 1 || block < 1 || slot_bytes < 16) return SPEIGEN_ERR_ARG;
  emulate_half_upload(S, m, is_complex ? 2 : 1, world, block, static_cast<size_t>(slot_bytes), use_simd, A_out, diff, sabs);
  return SPEIGEN_OK;
}
int speigen_host_has_avx2(void) { return host_has_avx2() ? 1 : 0; }
int64_t speigen_last_upload_bytes(void) { return static_cast<int64_t>(last_upload_bytes()); }
int speigen_solver_upload_info(speigen_solver* s, int32_t* staged, int32_t* half, int32_t* host_verified) {
  if (!s) return SPEIGEN_ERR_ARG;
  if (staged) *staged = s->impl.upload_staged ? 1 : 0;
  if (half) *half = s->impl.upload_half ? 1 : 0;
  if (host_verified) *host_verified = s->impl.host_verified ? 1 : 0;
  return SPEIGEN_OK;
}
int speigen_solver_tail_diag(speigen_solver* s, double* out48) {
  DeviceRestore restore_device__;
  if (!s || !out48) return SPEIGEN_ERR_ARG;
  API_TRY(s->impl.sync_all());
  Dev& d = s->impl.devs[0];
  cudaSetDevice(d.ordinal);
  SPE_CUDA(cudaMemcpy(out48, d.tdiag, 48 * sizeof(double), cudaMemcpyDeviceToHost));
  return SPEIGEN_OK;
}
int speigen_solver_sync(speigen_solver* s) {
  DeviceRestore restore_device__; return s->impl.sync_all(); }
int speigen_solver_exchange_mode(speigen_solver* s) {
  DeviceRestore restore_device__; return s->impl.prob.world <= 1 ? 0 : (s->impl.p2p ? 2 : 1); }
int speigen_solver_kernel_timing(speigen_solver* s, int32_t enable) {
  DeviceRestore restore_device__; s->impl.enable_kernel_timing(enable != 0); return SPEIGEN_OK; }
int speigen_solver_kernel_time(speigen_solver* s, double* ms_total, int32_t* launches) {
  DeviceRestore restore_device__;
  int n = 0;
  const int rc = s->impl.kernel_time(ms_total, &n);
  *launches = n;
  return rc;
}

// ---- one-shot drop-in calls -------------------------------------------------------------------
// One solver of the last shape is kept alive between one-shot calls: creating and releasing the device context costs more than
// a small solve (~7 of 16 ms at m <= 800) and 20-300 ms for a 20 GB matrix (cudaMalloc / cudaFree of the shards, varying from
// call to call).  Small solvers (matrix <= 256 MB, SPEIGEN_CACHE_MAX_BYTES) stay until speigen_shutdown() or process exit; a
// LARGE one is kept for `SPEIGEN_KEEPALIVE_MS` (default 2000) after the call that used it - long enough for the next call of a
// loop over rho / q / bootstrap draws to find it - and is then released by a reaper thread, so an idle R session does not
// sit on tens of GB of device memory.  SPEIGEN_KEEPALIVE_MS=0: released inside the call (the round-1 behaviour).
static std::mutex g_oneshot_mtx;                 // serialises one-shot calls and guards everything below
static speigen_solver* g_cached = nullptr;
static speigen_problem g_cached_prob;
static speigen_cfg g_cached_cfg;
static double g_cached_deadline = 0.0;           // wall_s() at which the reaper releases g_cached; 0: no deadline
static std::condition_variable g_reaper_cv;
static std::thread g_reaper;
static bool g_reaper_stop = false;
static std::atomic<int> g_reaped{0};             // how many solvers the reaper has released (test / measurement side channel)

static void reaper_main() {
  std::unique_lock<std::mutex> lk(g_oneshot_mtx);
  while (!g_reaper_stop) {
    if (g_cached && g_cached_deadline > 0.0) {
      const double now = wall_s();
      if (now >= g_cached_deadline) {
        speigen_solver_destroy(g_cached);        // under the lock: a call that arrives right now waits for the memory, as it used to
        g_cached = nullptr;
        g_cached_deadline = 0.0;
        g_reaped.fetch_add(1);
        continue;
      }
      g_reaper_cv.wait_for(lk, std::chrono::duration<double>(g_cached_deadline - now));
    } else {
      g_reaper_cv.wait(lk);
    }
  }
}

static void stop_reaper() {
  {
    std::lock_guard<std::mutex> lk(g_oneshot_mtx);
    g_reaper_stop = true;
  }
  g_reaper_cv.notify_all();
  if (g_reaper.joinable()) g_reaper.join();
  g_reaper_stop = false;
}

// caller holds g_oneshot_mtx
static void arm_reaper(double keepalive_s) {
  g_cached_deadline = wall_s() + keepalive_s;
  if (!g_reaper.joinable()) {
    static bool hooked = false;
    if (!hooked) {
      hooked = true;
      atexit([]() { stop_reaper(); });           // registered after the CUDA runtime's own exit handlers, so it runs before them
    }
    g_reaper = std::thread(reaper_main);
  }
  g_reaper_cv.notify_all();
}

int speigen_shutdown(void) {
  DeviceRestore restore_device__;
  stop_reaper();
  {
    std::lock_guard<std::mutex> lk(g_oneshot_mtx);
    if (g_cached) { speigen_solver_destroy(g_cached); g_cached = nullptr; }
    g_cached_deadline = 0.0;
  }
  release_cov_context();
  release_stage_ring();
  return SPEIGEN_OK;
}

// wall-clock side channel of the last one-shot call: [0] wait for the deferred host verdict, [1] how long the host workers took for
// it (from their start to the last one finishing), [2] release of a solver that is not kept, [3] whole call, [4] 1 if the call
// found a kept solver, [5] solvers released by the reaper so far
static double g_last_call_times[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
int speigen_last_call_times(double* out6) {
  if (!out6) return SPEIGEN_ERR_ARG;
  for (int i = 0; i < 5; ++i) out6[i] = g_last_call_times[i];
  out6[5] = static_cast<double>(g_reaped.load());
  return SPEIGEN_OK;
}

static int run_oneshot(const double* X, int64_t nrow, int64_t ncol, int is_data, int is_complex, double q, double rho,
                       const double* d, const double* V0, double thres, const speigen_cfg* cfg_in, speigen_out* out) {
  DeviceRestore restore_device__;
  if (!X || !out) { set_error("null argument"); return SPEIGEN_ERR_ARG; }
  speigen_cfg cfg;
  if (cfg_in) cfg = *cfg_in; else speigen_cfg_default(&cfg);
  API_TRY(speigen_validate_args(nrow, ncol, q, rho));
  if (!is_data && nrow != ncol) { set_error("The covariance matrix is not symmetric."); return SPEIGEN_ERR_NOT_SYMMETRIC; }
  speigen_problem pb;
  std::memset(&pb, 0, sizeof(pb));
  pb.m = ncol;
  pb.n = nrow;
  pb.q = static_cast<int32_t>(q);
  pb.is_complex = is_complex;
  pb.is_data = is_data ? 1 : 0;
  pb.world = cfg.n_gpus > 0 ? cfg.n_gpus : 1;
  pb.rank = 0;
  pb.local_gpus = pb.world;
  // R calls are single-threaded; the mutex keeps other hosts safe (and the reaper out of a running call).
  std::lock_guard<std::mutex> lock(g_oneshot_mtx);
  const size_t e = is_complex ? 2 : 1;
  const size_t mat_bytes = static_cast<size_t>(nrow) * ncol * e * sizeof(double);
  size_t cache_max = 256u << 20;
  if (const char* cm = getenv("SPEIGEN_CACHE_MAX_BYTES")) { const long long v = atoll(cm); if (v >= 0) cache_max = static_cast<size_t>(v); }
  double keepalive_s = 2.0;
  if (const char* ka = getenv("SPEIGEN_KEEPALIVE_MS")) { const double v = atof(ka); if (v >= 0.0) keepalive_s = 1e-3 * v; }
  const bool small = mat_bytes <= cache_max;
  const bool cacheable = small || keepalive_s > 0.0;
  const double t_setup0 = wall_s();
  g_cached_deadline = 0.0;                       // whatever is cached is ours now (taken or replaced below)
  speigen_solver* s = nullptr;
  g_last_call_times[4] = 0.0;
  if (g_cached && cacheable && std::memcmp(&g_cached_prob, &pb, sizeof(pb)) == 0 && std::memcmp(&g_cached_cfg, &cfg, sizeof(cfg)) == 0) {
    s = g_cached;
    g_cached = nullptr;
    s->impl.reset_for_reuse();
    g_last_call_times[4] = 1.0;
  } else {
    if (g_cached) { speigen_solver_destroy(g_cached); g_cached = nullptr; }
    API_TRY(speigen_solver_create(&s, &pb, &cfg));
  }
  int rc = s->impl.comm_ready ? SPEIGEN_OK : s->impl.init_comm_all();
  const double t0 = wall_s();
  out->seconds_setup = t0 - t_setup0;
  {
    // isSymmetric / anyNA of a half-uploaded S are evaluated on the host while the GPUs work (SPEIGEN_DEFER_VERIFY=0: inside the upload)
    const char* dv = getenv("SPEIGEN_DEFER_VERIFY");
    s->impl.defer_host_verify = !(dv && dv[0] == '0');
  }
  if (rc == SPEIGEN_OK) rc = s->impl.load_host(X);
  const double t1 = wall_s();
  out->seconds_upload = t1 - t0;
  out->upload_staged = (s->impl.upload_staged ? 1 : 0) | (s->impl.upload_half ? 2 : 0);
  if (rc == SPEIGEN_OK) rc = s->impl.prepare();
  out->seconds_prepare = wall_s() - t1;
  if (rc == SPEIGEN_OK) rc = s->impl.init_vectors(V0, out);
  if (rc == SPEIGEN_OK) rc = s->impl.run(rho, d, thres, out);
  {
    // the deferred host verdict comes first in the reference's order of stop()s (R/spEigen.R:53,75 precede :77,:79)
    const double tv = wall_s();
    const int vrc = s->impl.host_verdict();
    if (vrc != SPEIGEN_OK) rc = vrc;
    g_last_call_times[0] = wall_s() - tv;
    g_last_call_times[1] = s->impl.host_verify_seconds;
  }
  const double td = wall_s();
  if (rc == SPEIGEN_OK && cacheable) {
    g_cached = s;
    g_cached_prob = pb;
    g_cached_cfg = cfg;
    if (!small) arm_reaper(keepalive_s);
  } else {
    speigen_solver_destroy(s);
  }
  g_last_call_times[2] = wall_s() - td;
  g_last_call_times[3] = wall_s() - t_setup0;
  return rc;
}

int speigen_run_f64(const double* X, int64_t nrow, int64_t ncol, int is_data, double q, double rho, const double* d,
                    const double* V0, double thres, const speigen_cfg* cfg, speigen_out* out) {
  return run_oneshot(X, nrow, ncol, is_data, 0, q, rho, d, V0, thres, cfg, out);
}
int speigen_run_c64(const double* X, int64_t nrow, int64_t ncol, int is_data, double q, double rho, const double* d,
                    const double* V0, double thres, const speigen_cfg* cfg, speigen_out* out) {
  return run_oneshot(X, nrow, ncol, is_data, 1, q, rho, d, V0, thres, cfg, out);
}

}  // extern "C"
