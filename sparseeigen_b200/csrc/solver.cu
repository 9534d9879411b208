// Host driver (see solver.cuh).  Reference semantics: R/spEigen.R:46-163.
#include "solver.cuh"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include "host_pack.h"
#include <mutex>
#include <string>
#include <thread>

namespace speigen {

// ---- error / counters ---------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
const char* last_error() { return g_err.c_str(); }
static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(); }

#define SPE_TRY(x)                      \
  do {                                  \
    int rc__ = (x);                     \
    if (rc__ != SPEIGEN_OK) return rc__; \
  } while (0)
#define SPE_NCCL(x)                                                                         \
  do {                                                                                      \
    int r__ = (x);                                                                          \
    if (r__ != NCCL_SUCCESS) {                                                              \
      set_error("NCCL error %d (%s) at %s:%d", r__, nccl_api()->GetErrorString(r__), __FILE__, __LINE__); \
      return SPEIGEN_ERR_NCCL;                                                              \
    }                                                                                       \
  } while (0)

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

StageConsts make_stage(double p, double eps) {
  StageConsts s;
  s.p = p;
  s.eps = eps;
  s.c1 = std::log(1.0 + 1.0 / p);          // R/spEigen.R:110
  s.c2 = 2.0 * (p + eps) * s.c1;           // :111
  s.w0 = 1.0 / (eps * s.c2);               // :112
  s.half_over_c1 = 0.5 / s.c1;             // :173
  return s;
}

static int64_t round_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- creation -------------------------------------------------------------------------------------
int Solver::create(const speigen_problem* p, const speigen_cfg* c) {
  prob = *p;
  cfg = *c;
  cplx = p->is_complex ? 1 : 0;
  data = p->is_data ? 1 : 0;
  m = p->m;
  n = data ? p->n : p->m;
  q = p->q;
  if (m < 2) { set_error("Data is univariate!"); return SPEIGEN_ERR_UNIVARIATE; }
  if (q < 1) { set_error("The input argument q should be a positive integer."); return SPEIGEN_ERR_Q; }
  if (q > m) { set_error("The number of estimated eigenvectors q should not be larger than rank(X)."); return SPEIGEN_ERR_RANK; }
  if (p->world < 1 || p->local_gpus < 1 || p->rank < 0 || p->rank + p->local_gpus > p->world) {
    set_error("bad shard description (world=%d rank=%d local=%d)", p->world, p->rank, p->local_gpus);
    return SPEIGEN_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    set_error("no CUDA device available: speigen_b200 has no CPU fallback");
    return SPEIGEN_ERR_NO_DEVICE;
  }
  if (cfg.first_device + p->local_gpus > ndev) {
    set_error("requested %d GPUs starting at %d but only %d visible", p->local_gpus, cfg.first_device, ndev);
    return SPEIGEN_ERR_NO_DEVICE;
  }
  // initialiser block: fill the DMMA n-tile (free oversampling), bounded by the attainable rank
  const int e = 1 + cplx;
  int b = cfg.init_block > 0 ? cfg.init_block : (cplx ? (q + 3) / 4 * 4 : (q + 7) / 8 * 8);
  int64_t rank_cap = m;
  if (data) rank_cap = std::min<int64_t>(m, std::max<int64_t>(n - 1, 1));
  if (b > rank_cap) b = static_cast<int>(rank_cap);
  if (b < q) b = q;
  bq = b;
  bq0 = b;
  if (bq * e > kMaxCols) {
    set_error("q = %d (%s) exceeds this build's limit of %d real columns per contraction", q, cplx ? "complex" : "real", kMaxCols);
    return SPEIGEN_ERR_UNSUPPORTED;
  }
  pdq = make_dims(m, q, cplx);
  pdb = make_dims(m, bq, cplx);
  const int64_t shard_dim = data ? n : m;
  rows_per_shard = (shard_dim + prob.world - 1) / prob.world;
  mrows = padded_rows(m);
  if (!data) mrows = std::max<int64_t>(mrows, padded_rows(rows_per_shard * prob.world));
  devs.resize(p->local_gpus);
  for (int i = 0; i < p->local_gpus; ++i) {
    Dev& d = devs[i];
    d.ordinal = cfg.first_device + i;
    d.shard = prob.rank + i;
    d.lo = std::min<int64_t>(shard_dim, rows_per_shard * d.shard);
    d.hi = std::min<int64_t>(shard_dim, d.lo + rows_per_shard);
    SPE_TRY(alloc_dev(d));
  }
  SPE_CUDA(cudaSetDevice(devs[0].ordinal));
  SPE_CUDA(cudaMallocHost(&h_scal, 512 * sizeof(double)));
  // result block of the MM loop: written by k_finish_objective straight into mapped pinned memory (no copy on the critical path)
  SPE_CUDA(cudaHostAlloc(&h_res, 512 * sizeof(double), cudaHostAllocMapped | cudaHostAllocPortable));
  std::memset(h_res, 0, 512 * sizeof(double));
  SPE_CUDA(cudaHostGetDevicePointer(&devs[0].res_host_dev, h_res, 0));
  if (prob.world == 1) {
    Dev& d0 = devs[0];
    d0.peer_arena[0] = static_cast<char*>(d0.arena);
  }
  return SPEIGEN_OK;
}

int Solver::alloc_dev(Dev& d) {
  SPE_CUDA(cudaSetDevice(d.ordinal));
  cudaDeviceProp pr;
  SPE_CUDA(cudaGetDeviceProperties(&pr, d.ordinal));
  if (pr.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", d.ordinal, pr.major, pr.minor);
    return SPEIGEN_ERR_NO_DEVICE;
  }
  d.num_sms = pr.multiProcessorCount;
  SPE_CUDA(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
  const int e = 1 + cplx;
  if (data) { d.a_rows = m; d.a_kd = (d.hi - d.lo) * e; }
  else { d.a_rows = d.hi - d.lo; d.a_kd = m * e; }
  d.a_ld = round_up64(std::max<int64_t>(d.a_kd, 1), 16);
  const size_t a_bytes = static_cast<size_t>(std::max<int64_t>(d.a_rows, 1)) * d.a_ld * sizeof(double);
  if (cudaMalloc(&d.A, a_bytes) != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory allocating the %.2f GB matrix shard", a_bytes / 1e9);
    return SPEIGEN_ERR_NOMEM;
  }
  // Everything except the matrix shard and the IPC-exported exchange buffers is carved out of ONE arena per device:
  // ~45 cudaMalloc/cudaFree pairs per device cost more than the whole 8-GPU solve (0.35 s of a 0.53 s call).
  const size_t pbytes = static_cast<size_t>(mrows) * pdb.pitch * sizeof(double);
  std::vector<std::pair<void**, size_t>> req;
  auto want = [&](auto** p, size_t bytes) { req.emplace_back(reinterpret_cast<void**>(p), (bytes + 255) / 256 * 256); };
  // peer-visible head of the arena: panels, then the mailbox - sizes depend on global parameters only, so the offsets are the
  // same on every rank and a peer's copy of panel b is peer_arena[r] + off_buf[b]
  for (int bb = 0; bb < NBUF; ++bb) want(&d.buf[bb], pbytes);
  want(&d.mail, sizeof(Mailbox));
  want(&d.cta_part, static_cast<size_t>(d.num_sms) * kTriMax * sizeof(double));
  want(&d.cta_stat, static_cast<size_t>(d.num_sms) * kStatLen * sizeof(double));
  want(&d.tail_ticket, 4 * sizeof(int));
  want(&d.gram_epoch, sizeof(unsigned long long));
  want(&d.a_cur, sizeof(double));
  want(&d.tdiag, 48 * sizeof(double));
  want(&d.tstatus, sizeof(int));
  want(&d.twarm, static_cast<size_t>(N_TAIL_SITES) * 2 * kWarmDoubles * sizeof(double));
  if (cplx && !data) want(&d.hat, 2 * pbytes);
  if (data) {
    want(&d.Tdata, static_cast<size_t>(padded_rows(std::max<int64_t>(d.a_kd, 1))) * pdb.pitch * sizeof(double));
    if (prob.world > 1) want(&d.gather, pbytes * prob.world);
  }
  // contraction workspace (both forms share it; sized for the larger)
  const int ncr_max = bq * e;
  size_t slots = contract_workspace_slots(d.a_rows, d.num_sms);
  const size_t slot_d = contract_slot_doubles(ncr_max);
  int ncount = static_cast<int>((d.a_rows + kTJ - 1) / kTJ);
  if (data) {
    slots = std::max(slots, contract_tr_workspace_slots(d.a_kd, d.num_sms));
    ncount = std::max(ncount, static_cast<int>((d.a_kd + 127) / 128));
  }
  d.ws.partial_slots = slots;
  d.ws.n_counters = ncount + 1;
  want(&d.ws.partials, slots * slot_d * sizeof(double));
  want(&d.ws.counters, d.ws.n_counters * sizeof(int));
  const size_t glen = static_cast<size_t>(bq * e) * (bq * e);   // real Gram / real hat form of the small matrix
  want(&d.ps.gram_part, kPanelBlocks * glen * sizeof(double));
  want(&d.gram_part2, kPanelBlocks * glen * sizeof(double));
  want(&d.ps.col_part, kPanelBlocks * 2 * kMaxCols * sizeof(double));
  want(&d.ps.ticket, sizeof(int));
  want(&d.absmin_a, (kPanelBlocks + 1) * kMaxCols * sizeof(double));
  want(&d.absmin_b, (kPanelBlocks + 1) * kMaxCols * sizeof(double));
  want(&d.Msmall, glen * sizeof(double));
  want(&d.evals, kMaxCols * sizeof(double));
  want(&d.warm, 4 * kWarmStride * sizeof(double));
  want(&d.status, sizeof(int));
  want(&d.dvec, kMaxCols * sizeof(double));
  want(&d.rho, kMaxCols * sizeof(double));
  const size_t ntiles = static_cast<size_t>((d.a_rows + kTJ - 1) / kTJ) + 1;
  want(&d.dots_part, std::max<size_t>(ntiles, kPanelBlocks) * kMaxCols * sizeof(double));
  want(&d.dots_gather, static_cast<size_t>(prob.world) * kMaxCols * sizeof(double));
  want(&d.done_count, sizeof(int));
  want(&d.scal, 512 * sizeof(double));
  want(&d.vecm, static_cast<size_t>(2 * m + 16) * sizeof(double));
  want(&d.vecm2, static_cast<size_t>(2 * m + 16) * sizeof(double));
  want(&d.scan_part, kScanBlocks * 8 * sizeof(double));
  want(&d.out_cm, static_cast<size_t>(m) * bq * e * sizeof(double));
  size_t total = 0;
  for (auto& r : req) total += r.second;
  if (cudaMalloc(&d.arena, total) != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory allocating %.2f GB of panel workspace", total / 1e9);
    return SPEIGEN_ERR_NOMEM;
  }
  SPE_CUDA(cudaMemsetAsync(d.arena, 0, total, d.st));
  size_t off = 0;
  for (auto& r : req) { *r.first = static_cast<char*>(d.arena) + off; off += r.second; }
  for (int bb = 0; bb < NBUF; ++bb) d.off_buf[bb] = reinterpret_cast<char*>(d.buf[bb]) - static_cast<char*>(d.arena);
  d.off_mail = reinterpret_cast<char*>(d.mail) - static_cast<char*>(d.arena);
  SPE_TRY(tail_prepare_device());   // load the module / set attributes now: later launches must never block behind a spinning peer
  if (!data && prob.world > 1 && prob.world <= kMaxPeers) {   // separate allocations: exported with cudaIpcGetMemHandle
    const size_t xb_bytes = 4 * pbytes;
    SPE_CUDA(cudaMalloc(&d.xb, xb_bytes));
    SPE_CUDA(cudaMemsetAsync(d.xb, 0, xb_bytes, d.st));
    SPE_CUDA(cudaMalloc(&d.xdots, 2 * static_cast<size_t>(prob.world) * kMaxCols * sizeof(double)));
    SPE_CUDA(cudaMemsetAsync(d.xdots, 0, 2 * static_cast<size_t>(prob.world) * kMaxCols * sizeof(double), d.st));
    SPE_CUDA(cudaMalloc(&d.xflags, kMaxPeers * sizeof(unsigned long long)));
    SPE_CUDA(cudaMemsetAsync(d.xflags, 0, kMaxPeers * sizeof(unsigned long long), d.st));
  }
  SPE_CUDA(cudaStreamSynchronize(d.st));
  return SPEIGEN_OK;
}

void Solver::destroy() {
  finish_host_verify();
  for (Dev& d : devs) {
    cudaSetDevice(d.ordinal);
    if (d.st) cudaStreamSynchronize(d.st);
    for (int i = 0; i < d.n_ipc_opened; ++i) cudaIpcCloseMemHandle(d.ipc_opened[i]);
    if (d.comm && nccl_api()) nccl_api()->CommDestroy(d.comm);
    void* ptrs[] = {d.A, d.arena, d.xb, d.xdots, d.xflags, d.flush};
    for (void* p : ptrs)
      if (p) cudaFree(p);
    if (d.st) cudaStreamDestroy(d.st);
  }
  devs.clear();
  if (h_scal) cudaFreeHost(h_scal);
  h_scal = nullptr;
  if (h_res) cudaFreeHost(h_res);
  h_res = nullptr;

}

void Solver::reset_for_reuse() {
  finish_host_verify();
  loaded = prepared = inited = false;
  host_verified = upload_half = false;
  scale_max = 0.0;
  sv2.clear();
  contractions = 0;
  n_dot_parts = 0;
  bq = bq0;      // the initialiser may have shrunk its block on a rank-deficient input
  pdb = make_dims(m, bq, cplx);
  for (Dev& d : devs) {
    cudaSetDevice(d.ordinal);
    cudaMemsetAsync(d.status, 0, sizeof(int), d.st);
    d.Astd.ptr = nullptr;
  }
}

int Solver::init_comm_all() {
  if (prob.world == 1) { comm_ready = true; return SPEIGEN_OK; }
  if (static_cast<int>(devs.size()) != prob.world) {
    set_error("single-process communicator needs all %d shards local (have %zu)", prob.world, devs.size());
    return SPEIGEN_ERR_ARG;
  }
  // Row-sharded covariance inside one process: the fused peer-store exchange needs no NCCL communicator at all
  // (ncclCommInitAll costs hundreds of milliseconds, more than the whole solve at m = 50k on 8 GPUs).
  SPE_TRY(init_p2p());
  if (p2p && !data) { comm_ready = true; return SPEIGEN_OK; }
  const bool had_peers = peer_ok;
  const NcclApi* nc = nccl_api();
  if (!nc) return SPEIGEN_ERR_NCCL;
  std::vector<int> ords;
  for (Dev& d : devs) ords.push_back(d.ordinal);
  std::vector<ncclComm_t> comms(devs.size());
  SPE_NCCL(nc->CommInitAll(comms.data(), static_cast<int>(devs.size()), ords.data()));
  for (size_t i = 0; i < devs.size(); ++i) devs[i].comm = comms[i];
  comm_ready = true;
  (void)had_peers;
  return init_p2p();
}

int Solver::init_comm_rank(const void* id128) {
  if (prob.world == 1) { comm_ready = true; return SPEIGEN_OK; }
  const NcclApi* nc = nccl_api();
  if (!nc) return SPEIGEN_ERR_NCCL;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  SPE_NCCL(nc->GroupStart());
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_NCCL(nc->CommInitRank(&d.comm, prob.world, id, d.shard));
  }
  SPE_NCCL(nc->GroupEnd());
  comm_ready = true;
  return init_p2p();
}

// Map every shard's exchange buffers into every local device: direct peer access inside one process, CUDA IPC
// handles (allgathered over the NCCL communicator) across processes.  Any failure leaves the NCCL allgather path on.
int Solver::init_p2p() {
  p2p = false;
  peer_ok = false;
  if (prob.world <= 1 || data || prob.world > kMaxPeers) return SPEIGEN_OK;
  // the fused exchange needs every shard non-empty (an empty one would never raise its flag); the peer mappings alone
  // (symmetry scan across shards) do not
  const bool want_p2p = cfg.p2p_exchange && !(rows_per_shard * (prob.world - 1) >= m);
  const bool single_process = static_cast<int>(devs.size()) == prob.world;
  if (single_process) {
    for (Dev& a : devs) {
      if (cudaSetDevice(a.ordinal) != cudaSuccess) return SPEIGEN_OK;
      for (Dev& b : devs) {
        if (&a != &b) {
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, a.ordinal, b.ordinal) != cudaSuccess || !can) { cudaGetLastError(); return SPEIGEN_OK; }
          const cudaError_t e = cudaDeviceEnablePeerAccess(b.ordinal, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SPEIGEN_OK; }
          cudaGetLastError();
        }
        a.peer_xb[b.shard] = b.xb;
        a.peer_xdots[b.shard] = b.xdots;
        a.peer_xflags[b.shard] = b.xflags;
        a.peer_arena[b.shard] = static_cast<char*>(b.arena);
        a.peer_A[b.shard] = b.A;
      }
    }
    peer_ok = true;
    p2p = want_p2p;
    return SPEIGEN_OK;
  }
  // one process per GPU: exchange IPC handles through the communicator (3 handles of 64 bytes per rank)
  const NcclApi* nc = nccl_api();
  if (!nc || devs.size() != 1) return SPEIGEN_OK;
  Dev& d = devs[0];
  SPE_CUDA(cudaSetDevice(d.ordinal));
  constexpr int NH = 5;   // xb, xdots, xflags, arena (panels + mailbox), A (matrix shard: symmetry scan)
  cudaIpcMemHandle_t mine[NH];
  if (cudaIpcGetMemHandle(&mine[0], d.xb) != cudaSuccess || cudaIpcGetMemHandle(&mine[1], d.xdots) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine[2], d.xflags) != cudaSuccess || cudaIpcGetMemHandle(&mine[3], d.arena) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine[4], d.A) != cudaSuccess) { cudaGetLastError(); return SPEIGEN_OK; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  const size_t per = NH * 64 / sizeof(double);
  double* dev_h = nullptr;
  SPE_CUDA(cudaMalloc(&dev_h, per * prob.world * sizeof(double)));
  SPE_CUDA(cudaMemcpyAsync(dev_h + per * d.shard, mine, NH * 64, cudaMemcpyHostToDevice, d.st));
  SPE_NCCL(nc->AllGather(dev_h + per * d.shard, dev_h, per, NCCL_DOUBLE, d.comm, d.st));
  std::vector<cudaIpcMemHandle_t> all(NH * prob.world);
  SPE_CUDA(cudaMemcpyAsync(all.data(), dev_h, NH * 64 * prob.world, cudaMemcpyDeviceToHost, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  cudaFree(dev_h);
  bool ok = true;
  for (int r = 0; r < prob.world && ok; ++r) {
    if (r == d.shard) {
      d.peer_xb[r] = d.xb; d.peer_xdots[r] = d.xdots; d.peer_xflags[r] = d.xflags;
      d.peer_arena[r] = static_cast<char*>(d.arena); d.peer_A[r] = d.A;
      continue;
    }
    void* ptr[NH] = {};
    for (int k = 0; k < NH && ok; ++k) {
      if (cudaIpcOpenMemHandle(&ptr[k], all[NH * r + k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      d.ipc_opened[d.n_ipc_opened++] = ptr[k];
    }
    d.peer_xb[r] = static_cast<double*>(ptr[0]);
    d.peer_xdots[r] = static_cast<double*>(ptr[1]);
    d.peer_xflags[r] = static_cast<unsigned long long*>(ptr[2]);
    d.peer_arena[r] = static_cast<char*>(ptr[3]);
    d.peer_A[r] = static_cast<const double*>(ptr[4]);
  }
  // every rank must agree, otherwise some would wait for flags that never come
  double agree = ok ? 1.0 : 0.0;
  double* dflag = nullptr;
  SPE_CUDA(cudaMalloc(&dflag, sizeof(double)));
  SPE_CUDA(cudaMemcpyAsync(dflag, &agree, sizeof(double), cudaMemcpyHostToDevice, d.st));
  SPE_NCCL(nc->AllReduce(dflag, dflag, 1, NCCL_DOUBLE, NCCL_SUM, d.comm, d.st));
  SPE_CUDA(cudaMemcpyAsync(&agree, dflag, sizeof(double), cudaMemcpyDeviceToHost, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  cudaFree(dflag);
  peer_ok = (agree == static_cast<double>(prob.world));
  p2p = peer_ok && want_p2p;
  return SPEIGEN_OK;
}

int Solver::sync_all() {
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaStreamSynchronize(d.st));
  }
  return SPEIGEN_OK;
}

int Solver::read_scalars(int nd) {
  Dev& d = devs[0];
  SPE_CUDA(cudaSetDevice(d.ordinal));
  SPE_CUDA(cudaMemcpyAsync(h_scal, d.scal, nd * sizeof(double), cudaMemcpyDeviceToHost, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  return SPEIGEN_OK;
}

// ---- matrix upload / tensor maps ---------------------------------------------------------------
// The caller's matrix is ordinary pageable memory when the caller is R (or NumPy): a plain cudaMemcpyAsync from it is staged by
// the driver through a small internal bounce buffer at a fraction of the PCIe rate.  Pageable sources above a few MB are
// therefore staged here: worker threads copy 64 MB chunks into a ring of pinned slots while the previous chunks are in flight
// (one async DMA per chunk), and consecutive chunks go to different GPUs so that every GPU's PCIe link is busy at once.
// Process-wide pinned staging ring (allocated on first use, released by speigen_shutdown()): pinning memory costs ~0.3 ms per MB,
// more than a whole small solve, so it is not re-done per call.
struct StageRing {
  static constexpr size_t kSlotBytes = 8u << 20;
  static constexpr int kSlots = 48;
  static constexpr int kMaxDev = 16;
  void* base = nullptr;
  cudaEvent_t ev[kSlots][kMaxDev] = {};     // an event belongs to the device whose stream records it: one per (slot, device ordinal)
  int ensure() {
    if (base) return SPEIGEN_OK;
    SPE_CUDA(cudaHostAlloc(&base, kSlots * kSlotBytes, cudaHostAllocPortable));
    return SPEIGEN_OK;
  }
  // event of `slot` on device `ordinal` (current device must be `ordinal`); created on first use
  cudaError_t event(int slot, int ordinal, cudaEvent_t* out) {
    cudaEvent_t& e = ev[slot][ordinal % kMaxDev];
    if (!e) {
      const cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (rc != cudaSuccess) return rc;
    }
    *out = e;
    return cudaSuccess;
  }
  void release() {
    if (!base) return;
    for (auto& row : ev) for (auto& e : row) if (e) { cudaEventDestroy(e); e = nullptr; }
    cudaFreeHost(base);
    base = nullptr;
  }
};
static StageRing& stage_ring_global() { static StageRing r; return r; }
void release_stage_ring() { stage_ring_global().release(); }

namespace {
struct UploadChunk {
  int dev;                 // index into devs
  const char* src;         // first byte in the caller's matrix
  size_t spitch;           // bytes between consecutive rows in the source
  char* dst;               // device destination
  size_t dpitch;
  size_t width;            // bytes per row
  size_t rows;
};

}  // namespace

// Hermitian-half upload (host_pack.h): applies to a covariance matrix driven by one process whose GPUs can read each other's
// shards.  SPEIGEN_HALF_UPLOAD=0 disables it; SPEIGEN_HALF_BLOCK / SPEIGEN_HALF_MIN_BYTES override the super-tile edge and the
// size threshold (tests run it on small matrices that way).
static int64_t env_i64(const char* name, int64_t dflt) {
  const char* v = getenv(name);
  if (!v || !v[0]) return dflt;
  const long long x = atoll(v);
  return x > 0 ? static_cast<int64_t>(x) : dflt;
}

static std::atomic<long long> g_last_upload_bytes{0};
long long last_upload_bytes() { return g_last_upload_bytes.load(); }

struct Solver::HostVerify {
  std::vector<HalfUnit> units;
  std::vector<double> diff, sabs;
  std::atomic<size_t> next{0};
  std::vector<std::thread> threads;
  std::chrono::steady_clock::time_point t_start;
  std::atomic<long long> done_ns{0};          // latest worker finish, ns after t_start
};

void Solver::start_host_verify(const double* X, std::vector<HalfUnit>&& units) {
  finish_host_verify();
  HostVerify* hv = new HostVerify();
  hv->units = std::move(units);
  hv->diff.assign(hv->units.size(), 0.0);
  hv->sabs.assign(hv->units.size(), 0.0);
  const int hw = static_cast<int>(std::thread::hardware_concurrency());
  const int nthreads = std::max(1, std::min(24, hw - 1));       // one core stays with the thread that drives the GPUs
  const int e = 1 + cplx;
  const size_t lds = static_cast<size_t>(m) * e;
  const int64_t mm = m;
  hv->t_start = std::chrono::steady_clock::now();
  for (int w = 0; w < nthreads; ++w)
    hv->threads.emplace_back([hv, X, e, lds, mm]() {
      std::vector<double> scratch(2 * 128 * 128);
      for (;;) {
        const size_t i = hv->next.fetch_add(1);
        if (i >= hv->units.size()) break;
        const HalfUnit& u = hv->units[i];
        double acc[3] = {0.0, 0.0, 0.0};
        half_unit_pack_and_sums(X + (static_cast<size_t>(u.r0) * mm + u.c0) * e, lds, X + (static_cast<size_t>(u.c0) * mm + u.r0) * e, lds,
                                u.nr, u.nc, e, nullptr, true, scratch.data(), acc);
        half_unit_contribution(u, acc, &hv->diff[i], &hv->sabs[i]);
      }
      const long long ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - hv->t_start).count();
      long long seen = hv->done_ns.load();
      while (ns > seen && !hv->done_ns.compare_exchange_weak(seen, ns)) {}
    });
  verify = hv;
}

void Solver::finish_host_verify() {
  if (!verify) return;
  for (auto& t : verify->threads) t.join();
  host_verify_seconds = 1e-9 * static_cast<double>(verify->done_ns.load());
  host_diff = 0.0;
  host_sabs = 0.0;
  for (size_t i = 0; i < verify->units.size(); ++i) { host_diff += verify->diff[i]; host_sabs += verify->sabs[i]; }
  host_verified = true;
  delete verify;
  verify = nullptr;
}

// anyNA (R/spEigen.R:53) and isSymmetric.matrix (:75) on the host sums of the Hermitian-half upload
static int judge_host_sums(const speigen_cfg& cfg, double diff, double sabs) {
  if (cfg.check_na && (std::isnan(sabs) || std::isnan(diff))) {
    set_error("This function cannot handle NAs.");
    return SPEIGEN_ERR_NA;
  }
  if (cfg.check_symmetric) {
    const double tol = 100.0 * 2.220446049250313e-16;
    const bool sym = (sabs > 0.0) ? (diff / sabs <= tol) : (diff == 0.0);
    if (!sym) {
      set_error("The covariance matrix is not symmetric.");
      return SPEIGEN_ERR_NOT_SYMMETRIC;
    }
  }
  return SPEIGEN_OK;
}

int Solver::host_verdict() {
  if (!verify) return SPEIGEN_OK;
  finish_host_verify();
  return judge_host_sums(cfg, host_diff, host_sabs);
}

int Solver::load_host_half(const double* X, bool pageable) {
  const int e = 1 + cplx;
  const int64_t block = env_i64("SPEIGEN_HALF_BLOCK", kHalfBlock);
  const size_t slot_bytes = StageRing::kSlotBytes;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  for (Dev& d : devs) ranges.emplace_back(d.lo, d.hi);
  const std::vector<HalfUnit> units = plan_half_upload(m, e, ranges, block, slot_bytes);
  const bool want_compare = cfg.check_symmetric || cfg.check_na;
  // Deferred only for a staged (pageable) source, where the pack-and-copy pass is bound by host memory traffic and dropping the
  // partner reads shortens it (0.27 -> 0.24 s for 20 GB); a pinned source is DMA'd in place and its compare already overlaps the DMA.
  const bool defer = defer_host_verify && want_compare && pageable;
  const bool compare = want_compare && !defer;
  upload_staged = pageable;
  {
    long long bytes = 0;
    for (const HalfUnit& u : units) bytes += static_cast<long long>(u.nr) * u.nc * e * static_cast<long long>(sizeof(double));
    g_last_upload_bytes.store(bytes);
  }
  static std::mutex ring_mutex;
  std::unique_lock<std::mutex> ring_lock(ring_mutex, std::defer_lock);
  StageRing& ring = stage_ring_global();
  if (pageable) {
    ring_lock.lock();
    SPE_CUDA(cudaSetDevice(devs[0].ordinal));
    SPE_TRY(ring.ensure());
  }
  const int hw = static_cast<int>(std::thread::hardware_concurrency());
  const int nthreads = std::max(1, std::min<int>(StageRing::kSlots / 2, hw > 0 ? hw : 1));
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  std::vector<int> dev_ord(devs.size());
  std::vector<cudaStream_t> dev_st(devs.size());
  for (size_t i = 0; i < devs.size(); ++i) { dev_ord[i] = devs[i].ordinal; dev_st[i] = devs[i].st; }
  std::vector<double> unit_diff(units.size(), 0.0), unit_sabs(units.size(), 0.0);   // summed in unit order: deterministic
  const size_t spitch = static_cast<size_t>(m) * e * sizeof(double);
  auto worker = [&](int w) {
    int use = 0;
    const int my_slots = (StageRing::kSlots - w + nthreads - 1) / nthreads;
    std::vector<int> slot_dev(my_slots, -1);
    std::vector<double> scratch(2 * 128 * 128);
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= units.size() || failed.load()) break;
      const HalfUnit& u = units[i];
      Dev& d = devs[u.dev];
      const double* src = X + (static_cast<size_t>(u.r0) * m + u.c0) * e;           // A row r = column r of S (contiguous)
      const double* partner = X + (static_cast<size_t>(u.c0) * m + u.r0) * e;
      char* dst = reinterpret_cast<char*>(d.A + (u.r0 - d.lo) * d.a_ld + u.c0 * e);
      const size_t dpitch = static_cast<size_t>(d.a_ld) * sizeof(double);
      const size_t width = static_cast<size_t>(u.nc) * e * sizeof(double);
      const size_t lds = static_cast<size_t>(m) * e;
      double acc[3] = {0.0, 0.0, 0.0};
      cudaError_t err = cudaSuccess;
      if (pageable) {
        const int k = use % my_slots;
        const int slot = w + k * nthreads;
        ++use;
        if (slot_dev[k] >= 0) {
          cudaEvent_t pe = nullptr;
          if (cudaSetDevice(dev_ord[slot_dev[k]]) != cudaSuccess || ring.event(slot, dev_ord[slot_dev[k]], &pe) != cudaSuccess ||
              cudaEventSynchronize(pe) != cudaSuccess) { failed.store(1); break; }
        }
        char* stage = static_cast<char*>(ring.base) + static_cast<size_t>(slot) * slot_bytes;
        // one pass over the unit: every 128 x 128 sub-tile is streamed into the slot and compared with its partner while in cache
        half_unit_pack_and_sums(src, lds, partner, lds, u.nr, u.nc, e, reinterpret_cast<double*>(stage), compare, scratch.data(), acc);
        err = cudaSetDevice(dev_ord[u.dev]);
        if (err == cudaSuccess)
          err = cudaMemcpy2DAsync(dst, dpitch, stage, width, width, static_cast<size_t>(u.nr), cudaMemcpyHostToDevice, dev_st[u.dev]);
        cudaEvent_t me = nullptr;
        if (err == cudaSuccess) err = ring.event(slot, dev_ord[u.dev], &me);
        if (err == cudaSuccess) err = cudaEventRecord(me, dev_st[u.dev]);
        slot_dev[k] = u.dev;
      } else {
        err = cudaSetDevice(dev_ord[u.dev]);
        if (err == cudaSuccess)
          err = cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, static_cast<size_t>(u.nr), cudaMemcpyHostToDevice, dev_st[u.dev]);
        if (err == cudaSuccess && compare)      // the DMA reads the pinned source while this thread compares it with its partner
          half_unit_pack_and_sums(src, lds, partner, lds, u.nr, u.nc, e, nullptr, true, scratch.data(), acc);
      }
      if (err != cudaSuccess) { failed.store(1); break; }
      if (compare) half_unit_contribution(u, acc, &unit_diff[i], &unit_sabs[i]);
    }
  };
  {
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; ++w) th.emplace_back(worker, w);
    for (auto& t : th) t.join();
  }
  if (failed.load()) {
    cudaGetLastError();
    set_error("upload of the host matrix failed (CUDA error in a copy worker)");
    return SPEIGEN_ERR_CUDA;
  }
  SPE_TRY(sync_all());            // every shard's uploaded tiles are in place before any GPU reads a peer's
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    std::vector<const double*> shard_ptrs(prob.world, nullptr);
    for (int r = 0; r < prob.world; ++r) shard_ptrs[r] = (prob.world == 1) ? d.A : d.peer_A[r];
    SPE_TRY(launch_mirror_cov(d.A, d.a_rows, m, d.a_ld, d.lo, cplx, block, prob.world > 1 ? shard_ptrs.data() : nullptr, prob.world,
                              rows_per_shard, d.st));
  }
  SPE_TRY(sync_all());
  host_diff = 0.0;
  host_sabs = 0.0;
  for (size_t i = 0; i < units.size(); ++i) { host_diff += unit_diff[i]; host_sabs += unit_sabs[i]; }
  host_verified = compare;
  upload_half = true;
  loaded = true;
  if (defer) start_host_verify(X, std::vector<HalfUnit>(units));
  return SPEIGEN_OK;
}

int Solver::load_host(const double* X) {
  const int e = 1 + cplx;
  finish_host_verify();
  host_verified = false;
  upload_half = false;
  {
    const char* off = getenv("SPEIGEN_HALF_UPLOAD");
    const bool enabled = !(off && off[0] == '0');
    const bool single_process = static_cast<int>(devs.size()) == prob.world;
    const size_t bytes = static_cast<size_t>(m) * m * e * sizeof(double);
    const size_t min_bytes = static_cast<size_t>(env_i64("SPEIGEN_HALF_MIN_BYTES", 32ll << 20));
    if (enabled && !data && single_process && (prob.world == 1 || peer_ok) && prob.world <= 8 && bytes >= min_bytes) {
      cudaPointerAttributes pa{};
      bool pageable = true;
      if (cudaPointerGetAttributes(&pa, X) == cudaSuccess) pageable = (pa.type == cudaMemoryTypeUnregistered);
      cudaGetLastError();
      // A pinned source on several GPUs is the one case where the full upload wins: every GPU pulls its rows over its own PCIe
      // link (0.36 s / N for 20 GB) while the host compare of the half upload takes ~0.2 s whatever N (measured, 16 host cores).
      const char* force = getenv("SPEIGEN_HALF_UPLOAD");
      if (pageable || prob.world == 1 || (force && force[0] == '2')) return load_host_half(X, pageable && cfg_stage_uploads());
    }
  }
  // describe every device's shard as a 2-D copy (rows x width bytes)
  std::vector<UploadChunk> shards;
  for (size_t i = 0; i < devs.size(); ++i) {
    Dev& d = devs[i];
    if (d.hi <= d.lo) continue;
    UploadChunk c;
    c.dev = static_cast<int>(i);
    c.dst = reinterpret_cast<char*>(d.A);
    c.dpitch = static_cast<size_t>(d.a_ld) * sizeof(double);
    if (data) {
      // X is n x m column-major: row j of A = samples lo..hi of variable j
      c.src = reinterpret_cast<const char*>(X + d.lo * e);
      c.spitch = static_cast<size_t>(n) * e * sizeof(double);
      c.width = static_cast<size_t>(d.hi - d.lo) * e * sizeof(double);
      c.rows = static_cast<size_t>(m);
    } else {
      // S is m x m column-major and Hermitian: row j of A = column lo+j of S
      c.src = reinterpret_cast<const char*>(X + d.lo * m * e);
      c.spitch = static_cast<size_t>(m) * e * sizeof(double);
      c.width = static_cast<size_t>(m) * e * sizeof(double);
      c.rows = static_cast<size_t>(d.hi - d.lo);
    }
    shards.push_back(c);
  }
  size_t total_bytes = 0;
  for (auto& c : shards) total_bytes += c.rows * c.width;
  g_last_upload_bytes.store(static_cast<long long>(total_bytes));
  cudaPointerAttributes pa{};
  bool pageable = true;
  if (cudaPointerGetAttributes(&pa, X) == cudaSuccess) pageable = (pa.type == cudaMemoryTypeUnregistered);
  cudaGetLastError();
  upload_staged = pageable && total_bytes >= (32u << 20) && cfg_stage_uploads();
  if (!upload_staged) {
    for (auto& c : shards) {
      Dev& d = devs[c.dev];
      SPE_CUDA(cudaSetDevice(d.ordinal));
      if (c.spitch == c.width && c.dpitch == c.width)
        SPE_CUDA(cudaMemcpyAsync(c.dst, c.src, c.rows * c.width, cudaMemcpyHostToDevice, d.st));
      else
        SPE_CUDA(cudaMemcpy2DAsync(c.dst, c.dpitch, c.src, c.spitch, c.width, c.rows, cudaMemcpyHostToDevice, d.st));
    }
    SPE_TRY(sync_all());
    loaded = true;
    return SPEIGEN_OK;
  }
  // ---- staged upload: worker threads each take whole chunks (copy into a pinned slot, enqueue the DMA, move on)
  static std::mutex ring_mutex;                 // one staged upload at a time per process (the ring is process-wide)
  std::lock_guard<std::mutex> ring_lock(ring_mutex);
  StageRing& ring = stage_ring_global();
  SPE_CUDA(cudaSetDevice(devs[0].ordinal));
  SPE_TRY(ring.ensure());
  const size_t slot_bytes = StageRing::kSlotBytes;
  // cut the shards into chunks of whole rows (<= slot size) and interleave the devices
  std::vector<std::vector<UploadChunk>> per_dev(shards.size());
  for (size_t s = 0; s < shards.size(); ++s) {
    const UploadChunk& c = shards[s];
    if (c.width > slot_bytes) {
      // very long rows (m*e*8 > slot): split rows into column pieces
      for (size_t r = 0; r < c.rows; ++r)
        for (size_t off = 0; off < c.width; off += slot_bytes) {
          UploadChunk k = c;
          k.src = c.src + r * c.spitch + off; k.dst = c.dst + r * c.dpitch + off;
          k.width = std::min(slot_bytes, c.width - off); k.rows = 1; k.spitch = k.width; k.dpitch = k.width;
          per_dev[s].push_back(k);
        }
      continue;
    }
    const size_t rows_per = std::max<size_t>(1, slot_bytes / c.width);
    for (size_t r = 0; r < c.rows; r += rows_per) {
      UploadChunk k = c;
      k.src = c.src + r * c.spitch; k.dst = c.dst + r * c.dpitch; k.rows = std::min(rows_per, c.rows - r);
      per_dev[s].push_back(k);
    }
  }
  std::vector<UploadChunk> order;
  for (size_t i = 0;; ++i) {
    bool any = false;
    for (auto& v : per_dev) if (i < v.size()) { order.push_back(v[i]); any = true; }
    if (!any) break;
  }
  const int nthreads = std::max(1, std::min<int>(StageRing::kSlots / 2, static_cast<int>(std::thread::hardware_concurrency())));
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  std::vector<int> dev_ord(devs.size());
  std::vector<cudaStream_t> dev_st(devs.size());
  for (size_t i = 0; i < devs.size(); ++i) { dev_ord[i] = devs[i].ordinal; dev_st[i] = devs[i].st; }
  auto worker = [&](int w) {
    // worker w owns the slots w, w + nthreads, ... : no contention on a slot, its DMA is awaited by the thread that reuses it
    int use = 0;
    const int my_slots = (StageRing::kSlots - w + nthreads - 1) / nthreads;
    std::vector<int> slot_dev(my_slots, -1);
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= order.size() || failed.load()) break;
      const UploadChunk& c = order[i];
      const int k = use % my_slots;
      const int slot = w + k * nthreads;
      ++use;
      if (slot_dev[k] >= 0) {       // the DMA that last read this slot (recorded on that device's stream) must be done
        cudaEvent_t pe = nullptr;
        if (cudaSetDevice(dev_ord[slot_dev[k]]) != cudaSuccess || ring.event(slot, dev_ord[slot_dev[k]], &pe) != cudaSuccess ||
            cudaEventSynchronize(pe) != cudaSuccess) { failed.store(1); break; }
      }
      char* stage = static_cast<char*>(ring.base) + static_cast<size_t>(slot) * slot_bytes;
      if (c.spitch == c.width) std::memcpy(stage, c.src, c.rows * c.width);
      else for (size_t r = 0; r < c.rows; ++r) std::memcpy(stage + r * c.width, c.src + r * c.spitch, c.width);
      cudaError_t e = cudaSetDevice(dev_ord[c.dev]);
      if (e == cudaSuccess) {
        if (c.dpitch == c.width || c.rows == 1) e = cudaMemcpyAsync(c.dst, stage, c.rows * c.width, cudaMemcpyHostToDevice, dev_st[c.dev]);
        else e = cudaMemcpy2DAsync(c.dst, c.dpitch, stage, c.width, c.width, c.rows, cudaMemcpyHostToDevice, dev_st[c.dev]);
      }
      cudaEvent_t me = nullptr;
      if (e == cudaSuccess) e = ring.event(slot, dev_ord[c.dev], &me);
      if (e == cudaSuccess) e = cudaEventRecord(me, dev_st[c.dev]);
      if (e != cudaSuccess) { failed.store(1); break; }
      slot_dev[k] = c.dev;
    }
  };
  {
    std::vector<std::thread> th;
    for (int w = 0; w < nthreads; ++w) th.emplace_back(worker, w);
    for (auto& t : th) t.join();
  }
  if (failed.load()) {
    cudaGetLastError();
    set_error("staged upload of the host matrix failed (CUDA error in a copy worker)");
    return SPEIGEN_ERR_CUDA;
  }
  SPE_TRY(sync_all());
  loaded = true;
  return SPEIGEN_OK;
}

int Solver::build_tensor_maps() {
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_TRY(make_stream_mat(&d.Astd, d.A, d.a_rows, d.a_kd, d.a_ld, kTJ));
    if (data) {
      SPE_TRY(make_stream_mat(&d.Atr, d.A, d.a_rows, d.a_kd, d.a_ld, 64));
      d.has_tr = true;
    }
  }
  return SPEIGEN_OK;
}

// ---- validation scan, centring, rho scale -------------------------------------------------------
int Solver::prepare() {
  const bool need_nccl = prob.world > 1 && (data || !p2p || devs[0].comm != nullptr);
  const NcclApi* nc = need_nccl ? nccl_api() : nullptr;
  if (prob.world > 1 && ((need_nccl && !nc) || !comm_ready)) {
    set_error("multi-GPU solver used before the communicator was initialised");
    return SPEIGEN_ERR_NCCL;
  }
  SPE_TRY(build_tensor_maps());
  if (!data) {
    // isSymmetric.matrix needs S[i,j] next to S[j,i]: the transposed tile is read from the shard that owns it - the own one on
    // one GPU, a peer's over NVLink otherwise (peer access inside one process, CUDA IPC mappings across processes).  Without
    // peer mappings the check cannot be honoured: refuse rather than silently contract S^T (check_symmetric = 0 trusts the caller).
    const bool single_process = static_cast<int>(devs.size()) == prob.world;
    const bool can_check = prob.world == 1 || peer_ok;
    if (cfg.check_symmetric && !can_check) {
      set_error("isSymmetric(S) cannot be verified without peer access between the GPUs (p2p exchange unavailable); "
                "pass check_symmetric = 0 to trust the caller");
      return SPEIGEN_ERR_UNSUPPORTED;
    }
    // host_verified: load_host_half() already evaluated isSymmetric's sums on the host against the tiles that were not uploaded
    // (the device copy is Hermitian by construction); the device scan then only has to deliver NaN count, sum|S| and the diagonal.
    if (verify && !defer_host_verify) finish_host_verify();
    const bool host_side = host_verified || verify != nullptr;      // pending: the verdict is taken by host_verdict() later
    const int full = (!host_side && can_check && (cfg.check_symmetric || cfg.check_na)) ? 1 : 0;
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      std::vector<const double*> shard_ptrs(prob.world, nullptr);
      for (int r = 0; r < prob.world; ++r) shard_ptrs[r] = (prob.world == 1) ? d.A : d.peer_A[r];
      SPE_TRY(launch_scan_cov(d.A, d.a_rows, m, d.a_ld, d.lo, cplx, full, prob.world > 1 ? shard_ptrs.data() : nullptr,
                              prob.world, rows_per_shard, d.scan_part, d.ps.ticket, d.scal, d.st));
    }
    if (prob.world > 1 && devs[0].comm) {
      SPE_NCCL(nc->GroupStart());
      for (Dev& d : devs) {
        if (!single_process) {
          SPE_NCCL(nc->AllReduce(d.scal + 0, d.scal + 0, 2, NCCL_DOUBLE, NCCL_SUM, d.comm, d.st));
        }
        SPE_NCCL(nc->AllReduce(d.scal + 2, d.scal + 2, 1, NCCL_DOUBLE, NCCL_MAX, d.comm, d.st));
        SPE_NCCL(nc->AllReduce(d.scal + 3, d.scal + 3, 1, NCCL_DOUBLE, NCCL_SUM, d.comm, d.st));
        SPE_NCCL(nc->AllReduce(d.scal + 4, d.scal + 4, 1, NCCL_DOUBLE, NCCL_MIN, d.comm, d.st));
      }
      SPE_NCCL(nc->GroupEnd());
    }
    SPE_TRY(read_scalars(5));
    if (prob.world > 1 && single_process) {   // combine the shards' sums on the host (fixed device order)
      double df = h_scal[0], sa = h_scal[1], mx = h_scal[2], nn = h_scal[3], mn = h_scal[4];
      for (size_t i = 1; i < devs.size(); ++i) {
        double tmp[5];
        SPE_CUDA(cudaSetDevice(devs[i].ordinal));
        SPE_CUDA(cudaMemcpyAsync(tmp, devs[i].scal, sizeof(tmp), cudaMemcpyDeviceToHost, devs[i].st));
        SPE_CUDA(cudaStreamSynchronize(devs[i].st));
        df += tmp[0];
        sa += tmp[1];
        if (!devs[0].comm) { mx = std::max(mx, tmp[2]); nn += tmp[3]; mn = std::min(mn, tmp[4]); }
      }
      h_scal[0] = df;
      h_scal[1] = sa;
      h_scal[2] = mx;
      h_scal[3] = nn;
      h_scal[4] = mn;
    }
    if (host_verified) { h_scal[0] = host_diff; h_scal[1] = host_sabs; }   // NaN in a tile that stayed on the host shows up here
    if (cfg.check_na && (h_scal[3] > 0.0 || std::isnan(h_scal[1]))) {
      set_error("This function cannot handle NAs.");
      return SPEIGEN_ERR_NA;
    }
    if ((full || host_verified) && cfg.check_symmetric) {
      // isSymmetric.matrix: all.equal(X, Conj(t(X)), tolerance = 100 * .Machine$double.eps)   R/spEigen.R:75
      const double tol = 100.0 * 2.220446049250313e-16;
      const double diff = h_scal[0], sabs = h_scal[1];
      const bool sym = (sabs > 0.0) ? (diff / sabs <= tol) : (diff == 0.0);
      if (!sym) {
        set_error("The covariance matrix is not symmetric.");
        return SPEIGEN_ERR_NOT_SYMMETRIC;
      }
    }
    scale_max = h_scal[2];   // max(Re(diag(X)))   R/spEigen.R:82
    // necessary condition of R/spEigen.R:77 that costs nothing: x^T S x >= lambda_min for x = e_i, so a diagonal entry below
    // pd_tol proves an eigenvalue below pd_tol (the Ritz-value probe in init_vectors covers what the diagonal cannot show)
    if (h_scal[4] < cfg.pd_tol) {
      set_error("The covariance matrix is not PD.");
      return SPEIGEN_ERR_NOT_PD;
    }
  } else {
    const int e = 1 + cplx;
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_TRY(launch_row_sums(d.A, d.a_rows, d.hi - d.lo, d.a_ld, cplx, d.vecm, d.st));
    }
    if (prob.world > 1) {
      SPE_NCCL(nc->GroupStart());
      for (Dev& d : devs) SPE_NCCL(nc->AllReduce(d.vecm, d.vecm, static_cast<size_t>(m) * e, NCCL_DOUBLE, NCCL_SUM, d.comm, d.st));
      SPE_NCCL(nc->GroupEnd());
    }
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      // scale(X, center = TRUE, scale = FALSE)   R/spEigen.R:60
      SPE_TRY(launch_center_sqnorm(d.A, d.a_rows, d.hi - d.lo, d.a_ld, cplx, d.vecm, 1.0 / static_cast<double>(n), d.vecm2, d.st));
    }
    if (prob.world > 1) {
      SPE_NCCL(nc->GroupStart());
      for (Dev& d : devs) SPE_NCCL(nc->AllReduce(d.vecm2, d.vecm2, static_cast<size_t>(m), NCCL_DOUBLE, NCCL_SUM, d.comm, d.st));
      SPE_NCCL(nc->GroupEnd());
    }
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_TRY(launch_vec_max(d.vecm2, m, d.scal, d.st));
    }
    SPE_TRY(read_scalars(2));
    if (cfg.check_na && (h_scal[1] > 0.0 || std::isnan(h_scal[0]))) {
      set_error("This function cannot handle NAs.");
      return SPEIGEN_ERR_NA;
    }
    scale_max = h_scal[0];   // max(colSums(abs(X)^2))   R/spEigen.R:72
  }
  SPE_TRY(sync_all());
  prepared = true;
  return SPEIGEN_OK;
}

// ---- panel transfer -----------------------------------------------------------------------------
int Solver::op_upload_panel(const double* host_cm, Buf b, const PanelDims& pd) {
  const size_t bytes = static_cast<size_t>(pd.m) * pd.ncr * sizeof(double);
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemcpyAsync(d.out_cm, host_cm, bytes, cudaMemcpyHostToDevice, d.st));
    SPE_CUDA(cudaMemsetAsync(d.buf[b], 0, static_cast<size_t>(mrows) * pd.pitch * sizeof(double), d.st));
    SPE_TRY(launch_colmajor_to_panel(pd, d.out_cm, d.buf[b], d.st));
  }
  return sync_all();
}

int Solver::op_download_panel(Buf b, const PanelDims& pd, double* host_cm) {
  Dev& d = devs[0];
  SPE_CUDA(cudaSetDevice(d.ordinal));
  SPE_TRY(launch_panel_to_colmajor(pd, d.buf[b], d.out_cm, d.st));
  SPE_CUDA(cudaMemcpyAsync(host_cm, d.out_cm, static_cast<size_t>(pd.m) * pd.ncr * sizeof(double), cudaMemcpyDeviceToHost, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  return SPEIGEN_OK;
}

int Solver::set_qvecs(const double* dv, const double* rv) {
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemcpyAsync(d.dvec, dv, q * sizeof(double), cudaMemcpyHostToDevice, d.st));
    SPE_CUDA(cudaMemcpyAsync(d.rho, rv, q * sizeof(double), cudaMemcpyHostToDevice, d.st));
  }
  return sync_all();
}

// ---- exchange step -------------------------------------------------------------------------------
int Solver::exchange_allgather(Buf b, const PanelDims& pd) {
  const NcclApi* nc = nccl_api();
  if (!nc) return SPEIGEN_ERR_NCCL;
  const size_t count = static_cast<size_t>(rows_per_shard) * pd.pitch;
  SPE_NCCL(nc->GroupStart());
  for (Dev& d : devs) SPE_NCCL(nc->AllGather(d.buf[b] + d.shard * count, d.buf[b], count, NCCL_DOUBLE, d.comm, d.st));
  SPE_NCCL(nc->GroupEnd());
  return SPEIGEN_OK;
}

// ---- the contraction (K1 + exchange) -----------------------------------------------------------
// outY (raw S*V or X^H X V), outB (fused reweighting), dots -> d.dots_part / n_dot_parts
int Solver::op_contract(Buf in, const PanelDims& pd, Buf outY, Buf outB, Buf vepi, bool want_dots, const StageConsts& sc) {
  const bool multi = prob.world > 1;
  const bool want_y = outY != NBUF, want_b = outB != NBUF;
  const NcclApi* nc = multi ? nccl_api() : nullptr;
  if (multi && !nc) return SPEIGEN_ERR_NCCL;
  const bool post_unfused = data && multi;   // partial sums need the exchange before any epilogue math
  const bool use_p2p = p2p && multi && !data;
  const int parity = static_cast<int>(epoch & 1);
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    const double* operand = d.buf[in];
    if (!data && cplx) {
      SPE_TRY(launch_hat(pd, d.buf[in], d.hat, /*conj_form=*/1, d.st));
      operand = d.hat;
    }
    if (data) {
      // T = X * V (local samples), then the standard form contracts over those samples
      SPE_TRY(launch_contract_tr(d.Atr, d.buf[in], pd.ncr, d.Tdata, pd.pitch, d.ws, d.num_sms, d.st));
      if (cplx) SPE_TRY(launch_tr_combine(d.hi - d.lo, pd.q, d.Tdata, d.Tdata, d.st));
      operand = d.Tdata;
    }
    EpiArgs ep;
    ep.q = pd.q;
    ep.cplx = cplx;
    ep.pitch = pd.pitch;
    ep.sc = sc;
    ep.row_off = data ? 0 : d.lo;
    if (post_unfused) {
      ep.Y = d.gather + static_cast<size_t>(d.shard) * mrows * pd.pitch;
    } else {
      ep.Y = want_y ? d.buf[outY] : nullptr;
      if (want_b) { ep.B = d.buf[outB]; ep.d = d.dvec; ep.rho = d.rho; ep.wmax = d.wmax; }
      if (want_b || want_dots) ep.V = d.buf[vepi];
      if (want_dots) ep.dots = d.dots_part;
    }
    if (use_p2p) {
      // epilogue stores go to the exchange slots of every GPU (own one included); local Y/B/dots are filled by the waiter
      const size_t slot = static_cast<size_t>(mrows) * pdb.pitch;
      double* const loc_y = ep.Y;
      double* const loc_b = ep.B;
      ep.npeers = prob.world;
      ep.my_shard = d.shard;
      for (int r = 0; r < prob.world; ++r) {
        double* base = d.peer_xb[r] + (parity * 2) * slot;
        ep.peerY[r] = loc_y ? base : nullptr;
        ep.peerB[r] = loc_b ? base + slot : nullptr;
        ep.peer_dots[r] = d.peer_xdots[r] + static_cast<size_t>(parity) * prob.world * kMaxCols;
        ep.peer_flag[r] = d.peer_xflags[r];
      }
      ep.epoch = epoch + 1;
      ep.done_count = d.done_count;
    }
    const bool tme = ktiming && (&d == &devs[0]) && kev_used + 2 <= kev.size();
    if (tme) SPE_CUDA(cudaEventRecord(kev[kev_used], d.st));
    SPE_TRY(launch_contract_std(d.Astd, operand, pd.ncr, ep, d.ws, d.num_sms, d.st));
    if (tme) { SPE_CUDA(cudaEventRecord(kev[kev_used + 1], d.st)); kev_used += 2; }
  }
  contractions++;
  const int nout = pd.q;
  if (!multi) {
    n_dot_parts = static_cast<int>((devs[0].a_rows + kTJ - 1) / kTJ);
    return SPEIGEN_OK;
  }
  if (use_p2p) {
    ++epoch;
    const size_t slot = static_cast<size_t>(mrows) * pdb.pitch;
    const size_t nd = static_cast<size_t>(padded_rows(m)) * pd.pitch;   // doubles actually used by this panel shape
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      const double* sy = want_y ? d.xb + (parity * 2) * slot : nullptr;
      const double* sb = want_b ? d.xb + (parity * 2 + 1) * slot : nullptr;
      SPE_TRY(launch_exchange_wait_copy(d.xflags, prob.world, epoch, sy, want_y ? d.buf[outY] : nullptr, sb,
                                        want_b ? d.buf[outB] : nullptr, std::max(nd, static_cast<size_t>(rows_per_shard) * prob.world * pd.pitch),
                                        want_dots ? d.xdots + static_cast<size_t>(parity) * prob.world * kMaxCols : nullptr,
                                        d.dots_part, prob.world * nout, d.tstatus,
                                        static_cast<long long>(cfg.exchange_timeout_ms > 0 ? cfg.exchange_timeout_ms : 10000) * 1000000ll, d.st));
    }
    if (want_dots) n_dot_parts = prob.world;
    return SPEIGEN_OK;
  }
  if (!data) {
    // row-sharded covariance: allgather the finished row blocks (and the per-shard dot sums)
    if (want_y) SPE_TRY(exchange_allgather(outY, pd));
    if (want_b) SPE_TRY(exchange_allgather(outB, pd));
    if (want_dots) {
      for (Dev& d : devs) {
        SPE_CUDA(cudaSetDevice(d.ordinal));
        const int nt = static_cast<int>((d.a_rows + kTJ - 1) / kTJ);
        SPE_TRY(launch_sum_parts(d.dots_part, nt, nout, d.dots_gather + d.shard * nout, d.st));
      }
      SPE_NCCL(nc->GroupStart());
      for (Dev& d : devs)
        SPE_NCCL(nc->AllGather(d.dots_gather + d.shard * nout, d.dots_gather, nout, NCCL_DOUBLE, d.comm, d.st));
      SPE_NCCL(nc->GroupEnd());
      for (Dev& d : devs) {
        SPE_CUDA(cudaSetDevice(d.ordinal));
        SPE_CUDA(cudaMemcpyAsync(d.dots_part, d.dots_gather, static_cast<size_t>(prob.world) * nout * sizeof(double),
                                 cudaMemcpyDeviceToDevice, d.st));
      }
      n_dot_parts = prob.world;
    }
    return SPEIGEN_OK;
  }
  // sample-sharded data matrix: allgather the partial panels, sum them in rank order (bit-stable allreduce)
  const size_t pcount = static_cast<size_t>(mrows) * pd.pitch;
  SPE_NCCL(nc->GroupStart());
  for (Dev& d : devs) SPE_NCCL(nc->AllGather(d.gather + d.shard * pcount, d.gather, pcount, NCCL_DOUBLE, d.comm, d.st));
  SPE_NCCL(nc->GroupEnd());
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    double* ysum = want_y ? d.buf[outY] : d.buf[BT2];
    SPE_TRY(launch_sum_panels(pd, d.gather, prob.world, pcount, ysum, d.st));
    if (want_b)
      SPE_TRY(launch_reweight(pd, ysum, d.buf[vepi], d.dvec, d.rho, d.wmax, sc, d.buf[outB], nullptr, d.st));
    if (want_dots) SPE_TRY(launch_coldots(pd, d.buf[vepi], ysum, d.dots_part, d.st));
  }
  if (want_dots) n_dot_parts = pd.nblk;
  return SPEIGEN_OK;
}

// ---- polar factor (K3 + applies):  out = in (in^H in)^{-1/2}, two Gram passes ------
int Solver::op_polar(Buf in, Buf out, const PanelDims& pd, bool /*have_gram*/, double* /*unused*/, int /*which_absmin*/, int warm_slot) {
  // the fused tail kernel (tail.cu): Gram passes + q x q solve + applies in one launch, on the own rows when the panels are
  // row-sharded; the finished rows go to every rank, and the stream waits for the peers' rows (the callers read whole panels)
  TailSpec ts;
  ts.producer = TAIL_PANEL; ts.T = in; ts.out = out; ts.out_all = true;
  ts.site = (warm_slot >= 0) ? SITE_MISC : -1;
  SPE_TRY(op_tail(ts, pd, make_stage(0.1, 0.1)));
  return wait_latest_tail();
}

// ---- round 2: fused, row-sharded tail + MM-loop contraction ----------------------------------------
// Row ownership of the m x q panels during the MM loop: sharded() -> every GPU computes the rows of S it owns (and the tail
// kernels exchange q x q partials / statistics through the mailboxes and store finished operand rows into every peer);
// otherwise (data-matrix mode, or no peer access) every GPU runs the tail redundantly on all m rows with no exchange.
int Solver::op_tail(const TailSpec& ts, const PanelDims& pd, const StageConsts& sc) {
  const bool sh = sharded();
  const int nr = sh ? prob.world : 1;
  const int wpar = ts.site >= 0 ? warm_parity[ts.site] : 0;
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    TailArgs a;
    a.row0 = sh ? d.lo : 0;
    a.row1 = sh ? d.hi : m;
    a.q = pd.q; a.cplx = pd.cplx; a.ncr = pd.ncr; a.pitch = pd.pitch;
    a.nranks = nr;
    a.my_rank = sh ? d.shard : 0;
    a.producer = ts.producer;
    a.in0 = ts.in0 >= 0 ? d.buf[ts.in0] : nullptr;
    a.in1 = ts.in1 >= 0 ? d.buf[ts.in1] : nullptr;
    a.in2 = ts.in2 >= 0 ? d.buf[ts.in2] : nullptr;
    a.T = d.buf[ts.T];
    a.d = d.dvec; a.rho = d.rho; a.sc = sc;
    if (ts.use_prev) {
      a.prev_flags = d.mail->stat_flag;
      a.prev_epoch = stat_epoch;
      a.prev_stat = &d.mail->stat[stat_epoch & 1ull][0][0];
    }
    a.a_cur = d.a_cur;
    a.retry = ts.retry;
    const bool all = ts.out_all && sh && nr > 1;
    for (int r = 0; r < nr; ++r) {
      char* base = sh ? d.peer_arena[r] : static_cast<char*>(d.arena);
      a.out_peer[r] = reinterpret_cast<double*>(base + d.off_buf[ts.out]);
      a.mail[r] = reinterpret_cast<Mailbox*>(base + d.off_mail);
    }
    a.n_out = all ? nr : 1;
    a.skip_polar = ts.skip_polar ? 1 : 0;
    a.want_absmin = ts.absmin; a.want_gsum = ts.gsum; a.want_squarem = ts.squarem;
    a.sq_V = ts.sqV >= 0 ? d.buf[ts.sqV] : nullptr;
    a.sq_V1 = ts.sqV1 >= 0 ? d.buf[ts.sqV1] : nullptr;
    a.gram_epoch = d.gram_epoch;
    a.stat_epoch = stat_epoch + 1;
    a.cta_part = d.cta_part; a.cta_stat = d.cta_stat; a.ticket = d.tail_ticket;
    if (ts.site >= 0) {
      double* wbase = d.twarm + static_cast<size_t>(ts.site) * 2 * kWarmDoubles;
      a.warm_in = wbase + static_cast<size_t>(wpar) * kWarmDoubles;
      a.warm_out = wbase + static_cast<size_t>(wpar ^ 1) * kWarmDoubles;
    }
    a.status = d.tstatus;
    a.diag = d.tdiag;
    a.spin_budget_ns = static_cast<long long>(cfg.exchange_timeout_ms > 0 ? cfg.exchange_timeout_ms : 10000) * 1000000ll;
    SPE_TRY(launch_tail(a, d.num_sms, d.st));
  }
  ++stat_epoch;
  if (!ts.skip_polar && ts.site >= 0) warm_parity[ts.site] ^= 1;
  return SPEIGEN_OK;
}

int Solver::op_contract_loop(Buf in, const PanelDims& pd, Buf outY, Buf outB, Buf vepi, bool want_dots, const StageConsts& sc) {
  if (!sharded()) {
    // replicated tails: the contraction gathers full panels as before; the column minima come from the own mailbox
    for (Dev& d : devs) d.wmax = &d.mail->stat[stat_epoch & 1ull][0][kStatAbsmin];
    return op_contract(in, pd, outY, outB, vepi, want_dots, sc);
  }
  const bool multi = prob.world > 1;
  const bool want_y = outY != NBUF, want_b = outB != NBUF;
  const int parity = static_cast<int>(epoch & 1);
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    const double* operand = d.buf[in];
    const long long budget = static_cast<long long>(cfg.exchange_timeout_ms > 0 ? cfg.exchange_timeout_ms : 10000) * 1000000ll;
    if (cplx) {
      // the complex operand expansion reads the whole panel: wait for the peers' rows first (tiny kernel on the same stream)
      if (multi) SPE_TRY(launch_wait_flags(d.mail->stat_flag, kFlagStride, prob.world, stat_epoch, d.tstatus, budget, d.st));
      SPE_TRY(launch_hat(pd, d.buf[in], d.hat, /*conj_form=*/1, d.st));
      operand = d.hat;
    }
    EpiArgs ep;
    ep.q = pd.q; ep.cplx = cplx; ep.pitch = pd.pitch; ep.sc = sc;
    ep.row_off = d.lo;
    ep.Y = want_y ? d.buf[outY] : nullptr;
    if (want_b) { ep.B = d.buf[outB]; ep.d = d.dvec; ep.rho = d.rho; }
    if (want_b || want_dots) ep.V = d.buf[vepi];
    if (want_dots) ep.dots = d.dots_part;
    ep.wait_flags = d.mail->stat_flag;
    ep.wait_n = prob.world;
    ep.wait_epoch = stat_epoch;
    ep.stat_absmin = &d.mail->stat[stat_epoch & 1ull][0][kStatAbsmin];
    ep.stat_n = prob.world;
    ep.stat_stride = kStatLen;
    ep.status = d.tstatus;
    ep.spin_budget_ns = budget;
    if (multi && want_dots) {
      ep.npeers = prob.world;
      ep.my_shard = d.shard;
      for (int r = 0; r < prob.world; ++r) {
        ep.peer_dots[r] = d.peer_xdots[r] + static_cast<size_t>(parity) * prob.world * kMaxCols;
        ep.peer_flag[r] = d.peer_xflags[r];
      }
      ep.epoch = epoch + 1;
      ep.done_count = d.done_count;
    }
    const bool tme = ktiming && (&d == &devs[0]) && kev_used + 2 <= kev.size();
    if (tme) SPE_CUDA(cudaEventRecord(kev[kev_used], d.st));
    SPE_TRY(launch_contract_std(d.Astd, operand, pd.ncr, ep, d.ws, d.num_sms, d.st));
    if (tme) { SPE_CUDA(cudaEventRecord(kev[kev_used + 1], d.st)); kev_used += 2; }
  }
  contractions++;
  if (multi && want_dots) { ++epoch; dots_parity = parity; }
  n_dot_parts = multi ? prob.world : static_cast<int>((devs[0].a_rows + kTJ - 1) / kTJ);
  return SPEIGEN_OK;
}

// F = sum_c d_c quad_c - sum_c rho_c gsum_c (R/spEigen.R:136-138) from the last contraction's dot partials and the last
// tail's penalty sums; lands in h_res (mapped pinned memory) - the caller synchronises device 0's stream and reads it.
int Solver::op_finish_objective(const PanelDims& pd) {
  const bool sh = sharded();
  const bool multi_sh = sh && prob.world > 1;
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    FinishArgs f;
    f.q = pd.q;
    f.nranks = sh ? prob.world : 1;
    if (multi_sh) {
      f.dots = d.xdots + static_cast<size_t>(dots_parity) * prob.world * kMaxCols;
      f.n_dot_parts = prob.world;
      f.dots_flags = d.xflags;
      f.dots_epoch = epoch;
    } else {
      f.dots = d.dots_part;
      f.n_dot_parts = n_dot_parts;
    }
    f.stat_flags = d.mail->stat_flag;
    f.stat_epoch = stat_epoch;
    f.stat = &d.mail->stat[stat_epoch & 1ull][0][0];
    f.d = d.dvec; f.rho = d.rho;
    f.diag = d.tdiag;
    f.status = d.tstatus;
    f.out = d.scal + 8;
    f.out_host = (&d == &devs[0]) ? d.res_host_dev : nullptr;
    f.spin_budget_ns = static_cast<long long>(cfg.exchange_timeout_ms > 0 ? cfg.exchange_timeout_ms : 10000) * 1000000ll;
    SPE_TRY(launch_finish_objective(f, d.st));
  }
  return SPEIGEN_OK;
}

int Solver::check_tail_status(double sv) {
  const int st = static_cast<int>(sv);
  if (st == 0) return SPEIGEN_OK;
  for (Dev& d : devs) {          // the flag is sticky on the device: clear it so that the solver stays usable
    cudaSetDevice(d.ordinal);
    cudaMemsetAsync(d.tstatus, 0, sizeof(int), d.st);
  }
  if (st & 4) {
    set_error("a GPU waited more than %d ms for a peer in the fused exchange (a rank failed or left early)", cfg.exchange_timeout_ms);
    return SPEIGEN_ERR_TIMEOUT;
  }
  set_error("a Procrustes target (G - H of R/spEigen.R:180) is %s: its polar factor is not defined",
            (st & 1) ? "non-finite or zero" : "rank deficient beyond double precision");
  return SPEIGEN_ERR_NUMERIC;
}

// ---- initialiser: block subspace iteration with Rayleigh-Ritz (replaces eigen()/svd()) ----------
int Solver::init_vectors(const double* V0_host, speigen_out* out) {
  if (!prepared) SPE_TRY(prepare());
  const double t0 = now_s();
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemsetAsync(d.warm, 0, 4 * kWarmStride * sizeof(double), d.st));
    SPE_CUDA(cudaMemsetAsync(d.twarm, 0, static_cast<size_t>(N_TAIL_SITES) * 2 * kWarmDoubles * sizeof(double), d.st));
    SPE_CUDA(cudaMemsetAsync(d.tstatus, 0, sizeof(int), d.st));
  }
  const StageConsts sc0 = make_stage(0.1, 0.1);
  int passes = 0;
  bool converged = false;
  double resid_rel = 0.0;
  for (int attempt = 0; attempt < 2 && !converged; ++attempt) {
    if (attempt == 1) {   // rank-deficient block: retry with exactly q columns
      bq = q;
      pdb = make_dims(m, bq, cplx);
    }
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_CUDA(cudaMemsetAsync(d.status, 0, sizeof(int), d.st));
      SPE_CUDA(cudaMemsetAsync(d.buf[BQ], 0, static_cast<size_t>(mrows) * pdb.pitch * sizeof(double), d.st));
      SPE_TRY(launch_random_panel(pdb, cfg.init_seed, d.buf[BQ], d.st));
    }
    SPE_TRY(op_polar(BQ, BQ, pdb, false, nullptr, 0));
    double prev = 1e300;
    int stall = 0;
    bool rank_bad = false;
    int cheb_deg = 1;
    double lam_low = 0.0;
    sv2.assign(bq, 0.0);
    // Rayleigh-Ritz (cross-Gram, q x q eigen-solve, two applies, residual norms, one host read-back) is only needed to TEST
    // convergence and to steer the filter; a plain pass is Q <- polar(S Q) - one contraction and one tail kernel, no read-back.
    // With a measured geometric rate the next test is scheduled for the pass that is predicted to reach init_tol.
    int rr_skip = 0, last_rr_pass = 0;
    double last_rr_resid = 0.0;
    for (passes = 1; passes <= cfg.init_max_passes; ++passes) {
      SPE_TRY(op_contract(BQ, pdb, BZ, NBUF, NBUF, false, sc0));
      if (rr_skip > 0 && passes < cfg.init_max_passes) {
        --rr_skip;
        SPE_TRY(op_polar(BZ, BQ, pdb, false, nullptr, 0, WARM_INIT));
        continue;
      }
      for (Dev& d : devs) {
        SPE_CUDA(cudaSetDevice(d.ordinal));
        SPE_TRY(launch_gram(pdb, d.buf[BQ], d.buf[BZ], d.ps.gram_part, d.st));
        SPE_TRY(launch_small(pdb, d.ps.gram_part, pdb.nblk, SMALL_EIGVECS, d.Msmall, d.evals, d.status, nullptr, d.st));
        SPE_TRY(launch_apply(pdb, d.buf[BQ], d.Msmall, d.buf[BV1], nullptr, nullptr, d.ps.ticket, d.st));   // Ritz vectors
        SPE_TRY(launch_apply(pdb, d.buf[BZ], d.Msmall, d.buf[BV2], nullptr, nullptr, d.ps.ticket, d.st));   // S * Ritz vectors
        SPE_TRY(launch_residual_norms(pdb, d.buf[BV2], d.buf[BV1], d.evals, d.ps, d.scal + 64, d.st));
        SPE_CUDA(cudaMemcpyAsync(d.scal, d.evals, bq * sizeof(double), cudaMemcpyDeviceToDevice, d.st));
      }
      SPE_TRY(read_scalars(64 + bq));
      for (int c = 0; c < bq; ++c) sv2[c] = h_scal[c];
      const double th0 = std::max(std::fabs(sv2[0]), 1e-300);
      double r = 0.0;
      for (int c = 0; c < q; ++c) r = std::max(r, h_scal[64 + c]);
      resid_rel = r / th0;
      if (cfg.verbose) fprintf(stderr, "[speigen] init pass %d: theta1=%.6e theta_q=%.6e resid=%.3e\n", passes, sv2[0], sv2[q - 1], resid_rel);
      if (!(resid_rel == resid_rel)) { set_error("This function cannot handle NAs."); return SPEIGEN_ERR_NA; }
      if (resid_rel <= cfg.init_tol) { converged = true; break; }
      // round-off floor: residual stopped improving at a tiny level
      if (resid_rel > 0.5 * prev) ++stall; else stall = 0;
      if (stall >= 3 && resid_rel <= 1e-11) { converged = true; break; }
      // Chebyshev acceleration for slowly separating spectra: when plain power steps gain less than 2x per pass, the
      // Ritz vectors X are filtered with T_deg((S - c)/e), [c - e, c + e] = [lower bound, smallest Ritz value of the
      // block] (the unwanted interval), each column rescaled by 1 / T_deg(x_c) so the block stays well conditioned.
      // The degree doubles while convergence stays slow, capped so that wanted directions differ by <= 1e6 in gain.
      // per-pass rate since the previous Rayleigh-Ritz test
      double rate = 0.0;
      if (last_rr_pass > 0 && last_rr_resid > 0.0) rate = std::pow(resid_rel / last_rr_resid, 1.0 / static_cast<double>(passes - last_rr_pass));
      last_rr_pass = passes;
      last_rr_resid = resid_rel;
      prev = std::min(prev, resid_rel);
      lam_low = std::min(lam_low, sv2[bq - 1]);
      if (cheb_deg == 1 && rate > 0.0 && rate < 0.5 && resid_rel > cfg.init_tol) {
        const double need = std::ceil(std::log(cfg.init_tol / resid_rel) / std::log(rate));     // passes predicted to reach the tolerance
        // test one pass AFTER the predicted one: the MM trajectory is sensitive to the start vectors at the 1e-15 level (its
        // late SQUAREM steps amplify by |a| ~ 1e3..1e5), so the initialiser should end below init_tol, not at it
        // Only where a pass is cheap (a shard below 4 GB streams in < 0.7 ms, about one test): on a 20 GB shard a pass costs
        // ten tests, a mispredicted pass would cost more than all the skipped tests save, so every pass is tested there.
        const bool cheap_pass = static_cast<double>(devs[0].a_rows) * static_cast<double>(devs[0].a_ld) * sizeof(double) < 4e9;
        rr_skip = cheap_pass ? static_cast<int>(std::min(6.0, std::max(0.0, need))) : 0;
      }
      // Locking: Ritz pairs of the leading columns that have converged (e.g. the spikes of a spiked model when q reaches into the
      // bulk) are deflated - the filter then works on (I - X_L X_L^H) S (I - X_L X_L^H), whose wanted eigenvalues are close
      // together, so its degree is no longer capped by the gain ratio between a spike and a bulk-edge eigenvalue.
      int nlock = 0;
      while (nlock < q && h_scal[64 + nlock] <= cfg.init_tol * th0) ++nlock;
      const int deg_cap = (nlock > 0) ? 96 : 24;
      if (passes >= 3 && rate > 0.5 && bq > q) cheb_deg = std::min(cheb_deg * 2, deg_cap);
      int deg = std::min(cheb_deg, deg_cap);
      const double ub = sv2[bq - 1], lb = std::min(0.0, lam_low);
      const double cc = 0.5 * (ub + lb), ee = 0.5 * (ub - lb);
      auto cheb = [](int d, double x) { return x > 1.0 ? std::cosh(d * std::acosh(x)) : 1.0; };
      if (!(ub > 0.0)) deg = 1;   // the filter interval [lb, ub] assumes a PSD spectrum below the block
      if (deg > 1) {
        const double x0 = (sv2[nlock] - cc) / ee, xq = (sv2[q - 1] - cc) / ee;     // gain ratio among the ACTIVE wanted directions
        while (deg > 1 && !(cheb(deg, x0) <= 1e6 * cheb(deg, xq))) --deg;   // also steps down on overflow (inf/NaN)
      }
      if (deg <= 1 || !(ee > 0.0)) {
        SPE_TRY(op_polar(BV2, BQ, pdb, false, nullptr, 0, WARM_INIT));
      } else {
        Buf yprev = BV1, ycur = BT1, yspare = BV0;
        // Y <- Y - X_L (X_L^H Y): three small panel kernels per contraction when something is locked
        auto project = [&](Buf y) -> int {
          if (nlock == 0) return SPEIGEN_OK;
          for (Dev& d : devs) {
            SPE_CUDA(cudaSetDevice(d.ordinal));
            SPE_TRY(launch_gram(pdb, d.buf[BVX], d.buf[y], d.ps.gram_part, d.st));
            SPE_TRY(launch_gram_sum_hat(pdb, d.ps.gram_part, pdb.nblk, d.Msmall, d.st));
            SPE_TRY(launch_apply(pdb, d.buf[BVX], d.Msmall, d.buf[BT2], nullptr, nullptr, d.ps.ticket, d.st));
            SPE_TRY(launch_axpby(pdb, d.buf[y], d.buf[BT2], 1.0, -1.0, d.buf[y], d.st));
          }
          return SPEIGEN_OK;
        };
        for (Dev& d : devs) {   // X_L (locked Ritz vectors, zeros beyond) ; Y1 = (S X - c X) / e
          SPE_CUDA(cudaSetDevice(d.ordinal));
          if (nlock > 0) SPE_TRY(launch_mask_columns(pdb, d.buf[BV1], nlock, d.buf[BVX], d.st));
          SPE_TRY(launch_axpby(pdb, d.buf[BV2], d.buf[BV1], 1.0 / ee, -cc / ee, d.buf[ycur], d.st));
        }
        SPE_TRY(project(ycur));
        for (int kk = 1; kk < deg; ++kk) {   // Y_{k+1} = (2/e)(S Y_k - c Y_k) - Y_{k-1}
          SPE_TRY(op_contract(ycur, pdb, BZ, NBUF, NBUF, false, sc0));
          for (Dev& d : devs) {
            SPE_CUDA(cudaSetDevice(d.ordinal));
            SPE_TRY(launch_axpbypcz(pdb, d.buf[BZ], d.buf[ycur], d.buf[yprev], 2.0 / ee, -2.0 * cc / ee, -1.0, d.buf[yspare], d.st));
          }
          SPE_TRY(project(yspare));
          const Buf t = yprev; yprev = ycur; ycur = yspare; yspare = t;
        }
        std::vector<double> scale(kMaxCols, 1.0);
        for (int c = nlock; c < bq; ++c) scale[c] = 1.0 / cheb(deg, (sv2[c] - cc) / ee);
        for (Dev& d : devs) {
          SPE_CUDA(cudaSetDevice(d.ordinal));
          SPE_CUDA(cudaMemcpyAsync(d.rho, scale.data(), bq * sizeof(double), cudaMemcpyHostToDevice, d.st));
          SPE_TRY(launch_scale_columns(pdb, d.buf[ycur], d.rho, d.st));
          if (nlock > 0) SPE_TRY(launch_blend_columns(pdb, d.buf[ycur], d.buf[BVX], nlock, d.st));   // keep the locked vectors in the block
        }
        SPE_TRY(sync_all());   // `scale` is a host temporary
        SPE_TRY(op_polar(ycur, BQ, pdb, false, nullptr, 0, WARM_INIT));
        if (cfg.verbose) fprintf(stderr, "[speigen] init pass %d: Chebyshev degree %d on [%.3e, %.3e], %d locked\n", passes, deg, lb, ub, nlock);
      }
      {
        Dev& d = devs[0];
        int st = 0, tst = 0;
        SPE_CUDA(cudaSetDevice(d.ordinal));
        SPE_CUDA(cudaMemcpyAsync(&st, d.status, sizeof(int), cudaMemcpyDeviceToHost, d.st));
        SPE_CUDA(cudaMemcpyAsync(&tst, d.tstatus, sizeof(int), cudaMemcpyDeviceToHost, d.st));
        SPE_CUDA(cudaStreamSynchronize(d.st));
        if (tst & 4) return check_tail_status(static_cast<double>(tst));
        if (st || tst) {          // a block that cannot be orthonormalised (zero / dependent columns): rank-deficient input
          for (Dev& dd : devs) { cudaSetDevice(dd.ordinal); cudaMemsetAsync(dd.tstatus, 0, sizeof(int), dd.st); }
          rank_bad = true;
          break;
        }
      }
    }
    if (rank_bad) {
      if (bq > q && attempt == 0) continue;
      set_error("The number of estimated eigenvectors q should not be larger than rank(X).");
      return SPEIGEN_ERR_RANK;
    }
    break;
  }
  if (passes > cfg.init_max_passes) passes = cfg.init_max_passes;
  if (!converged && !cfg.allow_unconverged_init) {
    // eigen()/svd() of the reference always converge; carrying on with inexact sv2 / Vx would silently change the rho scaling, the
    // returned `values` / `standard_vectors` and the start point (and would skip the PD probe)
    set_error("the block subspace iteration that replaces eigen()/svd() did not converge: residual %.3e after %d passes "
              "(init_tol %.1e); raise init_max_passes or set allow_unconverged_init", resid_rel, passes, cfg.init_tol);
    return SPEIGEN_ERR_NOT_CONVERGED;
  }
  // R/spEigen.R:77 tests the full spectrum.  Without a full EVD we bound lambda_min from above with Ritz values: a few
  // block power passes on (theta_1 I - S) pull the block towards the bottom of the spectrum; any Ritz value of S below
  // pd_tol proves "not PD".  (A tiny negative eigenvalue hidden in a dense bottom cluster can escape this test.)
  if (!data && cfg.pd_check_passes > 0) {
    const double sigma = sv2[0];
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_CUDA(cudaMemsetAsync(d.buf[BQ], 0, static_cast<size_t>(mrows) * pdb.pitch * sizeof(double), d.st));
      SPE_TRY(launch_random_panel(pdb, cfg.init_seed + 7919, d.buf[BQ], d.st));
    }
    SPE_TRY(op_polar(BQ, BQ, pdb, false, nullptr, 0));
    double lam_min_est = sigma;
    for (int p = 0; p < cfg.pd_check_passes; ++p) {
      SPE_TRY(op_contract(BQ, pdb, BZ, NBUF, NBUF, false, sc0));
      for (Dev& d : devs) {
        SPE_CUDA(cudaSetDevice(d.ordinal));
        SPE_TRY(launch_gram(pdb, d.buf[BQ], d.buf[BZ], d.ps.gram_part, d.st));     // T = Q^H S Q
        SPE_TRY(launch_small(pdb, d.ps.gram_part, pdb.nblk, SMALL_EIGVECS, d.Msmall, d.evals, d.status, nullptr, d.st));
        SPE_CUDA(cudaMemcpyAsync(d.scal, d.evals, bq * sizeof(double), cudaMemcpyDeviceToDevice, d.st));
        SPE_TRY(launch_axpby(pdb, d.buf[BQ], d.buf[BZ], sigma, -1.0, d.buf[BT1], d.st));   // (sigma I - S) Q
      }
      SPE_TRY(read_scalars(bq));
      lam_min_est = std::min(lam_min_est, h_scal[bq - 1]);
      if (cfg.verbose) fprintf(stderr, "[speigen] PD probe pass %d: smallest Ritz value %.6e\n", p + 1, h_scal[bq - 1]);
      if (lam_min_est < cfg.pd_tol) break;
      SPE_TRY(op_polar(BT1, BQ, pdb, false, nullptr, 0));
    }
    if (lam_min_est < cfg.pd_tol) { set_error("The covariance matrix is not PD."); return SPEIGEN_ERR_NOT_PD; }
    for (Dev& d : devs) {           // the probe may have tripped the rank flag on a singular block; it is not an error here
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_CUDA(cudaMemsetAsync(d.status, 0, sizeof(int), d.st));
      SPE_CUDA(cudaMemsetAsync(d.tstatus, 0, sizeof(int), d.st));
    }
  }
  // checks of R/spEigen.R:69,77,79 on the Ritz values that exist
  for (int c = 0; c < bq; ++c)
    if (!data && sv2[c] < cfg.pd_tol) { set_error("The covariance matrix is not PD."); return SPEIGEN_ERR_NOT_PD; }
  for (int c = 0; c < bq; ++c)
    if (sv2[c] < 0.0) sv2[c] = 0.0;                           // :78
  const double rank_thr = data ? cfg.rank_tol * cfg.rank_tol : cfg.rank_tol;   // :69 tests d > 1e-9 on singular values
  if (!(sv2[q - 1] > rank_thr)) {
    set_error("The number of estimated eigenvectors q should not be larger than rank(X).");
    return SPEIGEN_ERR_RANK;
  }
  // Vx[,1:q] -> BVX ; start point V
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemsetAsync(d.buf[BVX], 0, static_cast<size_t>(mrows) * pdb.pitch * sizeof(double), d.st));
    SPE_TRY(launch_take_columns(pdb, d.buf[BV1], pdq, d.buf[BVX], d.st));
  }
  if (V0_host) {
    SPE_TRY(op_upload_panel(V0_host, BV, pdq));
  } else {
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_CUDA(cudaMemcpyAsync(d.buf[BV], d.buf[BVX], static_cast<size_t>(mrows) * pdq.pitch * sizeof(double), cudaMemcpyDeviceToDevice, d.st));
    }
  }
  SPE_TRY(sync_all());
  inited = true;
  if (out) {
    out->init_passes = passes;
    out->init_converged = converged ? 1 : 0;
    out->init_residual = resid_rel;
    out->seconds_init = now_s() - t0;
  }
  return SPEIGEN_OK;
}

// ---- the MM loop (R/spEigen.R:105-160) ----------------------------------------------------------
// Per outer iteration k (5 launches + 3 per backtrack, one host read-back per try):
//   tail   V1 = polar(reweight(S V, V))                         -> rows of V1 on every GPU, min|V1|
//   K1     B  = reweight(S V1, V1)   (fused epilogue, own rows)
//   tail   V2 = polar(B)                                        -> ||R||^2, ||U||^2
//   tail   V0 = polar(V - 2aR + a^2 U), a on the device         -> rows of V0 on every GPU, min|V0|, penalty sums
//   K1     Y0 = S V0 (own rows) + column dots Re(V0^H Y0)
//   finish F  -> mapped pinned memory
int Solver::run(double rho, const double* d_in, double thres, speigen_out* out) {
  if (!inited) { set_error("solver_run before solver_init"); return SPEIGEN_ERR_ARG; }
  const double t0 = now_s();
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemsetAsync(d.twarm, 0, static_cast<size_t>(N_TAIL_SITES) * 2 * kWarmDoubles * sizeof(double), d.st));
    SPE_CUDA(cudaMemsetAsync(d.tstatus, 0, sizeof(int), d.st));
    SPE_CUDA(cudaMemsetAsync(d.tdiag, 0, 48 * sizeof(double), d.st));
  }
  std::vector<double> dv(q), rv(q);
  if (d_in) for (int c = 0; c < q; ++c) dv[c] = d_in[c];
  else speigen_default_d(q, dv.data());
  // rho <- rho * max(...) * (sv2[1:q]/sv2[1]) * d/d[1]      R/spEigen.R:72,82
  for (int c = 0; c < q; ++c) rv[c] = rho * scale_max * (sv2[c] / sv2[0]) * dv[c] / dv[0];
  SPE_TRY(set_qvecs(dv.data(), rv.data()));

  const int K = cfg.n_continuation;
  std::vector<double> pp(K + 1), ee(K + 1), tl(K + 1);
  speigen_schedule(&cfg, pp.data(), ee.data(), tl.data());
  const bool fused = cfg.fused_epilogue != 0 && !(data && prob.world > 1);
  const int64_t c_before = contractions;

  StageConsts sc = make_stage(pp[0], ee[0]);
  Buf bV = BV, bV0 = BV0, bY = BY, bY0 = BY0;
  // start point: its column minima (and, row-sharded, nothing else to distribute: V is replicated by the initialiser)
  {
    TailSpec ts;
    ts.T = bV; ts.out = bV; ts.skip_polar = true; ts.absmin = true;
    SPE_TRY(op_tail(ts, pdq, sc));
  }
  SPE_TRY(op_contract_loop(bV, pdq, bY, NBUF, NBUF, false, sc));      // Y = S V
  std::vector<double> Fv;
  int k = 0, total_bt = 0;
  const int o_diag = 1 + 2 * q;
  for (int st = 0; st <= K; ++st) {
    sc = make_stage(pp[st], ee[st]);
    int flg = 1;
    while (true) {
      ++k;
      // ---- V1 = MMupdate(V): reweighting from the inherited Y = S V, then the polar factor ----
      {
        TailSpec ts;
        ts.producer = TAIL_REWEIGHT; ts.in0 = bY; ts.in1 = bV; ts.T = BB; ts.out = BV1; ts.out_all = true;
        ts.absmin = true; ts.use_prev = true; ts.site = SITE_MM1;
        SPE_TRY(op_tail(ts, pdq, sc));
      }
      // ---- V2 = MMupdate(V1): contraction with the reweighting fused into its epilogue ----
      if (fused) {
        SPE_TRY(op_contract_loop(BV1, pdq, NBUF, BB, BV1, false, sc));
        TailSpec ts;
        ts.producer = TAIL_PANEL; ts.T = BB; ts.out = BV2; ts.squarem = true; ts.sqV = bV; ts.sqV1 = BV1; ts.site = SITE_MM2;
        SPE_TRY(op_tail(ts, pdq, sc));
      } else {
        SPE_TRY(op_contract_loop(BV1, pdq, BT1, NBUF, NBUF, false, sc));
        TailSpec ts;
        ts.producer = TAIL_REWEIGHT; ts.in0 = BT1; ts.in1 = BV1; ts.T = BB; ts.out = BV2; ts.use_prev = true;
        ts.squarem = true; ts.sqV = bV; ts.sqV1 = BV1; ts.site = SITE_MM2;
        SPE_TRY(op_tail(ts, pdq, sc));
      }
      int nbt = 0;
      double Fk = 0.0, a = -1.0, nR = 0.0, nU = 0.0;
      while (true) {                            // backtracking loop, R/spEigen.R:126-146
        // SQUAREM step length a = min(-||R||/||U||, -1) (:121-123) is formed on the device from the statistics of the V2 tail
        // (first try) or halved from the previous try (:141); extrapolation + projection :127-131
        TailSpec ts;
        ts.producer = TAIL_EXTRAP; ts.in0 = bV; ts.in1 = BV1; ts.in2 = BV2; ts.T = BT1; ts.out = bV0; ts.out_all = true;
        ts.absmin = true; ts.gsum = true; ts.use_prev = (nbt == 0); ts.retry = (nbt > 0) ? 1 : 0; ts.site = SITE_EXTRAP;
        SPE_TRY(op_tail(ts, pdq, sc));
        SPE_TRY(op_contract_loop(bV0, pdq, bY0, NBUF, bV0, true, sc));
        SPE_TRY(op_finish_objective(pdq));
        SPE_CUDA(cudaSetDevice(devs[0].ordinal));
        SPE_CUDA(cudaStreamSynchronize(devs[0].st));
        SPE_TRY(check_tail_status(h_res[o_diag + 5]));
        Fk = h_res[0];
        a = h_res[o_diag + 0];
        if (nbt == 0) { nR = h_res[o_diag + 1]; nU = h_res[o_diag + 2]; }
        const double Fprev = Fv.empty() ? Fk : Fv[std::max(k - 1, 1) - 1];
        const double sgn = (Fk > 0) - (Fk < 0);
        if (flg == 0 && Fk * (1.0 + sgn * cfg.backtrack_slack) <= Fprev && nbt < cfg.max_backtracks) {
          ++nbt;                                // a <- (a - 1) / 2   :141 (on the device, next try)
        } else {
          std::swap(bV, bV0);                   // V <- V0   :143
          std::swap(bY, bY0);
          break;
        }
      }
      total_bt += nbt;
      Fv.push_back(Fk);
      if (out && k <= out->trace_capacity) {
        if (out->F) out->F[k - 1] = Fk;
        if (out->step_a) out->step_a[k - 1] = a;
        if (out->backtracks_per_k) out->backtracks_per_k[k - 1] = nbt;
        if (out->stage_of_k) out->stage_of_k[k - 1] = st + 1;
      }
      if (cfg.verbose) fprintf(stderr, "[speigen] stage %d k %d: nR=%.3e nU=%.3e a=%.4g bt=%d F=%.15g\n", st + 1, k, nR, nU, a, nbt, Fk);
      if (flg == 0) {                           // stopping criterion  :149-153 (k >= max_iter ends this stage only, as in R)
        const double Fp = Fv[k - 2];
        const double rel = std::fabs(Fk - Fp) / std::max(1.0, std::fabs(Fp));
        if (rel <= tl[st] || k >= cfg.max_iter) break;
      }
      flg = 0;
    }
  }
  // threshold + normalise + outputs   :158-162
  if (out) {
    // every rank holds every row of the accepted iterate (the extrapolation tail stores its rows to all ranks); make sure the
    // peers' stores have landed before device 0 reads the panel
    SPE_TRY(wait_latest_tail());
    Dev& d = devs[0];
    SPE_CUDA(cudaSetDevice(d.ordinal));
    const size_t bytes = static_cast<size_t>(m) * pdq.ncr * sizeof(double);
    if (out->vectors) {
      SPE_TRY(launch_finalize(pdq, d.buf[bV], thres, d.ps, d.scal + 128, d.out_cm, d.st));
      SPE_CUDA(cudaMemcpyAsync(out->vectors, d.out_cm, bytes, cudaMemcpyDeviceToHost, d.st));
      SPE_CUDA(cudaStreamSynchronize(d.st));
    }
    if (out->standard_vectors) SPE_TRY(op_download_panel(BVX, pdq, out->standard_vectors));
    if (out->values)
      for (int c = 0; c < q; ++c) out->values[c] = data ? sv2[c] / static_cast<double>(n - 1) : sv2[c];   // :162
    out->iterations = k;
    out->backtracks = total_bt;
    out->contractions = static_cast<int>(contractions - c_before);
    double dg[8] = {};
    SPE_CUDA(cudaMemcpyAsync(dg, d.tdiag, sizeof(dg), cudaMemcpyDeviceToHost, d.st));
    SPE_CUDA(cudaStreamSynchronize(d.st));
    out->polar_passes_max = static_cast<int>(dg[5]);
    out->seconds_loop = now_s() - t0;
  }
  SPE_TRY(sync_all());
  {
    int stt = 0;
    Dev& d = devs[0];
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemcpy(&stt, d.tstatus, sizeof(int), cudaMemcpyDeviceToHost));
    SPE_TRY(check_tail_status(static_cast<double>(stt)));
  }
  return SPEIGEN_OK;
}

// all ranks' rows of the latest tail's output panel have arrived on every local device (host-side wait on the stat flags)
int Solver::wait_latest_tail() {
  if (!sharded() || prob.world == 1) return SPEIGEN_OK;
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_TRY(launch_wait_flags(d.mail->stat_flag, kFlagStride, prob.world, stat_epoch, d.tstatus,
                              static_cast<long long>(cfg.exchange_timeout_ms > 0 ? cfg.exchange_timeout_ms : 10000) * 1000000ll, d.st));
  }
  return SPEIGEN_OK;
}

// ---- bench helpers ------------------------------------------------------------------------------
void Solver::enable_kernel_timing(bool on) {
  if (on && kev.empty()) {
    cudaSetDevice(devs[0].ordinal);
    kev.resize(8192);
    for (auto& e : kev) cudaEventCreate(&e);
  }
  ktiming = on;
  kev_used = 0;
}
int Solver::kernel_time(double* ms_total, int* launches) {
  SPE_TRY(sync_all());
  double t = 0.0;
  for (size_t i = 0; i + 1 < kev_used; i += 2) {
    float ms = 0.f;
    SPE_CUDA(cudaEventElapsedTime(&ms, kev[i], kev[i + 1]));
    t += ms;
  }
  *ms_total = t;
  *launches = static_cast<int>(kev_used / 2);
  kev_used = 0;
  return SPEIGEN_OK;
}

int Solver::load_device_rows(int li, int64_t row0, int64_t nrows, const double* src, int64_t src_ld) {
  if (li < 0 || li >= static_cast<int>(devs.size())) return SPEIGEN_ERR_ARG;
  Dev& d = devs[li];
  if (row0 < 0 || nrows < 0 || row0 + nrows > d.a_rows) { set_error("load_device_rows: rows %lld+%lld outside 0..%lld", (long long)row0, (long long)nrows, (long long)d.a_rows); return SPEIGEN_ERR_ARG; }
  SPE_CUDA(cudaSetDevice(d.ordinal));
  SPE_CUDA(cudaMemcpy2DAsync(d.A + row0 * d.a_ld, d.a_ld * sizeof(double), src, src_ld * sizeof(double), d.a_kd * sizeof(double),
                             static_cast<size_t>(nrows), cudaMemcpyDeviceToDevice, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  loaded = true;
  host_verified = false;
  return SPEIGEN_OK;
}

int Solver::load_device(int li, const double* src, int64_t src_ld) {
  if (li < 0 || li >= static_cast<int>(devs.size())) return SPEIGEN_ERR_ARG;
  Dev& d = devs[li];
  SPE_CUDA(cudaSetDevice(d.ordinal));
  SPE_CUDA(cudaMemcpy2DAsync(d.A, d.a_ld * sizeof(double), src, src_ld * sizeof(double), d.a_kd * sizeof(double),
                             static_cast<size_t>(d.a_rows), cudaMemcpyDeviceToDevice, d.st));
  SPE_CUDA(cudaStreamSynchronize(d.st));
  loaded = true;
  host_verified = false;
  return SPEIGEN_OK;
}

int Solver::mm_steps(int nsteps, int stage, double rho, const double* d_in, float* ms_out) {
  if (!inited) { set_error("mm_steps before solver_init"); return SPEIGEN_ERR_ARG; }
  std::vector<double> dv(q), rv(q);
  if (d_in) for (int c = 0; c < q; ++c) dv[c] = d_in[c];
  else speigen_default_d(q, dv.data());
  for (int c = 0; c < q; ++c) rv[c] = rho * scale_max * (sv2[c] / sv2[0]) * dv[c] / dv[0];
  SPE_TRY(set_qvecs(dv.data(), rv.data()));
  const int K = cfg.n_continuation;
  std::vector<double> pp(K + 1), ee(K + 1), tl(K + 1);
  speigen_schedule(&cfg, pp.data(), ee.data(), tl.data());
  if (stage < 0 || stage > K) stage = 0;
  const StageConsts sc = make_stage(pp[stage], ee[stage]);
  const bool fused = !(data && prob.world > 1);
  Buf bA = BV, bB = BV1;
  {
    TailSpec ts;   // column minima of the resident iterate
    ts.T = bA; ts.out = bA; ts.skip_polar = true; ts.absmin = true;
    SPE_TRY(op_tail(ts, pdq, sc));
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ms_out) {
    SPE_TRY(sync_all());
    SPE_CUDA(cudaSetDevice(devs[0].ordinal));
    SPE_CUDA(cudaEventCreate(&e0));
    SPE_CUDA(cudaEventCreate(&e1));
    SPE_CUDA(cudaEventRecord(e0, devs[0].st));
  }
  for (int i = 0; i < nsteps; ++i) {
    TailSpec ts;
    ts.out = bB; ts.out_all = true; ts.absmin = true; ts.site = SITE_MM1; ts.T = BB;
    if (fused) {
      SPE_TRY(op_contract_loop(bA, pdq, NBUF, BB, bA, false, sc));     // K1 with the fused reweighting epilogue
      ts.producer = TAIL_PANEL;
    } else {
      SPE_TRY(op_contract_loop(bA, pdq, BT1, NBUF, NBUF, false, sc));
      ts.producer = TAIL_REWEIGHT; ts.in0 = BT1; ts.in1 = bA; ts.use_prev = true;
    }
    SPE_TRY(op_tail(ts, pdq, sc));                                     // fused polar tail (+ min|V| for the next weights)
    std::swap(bA, bB);
  }
  if (ms_out) {
    SPE_CUDA(cudaSetDevice(devs[0].ordinal));
    SPE_CUDA(cudaEventRecord(e1, devs[0].st));
    SPE_TRY(sync_all());
    SPE_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  if (bA != BV) {        // leave the final iterate in BV (all rows are present on every rank once the peers' flags are up)
    SPE_TRY(wait_latest_tail());
    for (Dev& d : devs) {
      SPE_CUDA(cudaSetDevice(d.ordinal));
      SPE_CUDA(cudaMemcpyAsync(d.buf[BV], d.buf[bA], static_cast<size_t>(mrows) * pdq.pitch * sizeof(double), cudaMemcpyDeviceToDevice, d.st));
    }
  }
  SPE_TRY(sync_all());
  {
    int stt = 0;
    Dev& d = devs[0];
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemcpy(&stt, d.tstatus, sizeof(int), cudaMemcpyDeviceToHost));
    SPE_TRY(check_tail_status(static_cast<double>(stt)));
  }
  return SPEIGEN_OK;
}

// ---- timing helper for bench.py (CUDA events on the launching stream) ---------------------------
int Solver::time_contract(int ncols, int reps, int flush_l2, float* ms_each) {
  PanelDims pd = make_dims(m, ncols, cplx);
  if (pd.ncr > kMaxCols || ncols > bq) { set_error("time_contract: ncols %d > block %d", ncols, bq); return SPEIGEN_ERR_ARG; }
  const size_t flush_bytes = 512ull << 20;
  for (Dev& d : devs) {
    SPE_CUDA(cudaSetDevice(d.ordinal));
    SPE_CUDA(cudaMemsetAsync(d.buf[BQ], 0, static_cast<size_t>(mrows) * pdb.pitch * sizeof(double), d.st));
    SPE_TRY(launch_random_panel(pd, 1234, d.buf[BQ], d.st));
    if (flush_l2 && !d.flush) SPE_CUDA(cudaMalloc(&d.flush, flush_bytes));
  }
  SPE_TRY(sync_all());
  const StageConsts sc0 = make_stage(0.1, 0.1);
  Dev& d0 = devs[0];
  SPE_CUDA(cudaSetDevice(d0.ordinal));
  cudaEvent_t e0, e1;
  SPE_CUDA(cudaEventCreate(&e0));
  SPE_CUDA(cudaEventCreate(&e1));
  for (int r = 0; r < reps; ++r) {
    if (flush_l2)
      for (Dev& d : devs) {
        SPE_CUDA(cudaSetDevice(d.ordinal));
        SPE_CUDA(cudaMemsetAsync(d.flush, r & 0xff, flush_bytes, d.st));
      }
    SPE_TRY(sync_all());
    SPE_CUDA(cudaSetDevice(d0.ordinal));
    SPE_CUDA(cudaEventRecord(e0, d0.st));
    SPE_TRY(op_contract(BQ, pd, BZ, NBUF, NBUF, false, sc0));
    SPE_CUDA(cudaSetDevice(d0.ordinal));
    SPE_CUDA(cudaEventRecord(e1, d0.st));
    SPE_TRY(sync_all());
    SPE_CUDA(cudaEventElapsedTime(&ms_each[r], e0, e1));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return SPEIGEN_OK;
}

}  // namespace speigen
