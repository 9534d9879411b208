// K3 (round 2): the whole m x q "tail" of one MM update as ONE persistent kernel per GPU.
//
//   T   = producer(...)                      reweighting B = d*Y - (w(|V|) - wmax) V rho   R/spEigen.R:170-177
//                                            | SQUAREM extrapolation V - 2aR + a^2 U        R/spEigen.R:121-127
//                                            | a panel already in memory (K1's fused-reweight output)
//   Out = T (T^H T)^{-1/2}                   Procrustes / polar factor                      R/spEigen.R:130-131,180-182
//   stats(Out)                               column min|.| (next weights' column max, :176), penalty sums (:133-134),
//                                            SQUAREM norms ||V1-V||^2, ||V2-2V1+V||^2 (:121-123)
//
// Every GPU runs the kernel on ITS OWN rows of the panel only (row-sharded covariance: rows [row0,row1) = the rows of S
// it owns).  What crosses GPUs is three tiny exchanges through peer-mapped mailboxes (NVLink P2P stores + epoch flags,
// no host, no NCCL launch): the q x q Gram partial of every polar pass and the final statistics; the finished rows of
// Out are stored straight into every peer's copy of the panel, so the next contraction finds its operand in place.
// Inside a GPU the kernel is a persistent grid (<= one CTA per SM, all co-resident) with ticket/flag grid barriers; the
// small q x q solve (parallel-ordered two-sided Jacobi or a Taylor series when the Gram matrix is already I + E) is
// executed redundantly by every CTA so that no broadcast step follows it.  All reductions have a fixed order.
#pragma once
#include "common.cuh"

namespace speigen {

constexpr int kTailThreads = 512;
constexpr int kTailMaxRanks = 8;
constexpr int kTriMax = kMaxCols * (kMaxCols + 1) / 2;     // packed upper triangle of the largest Gram matrix
constexpr int kTailMaxPasses = 6;
// statistics vector published by every tail call: [0,64) column min|Out|, [64,128) penalty column sums,
// 128 ||R||_F^2, 129 ||U||_F^2 (partial sums over the rank's rows)
constexpr int kStatAbsmin = 0;
constexpr int kStatGsum = kMaxCols;
constexpr int kStatSqR = 2 * kMaxCols;
constexpr int kStatSqU = 2 * kMaxCols + 1;
constexpr int kStatLen = 2 * kMaxCols + 8;
constexpr int kFlagStride = 16;   // one 128-byte line per flag

// Peer-mapped mailbox (one per GPU, same layout everywhere).  Writers: every rank (its own slot r); readers: the owner.
// The Gram partials travel as self-validating words (no fence, no flag on the critical path): every double is sent as two
// 8-byte words {32 data bits | 32-bit epoch tag}; an aligned 8-byte store is atomic, so a reader that sees the current tag in
// both words has the value - one NVLink store latency per exchange instead of store + fence.sys + flag + acquire.
struct Mailbox {
  unsigned long long stat_flag[kTailMaxRanks * kFlagStride];   // epoch of the last statistics vector (and rows) from rank r
  unsigned long long gram_ll[2][kTailMaxRanks][kTriMax][2];    // [epoch parity][rank][packed upper triangle][lo, hi word]
  double stat[2][kTailMaxRanks][kStatLen];
};

enum TailProducer { TAIL_PANEL = 0, TAIL_REWEIGHT = 1, TAIL_EXTRAP = 2 };

struct TailArgs {
  // geometry: this GPU's rows of the m x q panel (global row indices), panel layout
  int64_t row0 = 0, row1 = 0;
  int q = 0, cplx = 0, ncr = 0, pitch = 0;
  int nranks = 1, my_rank = 0;
  // producer
  int producer = TAIL_PANEL;
  const double* in0 = nullptr;    // REWEIGHT: Y ; EXTRAP: V
  const double* in1 = nullptr;    // REWEIGHT: V ; EXTRAP: V1
  const double* in2 = nullptr;    //               EXTRAP: V2
  double* T = nullptr;            // work panel (PANEL: holds the input; overwritten by the intermediate passes)
  const double* d = nullptr;      // [q]
  const double* rho = nullptr;    // [q]
  StageConsts sc{};
  // statistics of the previous tail call (absmin for REWEIGHT, squarem sums for EXTRAP): wait, then read
  const unsigned long long* prev_flags = nullptr;   // own mailbox stat_flag
  unsigned long long prev_epoch = 0;
  const double* prev_stat = nullptr;                // own mailbox stat[parity] : [nranks][kStatLen]
  // SQUAREM step (EXTRAP): retry == 0: a = min(-||R||/||U||, -1) from prev_stat; retry != 0: a = (a_cur - 1) / 2   (:123,:141)
  double* a_cur = nullptr;
  int retry = 0;
  // outputs
  double* out_peer[kTailMaxRanks] = {};   // the destination panel on every rank (entry my_rank = the local copy)
  int n_out = 1;                          // 1: local store only (out_peer[my_rank]); nranks: store to every rank
  int skip_polar = 0;                     // 1: Out = T (statistics / row distribution only)
  // statistics wanted
  int want_absmin = 0, want_gsum = 0, want_squarem = 0;
  const double* sq_V = nullptr;           // squarem: V, V1 (Out plays V2)
  const double* sq_V1 = nullptr;
  // exchange
  Mailbox* mail[kTailMaxRanks] = {};      // mailbox of every rank as mapped here (entry my_rank = own)
  unsigned long long* gram_epoch = nullptr;   // device counter of Gram exchanges (advanced by the kernel)
  unsigned long long stat_epoch = 0;          // epoch of THIS call's statistics
  // scratch (device local)
  double* cta_part = nullptr;     // [grid][kTriMax]
  double* cta_stat = nullptr;     // [grid][kStatLen]
  int* ticket = nullptr;          // [2] grid tickets (self-resetting)
  const double* warm_in = nullptr;    // Jacobi warm start: {n, count, W[n*n]} of the previous call from this call site
  double* warm_out = nullptr;
  int* status = nullptr;          // bit 0: non-finite / non-positive Gram, bit 1: polar did not converge (rank deficient), bit 2: exchange timeout
  double* diag = nullptr;         // [48] a, ||R||, ||U||, passes, jacobi sweeps, -, -, #clock stamps, then CTA 0's phase clock (ns)
  long long spin_budget_ns = 10000000000ll;
};

constexpr size_t kWarmDoubles = 2 + static_cast<size_t>(kMaxCols) * kMaxCols;

// grid size used for `rows` own rows on a device with num_sms SMs
int tail_grid(int64_t rows, int num_sms);
// one-time per-device set-up (function attributes; forces the module to load before any kernel spins on a peer)
int tail_prepare_device();
int launch_tail(const TailArgs& a, int num_sms, cudaStream_t st);

// Objective assembly (R/spEigen.R:136-138): F = sum_c d_c quad_c - sum_c rho_c gsum_c.
//   quad_c: fixed-order sum of the contraction's dot partials (dots[n_dot_parts][q]; multi-GPU: one partial per rank, after
//           waiting for the contraction's epoch flags), gsum_c: sum over ranks of the tail statistics.
// out[0] = F, out[1..q] = quad, out[1+q..2q] = gsum, out[2q+1 .. 2q+4] = diag (a, ||R||, ||U||, passes), status.
struct FinishArgs {
  int q = 0, nranks = 1;
  const double* dots = nullptr;
  int n_dot_parts = 0;
  const unsigned long long* dots_flags = nullptr;   // [nranks] (nullptr: no wait)
  unsigned long long dots_epoch = 0;
  const unsigned long long* stat_flags = nullptr;   // own mailbox stat_flag
  unsigned long long stat_epoch = 0;
  const double* stat = nullptr;                     // own mailbox stat[parity]
  const double* d = nullptr;
  const double* rho = nullptr;
  const double* diag = nullptr;
  int* status = nullptr;
  double* out = nullptr;            // device scalars
  double* out_host = nullptr;       // mapped pinned host copy (nullable)
  long long spin_budget_ns = 10000000000ll;
};
int launch_finish_objective(const FinishArgs& a, cudaStream_t st);
// stream-ordered wait until flags[r * stride] >= epoch for r < n (bounded; sets bit 2 of *status on time-out)
int launch_wait_flags(const unsigned long long* flags, int stride, int n, unsigned long long epoch, int* status,
                      long long budget_ns, cudaStream_t st);

// device-side helpers shared with contract.cu (flag waits of the fused exchange)
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// self-validating 2 x 8-byte transfer of one double (see Mailbox)
__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned tag) {
  const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
  const unsigned long long t = static_cast<unsigned long long>(tag) << 32;
  const unsigned long long w0 = (b & 0xffffffffull) | t, w1 = (b >> 32) | t;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const unsigned long long* slot, unsigned tag, double& v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
  if (static_cast<unsigned>(w0 >> 32) != tag || static_cast<unsigned>(w1 >> 32) != tag) return false;
  v = __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
  return true;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait for *f >= target (relaxed polling, one acquire fence at the end; `sys`: the flag is written by a peer GPU).
// On time-out sets bit 2 of *status and returns false (the caller carries on with whatever is there; the host turns the
// status bit into SPEIGEN_ERR_TIMEOUT at its next read-back).
__device__ __forceinline__ bool spin_until_ge(const unsigned long long* f, unsigned long long target, long long budget_ns,
                                              int* status, bool sys = true) {
  bool ok = true;
  if ((sys ? ld_relaxed_sys_u64(f) : ld_relaxed_gpu_u64(f)) < target) {
    const unsigned long long t0 = global_timer_ns();
    unsigned it = 0;
    while ((sys ? ld_relaxed_sys_u64(f) : ld_relaxed_gpu_u64(f)) < target) {
      if ((++it & 63u) == 0u) {
        __nanosleep(20);
        if (static_cast<long long>(global_timer_ns() - t0) > budget_ns) {
          if (status) atomicOr(status, 4);
          ok = false;
          break;
        }
      }
    }
  }
  if (sys) fence_acq_rel_sys(); else fence_acq_rel_gpu();
  return ok;
}
#endif

}  // namespace speigen
