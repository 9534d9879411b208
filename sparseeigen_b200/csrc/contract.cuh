// K1: the tall-skinny contraction  Y = A * P  (and its transposed form  T = A^T * P)  where A is the
// streamed matrix (covariance slab / data shard, the HBM-bound operand) and P is an m x q panel.
// Replaces the two dgemm/zgemm calls of R/spEigen.R:177 and the X %*% V0 products of :136,:138.
#pragma once
#include "common.cuh"

namespace speigen {

// Streamed matrix resident in HBM, viewed as a row-major real matrix of `rows` x `kd` doubles with
// row pitch `ld` doubles (ld % 16 == 0, base 128-byte aligned).  R's column-major S (m x m) is
// exactly this with rows = columns of S (S = S^H); complex entries are interleaved
// {re,im} so kd = 2 * (complex length).
struct StreamMat {
  const double* ptr = nullptr;
  int64_t rows = 0;   // number of rows (outputs of the standard form)
  int64_t kd = 0;     // contiguous (contracted) dimension in doubles
  int64_t ld = 0;     // row pitch in doubles
  CUtensorMap tmap;   // box = 16 doubles x TJ rows, SWIZZLE_128B
  int box_rows = 0;
};

// Epilogue selector/arguments for the standard form.
struct EpiArgs {
  double* Y = nullptr;             // raw product rows -> Y[(row_off + j) * pitch + c]
  double* B = nullptr;             // fused reweighting: B = d*Y - (w(|V|) - wmax) * V * rho   (R/spEigen.R:170-177)
  const double* V = nullptr;       // panel the weights / column dots are taken from (global rows)
  const double* d = nullptr;       // [q]
  const double* rho = nullptr;     // [q]
  const double* wmax = nullptr;    // [q] min_j |V[j,c]|: the kernel derives the column max of the weights w(min) (R/spEigen.R:176)
  double* dots = nullptr;          // per-row-tile partial column dots Re(conj(V) .* Y): [n_row_tiles][q]
  StageConsts sc{};
  int q = 0;
  int cplx = 0;
  int pitch = 0;                   // panel pitch (doubles)
  int64_t row_off = 0;             // global row of local output row 0 (row-sharded covariance)
  // ---- fused exchange (row-sharded covariance, P > 1): the epilogue stores its row block straight into every GPU's
  // exchange buffer over NVLink (peer pointers, own GPU included) instead of a local store + NCCL allgather; the CTA
  // that finishes last publishes this rank's column-dot sums and raises an epoch flag on every peer.
  int npeers = 0;                  // 0 = local stores only
  int my_shard = 0;
  double* peerY[8] = {};           // exchange slot for Y on each GPU (as mapped into this GPU's address space)
  double* peerB[8] = {};
  double* peer_dots[8] = {};       // [world][q] on each GPU
  unsigned long long* peer_flag[8] = {};   // [world] epoch flags on each GPU
  unsigned long long epoch = 0;
  int* done_count = nullptr;       // CTA completion counter (self-resetting)
  // ---- row-sharded MM loop (round 2): the operand panel is assembled in place by the peers' tail kernels (tail.cuh), which
  // also publish the column minima of |V|.  The kernel waits for their epoch flags before it touches either.
  const unsigned long long* wait_flags = nullptr;   // own mailbox stat_flag (stride kFlagStride)
  int wait_n = 0;
  unsigned long long wait_epoch = 0;
  const double* stat_absmin = nullptr;   // [stat_n][stat_stride] column minima of |V| per rank (overrides wmax)
  int stat_n = 0;
  int stat_stride = 0;
  int* status = nullptr;           // bit 2: exchange wait timed out
  long long spin_budget_ns = 10000000000ll;
};

struct ContractWorkspace {
  double* partials = nullptr;      // split-K partial tiles
  int* counters = nullptr;         // per row tile arrival counters (self-resetting)
  size_t partial_slots = 0;
  int n_counters = 0;
};

// Tile configuration of the standard form.
constexpr int kTJ = 256;           // output rows per CTA tile
constexpr int kNBox = 2;           // TMA boxes (16 doubles wide) per pipeline stage
constexpr int kTK = 16 * kNBox;    // contracted doubles per stage
constexpr int kConsumerWarps = 8;                 // transposed form: consumer warps (16 outputs each)
#ifndef SPEIGEN_STD_WARPS
#define SPEIGEN_STD_WARPS 8
#endif
constexpr int kStdWarps = SPEIGEN_STD_WARPS;      // standard form: DMMA consumer warps per CTA
constexpr int kMT = kTJ / (8 * kStdWarps);        // m-tiles (8 rows) per warp
constexpr int kThreads = kConsumerWarps * 32 + 32;
constexpr int kMaxPeers = 8;

int make_stream_mat(StreamMat* sm, const double* ptr, int64_t rows, int64_t kd, int64_t ld, int box_rows);

// Y(rows x ncr) = A(rows x kd) * P(kd x ncr).  P is a panel with pitch panel_pitch(ncr) whose row count
// is padded (with zeros) to a multiple of kTK.  Deterministic (fixed-order split-K fix-up).
int launch_contract_std(const StreamMat& A, const double* P, int ncr, const EpiArgs& epi,
                        const ContractWorkspace& ws, int num_sms, cudaStream_t stream);
size_t contract_workspace_slots(int64_t rows, int num_sms);
size_t contract_slot_doubles(int ncr);

// T(kd x ncr) = A^T * P(rows x ncr)   (data-matrix mode first product, R/spEigen.R:136 `X %*% V0`).
int launch_contract_tr(const StreamMat& A, const double* P, int ncr, double* T, int t_pitch,
                       const ContractWorkspace& ws, int num_sms, cudaStream_t stream);
size_t contract_tr_workspace_slots(int64_t kd, int num_sms);

// Consumer side of the fused exchange: wait until every rank's epoch flag has arrived, then copy the exchange slots
// (panel rows [0, rows) and the per-rank dot sums) into their destination buffers.
int launch_exchange_wait_copy(const unsigned long long* flags, int world, unsigned long long epoch, const double* srcY,
                              double* dstY, const double* srcB, double* dstB, size_t panel_doubles, const double* src_dots,
                              double* dst_dots, int ndots, int* status, long long budget_ns, cudaStream_t stream);

}  // namespace speigen
