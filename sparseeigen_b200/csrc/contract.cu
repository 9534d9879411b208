// K1 — streamed tall-skinny contraction on sm_100a.
//
//   standard form   Y[j, :] = sum_i A[j][i] * P[i, :]      (A row-major, i contiguous)
//   transposed form T[i, :] = sum_j A[j][i] * P[j, :]
//
// A (the covariance slab or data shard) is read exactly once per launch through TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 16 doubles x TJ rows, L2 evict-first) into a
// multi-stage shared-memory ring guarded by mbarriers; the matching panel rows arrive with a 1-D
// bulk copy (L2 evict-last).  One producer warp issues the copies, eight consumer warps run
// FP64 tensor-core MMAs (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4; measured 37.0 TFLOP/s vs 33.7
// for DFMA on B200, tools/microbench/fp64_pipes.cu) straight out of the swizzled tiles.  tcgen05
// has no FP64 kind, so DMMA is the tensor path for this arithmetic type.
//
// Work decomposition is a flat stream-K split: the (row tile, k-stage) units are cut into one
// contiguous range per CTA (grid = #SMs), so every SM streams the same number of bytes no matter
// how m relates to 148.  Row tiles that straddle CTAs are finished by the last-arriving CTA, which
// sums the partial tiles in a FIXED order (by k) -> bit-reproducible results.
//
// Replaces: R/spEigen.R:177 (G), :136/:138 (X %*% V0) of the reference (base-R dgemm/zgemm).
#include "contract.cuh"
#include "tail.cuh"
#include "../../include/speigen_b200.h"
#include <cstdlib>

namespace speigen {

namespace {

struct KParams {
  const double* P;
  int p_pitch;
  int ncr;
  int64_t rows, kd;
  int n_row_tiles;
  int ksteps;
  int64_t total_units;
  int stages;
  uint32_t u_tile_bytes;
  uint32_t stage_bytes;
  double* partials;
  int* counters;
  int slot_doubles;
  EpiArgs epi;
};

constexpr uint32_t kBoxBytes = kTJ * 128;             // one TMA box: TJ rows x 16 doubles
constexpr uint32_t kAStageBytes = kNBox * kBoxBytes;  // 64 KB

__device__ __forceinline__ int64_t unit_begin(int64_t cta, int64_t total, int64_t nctas) { return cta * total / nctas; }
__device__ __forceinline__ int cta_of_unit(int64_t u, int64_t total, int64_t nctas) {
  return static_cast<int>(((u + 1) * nctas - 1) / total);
}

// ---- fused epilogue for one (row, column pair): raw store / reweighting / objective dots ------------------
__device__ __forceinline__ void epi_pair(const KParams& prm, const double* s_d, const double* s_rho, const double* s_wmax,
                                         int64_t jl, int c0, double y0, double y1, double& ds0, double& ds1) {
  const EpiArgs& ep = prm.epi;
  if (jl >= prm.rows || c0 >= prm.ncr) return;
  const size_t roff = static_cast<size_t>(ep.row_off + jl) * ep.pitch;
  const bool has1 = (c0 + 1) < prm.ncr;
  if (ep.npeers > 0 && ep.peerY[0]) {
    for (int p = 0; p < ep.npeers; ++p) {
      if (has1) *reinterpret_cast<double2*>(ep.peerY[p] + roff + c0) = make_double2(y0, y1);
      else ep.peerY[p][roff + c0] = y0;
    }
  } else if (ep.Y) {
    if (has1) *reinterpret_cast<double2*>(ep.Y + roff + c0) = make_double2(y0, y1);
    else ep.Y[roff + c0] = y0;
  }
  if (!ep.V) return;
  double v0, v1 = 0.0;
  if (has1) { const double2 vv = *reinterpret_cast<const double2*>(ep.V + roff + c0); v0 = vv.x; v1 = vv.y; }
  else v0 = ep.V[roff + c0];
  if (ep.cplx) {
    const int cc = c0 >> 1;
    if (ep.B || ep.peerB[0]) {
      const double t = weight_of_abs(hypot(v0, v1), ep.sc) - s_wmax[cc];
      const double2 bb = make_double2(reweight_elem(y0, v0, t, s_d[cc], s_rho[cc]), reweight_elem(y1, v1, t, s_d[cc], s_rho[cc]));
      if (ep.npeers > 0 && ep.peerB[0]) { for (int p = 0; p < ep.npeers; ++p) *reinterpret_cast<double2*>(ep.peerB[p] + roff + c0) = bb; }
      else *reinterpret_cast<double2*>(ep.B + roff + c0) = bb;
    }
    ds0 += v0 * y0 + v1 * y1;
  } else {
    if (ep.B || ep.peerB[0]) {
      const double t0 = weight_of_abs(fabs(v0), ep.sc) - s_wmax[c0];
      const double b0 = reweight_elem(y0, v0, t0, s_d[c0], s_rho[c0]);
      double b1 = 0.0;
      if (has1) {
        const double t1 = weight_of_abs(fabs(v1), ep.sc) - s_wmax[c0 + 1];
        b1 = reweight_elem(y1, v1, t1, s_d[c0 + 1], s_rho[c0 + 1]);
      }
      if (ep.npeers > 0 && ep.peerB[0]) {
        for (int p = 0; p < ep.npeers; ++p) {
          if (has1) *reinterpret_cast<double2*>(ep.peerB[p] + roff + c0) = make_double2(b0, b1);
          else ep.peerB[p][roff + c0] = b0;
        }
      } else {
        if (has1) *reinterpret_cast<double2*>(ep.B + roff + c0) = make_double2(b0, b1);
        else ep.B[roff + c0] = b0;
      }
    }
    ds0 += v0 * y0;
    ds1 += v1 * y1;
  }
}

constexpr int kFWarps = 4;                       // DFMA warps (when RF > 0): one per SM sub-partition
constexpr int kFMT = kTJ / (8 * kFWarps);        // m-tiles per DFMA warp = 8

// NTF = n-tiles (8 columns) on the FP64 tensor pipe (DMMA), RF = left-over columns (0 or 4) done with DFMA.
// DMMA and DFMA share one FP64 datapath on B200 (tools/microbench/fp64_concurrency.cu: 36.5 TFLOP/s combined), so
// padding q = 20 to 24 DMMA columns wastes 17 % of the binding resource.  The left-over columns are given to four
// dedicated DFMA warps (one per sub-partition) that read the same swizzled A fragments: the eight DMMA warps keep a
// pure tensor-pipe instruction stream and the DFMA warps fill the datapath whenever a DMMA warp waits on shared memory.
template <int NTF, int RF>
__global__ void __launch_bounds__((kStdWarps + (RF > 0 ? kFWarps : 0) + 1) * 32, 1)
k_contract_std(const __grid_constant__ CUtensorMap tmA, const KParams prm) {
  constexpr int NCW = kStdWarps + (RF > 0 ? kFWarps : 0);   // consumer warps
  constexpr int NTD = NTF > 0 ? NTF : 1;
  constexpr int DSLOT = kMT * NTF * 2 * 256;                     // doubles of a partial slot owned by the DMMA warps
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ int s_last;
  __shared__ double s_d[kMaxCols], s_rho[kMaxCols], s_wmax[kMaxCols];
  __shared__ double s_dots[kStdWarps + kFWarps][kMaxCols];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = prm.stages;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], NCW);
    }
    mbar_fence_init();
  }
  if (prm.epi.wait_flags) {
    // row-sharded MM loop: the operand panel's rows and the column minima arrive from the peers' tail kernels
    if (threadIdx.x < prm.epi.wait_n)
      spin_until_ge(prm.epi.wait_flags + threadIdx.x * kFlagStride, prm.epi.wait_epoch, prm.epi.spin_budget_ns, prm.epi.status,
                    prm.epi.wait_n > 1);
    __syncthreads();
    asm volatile("fence.proxy.async.global;" ::: "memory");   // peer (generic-proxy) stores -> this CTA's bulk/TMA loads
  }
  if (threadIdx.x < kMaxCols) {
    const int c = threadIdx.x;
    const bool ok = c < prm.epi.q;
    s_d[c] = (ok && prm.epi.d) ? prm.epi.d[c] : 1.0;
    s_rho[c] = (ok && prm.epi.rho) ? prm.epi.rho[c] : 0.0;
    // epi.wmax carries min_j |V[j,c]|; the column max of the weights is w(min) (w is non-increasing)   R/spEigen.R:176
    double wm = 0.0;
    if (ok && prm.epi.stat_absmin) {
      double v = 1.7976931348623157e308;
      for (int r = 0; r < prm.epi.stat_n; ++r) v = fmin(v, __ldcv(prm.epi.stat_absmin + static_cast<size_t>(r) * prm.epi.stat_stride + c));
      wm = weight_of_abs(v, prm.epi.sc);
    } else if (ok && prm.epi.wmax) {
      wm = weight_of_abs(prm.epi.wmax[c], prm.epi.sc);
    }
    s_wmax[c] = wm;
  }
  __syncthreads();

  const int64_t nctas = gridDim.x;
  const int64_t u0 = unit_begin(blockIdx.x, prm.total_units, nctas);
  const int64_t u1 = unit_begin(blockIdx.x + 1, prm.total_units, nctas);

  if (warp == NCW) {
    // ------------------------------- producer warp -------------------------------
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = u0; u < u1; ++u) {
        const int rt = static_cast<int>(u / prm.ksteps);
        const int ks = static_cast<int>(u - static_cast<int64_t>(rt) * prm.ksteps);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        unsigned char* sA = smem + static_cast<size_t>(stage) * prm.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], kAStageBytes + prm.u_tile_bytes);
#pragma unroll
        for (int b = 0; b < kNBox; ++b)
          tma_load_2d(sA + b * kBoxBytes, &tmA, ks * kTK + b * 16, rt * kTJ, &full_bar[stage], pol_stream);
        bulk_load_1d(sA + kAStageBytes, prm.P + static_cast<size_t>(ks) * kTK * prm.p_pitch, prm.u_tile_bytes,
                     &full_bar[stage], pol_keep);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // --------------------------------- consumer warps ---------------------------------
  const int ctid = threadIdx.x;                  // 0 .. NCW*32-1
  const int grp = lane >> 2;                     // fragment row / B column within the n-tile
  const int kk = lane & 3;                       // fragment k index
  const int prow = 2 * (grp & 3) + (grp >> 2);   // row permutation -> conflict-free swizzled LDS.64
  const int pitch = prm.p_pitch;
  const bool want_dots = prm.epi.dots != nullptr;
  const int nout = prm.epi.cplx ? prm.epi.q : prm.ncr;

  // segment-end protocol shared by both warp kinds: returns true if this CTA finishes the row tile
  auto arrive_segment = [&](int rt, int c_first, int c_last) -> bool {
    __threadfence();
    named_bar_sync(1, NCW * 32);
    if (ctid == 0) {
      const int old = atomicAdd(&prm.counters[rt], 1);
      const int last = (old == c_last - c_first);
      if (last) prm.counters[rt] = 0;   // self-reset for the next launch
      s_last = last;
    }
    named_bar_sync(1, NCW * 32);
    const bool fin = (s_last != 0);
    if (fin) __threadfence();
    return fin;
  };
  // final fixed-order sum of the per-warp column dots (called by all consumer threads)
  auto finish_dots = [&](int rt) {
    named_bar_sync(1, NCW * 32);
    if (ctid < nout) {
      const int col = prm.epi.cplx ? 2 * ctid : ctid;
      double s = 0.0;
      if (col < NTF * 8) {
#pragma unroll
        for (int w = 0; w < kStdWarps; ++w) s += s_dots[w][col];
      } else {
#pragma unroll
        for (int w = 0; w < kFWarps; ++w) s += s_dots[kStdWarps + w][col];
      }
      prm.epi.dots[static_cast<size_t>(rt) * nout + ctid] = s;
    }
    named_bar_sync(1, NCW * 32);
  };

  // fused exchange: after this CTA's last store, the CTA that finishes last (grid-wide) publishes the rank's dot sums
  // and raises the epoch flag on every peer.  fence.sys + counter make all peer stores of all CTAs visible first.
  auto finish_exchange = [&]() {
    if (prm.epi.npeers == 0) return;
    __threadfence_system();
    named_bar_sync(1, NCW * 32);
    if (ctid == 0) {
      const int old = atomicAdd(prm.epi.done_count, 1);
      const int last = (old == static_cast<int>(gridDim.x) - 1);
      if (last) *prm.epi.done_count = 0;
      s_last = last;
    }
    named_bar_sync(1, NCW * 32);
    if (!s_last) return;
    __threadfence_system();
    if (want_dots && ctid < nout) {          // this rank's column dots: fixed-order sum over its row tiles
      double s = 0.0;
      for (int t = 0; t < prm.n_row_tiles; ++t) s += __ldcg(prm.epi.dots + static_cast<size_t>(t) * nout + ctid);
      for (int p = 0; p < prm.epi.npeers; ++p) prm.epi.peer_dots[p][prm.epi.my_shard * nout + ctid] = s;
    }
    __threadfence_system();
    named_bar_sync(1, NCW * 32);
    if (ctid < prm.epi.npeers) {
      __threadfence_system();
      st_release_sys_u64(prm.epi.peer_flag[ctid] + prm.epi.my_shard, prm.epi.epoch);
    }
  };

  if (RF > 0 && warp >= kStdWarps) {
    // ================================ DFMA warps ================================
    const int fw = warp - kStdWarps;        // 0..3, rows [64 fw, 64 fw + 64)
    const int ftid = ctid - kStdWarps * 32; // 0..127
    double facc[kFMT][RF > 0 ? RF : 1];
#pragma unroll
    for (int mt = 0; mt < kFMT; ++mt)
#pragma unroll
      for (int c = 0; c < RF; ++c) facc[mt][c] = 0.0;
    uint32_t a_row_off[kFMT];
    int a_row7[kFMT];
#pragma unroll
    for (int mt = 0; mt < kFMT; ++mt) {
      const int jl = fw * (8 * kFMT) + mt * 8 + prow;
      a_row_off[mt] = jl * 128 + (kk & 1) * 8;
      a_row7[mt] = jl & 7;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t u = u0; u < u1; ++u) {
      const int rt = static_cast<int>(u / prm.ksteps);
      const int ks = static_cast<int>(u - static_cast<int64_t>(rt) * prm.ksteps);
      mbar_wait(&full_bar[stage], phase);
      const unsigned char* sA = smem + static_cast<size_t>(stage) * prm.stage_bytes;
      const double* sU = reinterpret_cast<const double*>(sA + kAStageBytes);
#pragma unroll
      for (int b = 0; b < kNBox; ++b) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int il = b * 16 + k4 * 4 + kk;
          const int chunk = (k4 * 4 + kk) >> 1;
          double uf[RF > 0 ? RF : 1];
#pragma unroll
          for (int c = 0; c < RF; c += 2) {
            const double2 t = *reinterpret_cast<const double2*>(sU + il * pitch + NTF * 8 + c);
            uf[c] = t.x;
            uf[c + 1] = t.y;
          }
#pragma unroll
          for (int mt = 0; mt < kFMT; ++mt) {
            const double af = *reinterpret_cast<const double*>(sA + b * kBoxBytes + a_row_off[mt] + ((chunk ^ a_row7[mt]) << 4));
#pragma unroll
            for (int c = 0; c < RF; ++c) facc[mt][c] = fma(af, uf[c], facc[mt][c]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1u; }
      const bool seg_end = (ks == prm.ksteps - 1) || (u == u1 - 1);
      if (!seg_end) continue;

      // quad butterfly over the 4 k-lanes (fixed order); lane kk keeps columns 8 NTF + 2kk, +1
      double y[kFMT][2];
#pragma unroll
      for (int mt = 0; mt < kFMT; ++mt) {
        double t[RF > 0 ? RF : 1];
#pragma unroll
        for (int c = 0; c < RF; ++c) {
          double v = facc[mt][c];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          t[c] = v;
          facc[mt][c] = 0.0;
        }
        y[mt][0] = y[mt][1] = 0.0;
#pragma unroll
        for (int c = 0; c < RF; c += 2)
          if (kk == (c >> 1)) { y[mt][0] = t[c]; y[mt][1] = t[c + 1]; }
      }
      const int64_t t_first = static_cast<int64_t>(rt) * prm.ksteps;
      const int c_first = cta_of_unit(t_first, prm.total_units, nctas);
      const int c_last = cta_of_unit(t_first + prm.ksteps - 1, prm.total_units, nctas);
      bool do_epi = true;
      if (c_last > c_first) {
        double* slot = prm.partials + static_cast<size_t>(rt + blockIdx.x) * prm.slot_doubles + DSLOT;
#pragma unroll
        for (int mt = 0; mt < kFMT; ++mt) {
          __stcg(slot + (mt * 2 + 0) * 128 + ftid, y[mt][0]);
          __stcg(slot + (mt * 2 + 1) * 128 + ftid, y[mt][1]);
        }
        do_epi = arrive_segment(rt, c_first, c_last);
        if (do_epi) {
#pragma unroll
          for (int mt = 0; mt < kFMT; ++mt) y[mt][0] = y[mt][1] = 0.0;
          for (int cc = c_first; cc <= c_last; ++cc) {   // fixed order (by k) => deterministic sums
            const double* ps = prm.partials + static_cast<size_t>(rt + cc) * prm.slot_doubles + DSLOT;
#pragma unroll
            for (int mt = 0; mt < kFMT; ++mt) {
              y[mt][0] += __ldcg(ps + (mt * 2 + 0) * 128 + ftid);
              y[mt][1] += __ldcg(ps + (mt * 2 + 1) * 128 + ftid);
            }
          }
        }
      }
      if (do_epi) {
        double ds0 = 0.0, ds1 = 0.0;
        const int c0 = NTF * 8 + 2 * kk;
        if (kk < RF / 2) {
#pragma unroll
          for (int mt = 0; mt < kFMT; ++mt) {
            const int64_t jl = static_cast<int64_t>(rt) * kTJ + fw * (8 * kFMT) + mt * 8 + prow;
            epi_pair(prm, s_d, s_rho, s_wmax, jl, c0, y[mt][0], y[mt][1], ds0, ds1);
          }
        }
        if (want_dots) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = e ? ds1 : ds0;
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (grp == 0 && kk < RF / 2) s_dots[warp][c0 + e] = v;
          }
          finish_dots(rt);
        }
      }
    }
    finish_exchange();
    return;
  }

  // ================================ DMMA warps ================================
  double acc[kMT][NTD][2];
#pragma unroll
  for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  uint32_t a_row_off[kMT];
  int a_row7[kMT];
#pragma unroll
  for (int mt = 0; mt < kMT; ++mt) {
    const int jl = warp * (8 * kMT) + mt * 8 + prow;
    a_row_off[mt] = jl * 128 + (kk & 1) * 8;
    a_row7[mt] = jl & 7;
  }

  int stage = 0;
  uint32_t phase = 0;
  for (int64_t u = u0; u < u1; ++u) {
    const int rt = static_cast<int>(u / prm.ksteps);
    const int ks = static_cast<int>(u - static_cast<int64_t>(rt) * prm.ksteps);
    mbar_wait(&full_bar[stage], phase);
    if (NTF > 0) {
      const unsigned char* sA = smem + static_cast<size_t>(stage) * prm.stage_bytes;
      const double* sU = reinterpret_cast<const double*>(sA + kAStageBytes);
#pragma unroll
      for (int b = 0; b < kNBox; ++b) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int il = b * 16 + k4 * 4 + kk;         // row of the staged panel tile
          const int chunk = (k4 * 4 + kk) >> 1;        // 16-byte chunk inside the 128-byte box row
          double bf[NTD];
#pragma unroll
          for (int nt = 0; nt < NTF; ++nt) bf[nt] = sU[il * pitch + nt * 8 + grp];
          double af[kMT];
#pragma unroll
          for (int mt = 0; mt < kMT; ++mt)
            af[mt] = *reinterpret_cast<const double*>(sA + b * kBoxBytes + a_row_off[mt] + ((chunk ^ a_row7[mt]) << 4));
#pragma unroll
          for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTF; ++nt) dmma_884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == stages) { stage = 0; phase ^= 1u; }

    const bool seg_end = (ks == prm.ksteps - 1) || (u == u1 - 1);
    if (!seg_end) continue;

    // ---------------- end of this CTA's segment of row tile rt ----------------
    const int64_t t_first = static_cast<int64_t>(rt) * prm.ksteps;
    const int c_first = cta_of_unit(t_first, prm.total_units, nctas);
    const int c_last = cta_of_unit(t_first + prm.ksteps - 1, prm.total_units, nctas);
    bool do_epi = true;
    if (c_last > c_first) {
      double* slot = prm.partials + static_cast<size_t>(rt + blockIdx.x) * prm.slot_doubles;
#pragma unroll
      for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTF; ++nt) {
          __stcg(slot + ((mt * NTF + nt) * 2 + 0) * 256 + ctid, acc[mt][nt][0]);
          __stcg(slot + ((mt * NTF + nt) * 2 + 1) * 256 + ctid, acc[mt][nt][1]);
        }
      do_epi = arrive_segment(rt, c_first, c_last);
      if (do_epi) {
#pragma unroll
        for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTF; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        for (int cc = c_first; cc <= c_last; ++cc) {   // fixed order (by k) => deterministic sums
          const double* ps = prm.partials + static_cast<size_t>(rt + cc) * prm.slot_doubles;
#pragma unroll
          for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTF; ++nt) {
              acc[mt][nt][0] += __ldcg(ps + ((mt * NTF + nt) * 2 + 0) * 256 + ctid);
              acc[mt][nt][1] += __ldcg(ps + ((mt * NTF + nt) * 2 + 1) * 256 + ctid);
            }
        }
      }
    }
    if (do_epi) {
      double dsum[NTD][2];
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) dsum[nt][0] = dsum[nt][1] = 0.0;
#pragma unroll
      for (int mt = 0; mt < kMT; ++mt) {
        const int64_t jl = static_cast<int64_t>(rt) * kTJ + warp * (8 * kMT) + mt * 8 + prow;
#pragma unroll
        for (int nt = 0; nt < NTF; ++nt)
          epi_pair(prm, s_d, s_rho, s_wmax, jl, nt * 8 + 2 * kk, acc[mt][nt][0], acc[mt][nt][1], dsum[nt][0], dsum[nt][1]);
      }
      if (want_dots) {
        // rows: 8 fragment groups per warp (shuffle), then the warps (shared memory), both in fixed order
#pragma unroll
        for (int nt = 0; nt < NTF; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double v = dsum[nt][e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (grp == 0) s_dots[warp][nt * 8 + 2 * kk + e] = v;
          }
        finish_dots(rt);
      }
    }
#pragma unroll
    for (int mt = 0; mt < kMT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  }
  finish_exchange();
}

// consumer side of the fused exchange
__global__ void __launch_bounds__(256) k_exchange_wait_copy(const unsigned long long* flags, int world, unsigned long long epoch,
                                                            const double* srcY, double* dstY, const double* srcB, double* dstB,
                                                            size_t panel_doubles, const double* src_dots, double* dst_dots,
                                                            int ndots, int* status, long long budget_ns) {
  if (threadIdx.x < world) spin_until_ge(flags + threadIdx.x, epoch, budget_ns, status, true);
  __syncthreads();
  __threadfence_system();
  const size_t n2 = panel_doubles / 2;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2; i += stride) {
    if (srcY) reinterpret_cast<double2*>(dstY)[i] = __ldcv(reinterpret_cast<const double2*>(srcY) + i);
    if (srcB) reinterpret_cast<double2*>(dstB)[i] = __ldcv(reinterpret_cast<const double2*>(srcB) + i);
  }
  if (src_dots && blockIdx.x == 0)
    for (int i = threadIdx.x; i < ndots; i += blockDim.x) dst_dots[i] = __ldcv(src_dots + i);
}

// ------------------------------------------------------------------------------------------------
// Transposed form: T[i, :] = sum_j A[j][i] * P[j, :].  CTA tile = 128 outputs (i) x 64 contracted rows
// (j) per stage (8 boxes of 16 doubles x 64 rows); warp w owns outputs [16w, 16w+16).
// ------------------------------------------------------------------------------------------------
constexpr int kTrRows = 64;                         // contracted rows (j) per stage
constexpr int kTrBoxes = 8;                         // 16-double boxes per stage -> 128 outputs per CTA tile
constexpr int kTrOut = 16 * kTrBoxes;               // 128
constexpr uint32_t kTrBoxBytes = kTrRows * 128;     // 8 KB
constexpr uint32_t kTrAStageBytes = kTrBoxes * kTrBoxBytes;   // 64 KB

struct KParamsTr {
  const double* P;
  int p_pitch;
  int ncr;
  int64_t rows, kd;          // A is rows x kd ; contraction over rows, outputs over kd
  int n_out_tiles;           // ceil(kd / 128)
  int ksteps;                // ceil(rows / 64)
  int64_t total_units;
  int stages;
  uint32_t u_tile_bytes;
  uint32_t stage_bytes;
  double* partials;
  int* counters;
  int slot_doubles;
  double* T;
  int t_pitch;
};

template <int NT>
__global__ void __launch_bounds__(kThreads, 1) k_contract_tr(const __grid_constant__ CUtensorMap tmA, const KParamsTr prm) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ int s_last;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = prm.stages;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int64_t nctas = gridDim.x;
  const int64_t u0 = unit_begin(blockIdx.x, prm.total_units, nctas);
  const int64_t u1 = unit_begin(blockIdx.x + 1, prm.total_units, nctas);

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = u0; u < u1; ++u) {
        const int ot = static_cast<int>(u / prm.ksteps);
        const int ks = static_cast<int>(u - static_cast<int64_t>(ot) * prm.ksteps);
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        unsigned char* sA = smem + static_cast<size_t>(stage) * prm.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], kTrAStageBytes + prm.u_tile_bytes);
#pragma unroll
        for (int b = 0; b < kTrBoxes; ++b)
          tma_load_2d(sA + b * kTrBoxBytes, &tmA, ot * kTrOut + b * 16, ks * kTrRows, &full_bar[stage], pol_stream);
        bulk_load_1d(sA + kTrAStageBytes, prm.P + static_cast<size_t>(ks) * kTrRows * prm.p_pitch, prm.u_tile_bytes,
                     &full_bar[stage], pol_keep);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  const int ctid = threadIdx.x;
  const int grp = lane >> 2;
  const int kk = lane & 3;
  const int pitch = prm.p_pitch;
  // Output (i) index inside the warp's 16-wide box for fragment row `grp` of m-tile mt:
  //   i = (grp&1) + 8*((grp>>1)&1) + 2*(grp>>2) + 4*mt   -> conflict-free LDS.64 on the swizzled box.
  int i_loc[2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) i_loc[mt] = (grp & 1) + 8 * ((grp >> 1) & 1) + 2 * (grp >> 2) + 4 * mt;

  double acc[2][NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  int stage = 0;
  uint32_t phase = 0;
  for (int64_t u = u0; u < u1; ++u) {
    const int ot = static_cast<int>(u / prm.ksteps);
    const int ks = static_cast<int>(u - static_cast<int64_t>(ot) * prm.ksteps);
    mbar_wait(&full_bar[stage], phase);
    const unsigned char* sA = smem + static_cast<size_t>(stage) * prm.stage_bytes + warp * kTrBoxBytes;
    const double* sU = reinterpret_cast<const double*>(smem + static_cast<size_t>(stage) * prm.stage_bytes + kTrAStageBytes);
#pragma unroll 4
    for (int k4 = 0; k4 < kTrRows / 4; ++k4) {
      const int jl = k4 * 4 + kk;   // contracted row inside the stage
      double bf[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bf[nt] = sU[jl * pitch + nt * 8 + grp];
      double af[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int il = i_loc[mt];
        af[mt] = *reinterpret_cast<const double*>(sA + jl * 128 + ((((il >> 1) ^ (jl & 7))) << 4) + (il & 1) * 8);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma_884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == stages) { stage = 0; phase ^= 1u; }

    const bool seg_end = (ks == prm.ksteps - 1) || (u == u1 - 1);
    if (!seg_end) continue;

    const int64_t t_first = static_cast<int64_t>(ot) * prm.ksteps;
    const int c_first = cta_of_unit(t_first, prm.total_units, nctas);
    const int c_last = cta_of_unit(t_first + prm.ksteps - 1, prm.total_units, nctas);
    bool do_epi = true;
    if (c_last > c_first) {
      double* slot = prm.partials + static_cast<size_t>(ot + blockIdx.x) * prm.slot_doubles;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          __stcg(slot + ((mt * NT + nt) * 2 + 0) * 256 + ctid, acc[mt][nt][0]);
          __stcg(slot + ((mt * NT + nt) * 2 + 1) * 256 + ctid, acc[mt][nt][1]);
        }
      __threadfence();
      named_bar_sync(1, kConsumerWarps * 32);
      if (ctid == 0) {
        const int old = atomicAdd(&prm.counters[ot], 1);
        const int last = (old == c_last - c_first);
        if (last) prm.counters[ot] = 0;
        s_last = last;
      }
      named_bar_sync(1, kConsumerWarps * 32);
      do_epi = (s_last != 0);
      if (do_epi) {
        __threadfence();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        for (int cc = c_first; cc <= c_last; ++cc) {
          const double* ps = prm.partials + static_cast<size_t>(ot + cc) * prm.slot_doubles;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
              acc[mt][nt][0] += __ldcg(ps + ((mt * NT + nt) * 2 + 0) * 256 + ctid);
              acc[mt][nt][1] += __ldcg(ps + ((mt * NT + nt) * 2 + 1) * 256 + ctid);
            }
        }
      }
    }
    if (do_epi) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int64_t i = static_cast<int64_t>(ot) * kTrOut + warp * 16 + i_loc[mt];
        if (i < prm.kd) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int c0 = nt * 8 + 2 * kk;
            if (c0 >= prm.ncr) continue;
            double* dst = prm.T + static_cast<size_t>(i) * prm.t_pitch + c0;
            if (c0 + 1 < prm.ncr) *reinterpret_cast<double2*>(dst) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            else *dst = acc[mt][nt][0];
          }
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Default: left-over columns are padded onto the tensor pipe (all-DMMA).  SPEIGEN_HYBRID_DFMA=1 selects the
// <NTF, RF=4> variant with dedicated DFMA warps instead.  Measured on B200 (tools/k1_probe.py ab, same process, same thermal
// state, m=50k q=20): all-DMMA 4.24 ms vs hybrid 4.54 ms per launch in the sustained, power-capped regime - DFMA
// costs more energy per flop than DMMA, so doing 17 % fewer flops on the same datapath does not pay.  Read per launch
// so a harness can A/B inside one process.
inline bool prm_force_dmma() {
  const char* e = getenv("SPEIGEN_HYBRID_DFMA");
  return !(e && e[0] == '1');
}

constexpr size_t kMaxDynSmem = 232448 - 10240;   // 227 KB per CTA minus (generous) static shared memory

inline uint32_t round_up_u32(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

template <int NTF, int RF>
int launch_std_nt(const StreamMat& A, const KParams& prm, int grid, size_t smem, cudaStream_t stream) {
  static bool attr_set_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];   // function attributes are per device
  if (!attr_set) {
    SPE_CUDA(cudaFuncSetAttribute(k_contract_std<NTF, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    attr_set = true;
  }
  k_contract_std<NTF, RF><<<grid, (kStdWarps + (RF > 0 ? kFWarps : 0) + 1) * 32, smem, stream>>>(A.tmap, prm);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
template <int NT>
int launch_tr_nt(const StreamMat& A, const KParamsTr& prm, int grid, size_t smem, cudaStream_t stream) {
  static bool attr_set_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool& attr_set = attr_set_dev[dev_ & 63];   // function attributes are per device
  if (!attr_set) {
    SPE_CUDA(cudaFuncSetAttribute(k_contract_tr<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
    attr_set = true;
  }
  k_contract_tr<NT><<<grid, kThreads, smem, stream>>>(A.tmap, prm);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

}  // namespace

int make_stream_mat(StreamMat* sm, const double* ptr, int64_t rows, int64_t kd, int64_t ld, int box_rows) {
  if ((ld % 16) != 0 || (reinterpret_cast<uintptr_t>(ptr) % 128) != 0) {
    set_error("stream matrix must be 128-byte aligned with a pitch that is a multiple of 16 doubles");
    return SPEIGEN_ERR_ARG;
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SPEIGEN_ERR_CUDA;
  }
  sm->ptr = ptr;
  sm->rows = rows;
  sm->kd = kd;
  sm->ld = ld;
  sm->box_rows = box_rows;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kd), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * sizeof(double)};
  cuuint32_t box[2] = {16u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(&sm->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld kd=%lld ld=%lld)", (int)r, (long long)rows,
              (long long)kd, (long long)ld);
    return SPEIGEN_ERR_CUDA;
  }
  return SPEIGEN_OK;
}

size_t contract_slot_doubles(int ncr) {
  const int nt = (ncr + 7) / 8;   // >= NTF (+1 covers the DFMA warps' 8 x 2 x 128 values)
  return static_cast<size_t>(kMT) * nt * 2 * 256;
}
size_t contract_workspace_slots(int64_t rows, int num_sms) {
  return static_cast<size_t>((rows + kTJ - 1) / kTJ) + num_sms;
}
size_t contract_tr_workspace_slots(int64_t kd, int num_sms) {
  return static_cast<size_t>((kd + kTrOut - 1) / kTrOut) + num_sms;
}

int launch_contract_std(const StreamMat& A, const double* P, int ncr, const EpiArgs& epi, const ContractWorkspace& ws,
                        int num_sms, cudaStream_t stream) {
  if (ncr < 1 || ncr > kMaxCols) {
    set_error("contract: %d real columns unsupported (1..%d)", ncr, kMaxCols);
    return SPEIGEN_ERR_ARG;
  }
  if (A.box_rows != kTJ) {
    set_error("contract: stream matrix tensor map built for box_rows=%d, need %d", A.box_rows, kTJ);
    return SPEIGEN_ERR_ARG;
  }
  KParams prm{};
  prm.P = P;
  prm.ncr = ncr;
  prm.p_pitch = panel_pitch(ncr);
  prm.rows = A.rows;
  prm.kd = A.kd;
  prm.n_row_tiles = static_cast<int>((A.rows + kTJ - 1) / kTJ);
  prm.ksteps = static_cast<int>((A.kd + kTK - 1) / kTK);
  prm.total_units = static_cast<int64_t>(prm.n_row_tiles) * prm.ksteps;
  prm.u_tile_bytes = static_cast<uint32_t>(kTK * prm.p_pitch * sizeof(double));
  prm.stage_bytes = kAStageBytes + round_up_u32(prm.u_tile_bytes + 512, 1024);
  prm.stages = static_cast<int>(kMaxDynSmem / prm.stage_bytes);
  if (prm.stages > 6) prm.stages = 6;
  if (prm.stages < 2) {
    set_error("contract: not enough shared memory for a 2-stage pipeline");
    return SPEIGEN_ERR_ARG;
  }
  prm.partials = ws.partials;
  prm.counters = ws.counters;
  prm.slot_doubles = static_cast<int>(contract_slot_doubles(ncr));
  prm.epi = epi;
  int grid = num_sms;
  if (prm.total_units < grid) grid = static_cast<int>(prm.total_units);
  if (grid < 1) return SPEIGEN_OK;
  if (static_cast<size_t>(prm.n_row_tiles) + grid > ws.partial_slots + 0 || prm.n_row_tiles > ws.n_counters) {
    set_error("contract: workspace too small (%d tiles, %zu slots)", prm.n_row_tiles, ws.partial_slots);
    return SPEIGEN_ERR_ARG;
  }
  const size_t smem = static_cast<size_t>(prm.stages) * prm.stage_bytes;
  // columns: ntf full DMMA n-tiles + (remainder 1..4 -> 4 DFMA columns | remainder 5..7 -> one more padded DMMA tile)
  int ntf = ncr / 8;
  const int rem = ncr % 8;
  int rf = 0;
  if (rem >= 1 && rem <= 4 && !prm_force_dmma()) rf = 4;
  else if (rem > 0) ntf += 1;
  switch (ntf * 8 + rf) {
    case 0 * 8 + 4: return launch_std_nt<0, 4>(A, prm, grid, smem, stream);
    case 1 * 8 + 0: return launch_std_nt<1, 0>(A, prm, grid, smem, stream);
    case 1 * 8 + 4: return launch_std_nt<1, 4>(A, prm, grid, smem, stream);
    case 2 * 8 + 0: return launch_std_nt<2, 0>(A, prm, grid, smem, stream);
    case 2 * 8 + 4: return launch_std_nt<2, 4>(A, prm, grid, smem, stream);
    case 3 * 8 + 0: return launch_std_nt<3, 0>(A, prm, grid, smem, stream);
    case 3 * 8 + 4: return launch_std_nt<3, 4>(A, prm, grid, smem, stream);
    case 4 * 8 + 0: return launch_std_nt<4, 0>(A, prm, grid, smem, stream);
    case 4 * 8 + 4: return launch_std_nt<4, 4>(A, prm, grid, smem, stream);
    case 5 * 8 + 0: return launch_std_nt<5, 0>(A, prm, grid, smem, stream);
    case 5 * 8 + 4: return launch_std_nt<5, 4>(A, prm, grid, smem, stream);
    case 6 * 8 + 0: return launch_std_nt<6, 0>(A, prm, grid, smem, stream);
    case 6 * 8 + 4: return launch_std_nt<6, 4>(A, prm, grid, smem, stream);
    case 7 * 8 + 0: return launch_std_nt<7, 0>(A, prm, grid, smem, stream);
    case 7 * 8 + 4: return launch_std_nt<7, 4>(A, prm, grid, smem, stream);
    default: return launch_std_nt<8, 0>(A, prm, grid, smem, stream);
  }
}

int launch_exchange_wait_copy(const unsigned long long* flags, int world, unsigned long long epoch, const double* srcY,
                              double* dstY, const double* srcB, double* dstB, size_t panel_doubles, const double* src_dots,
                              double* dst_dots, int ndots, int* status, long long budget_ns, cudaStream_t stream) {
  k_exchange_wait_copy<<<148, 256, 0, stream>>>(flags, world, epoch, srcY, dstY, srcB, dstB, panel_doubles, src_dots, dst_dots, ndots,
                                                status, budget_ns);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_contract_tr(const StreamMat& A, const double* P, int ncr, double* T, int t_pitch, const ContractWorkspace& ws,
                       int num_sms, cudaStream_t stream) {
  if (ncr < 1 || ncr > kMaxCols) {
    set_error("contract_tr: %d real columns unsupported (1..%d)", ncr, kMaxCols);
    return SPEIGEN_ERR_ARG;
  }
  if (A.box_rows != kTrRows) {
    set_error("contract_tr: stream matrix tensor map built for box_rows=%d, need %d", A.box_rows, kTrRows);
    return SPEIGEN_ERR_ARG;
  }
  KParamsTr prm{};
  prm.P = P;
  prm.ncr = ncr;
  prm.p_pitch = panel_pitch(ncr);
  prm.rows = A.rows;
  prm.kd = A.kd;
  prm.n_out_tiles = static_cast<int>((A.kd + kTrOut - 1) / kTrOut);
  prm.ksteps = static_cast<int>((A.rows + kTrRows - 1) / kTrRows);
  prm.total_units = static_cast<int64_t>(prm.n_out_tiles) * prm.ksteps;
  prm.u_tile_bytes = static_cast<uint32_t>(kTrRows * prm.p_pitch * sizeof(double));
  prm.stage_bytes = kTrAStageBytes + round_up_u32(prm.u_tile_bytes + 512, 1024);
  prm.stages = static_cast<int>(kMaxDynSmem / prm.stage_bytes);
  if (prm.stages > 6) prm.stages = 6;
  if (prm.stages < 2) {
    set_error("contract_tr: not enough shared memory for a 2-stage pipeline");
    return SPEIGEN_ERR_ARG;
  }
  prm.partials = ws.partials;
  prm.counters = ws.counters;
  prm.slot_doubles = 2 * ((ncr + 7) / 8) * 2 * 256;
  prm.T = T;
  prm.t_pitch = t_pitch;
  int grid = num_sms;
  if (prm.total_units < grid) grid = static_cast<int>(prm.total_units);
  if (grid < 1) return SPEIGEN_OK;
  if (static_cast<size_t>(prm.n_out_tiles) + grid > ws.partial_slots || prm.n_out_tiles > ws.n_counters) {
    set_error("contract_tr: workspace too small");
    return SPEIGEN_ERR_ARG;
  }
  const size_t smem = static_cast<size_t>(prm.stages) * prm.stage_bytes;
  const int nt = (ncr + 7) / 8;
  switch (nt) {
    case 1: return launch_tr_nt<1>(A, prm, grid, smem, stream);
    case 2: return launch_tr_nt<2>(A, prm, grid, smem, stream);
    case 3: return launch_tr_nt<3>(A, prm, grid, smem, stream);
    case 4: return launch_tr_nt<4>(A, prm, grid, smem, stream);
    case 5: return launch_tr_nt<5>(A, prm, grid, smem, stream);
    case 6: return launch_tr_nt<6>(A, prm, grid, smem, stream);
    case 7: return launch_tr_nt<7>(A, prm, grid, smem, stream);
    default: return launch_tr_nt<8>(A, prm, grid, smem, stream);
  }
}

}  // namespace speigen
