// Fused, row-sharded polar tail (see tail.cuh).  sm_100a; FP64 tensor-core MMAs (DMMA.8x8x4) for the Gram matrices and the
// panel x small-matrix products, fragments read straight from the L2-resident panel; one persistent co-resident grid.
#include "tail.cuh"
#include "../../include/speigen_b200.h"
#include <algorithm>
#include <cfloat>
#include <cstdlib>

namespace speigen {
namespace {

constexpr int TT = kTailThreads;
constexpr int TW = TT / 32;

struct SmemPlan {
  int n, kt4, pm, nb;
  size_t ms, a0, u1, lam, tri, total;   // offsets in doubles (tri: offset of the uint32 table), total bytes
};
__host__ __device__ inline SmemPlan plan_smem(int n, int nt) {
  SmemPlan s;
  s.n = n;
  s.kt4 = (n + 3) / 4 * 4;
  s.pm = panel_pitch(nt * 8);
  s.nb = nt < 4 ? nt : 4;
  size_t off = 0;
  s.ms = off; off += static_cast<size_t>(s.kt4) * s.pm;
  s.a0 = off; off += static_cast<size_t>(n) * n;
  size_t u1 = 3 * static_cast<size_t>(n) * n;
  const size_t red = static_cast<size_t>(TW / 2) * s.nb * s.nb * 64;
  const size_t st = static_cast<size_t>(TW) * (2 * kMaxCols + 2);
  const size_t tiles = static_cast<size_t>(TW) * 8 * s.pm;     // per-warp staging of the fused apply + Gram pass
  if (red > u1) u1 = red;
  if (st > u1) u1 = st;
  if (tiles > u1) u1 = tiles;
  s.u1 = off; off += u1;
  s.lam = off; off += kMaxCols + 8;
  s.tri = off; off += (static_cast<size_t>(n) * (n + 1) / 2 + 1) / 2 + 1;
  s.total = off * sizeof(double);
  return s;
}

__device__ __forceinline__ void tri_decode(const uint32_t* tri, int t, int& i, int& j) {
  const uint32_t v = tri[t];
  i = static_cast<int>(v >> 16);
  j = static_cast<int>(v & 0xffffu);
}

// ---- grid-wide "last CTA" ticket: returns true in the CTA that arrives last (all earlier global writes of every CTA are
// visible to it).  Self-resetting.
__device__ __forceinline__ bool grid_ticket_last(int* ticket, int* s_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                 // release: cumulative over the CTA's writes ordered before the barrier
    const int t = atomicAdd(ticket, 1);
    const int last = (t == static_cast<int>(gridDim.x) - 1);
    if (last) *ticket = 0;
    __threadfence();                 // acquire side for the last CTA
    *s_flag = last;
  }
  __syncthreads();
  return *s_flag != 0;
}

// fixed-order binary tree over the 16 warps ((w, w + half) pairs) of per-warp DMMA accumulator blocks, then warp 0 writes the
// NB x NB tile block at (tm0, tn0) of the symmetric n x n matrix G (mirrored when off-diagonal).  Ends with a CTA barrier.
template <int NB>
__device__ __forceinline__ void gram_tree_store(double (&acc)[NB][NB][2], int tm0, int tn0, int n, double* G, double* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, kk = lane & 3;
#pragma unroll 1
  for (int half = TW / 2; half >= 1; half >>= 1) {
    if (warp >= half && warp < 2 * half) {
      double* dst = red + static_cast<size_t>(warp - half) * NB * NB * 64;
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          dst[(i * NB + j) * 64 + lane * 2] = acc[i][j][0];
          dst[(i * NB + j) * 64 + lane * 2 + 1] = acc[i][j][1];
        }
    }
    __syncthreads();
    if (warp < half) {
      const double* src = red + static_cast<size_t>(warp) * NB * NB * 64;
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          acc[i][j][0] += src[(i * NB + j) * 64 + lane * 2];
          acc[i][j][1] += src[(i * NB + j) * 64 + lane * 2 + 1];
        }
    }
    __syncthreads();
  }
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = (tm0 + i) * 8 + grp, c = (tn0 + j) * 8 + 2 * kk + e;
          if (a < n && c < n) {
            G[a * n + c] = acc[i][j][e];
            if (tm0 != tn0) G[c * n + a] = acc[i][j][e];
          }
        }
  }
  __syncthreads();
}

// ---- Gram matrix of the CTA's rows of T (real view, n = ncr columns): full symmetric n x n result in smem `G`.
// Blocked over NB x NB n-tiles (accumulators in registers), DMMA, fixed-order tree reduction over the 16 warps.
template <int NB>
__device__ void cta_gram(const double* __restrict__ T, int pitch, int n, int64_t r0, int64_t r1, double* G, double* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, kk = lane & 3;
  const int ntiles = (n + 7) / 8;
  const int64_t nchunks = (r1 - r0 + 3) / 4;
  for (int tm0 = 0; tm0 < ntiles; tm0 += NB) {
    for (int tn0 = tm0; tn0 < ntiles; tn0 += NB) {
      double acc[NB][NB][2];
#pragma unroll
      for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      // chunks of 4 rows, 4 chunks in flight per warp (the panel rows come from L2: latency, not bandwidth, is the cost)
      for (int64_t ch0 = warp; ch0 < nchunks; ch0 += 4 * TW) {
        double af[4][NB], bf[4][NB];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t ch = ch0 + static_cast<int64_t>(u) * TW;
          const int64_t row = r0 + ch * 4 + kk;
          const bool rok = ch < nchunks && row < r1;
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int ca = (tm0 + i) * 8 + grp, cb = (tn0 + i) * 8 + grp;
            af[u][i] = (rok && ca < n) ? __ldcg(T + row * pitch + ca) : 0.0;
            bf[u][i] = (tm0 == tn0) ? af[u][i] : ((rok && cb < n) ? __ldcg(T + row * pitch + cb) : 0.0);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = 0; j < NB; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[u][i], bf[u][j]);
      }
      gram_tree_store<NB>(acc, tm0, tn0, n, G, red);
    }
  }
}

// ---- exchange of the Gram partials: CTA partial -> last CTA sums (fixed order) -> every rank's mailbox -> wait for all ranks
// -> A0 = sum over ranks (rank order), symmetric; complex panels: the Hermitian Gram matrix in its real "hat" embedding.
template <bool CPLX>
__device__ void exchange_gram(const TailArgs& p, const SmemPlan& sp, double* sm, unsigned long long epoch, int* s_flag) {
  const int n = sp.n, ntri = n * (n + 1) / 2;
  double* A0 = sm + sp.a0;
  double* A1 = sm + sp.u1;
  const uint32_t* tri = reinterpret_cast<const uint32_t*>(sm + sp.tri);
  const int tid = threadIdx.x;
  const int parity = static_cast<int>(epoch & 1ull);
  const unsigned tag = static_cast<unsigned>(epoch & 0xffffffffull);
  double* mine = p.cta_part + static_cast<size_t>(blockIdx.x) * kTriMax;
  for (int t = tid; t < ntri; t += TT) {
    int i, j;
    tri_decode(tri, t, i, j);
    __stcg(mine + t, A0[i * n + j]);
  }
  const bool last = grid_ticket_last(p.ticket, s_flag);
  if (last) {
    const int G = gridDim.x;
    // each packed entry: H threads take the partials h, h+H, ... with 16 loads in flight, combined in a fixed order
    int H = TT / ntri;
    if (H < 1) H = 1;
    if (H > 8) H = 8;
    double* piece = A1;                       // [H][per] staging
    const int per = TT / H;
    for (int base = 0; base < ntri; base += per) {
      const int loc = tid / H, h = tid - loc * H;
      const int e = base + loc;
      double v = 0.0;
      if (loc < per && e < ntri) {
        double s[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) s[u] = 0.0;
        const double* gp = p.cta_part + e;
        int k = h;
        for (; k + 15 * H < G; k += 16 * H) {
#pragma unroll
          for (int u = 0; u < 16; ++u) s[u] += __ldcg(gp + static_cast<size_t>(k + u * H) * kTriMax);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (k + u * H < G) s[u] += __ldcg(gp + static_cast<size_t>(k + u * H) * kTriMax);
        v = (((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]))) +
            (((s[8] + s[9]) + (s[10] + s[11])) + ((s[12] + s[13]) + (s[14] + s[15])));
      }
      if (loc < per) piece[h * per + loc] = v;
      __syncthreads();
      if (tid < per && base + tid < ntri) {
        double t = piece[tid];
        for (int hh = 1; hh < H; ++hh) t += piece[hh * per + tid];
#pragma unroll 1
        for (int r = 0; r < p.nranks; ++r) ll_store(&p.mail[r]->gram_ll[parity][p.my_rank][base + tid][0], t, tag);
      }
      __syncthreads();
    }
  }
  // every CTA: collect all ranks' partials from the own mailbox (self-validating words: spin per value, rank-order sum)
  Mailbox* own = p.mail[p.my_rank];
  double* R = CPLX ? A1 : A0;
  for (int t = tid; t < ntri; t += TT) {
    int i, j;
    tri_decode(tri, t, i, j);
    double s = 0.0;
#pragma unroll 1
    for (int r = 0; r < p.nranks; ++r) {
      const unsigned long long* slot = &own->gram_ll[parity][r][t][0];
      double v = 0.0;
      if (!ll_try_load(slot, tag, v)) {
        const unsigned long long t0 = global_timer_ns();
        unsigned it = 0;
        while (!ll_try_load(slot, tag, v)) {
          if ((++it & 63u) == 0u && static_cast<long long>(global_timer_ns() - t0) > p.spin_budget_ns) {
            atomicOr(p.status, 4);
            break;
          }
        }
      }
      s += v;
    }
    R[i * n + j] = s;
    R[j * n + i] = s;
  }
  __syncthreads();
  if (CPLX) {
    // C[a][b] = (R[2a][2b] + R[2a+1][2b+1]) + i (R[2a][2b+1] - R[2a+1][2b]);  hat(C) = [[cr, ci], [-ci, cr]]
    const int qq = n / 2;
    for (int e = tid; e < qq * qq; e += TT) {
      const int a = e / qq, b = e - a * qq;
      const double cr = R[(2 * a) * n + 2 * b] + R[(2 * a + 1) * n + 2 * b + 1];
      const double ci = R[(2 * a) * n + 2 * b + 1] - R[(2 * a + 1) * n + 2 * b];
      A0[(2 * a) * n + 2 * b] = cr;
      A0[(2 * a + 1) * n + 2 * b + 1] = cr;
      A0[(2 * a) * n + 2 * b + 1] = ci;
      A0[(2 * a + 1) * n + 2 * b] = -ci;
    }
    __syncthreads();
  }
}

// ---- small solve: Ms <- A0^{-1/2} (Taylor, returns true: the product with it is the final polar factor)  or
// Ms <- W max(lam, floor)^{-1/2} W^T from a parallel-ordered two-sided Jacobi (returns false: another pass follows).
struct Rot { double c, s; };
// Rotation that annihilates A[pa][pb] (the small one, |theta| <= pi/4) from two reciprocal square roots - no division, no
// square root: ir = 1/sqrt(dl^2 + 4 apq^2), cos 2t = |dl| ir, h = (1 + cos 2t)/2 = c^2, c = h rsqrt(h), s = +-apq ir rsqrt(h).
__device__ __forceinline__ Rot make_rot(const double* A, int n, int pa, int pb, double tol2, double loose2, int* s_any, int* s_big,
                                        bool set_flags) {
  Rot r{1.0, 0.0};
  if (pb >= n) return r;
  const double app = A[pa * n + pa], aqq = A[pb * n + pb], apq = A[pa * n + pb];
  const double g2 = apq * apq;
  const double dd = fabs(app * aqq);
  if (!(g2 > tol2 * dd)) return r;
  if (set_flags) {
    *s_any = 1;
    if (g2 > loose2 * dd) *s_big = 1;
  }
  const double dl = aqq - app;
  const double ir = rsqrt(fma(dl, dl, 4.0 * g2));
  const double h = fma(0.5 * fabs(dl), ir, 0.5);
  const double rh = rsqrt(h);
  r.c = h * rh;
  r.s = (dl >= 0.0 ? apq : -apq) * ir * rh;
  return r;
}

__device__ __forceinline__ void pair_of(int s, int j, int ne, int& pa, int& pb) {
  const int n1 = ne - 1;
  if (j == 0) { pa = s; pb = ne - 1; }
  else { pa = (s + j) % n1; pb = (s - j + n1) % n1; }
  if (pa > pb) { const int t = pa; pa = pb; pb = t; }
}

template <bool CPLX>
__device__ bool small_solve(const TailArgs& p, const SmemPlan& sp, double* sm, bool allow_warm, bool save_warm, int* s_ints,
                            int* sweeps_out) {
  const int n = sp.n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Ms = sm + sp.ms;
  double* A0 = sm + sp.a0;
  double* A1 = sm + sp.u1;
  double* W = A1 + n * n;
  double* X = W + n * n;
  double* lam = sm + sp.lam;
  int* s_any = s_ints + 1;
  int* s_big = s_ints + 2;

  // max |A - I| and max diagonal
  double emax = 0.0, dmax = 0.0;
  bool finite = true;
  for (int e = tid; e < n * n; e += TT) {
    const int a = e / n, b = e - a * n;
    const double v = A0[e];
    if (!isfinite(v)) finite = false;
    emax = fmax(emax, fabs(v - (a == b ? 1.0 : 0.0)));
    if (a == b) dmax = fmax(dmax, v);
  }
  for (int o = 16; o; o >>= 1) {
    emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  const int anybad = __syncthreads_or(finite ? 0 : 1);
  if (lane == 0) { X[warp] = emax; X[TW + warp] = dmax; }
  __syncthreads();
  emax = 0.0; dmax = 0.0;
#pragma unroll 1
  for (int w = 0; w < TW; ++w) { emax = fmax(emax, X[w]); dmax = fmax(dmax, X[TW + w]); }
  __syncthreads();
  const int kt4 = sp.kt4, pm = sp.pm;
  if (anybad || !(dmax > 0.0)) {
    // non-finite or non-positive Gram matrix: flag it, apply the identity (the host turns the flag into an error)
    if (tid == 0) atomicOr(p.status, 1);
    for (int idx = tid; idx < kt4 * pm; idx += TT) {
      const int k = idx / pm, c = idx - k * pm;
      Ms[idx] = (k == c && k < n) ? 1.0 : 0.0;
    }
    __syncthreads();
    return true;
  }

  if (emax * n <= 1e-4) {
    // C = I + E, tiny E:  C^{-1/2} = I - E/2 + 3/8 E^2 - 5/16 E^3 + O(E^4)   (truncation <= 0.3 (n max|E|)^4 <= 3e-17)
    for (int e = tid; e < n * n; e += TT) { const int a = e / n, b = e - a * n; W[e] = A0[e] - (a == b ? 1.0 : 0.0); }
    __syncthreads();
    for (int e = tid; e < n * n; e += TT) {
      const int a = e / n, b = e - a * n;
      double s = 0.0;
#pragma unroll 2
      for (int c = 0; c < n; ++c) s += W[a * n + c] * W[c * n + b];
      X[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += TT) {
      const int a = e / n, b = e - a * n;
      double s = 0.0;
#pragma unroll 2
      for (int c = 0; c < n; ++c) s += W[a * n + c] * X[c * n + b];
      A1[e] = (a == b ? 1.0 : 0.0) - 0.5 * W[e] + 0.375 * X[e] - 0.3125 * s;
    }
    __syncthreads();
  } else {
    if (p.diag && blockIdx.x == 0 && tid == 0) p.diag[41] = static_cast<double>(global_timer_ns() % 1000000000ull);   // start of the Jacobi branch
    // ---- two-sided Jacobi, round-robin parallel ordering: every round rotates n/2 disjoint pairs; the update
    // A <- J^T A J touches every 2x2 block (pair I x pair J) independently, so one thread owns one block per round
    // (upper blocks + mirror), reads the old matrix and writes the other buffer: ONE barrier per round, no reductions.
    const double scale = ldexp(1.0, -ilogb(dmax));            // exact power-of-two normalisation
    for (int e = tid; e < n * n; e += TT) A0[e] *= scale;
    bool use_warm = false;
    if (allow_warm && p.warm_in) use_warm = (p.warm_in[0] == static_cast<double>(n)) && (p.warm_in[1] < 16.0);
    if (use_warm) {
      for (int e = tid; e < n * n; e += TT) W[e] = p.warm_in[2 + e];
      __syncthreads();
      for (int e = tid; e < n * n; e += TT) {      // X = A W0
        const int a = e / n, b = e - a * n;
        double s = 0.0;
#pragma unroll 2
        for (int c = 0; c < n; ++c) s += A0[a * n + c] * W[c * n + b];
        X[e] = s;
      }
      __syncthreads();
      for (int e = tid; e < n * n; e += TT) {      // A1 = W0^T X  (upper), mirrored
        const int a = e / n, b = e - a * n;
        if (a <= b) {
          double s = 0.0;
#pragma unroll 2
          for (int c = 0; c < n; ++c) s += W[c * n + a] * X[c * n + b];
          A1[a * n + b] = s;
          A1[b * n + a] = s;
        }
      }
      __syncthreads();
      for (int e = tid; e < n * n; e += TT) A0[e] = A1[e];
    } else {
      for (int e = tid; e < n * n; e += TT) { const int a = e / n, b = e - a * n; W[e] = (a == b) ? 1.0 : 0.0; }
    }
    if (tid == 0) { *s_any = 0; *s_big = 0; }
    __syncthreads();
    const int ne = n + (n & 1), np = ne / 2, n1 = ne - 1;
    const double tol2 = 16.0 * DBL_EPSILON * DBL_EPSILON;
    // The eigenvectors must be converged to working precision although another Gram pass follows: with W^T G W = D^-1 (I + E') D^-1
    // the product B W D W^T is polar(B) rotated by I + E'_ij (s_i - s_j) / (2 (s_i + s_j)), which the second pass cannot undo.
    // Quadratic convergence: a sweep whose rotations all started below 1e-8 (relative) ends at ~1e-16 - no confirming sweep.
    const double loose2 = 1e-8 * 1e-8;
    // pair table of the round-robin ordering: ptab[s * np + j] = pa | pb << 8   (X is free during the sweeps)
    unsigned short* ptab = reinterpret_cast<unsigned short*>(X);
    for (int e = tid; e < n1 * np; e += TT) {
      int pa, pb;
      pair_of(e / np, e % np, ne, pa, pb);
      ptab[e] = static_cast<unsigned short>(pa | (pb << 8));
    }
    // block table: btab[b] = I | J << 8 for the np x np grid of 2x2 blocks (pair I x pair J)
    unsigned short* btab = ptab + ((n1 * np + 3) & ~3);
    for (int b = tid; b < np * np; b += TT) btab[b] = static_cast<unsigned short>((b / np) | ((b % np) << 8));
    __syncthreads();
    double* cur = A0;
    double* nxt = A1;
    int sweeps = 0;
    if (p.diag && blockIdx.x == 0 && tid == 0) p.diag[42] = static_cast<double>(global_timer_ns() % 1000000000ull);   // start of the sweeps
    // Four lanes per block: lane 0/2 set up the rotation of pair I, lane 1/3 that of pair J (exchanged with shuffles), then
    // each lane produces ONE entry of the rotated 2x2 blocks of A and W.  Short dependent chains, 3-4 warps per scheduler.
    const int sub = tid & 3, ri = sub >> 1, ci = sub & 1;
    const int lane_base = lane & ~3;
    const int nbt = np * np;
    for (int sweep = 0; sweep < 40; ++sweep) {
      for (int s = 0; s < n1; ++s) {
        const unsigned short* prow = ptab + s * np;
        for (int b0 = 0; b0 < nbt; b0 += TT / 4) {
          const int b = b0 + (tid >> 2);
          const bool act = b < nbt;
          const int e = act ? btab[b] : 0;
          const int I = e & 0xff, J = e >> 8;
          const int eI = prow[I], eJ = prow[J];
          const int pI = eI & 0xff, qI = eI >> 8, pJ = eJ & 0xff, qJ = eJ >> 8;
          // lane's rotation: even lanes pair I, odd lanes pair J (flags raised once per pair: diagonal block, lane 1)
          const Rot mine = make_rot(cur, n, ci ? pJ : pI, ci ? qJ : qI, tol2, loose2, s_any, s_big, act && I == J && sub == 1);
          const double cI = __shfl_sync(0xffffffffu, mine.c, lane_base), sI = __shfl_sync(0xffffffffu, mine.s, lane_base);
          const double cJ = __shfl_sync(0xffffffffu, mine.c, lane_base + 1), sJ = __shfl_sync(0xffffffffu, mine.s, lane_base + 1);
          // coefficients of this lane's entry: row combination (R_I^T) and column combination (R_J)
          const double uI0 = ri ? sI : cI, uI1 = ri ? cI : -sI;
          const double uJ0 = ci ? sJ : cJ, uJ1 = ci ? cJ : -sJ;
          const int row = ri ? qI : pI, col = ci ? qJ : pJ;
          const bool vq = (qI < n) && (qJ < n);
          const bool val = act && row < n && col < n;
          double w0 = 0.0, w1 = 0.0;
          if (val && qJ < n) { w0 = W[row * n + pJ]; w1 = W[row * n + qJ]; }
          if (act && I <= J && val) {
            const double a11 = cur[pI * n + pJ];
            const double a12 = (qJ < n) ? cur[pI * n + qJ] : 0.0;
            const double a21 = (qI < n) ? cur[qI * n + pJ] : 0.0;
            const double a22 = vq ? cur[qI * n + qJ] : 0.0;
            double v = uJ0 * (uI0 * a11 + uI1 * a21) + uJ1 * (uI0 * a12 + uI1 * a22);
            if (I == J && ri != ci && sJ != 0.0) v = 0.0;      // the annihilated entry, exactly
            nxt[row * n + col] = v;
            nxt[col * n + row] = v;
          }
          __syncwarp();
          if (val && qJ < n) W[row * n + col] = uJ0 * w0 + uJ1 * w1;      // W <- W J (in place: the four lanes own these entries)
        }
        __syncthreads();
        double* t = cur; cur = nxt; nxt = t;
      }
      ++sweeps;
      const int any = *s_any, big = *s_big;
      __syncthreads();
      if (tid == 0) { *s_any = 0; *s_big = 0; }
      __syncthreads();
      if (!any || !big) break;
    }
    if (sweeps_out) *sweeps_out += sweeps;
    if (p.diag && blockIdx.x == 0 && tid == 0) p.diag[40] = static_cast<double>(global_timer_ns() % 1000000000ull);   // end of the sweeps
    if (save_warm && p.warm_out && blockIdx.x == 0) {
      for (int e = tid; e < n * n; e += TT) p.warm_out[2 + e] = W[e];
      if (tid == 0) { p.warm_out[0] = static_cast<double>(n); p.warm_out[1] = use_warm ? p.warm_in[1] + 1.0 : 0.0; }
    }
    // eigenvalues (of the normalised matrix) on the diagonal; clamp the ones drowned in round-off (graded / nearly singular
    // Gram matrices): the product keeps the polar factor of T (any function of T^H T does) and the next pass sees cond^(1/2)
    if (warp == 0) {       // n <= 64: two diagonals per lane
      const double l0 = lane < n ? cur[lane * n + lane] : 0.0;
      const double l1 = lane + 32 < n ? cur[(lane + 32) * n + lane + 32] : 0.0;
      double lmax = fmax(l0, l1);
      for (int o = 16; o; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      const double floor_ = lmax * 1e-13;
      if (lane < n) lam[lane] = rsqrt(fmax(l0, floor_) / scale);            // back to the un-normalised matrix
      if (lane + 32 < n) lam[lane + 32] = rsqrt(fmax(l1, floor_) / scale);
    }
    __syncthreads();
    double* Mout = (cur == A0) ? A1 : A0;   // the buffer that is not `cur`
    for (int e = tid; e < n * n; e += TT) {
      const int a = e / n, b = e - a * n;
      if (a <= b) {
        double s = 0.0;
#pragma unroll 2
        for (int c = 0; c < n; ++c) s += W[a * n + c] * lam[c] * W[b * n + c];
        Mout[a * n + b] = s;
        Mout[b * n + a] = s;
      }
    }
    __syncthreads();
    if (Mout != A1) {
      for (int e = tid; e < n * n; e += TT) A1[e] = Mout[e];
      __syncthreads();
    }
  }
  const bool taylor = (emax * n <= 1e-4);
  // A1 holds M (symmetric).  Complex panels: project onto the exact hat structure [[mr, mi], [-mi, mr]].
  for (int idx = tid; idx < kt4 * pm; idx += TT) {
    const int k = idx / pm, c = idx - k * pm;
    double v = 0.0;
    if (k < n && c < n) {
      if (CPLX) {
        const int a = k >> 1, b = c >> 1;
        const double mr = 0.5 * (A1[(2 * a) * n + 2 * b] + A1[(2 * a + 1) * n + 2 * b + 1]);
        const double mi = 0.5 * (A1[(2 * a) * n + 2 * b + 1] - A1[(2 * a + 1) * n + 2 * b]);
        v = ((k & 1) == (c & 1)) ? mr : ((k & 1) ? -mi : mi);
      } else {
        v = A1[k * n + c];
      }
    }
    Ms[idx] = v;
  }
  __syncthreads();
  return taylor;
}

// ---- T rows <- producer
template <bool CPLX>
__device__ void produce_rows(const TailArgs& p, int64_t r0, int64_t r1, double a, const double* s_w) {
  const int tid = threadIdx.x;
  const int pitch = p.pitch;
  if (p.producer == TAIL_REWEIGHT) {
    // B = d*Y - (w(|V|) - wmax) * V * rho      R/spEigen.R:170-177   (s_w[c] = column max of the weights = w(min |V[,c]|))
    const int q = p.q;
    const int nel = static_cast<int>(r1 - r0) * q;          // the CTA's rows x q: far below 2^31
    for (int idx = tid; idx < nel; idx += TT) {
      const int64_t r = r0 + idx / q;
      const int c = idx % q;
      if (CPLX) {
        const size_t g = static_cast<size_t>(r) * pitch + 2 * c;
        const double2 v = *reinterpret_cast<const double2*>(p.in1 + g);
        const double2 y = *reinterpret_cast<const double2*>(p.in0 + g);
        const double tt = weight_of_abs(hypot(v.x, v.y), p.sc) - s_w[c];
        const double dc = p.d[c], rc = p.rho[c];
        *reinterpret_cast<double2*>(p.T + g) = make_double2(reweight_elem(y.x, v.x, tt, dc, rc), reweight_elem(y.y, v.y, tt, dc, rc));
      } else {
        const size_t g = static_cast<size_t>(r) * pitch + c;
        const double v = p.in1[g], y = p.in0[g];
        const double tt = weight_of_abs(fabs(v), p.sc) - s_w[c];
        p.T[g] = reweight_elem(y, v, tt, p.d[c], p.rho[c]);
      }
    }
  } else if (p.producer == TAIL_EXTRAP) {
    // V - 2 a R + a^2 U,  R = V1 - V, U = V2 - V1 - R      R/spEigen.R:121-127
    const int ncr = p.ncr;
    const double a2 = a * a;
    const int nel = static_cast<int>(r1 - r0) * ncr;
    for (int idx = tid; idx < nel; idx += TT) {
      const int64_t r = r0 + idx / ncr;
      const size_t g = static_cast<size_t>(r) * pitch + (idx % ncr);
      const double v = p.in0[g], v1 = p.in1[g], v2 = p.in2[g];
      const double R = v1 - v;
      const double U = v2 - v1 - R;
      p.T[g] = v - 2.0 * a * R + a2 * U;
    }
  }
}

// ---- Out rows = T rows * Ms (DMMA).  FINAL: stores to the destination panel(s) and accumulates the statistics.
// g of R/spEigen.R:133-134 (one out-of-line copy: log() is ~100 instructions and the kernel is instruction-fetch bound)
__device__ __noinline__ double penalty_g(double a, double p, double eps, double c1, double c2) {
  return (a <= eps) ? (a * a) / (eps * c2) : log((p + a) / (p + eps)) / c1 + eps / c2;
}

// GRAM (non-final passes, NT <= 3): the Gram matrix of the product is accumulated in the same pass (output tile staged per warp
// in shared memory, re-read as DMMA fragments) and left in the solve buffer A0 - no second trip over the rows.
// FINAL: stores to the destination panel(s); the column statistics are taken from the same staged tile by "column owner"
// lanes (lane c owns logical columns c and c + 32): one running minimum / sum per lane instead of per-fragment arrays.
template <int NT, bool CPLX, bool FINAL, bool GRAM = false>
__device__ void apply_rows(const TailArgs& p, const SmemPlan& sp, double* sm, int64_t r0, int64_t r1, double* s_stat) {
  const int n = sp.n, pm = sp.pm, pitch = p.pitch;
  const int kt = sp.kt4 / 4;
  const double* Ms = sm + sp.ms;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane >> 2, kk = lane & 3;
  const int64_t ntile = (r1 - r0 + 7) / 8;
  const bool col_stats = FINAL && (p.want_absmin || p.want_gsum);
  const bool stage = GRAM || col_stats;
  double cmin0 = DBL_MAX, cmin1 = DBL_MAX, cgs0 = 0.0, cgs1 = 0.0;
  double sR = 0.0, sU = 0.0;
  constexpr int NG = GRAM ? NT : 1;
  double gacc[NG][NG][2];
#pragma unroll
  for (int i = 0; i < NG; ++i)
#pragma unroll
    for (int j = 0; j < NG; ++j) gacc[i][j][0] = gacc[i][j][1] = 0.0;
  double* mytile = sm + sp.u1 + static_cast<size_t>(warp) * 8 * pm;     // per-warp staging of the 8 x ncr output tile
  const StageConsts sc = p.sc;
  // A-fragments of the next 8-row tile are fetched (L2) while the current one is multiplied
  double af[2 * NT];
  auto load_tile = [&](int64_t t, double* dst) {
    const int64_t row = r0 + t * 8 + grp;
    const bool rok = t < ntile && row < r1;
    const double* src = p.T + row * pitch + kk;
#pragma unroll
    for (int k4 = 0; k4 < 2 * NT; ++k4) dst[k4] = (k4 < kt && rok && k4 * 4 + kk < n) ? __ldcg(src + k4 * 4) : 0.0;
  };
  load_tile(warp, af);
  for (int64_t t = warp; t < ntile; t += TW) {
    const int64_t row = r0 + t * 8 + grp;
    const bool rok = row < r1;
    double acc[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
    double afn[2 * NT];
    load_tile(t + TW, afn);
#pragma unroll
    for (int k4 = 0; k4 < 2 * NT; ++k4) {
      if (k4 < kt) {
        const int k = k4 * 4 + kk;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma_884(acc[nt][0], acc[nt][1], af[k4], Ms[k * pm + nt * 8 + grp]);
      }
    }
#pragma unroll
    for (int k4 = 0; k4 < 2 * NT; ++k4) af[k4] = afn[k4];
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c0 = nt * 8 + 2 * kk;
      const bool ok = rok && c0 < n;
      const bool has1 = c0 + 1 < n;
      if (stage) {
        mytile[grp * pm + c0] = ok ? acc[nt][0] : 0.0;
        mytile[grp * pm + c0 + 1] = (ok && has1) ? acc[nt][1] : 0.0;
      }
      if (!ok) continue;
      const size_t g = static_cast<size_t>(row) * pitch + c0;
      if (!FINAL) {
        if (has1) *reinterpret_cast<double2*>(p.T + g) = make_double2(acc[nt][0], acc[nt][1]);
        else p.T[g] = acc[nt][0];
        continue;
      }
      if (p.n_out > 1) {
#pragma unroll 1
        for (int r = 0; r < p.n_out; ++r) {
          if (has1) *reinterpret_cast<double2*>(p.out_peer[r] + g) = make_double2(acc[nt][0], acc[nt][1]);
          else p.out_peer[r][g] = acc[nt][0];
        }
      } else {
        double* o = p.out_peer[p.my_rank];
        if (has1) *reinterpret_cast<double2*>(o + g) = make_double2(acc[nt][0], acc[nt][1]);
        else o[g] = acc[nt][0];
      }
      if (p.want_squarem) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (e == 1 && !has1) continue;
          const double v = p.sq_V[g + e], v1 = p.sq_V1[g + e];
          const double R = v1 - v;
          const double U = acc[nt][e] - v1 - R;
          sR += R * R;
          sU += U * U;
        }
      }
    }
    if (stage) __syncwarp();
    if (GRAM) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double f[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) f[i] = mytile[(h * 4 + kk) * pm + i * 8 + grp];
#pragma unroll
        for (int i = 0; i < NG; ++i)
#pragma unroll
          for (int j = 0; j < NG; ++j) dmma_884(gacc[i][j][0], gacc[i][j][1], f[i], f[j]);
      }
    }
    if (col_stats) {
      const int64_t left = r1 - (r0 + t * 8);
      const int nrow = left < 8 ? static_cast<int>(left) : 8;
      auto own = [&](int c, double& mn, double& gsum) {
#pragma unroll 1
        for (int r = 0; r < nrow; ++r) {
          const double a = CPLX ? hypot(mytile[r * pm + 2 * c], mytile[r * pm + 2 * c + 1]) : fabs(mytile[r * pm + c]);
          mn = fmin(mn, a);
          if (p.want_gsum) gsum += penalty_g(a, sc.p, sc.eps, sc.c1, sc.c2);
        }
      };
      if (lane < p.q) own(lane, cmin0, cgs0);
      if (lane + 32 < p.q) own(lane + 32, cmin1, cgs1);
    }
    if (stage) __syncwarp();
  }
  if (GRAM) {
    __syncthreads();          // every warp is done with its staging tile: the region becomes the reduction scratch
    gram_tree_store<NG>(gacc, 0, 0, n, sm + sp.a0, sm + sp.u1);
  }
  if (!FINAL) return;
  // fixed-order reductions: tiles of a warp in order (above), then the warps in order
  __syncthreads();            // the staging tiles are dead: the region becomes the statistics scratch
  double* s_min = s_stat;                         // [TW][kMaxCols]
  double* s_gs = s_stat + TW * kMaxCols;          // [TW][kMaxCols]
  double* s_sq = s_gs + TW * kMaxCols;            // [TW][2]
  s_min[warp * kMaxCols + lane] = cmin0; s_min[warp * kMaxCols + lane + 32] = cmin1;
  s_gs[warp * kMaxCols + lane] = cgs0;   s_gs[warp * kMaxCols + lane + 32] = cgs1;
  for (int o = 16; o; o >>= 1) {
    sR += __shfl_xor_sync(0xffffffffu, sR, o);
    sU += __shfl_xor_sync(0xffffffffu, sU, o);
  }
  if (lane == 0) { s_sq[warp * 2] = sR; s_sq[warp * 2 + 1] = sU; }
  __syncthreads();
  double* mine = p.cta_stat + static_cast<size_t>(blockIdx.x) * kStatLen;
  if (threadIdx.x < p.q) {
    double v = s_min[threadIdx.x], g = s_gs[threadIdx.x];
#pragma unroll 1
    for (int w = 1; w < TW; ++w) { v = fmin(v, s_min[w * kMaxCols + threadIdx.x]); g += s_gs[w * kMaxCols + threadIdx.x]; }
    __stcg(mine + kStatAbsmin + threadIdx.x, v);
    __stcg(mine + kStatGsum + threadIdx.x, g);
  }
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll 1
    for (int w = 0; w < TW; ++w) { a += s_sq[w * 2]; b += s_sq[w * 2 + 1]; }
    __stcg(mine + kStatSqR, a);
    __stcg(mine + kStatSqU, b);
  }
}

template <int NT, bool CPLX>
__global__ void __launch_bounds__(TT, 1) k_tail(const TailArgs p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ int s_ints[4];
  __shared__ double s_w[kMaxCols];
  __shared__ double s_a;
  const int tid = threadIdx.x;
  const int n = p.ncr;
  const SmemPlan sp = plan_smem(n, NT);
  constexpr int NB = NT < 4 ? NT : 4;
  {
    uint32_t* tri = reinterpret_cast<uint32_t*>(sm + sp.tri);
    for (int i = tid; i < n; i += TT) {
      // row i of the packed upper triangle starts at i*n - i*(i-1)/2
      const int base = i * n - i * (i - 1) / 2;
      for (int j = i; j < n; ++j) tri[base + (j - i)] = (static_cast<uint32_t>(i) << 16) | static_cast<uint32_t>(j);
    }
  }
  // this CTA's rows
  int64_t per = (p.row1 - p.row0 + gridDim.x - 1) / gridDim.x;
  per = (per + 7) / 8 * 8;
  int64_t r0 = p.row0 + static_cast<int64_t>(blockIdx.x) * per;
  int64_t r1 = r0 + per;
  if (r0 > p.row1) r0 = p.row1;
  if (r1 > p.row1) r1 = p.row1;

  // statistics of the previous tail (all ranks): column minima for the weights, SQUAREM norms for the step length
  if (p.prev_flags) {
    if (tid < p.nranks) spin_until_ge(p.prev_flags + tid * kFlagStride, p.prev_epoch, p.spin_budget_ns, p.status, p.nranks > 1);
    __syncthreads();
  }
  double nR = 0.0, nU = 0.0;
  if (p.producer == TAIL_REWEIGHT) {
    if (tid < p.q) {
      double v = DBL_MAX;
#pragma unroll 1
      for (int r = 0; r < p.nranks; ++r) v = fmin(v, __ldcv(p.prev_stat + static_cast<size_t>(r) * kStatLen + kStatAbsmin + tid));
      s_w[tid] = weight_of_abs(v, p.sc);        // column max of the weights = w(min |V|)   R/spEigen.R:176
    }
  } else if (p.producer == TAIL_EXTRAP) {
    if (tid == 0) {
      double a;
      if (p.retry) {
        a = (*p.a_cur - 1.0) / 2.0;             // :141
      } else {
        double sR = 0.0, sU = 0.0;
#pragma unroll 1
        for (int r = 0; r < p.nranks; ++r) {
          sR += __ldcv(p.prev_stat + static_cast<size_t>(r) * kStatLen + kStatSqR);
          sU += __ldcv(p.prev_stat + static_cast<size_t>(r) * kStatLen + kStatSqU);
        }
        nR = sqrt(sR);
        nU = sqrt(sU);
        // a <- min(-||R||/||U||, -1)  :123.  ||U|| == 0 gives -Inf in R and then a NaN iterate that makes svd() fail; here
        // that case takes the plain double MM step a = -1 instead.
        a = (nU > 0.0) ? fmin(-nR / nU, -1.0) : -1.0;
        if (a != a) a = -1.0;
      }
      s_a = a;
    }
  }
  if (tid == 0) s_ints[3] = 0;
  __syncthreads();
  const double a_step = (p.producer == TAIL_EXTRAP) ? s_a : 0.0;
  const unsigned long long ge0 = *p.gram_epoch;
  // phase clock of CTA 0 (ns, relative to kernel entry) -> diag[8..]: entry, produce, then per pass {gram, exchange, solve, apply}
  const bool clk = p.diag != nullptr && blockIdx.x == 0 && tid == 0;
  const unsigned long long t_entry = clk ? global_timer_ns() : 0ull;
  int clk_i = 8;
#define TAIL_CLK() do { if (clk && clk_i < 40) p.diag[clk_i++] = static_cast<double>(global_timer_ns() - t_entry); } while (0)

  produce_rows<CPLX>(p, r0, r1, a_step, s_w);
  __threadfence_block();
  __syncthreads();
  TAIL_CLK();

  int passes = 0, sweeps = 0;
  double* G = sm + sp.a0;
  double* red = sm + sp.u1;
  if (!p.skip_polar) {
    bool done = false;
    constexpr bool FUSE = NT <= 3;       // Gram of the product accumulated inside the apply pass
    cta_gram<NB>(p.T, p.pitch, n, r0, r1, G, red);
    TAIL_CLK();
    for (int pass = 0; pass < kTailMaxPasses; ++pass) {
      exchange_gram<CPLX>(p, sp, sm, ge0 + pass + 1, &s_ints[0]);
      TAIL_CLK();
      ++passes;
      const bool taylor = small_solve<CPLX>(p, sp, sm, pass == 0, pass == 0, s_ints, &sweeps);
      TAIL_CLK();
      if (taylor) { done = true; break; }
      apply_rows<NT, CPLX, false, FUSE>(p, sp, sm, r0, r1, nullptr);
      __threadfence_block();
      __syncthreads();
      TAIL_CLK();
      if (!FUSE) {
        cta_gram<NB>(p.T, p.pitch, n, r0, r1, G, red);
        TAIL_CLK();
      }
    }
    if (!done && tid == 0) atomicOr(p.status, 2);
  } else {
    double* Ms = sm + sp.ms;
    for (int idx = tid; idx < sp.kt4 * sp.pm; idx += TT) {
      const int k = idx / sp.pm, c = idx - k * sp.pm;
      Ms[idx] = (k == c && k < n) ? 1.0 : 0.0;
    }
    __syncthreads();
  }
  apply_rows<NT, CPLX, true>(p, sp, sm, r0, r1, sm + sp.u1);
  TAIL_CLK();
  if (clk) p.diag[7] = static_cast<double>(clk_i - 8);

  // ---- publish: rows (already stored to every rank) + statistics, then the epoch flags
  const bool multi = p.nranks > 1;
  if (p.n_out > 1) __threadfence_system();      // this thread's peer stores are performed before its CTA takes a ticket
  const bool last = grid_ticket_last(p.ticket + 1, &s_ints[0]);
  if (!last) return;
  const int G_ = gridDim.x;
  const int parity = static_cast<int>(p.stat_epoch & 1ull);
  if (tid < kStatLen) {
    double v = 0.0;
    const bool is_min = tid < kMaxCols;
    const bool used = (tid < p.q) || (tid >= kStatGsum && tid < kStatGsum + p.q) || tid == kStatSqR || tid == kStatSqU;
    if (used) {
      // 8 independent chains over the CTAs (loads in flight together), combined in a fixed order
      double acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = is_min ? DBL_MAX : 0.0;
      const double* src = p.cta_stat + tid;
      int c = 0;
      for (; c + 8 <= G_; c += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double x = __ldcg(src + static_cast<size_t>(c + u) * kStatLen);
          acc[u] = is_min ? fmin(acc[u], x) : acc[u] + x;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c + u < G_) {
          const double x = __ldcg(src + static_cast<size_t>(c + u) * kStatLen);
          acc[u] = is_min ? fmin(acc[u], x) : acc[u] + x;
        }
      v = is_min ? fmin(fmin(fmin(acc[0], acc[1]), fmin(acc[2], acc[3])), fmin(fmin(acc[4], acc[5]), fmin(acc[6], acc[7])))
                 : ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
#pragma unroll 1
    for (int r = 0; r < p.nranks; ++r) p.mail[r]->stat[parity][p.my_rank][tid] = v;
  }
  if (tid == 0) {
    *p.gram_epoch = ge0 + passes;
    if (p.producer == TAIL_EXTRAP) *p.a_cur = a_step;
    if (p.diag) {
      p.diag[0] = a_step;
      if (p.producer == TAIL_EXTRAP && !p.retry) { p.diag[1] = nR; p.diag[2] = nU; }
      p.diag[3] = passes;
      p.diag[4] = sweeps;
      p.diag[5] = fmax(p.diag[5], static_cast<double>(passes));     // running maximum (reset by the host per solve)
      p.diag[6] += sweeps;
    }
  }
  if (multi) __threadfence_system();
  __syncthreads();
  if (tid < p.nranks) {
    if (multi) st_release_sys_u64(&p.mail[tid]->stat_flag[p.my_rank * kFlagStride], p.stat_epoch);
    else st_release_gpu_u64(&p.mail[tid]->stat_flag[p.my_rank * kFlagStride], p.stat_epoch);
  }
}

// ---- objective assembly ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_finish_objective(const FinishArgs p) {
  __shared__ double s_quad[kMaxCols], s_gs[kMaxCols];
  const int tid = threadIdx.x;
  if (p.dots_flags && tid < p.nranks) spin_until_ge(p.dots_flags + tid, p.dots_epoch, p.spin_budget_ns, p.status, true);
  if (p.stat_flags && tid >= 32 && tid < 32 + p.nranks)
    spin_until_ge(p.stat_flags + (tid - 32) * kFlagStride, p.stat_epoch, p.spin_budget_ns, p.status, p.nranks > 1);
  __syncthreads();
  if (tid < p.q) {
    double qd = 0.0, g = 0.0;
    for (int k = 0; k < p.n_dot_parts; ++k) qd += __ldcv(p.dots + static_cast<size_t>(k) * p.q + tid);
    for (int r = 0; r < p.nranks; ++r) g += __ldcv(p.stat + static_cast<size_t>(r) * kStatLen + kStatGsum + tid);
    s_quad[tid] = qd;
    s_gs[tid] = g;
    p.out[1 + tid] = qd;
    p.out[1 + p.q + tid] = g;
    if (p.out_host) { p.out_host[1 + tid] = qd; p.out_host[1 + p.q + tid] = g; }
  }
  __syncthreads();
  if (tid == 0) {
    double fq = 0.0, fg = 0.0;
    for (int c = 0; c < p.q; ++c) { fq += s_quad[c] * p.d[c]; fg += s_gs[c] * p.rho[c]; }
    const double F = fq - fg;      // R/spEigen.R:136,138
    p.out[0] = F;
    const int st = p.status ? *reinterpret_cast<volatile int*>(p.status) : 0;
    const int o = 1 + 2 * p.q;
    for (int i = 0; i < 5; ++i) p.out[o + i] = p.diag ? p.diag[i] : 0.0;
    p.out[o + 5] = static_cast<double>(st);
    if (p.out_host) {
      for (int i = 0; i < 5; ++i) p.out_host[o + i] = p.diag ? p.diag[i] : 0.0;
      p.out_host[o + 5] = static_cast<double>(st);
      p.out_host[0] = F;
    }
  }
}

__global__ void k_wait_flags(const unsigned long long* flags, int stride, int n, unsigned long long epoch, int* status,
                             long long budget_ns) {
  if (static_cast<int>(threadIdx.x) < n) spin_until_ge(flags + threadIdx.x * stride, epoch, budget_ns, status, true);
  __syncthreads();
}

template <int NT, bool CPLX>
int launch_tail_nt(const TailArgs& a, int grid, cudaStream_t st) {
  const SmemPlan sp = plan_smem(a.ncr, NT);
  TailArgs args = a;
  void* params[] = {&args};
  // Cooperative launch = the driver guarantees that all CTAs are co-resident (the kernel waits inside).  SPEIGEN_TAIL_COOP=0
  // (measurement only) uses a plain launch, which is co-resident in practice on an otherwise idle GPU (grid <= #SMs, 1 CTA/SM).
  static const bool coop = []() { const char* v = getenv("SPEIGEN_TAIL_COOP"); return !(v && v[0] == '0'); }();
  cudaError_t e;
  if (coop) {
    e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_tail<NT, CPLX>), dim3(grid), dim3(TT), params, sp.total, st);
  } else {
    k_tail<NT, CPLX><<<grid, TT, sp.total, st>>>(args);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    set_error("tail kernel launch failed: %s (grid %d, %zu bytes of shared memory)", cudaGetErrorString(e), grid, sp.total);
    return SPEIGEN_ERR_CUDA;
  }
  note_launch(1);
  return SPEIGEN_OK;
}

template <int NT, bool CPLX>
int prepare_nt() {
  SPE_CUDA(cudaFuncSetAttribute(k_tail<NT, CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return SPEIGEN_OK;
}

}  // namespace

int tail_grid(int64_t rows, int num_sms) {
  int64_t g = (rows + 127) / 128;
  if (g < 1) g = 1;
  if (g > num_sms) g = num_sms;
  return static_cast<int>(g);
}

int tail_prepare_device() {
  int rc;
#define PREP(NT)                                         \
  if ((rc = prepare_nt<NT, false>()) != SPEIGEN_OK) return rc; \
  if ((rc = prepare_nt<NT, true>()) != SPEIGEN_OK) return rc;
  PREP(1) PREP(2) PREP(3) PREP(4) PREP(5) PREP(6) PREP(7) PREP(8)
#undef PREP
  cudaFuncAttributes fa;
  SPE_CUDA(cudaFuncGetAttributes(&fa, k_finish_objective));
  return SPEIGEN_OK;
}

int launch_tail(const TailArgs& a, int num_sms, cudaStream_t st) {
  if (a.ncr < 1 || a.ncr > kMaxCols || a.nranks < 1 || a.nranks > kTailMaxRanks) {
    set_error("tail: unsupported shape (ncr %d, ranks %d)", a.ncr, a.nranks);
    return SPEIGEN_ERR_ARG;
  }
  const int grid = tail_grid(a.row1 - a.row0, num_sms);
  const int nt = (a.ncr + 7) / 8;
#define CASE(NT) \
  case NT: return a.cplx ? launch_tail_nt<NT, true>(a, grid, st) : launch_tail_nt<NT, false>(a, grid, st);
  switch (nt) {
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
    default: return a.cplx ? launch_tail_nt<8, true>(a, grid, st) : launch_tail_nt<8, false>(a, grid, st);
  }
#undef CASE
}

int launch_wait_flags(const unsigned long long* flags, int stride, int n, unsigned long long epoch, int* status,
                      long long budget_ns, cudaStream_t st) {
  k_wait_flags<<<1, 32, 0, st>>>(flags, stride, n, epoch, status, budget_ns);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_finish_objective(const FinishArgs& a, cudaStream_t st) {
  k_finish_objective<<<1, 128, 0, st>>>(a);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

}  // namespace speigen
