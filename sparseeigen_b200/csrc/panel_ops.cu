// m x q panel kernels (see panel_ops.cuh).  sm_100a; panels are L2-resident (8 MB at m=50k, q=20), so these
// kernels are latency/launch bound, not HBM bound: one CTA per SM, fixed-order two-stage reductions.
#include "panel_ops.cuh"
#include "../../include/speigen_b200.h"
#include <cfloat>

namespace speigen {
namespace {

constexpr int kPT = 256;   // threads per panel block

__device__ __forceinline__ void block_rows(int64_t m, int64_t& r0, int64_t& r1) {
  int64_t per = (m + gridDim.x - 1) / gridDim.x;
  per = (per + kTileRows - 1) / kTileRows * kTileRows;
  r0 = static_cast<int64_t>(blockIdx.x) * per;
  r1 = r0 + per;
  if (r0 > m) r0 = m;
  if (r1 > m) r1 = m;
}

__device__ __forceinline__ bool last_block_done(int* ticket) {
  __shared__ int s_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    s_is_last = (t == static_cast<int>(gridDim.x) - 1);
    if (s_is_last) *ticket = 0;
  }
  __syncthreads();
  if (s_is_last) __threadfence();
  return s_is_last != 0;
}

// load `nrows` rows starting at r0 of a pitched panel into a shared tile (rows beyond r1 zero-filled)
__device__ __forceinline__ void load_tile(double* tile, const double* P, int64_t r0, int64_t r1, int pitch) {
  const int n = kTileRows * pitch;
  const int64_t avail = (r1 - r0) * pitch;
  const double* src = P + r0 * pitch;
  for (int i = threadIdx.x; i < n; i += kPT) tile[i] = (i < avail) ? __ldcg(src + i) : 0.0;
}

// accumulate Gram entries of conj(TA)^T * TB over the tile rows (fixed row order)
template <bool CPLX>
__device__ __forceinline__ void gram_accumulate(double* gacc, const double* TA, const double* TB, int q, int pitch,
                                                int nrows) {
  const int nent = q * q;
  if (CPLX) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = threadIdx.x + k * kPT;
      if (e < nent) {
        const int a = e / q, b = e - a * q;
        double sr = gacc[2 * k], si = gacc[2 * k + 1];
        for (int r = 0; r < nrows; ++r) {
          const double ar = TA[r * pitch + 2 * a], ai = TA[r * pitch + 2 * a + 1];
          const double br = TB[r * pitch + 2 * b], bi = TB[r * pitch + 2 * b + 1];
          sr += ar * br + ai * bi;
          si += ar * bi - ai * br;
        }
        gacc[2 * k] = sr;
        gacc[2 * k + 1] = si;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int e = threadIdx.x + k * kPT;
      if (e < nent) {
        const int a = e / q, b = e - a * q;
        double s = gacc[k];
        for (int r = 0; r < nrows; ++r) s += TA[r * pitch + a] * TB[r * pitch + b];
        gacc[k] = s;
      }
    }
  }
}
template <bool CPLX>
__device__ __forceinline__ void gram_store(const double* gacc, double* gram_part, int q) {
  const int nent = q * q;
  if (CPLX) {
    double* dst = gram_part + static_cast<size_t>(blockIdx.x) * nent * 2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = threadIdx.x + k * kPT;
      if (e < nent) { dst[2 * e] = gacc[2 * k]; dst[2 * e + 1] = gacc[2 * k + 1]; }
    }
  } else {
    double* dst = gram_part + static_cast<size_t>(blockIdx.x) * nent;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int e = threadIdx.x + k * kPT;
      if (e < nent) dst[e] = gacc[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <bool CPLX>
__global__ void __launch_bounds__(kPT) k_gram(const double* A, const double* B, int64_t m, int q, int pitch,
                                              double* gram_part) {
  extern __shared__ double sm[];
  double* TA = sm;
  double* TB = (A == B) ? TA : sm + kTileRows * pitch;
  int64_t r0, r1;
  block_rows(m, r0, r1);
  double gacc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) gacc[k] = 0.0;
  for (int64_t t = r0; t < r1; t += kTileRows) {
    load_tile(TA, A, t, r1, pitch);
    if (A != B) load_tile(TB, B, t, r1, pitch);
    __syncthreads();
    const int nrows = static_cast<int>((r1 - t) < kTileRows ? (r1 - t) : kTileRows);
    gram_accumulate<CPLX>(gacc, TA, TB, q, pitch, nrows);
    __syncthreads();
  }
  gram_store<CPLX>(gacc, gram_part, q);
}

// ------------------------------------------------------------------------------------------------
// Producer kernels: Out tile = op(inputs); fused Gram / min|.| stage-1 reductions of Out.
enum ProdOp { OP_APPLY = 0, OP_REWEIGHT = 1, OP_EXTRAP = 2 };
struct ProdArgs {
  const double* in0;
  const double* in1;
  const double* in2;
  const double* Msmall;
  const double* d;
  const double* rho;
  const double* wmax;
  StageConsts sc;
  double a;
  double* Out;
  double* gram_part;
  double* absmin_part;
  int64_t m;
  int q, ncr, pitch;
};

template <int OP, bool CPLX>
__global__ void __launch_bounds__(kPT) k_produce(const ProdArgs p) {
  extern __shared__ double sm[];
  const int pitch = p.pitch, q = p.q, ncr = p.ncr;
  double* Tin = sm;                              // [kTileRows][pitch]  (APPLY input)
  double* Tout = sm + kTileRows * pitch;         // [kTileRows][pitch]
  double* Ms = Tout + kTileRows * pitch;         // APPLY: q*q*(1+CPLX) ; REWEIGHT: d,rho,wmax
  if (OP == OP_APPLY) {
    const int n = q * q * (CPLX ? 2 : 1);
    for (int i = threadIdx.x; i < n; i += kPT) Ms[i] = p.Msmall[i];
  } else if (OP == OP_REWEIGHT) {
    for (int i = threadIdx.x; i < q; i += kPT) { Ms[i] = p.d[i]; Ms[kMaxCols + i] = p.rho[i]; Ms[2 * kMaxCols + i] = p.wmax[i]; }
  }
  int64_t r0, r1;
  block_rows(p.m, r0, r1);
  double gacc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) gacc[k] = 0.0;
  double amin = DBL_MAX;
  __syncthreads();
  for (int64_t t = r0; t < r1; t += kTileRows) {
    const int nrows = static_cast<int>((r1 - t) < kTileRows ? (r1 - t) : kTileRows);
    if (OP == OP_APPLY) {
      load_tile(Tin, p.in0, t, r1, pitch);
      __syncthreads();
      const int nout = nrows * q;
      for (int idx = threadIdx.x; idx < nout; idx += kPT) {
        const int r = idx / q, c = idx - r * q;
        if (CPLX) {
          double sr = 0.0, si = 0.0;
          for (int a = 0; a < q; ++a) {
            const double ar = Tin[r * pitch + 2 * a], ai = Tin[r * pitch + 2 * a + 1];
            const double mr = Ms[(a * q + c) * 2], mi = Ms[(a * q + c) * 2 + 1];
            sr += ar * mr - ai * mi;
            si += ar * mi + ai * mr;
          }
          Tout[r * pitch + 2 * c] = sr;
          Tout[r * pitch + 2 * c + 1] = si;
        } else {
          double s = 0.0;
          for (int a = 0; a < q; ++a) s += Tin[r * pitch + a] * Ms[a * q + c];
          Tout[r * pitch + c] = s;
        }
      }
    } else if (OP == OP_REWEIGHT) {
      // B = d*Y - (w(|V|) - wmax) * V * rho     (R/spEigen.R:170-177)
      const int nout = nrows * q;
      for (int idx = threadIdx.x; idx < nout; idx += kPT) {
        const int r = idx / q, c = idx - r * q;
        const size_t g = static_cast<size_t>(t + r) * pitch;
        if (CPLX) {
          const double vr = __ldcg(p.in1 + g + 2 * c), vi = __ldcg(p.in1 + g + 2 * c + 1);
          const double yr = __ldcg(p.in0 + g + 2 * c), yi = __ldcg(p.in0 + g + 2 * c + 1);
          const double tt = weight_of_abs(hypot(vr, vi), p.sc) - Ms[2 * kMaxCols + c];
          Tout[r * pitch + 2 * c] = Ms[c] * yr - tt * vr * Ms[kMaxCols + c];
          Tout[r * pitch + 2 * c + 1] = Ms[c] * yi - tt * vi * Ms[kMaxCols + c];
        } else {
          const double v = __ldcg(p.in1 + g + c), y = __ldcg(p.in0 + g + c);
          const double tt = weight_of_abs(fabs(v), p.sc) - Ms[2 * kMaxCols + c];
          Tout[r * pitch + c] = Ms[c] * y - tt * v * Ms[kMaxCols + c];
        }
      }
    } else {
      // V0 = V - 2 a R + a^2 U      (R/spEigen.R:121-127)
      const int nout = nrows * ncr;
      const double a = p.a, a2 = p.a * p.a;
      for (int idx = threadIdx.x; idx < nout; idx += kPT) {
        const int r = idx / ncr, c = idx - r * ncr;
        const size_t g = static_cast<size_t>(t + r) * pitch + c;
        const double v = __ldcg(p.in0 + g), v1 = __ldcg(p.in1 + g), v2 = __ldcg(p.in2 + g);
        const double R = v1 - v;
        const double U = v2 - v1 - R;
        Tout[r * pitch + c] = v - 2.0 * a * R + a2 * U;
      }
    }
    __syncthreads();
    // write the tile
    {
      const int nout = nrows * ncr;
      for (int idx = threadIdx.x; idx < nout; idx += kPT) {
        const int r = idx / ncr, c = idx - r * ncr;
        p.Out[static_cast<size_t>(t + r) * pitch + c] = Tout[r * pitch + c];
      }
    }
    if (p.gram_part) gram_accumulate<CPLX>(gacc, Tout, Tout, q, pitch, nrows);
    if (p.absmin_part && threadIdx.x < q) {
      const int c = threadIdx.x;
      for (int r = 0; r < nrows; ++r) {
        const double v = CPLX ? hypot(Tout[r * pitch + 2 * c], Tout[r * pitch + 2 * c + 1]) : fabs(Tout[r * pitch + c]);
        amin = fmin(amin, v);
      }
    }
    __syncthreads();
  }
  if (p.gram_part) gram_store<CPLX>(gacc, p.gram_part, q);
  if (p.absmin_part && threadIdx.x < q) p.absmin_part[static_cast<size_t>(blockIdx.x) * kMaxCols + threadIdx.x] = amin;
}

template <bool CPLX>
__global__ void __launch_bounds__(kPT) k_absmin(const double* V, int64_t m, int q, int pitch, double* absmin_part) {
  // thread layout: 8 row lanes x 32 column slots
  __shared__ double s_min[8][kMaxCols];
  int64_t r0, r1;
  block_rows(m, r0, r1);
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  for (int cc = c; cc < q; cc += 32) {
    double amin = DBL_MAX;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const double* row = V + r * pitch;
      const double v = CPLX ? hypot(__ldcg(row + 2 * cc), __ldcg(row + 2 * cc + 1)) : fabs(__ldcg(row + cc));
      amin = fmin(amin, v);
    }
    s_min[rl][cc] = amin;
  }
  __syncthreads();
  if (threadIdx.x < q) {
    double amin = s_min[0][threadIdx.x];
    for (int k = 1; k < 8; ++k) amin = fmin(amin, s_min[k][threadIdx.x]);
    absmin_part[static_cast<size_t>(blockIdx.x) * kMaxCols + threadIdx.x] = amin;
  }
}

__global__ void k_wmax(const double* absmin_part, int nparts, int q, StageConsts sc, double* wmax) {
  const int c = threadIdx.x;
  if (c >= q) return;
  double amin = DBL_MAX;
  for (int b = 0; b < nparts; ++b) amin = fmin(amin, absmin_part[static_cast<size_t>(b) * kMaxCols + c]);
  wmax[c] = weight_of_abs(amin, sc);
}

// ------------------------------------------------------------------------------------------------
// block-level fixed-order sum of one value per thread (all threads get nothing; result in s_out by thread 0)
__device__ __forceinline__ double block_sum_fixed(double v, double* s_buf) {
  // warp butterfly (deterministic), then thread 0 sums the 8 warp totals in order
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kPT / 32; ++w) s += s_buf[w];
  __syncthreads();
  return s;   // valid in thread 0
}

__global__ void __launch_bounds__(kPT) k_squarem_norms(const double* V, const double* V1, const double* V2, int64_t m,
                                                       int ncr, int pitch, double* col_part, int* ticket, double* norms) {
  __shared__ double s_buf[kPT / 32];
  int64_t r0, r1;
  block_rows(m, r0, r1);
  double sR = 0.0, sU = 0.0;
  const int64_t n = (r1 - r0) * ncr;
  for (int64_t idx = threadIdx.x; idx < n; idx += kPT) {
    const int64_t r = idx / ncr;
    const int c = static_cast<int>(idx - r * ncr);
    const size_t g = static_cast<size_t>(r0 + r) * pitch + c;
    const double v = __ldcg(V + g), v1 = __ldcg(V1 + g), v2 = __ldcg(V2 + g);
    const double R = v1 - v;
    const double U = v2 - v1 - R;
    sR += R * R;
    sU += U * U;
  }
  const double tR = block_sum_fixed(sR, s_buf);
  const double tU = block_sum_fixed(sU, s_buf);
  if (threadIdx.x == 0) {
    col_part[blockIdx.x * 2] = tR;
    col_part[blockIdx.x * 2 + 1] = tU;
  }
  if (last_block_done(ticket) && threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) { a += __ldcg(col_part + 2 * k); b += __ldcg(col_part + 2 * k + 1); }
    norms[0] = sqrt(a);
    norms[1] = sqrt(b);
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(kPT) k_objective(const double* V0, int64_t m, int q, int pitch, StageConsts sc,
                                                   const double* dots_part, int n_dot_parts, const double* d,
                                                   const double* rho, double* col_part, int* ticket, double* out) {
  __shared__ double s_g[8][kMaxCols];
  int64_t r0, r1;
  block_rows(m, r0, r1);
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const double inv_eps_c2 = 1.0 / (sc.eps * sc.c2);   // R: abs(V0)^2 / (epsi*c2)
  for (int cc = c; cc < q; cc += 32) {
    double s = 0.0;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const double* row = V0 + r * pitch;
      const double a = CPLX ? hypot(__ldcg(row + 2 * cc), __ldcg(row + 2 * cc + 1)) : fabs(__ldcg(row + cc));
      // g of R/spEigen.R:133-134
      s += (a <= sc.eps) ? (a * a) / (sc.eps * sc.c2) : log((sc.p + a) / (sc.p + sc.eps)) / sc.c1 + sc.eps / sc.c2;
    }
    s_g[rl][cc] = s;
  }
  (void)inv_eps_c2;
  __syncthreads();
  if (threadIdx.x < q) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += s_g[k][threadIdx.x];
    col_part[static_cast<size_t>(blockIdx.x) * kMaxCols + threadIdx.x] = s;
  }
  if (last_block_done(ticket)) {
    __shared__ double s_quad[kMaxCols], s_gs[kMaxCols];
    if (threadIdx.x < q) {
      const int cc = threadIdx.x;
      double g = 0.0;
      for (unsigned k = 0; k < gridDim.x; ++k) g += __ldcg(col_part + static_cast<size_t>(k) * kMaxCols + cc);
      double qd = 0.0;
      for (int k = 0; k < n_dot_parts; ++k) qd += __ldcg(dots_part + static_cast<size_t>(k) * q + cc);
      s_quad[cc] = qd;
      s_gs[cc] = g;
      out[1 + cc] = qd;
      out[1 + q + cc] = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double fq = 0.0, fg = 0.0;
      for (int cc = 0; cc < q; ++cc) { fq += s_quad[cc] * d[cc]; fg += s_gs[cc] * rho[cc]; }
      out[0] = fq - fg;   // R/spEigen.R:136,138
    }
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(kPT) k_thres_norms(const double* V, int64_t m, int q, int pitch, double thres,
                                                     double* col_part, int* ticket, double* norms) {
  __shared__ double s_g[8][kMaxCols];
  int64_t r0, r1;
  block_rows(m, r0, r1);
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  for (int cc = c; cc < q; cc += 32) {
    double s = 0.0;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const double* row = V + r * pitch;
      if (CPLX) {
        const double vr = __ldcg(row + 2 * cc), vi = __ldcg(row + 2 * cc + 1);
        const double a = hypot(vr, vi);
        if (!(a < thres)) s += a * a;     // abs(V)^2 as R computes it (Mod then square)
      } else {
        const double v = __ldcg(row + cc);
        if (!(fabs(v) < thres)) s += v * v;
      }
    }
    s_g[rl][cc] = s;
  }
  __syncthreads();
  if (threadIdx.x < q) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += s_g[k][threadIdx.x];
    col_part[static_cast<size_t>(blockIdx.x) * kMaxCols + threadIdx.x] = s;
  }
  if (last_block_done(ticket) && threadIdx.x < q) {
    double s = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) s += __ldcg(col_part + static_cast<size_t>(k) * kMaxCols + threadIdx.x);
    norms[threadIdx.x] = 1.0 / sqrt(s);   // nrm of R/spEigen.R:159
  }
}

template <bool CPLX>
__global__ void k_thres_scale_colmajor(const double* V, int64_t m, int q, int pitch, double thres, const double* nrm,
                                       double* out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= m * q) return;
  const int c = static_cast<int>(idx / m);
  const int64_t r = idx - static_cast<int64_t>(c) * m;
  const double* row = V + r * pitch;
  if (CPLX) {
    double vr = row[2 * c], vi = row[2 * c + 1];
    if (hypot(vr, vi) < thres) { vr = 0.0; vi = 0.0; }
    out[2 * idx] = nrm[c] * vr;
    out[2 * idx + 1] = nrm[c] * vi;
  } else {
    double v = row[c];
    if (fabs(v) < thres) v = 0.0;
    out[idx] = nrm[c] * v;
  }
}

__global__ void k_panel_to_colmajor(const double* P, int64_t m, int q, int e, int pitch, double* out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= m * q) return;
  const int c = static_cast<int>(idx / m);
  const int64_t r = idx - static_cast<int64_t>(c) * m;
  for (int k = 0; k < e; ++k) out[e * idx + k] = P[r * pitch + e * c + k];
}
__global__ void k_colmajor_to_panel(const double* in, int64_t m, int q, int e, int pitch, double* P) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= m * q) return;
  const int64_t r = idx / q;
  const int c = static_cast<int>(idx - r * q);
  for (int k = 0; k < e; ++k) P[r * pitch + e * c + k] = in[e * (static_cast<int64_t>(c) * m + r) + k];
}

__global__ void k_hat(const double* V, int64_t m, int q, int pitch, int conj_form, double* Vhat) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= m * q) return;
  const int64_t r = idx / q;
  const int c = static_cast<int>(idx - r * q);
  const double vr = V[r * pitch + 2 * c], vi = V[r * pitch + 2 * c + 1];
  double* h0 = Vhat + (2 * r) * pitch + 2 * c;
  double* h1 = Vhat + (2 * r + 1) * pitch + 2 * c;
  h0[0] = vr;
  h0[1] = vi;
  if (conj_form) { h1[0] = vi; h1[1] = -vr; }    // conj(a) * u : imaginary part of a multiplies -i*u
  else           { h1[0] = -vi; h1[1] = vr; }    // a * u       : imaginary part of a multiplies +i*u
}

__global__ void k_tr_combine(const double* Pm, int64_t n, int q, int pitch, double* That) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * q) return;
  const int64_t i = idx / q;
  const int c = static_cast<int>(idx - i * q);
  const double* p0 = Pm + (2 * i) * pitch + 2 * c;
  const double* p1 = Pm + (2 * i + 1) * pitch + 2 * c;
  const double rr = p0[0], ri = p0[1], ir = p1[0], ii = p1[1];
  const double tr = rr - ii, ti = ri + ir;
  double* h0 = That + (2 * i) * pitch + 2 * c;
  double* h1 = That + (2 * i + 1) * pitch + 2 * c;
  h0[0] = tr; h0[1] = ti;
  h1[0] = ti; h1[1] = -tr;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_random_panel(int64_t m, int ncr, int pitch, uint64_t seed, double* P) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= m * ncr) return;
  const int64_t r = idx / ncr;
  const int c = static_cast<int>(idx - r * ncr);
  const uint64_t h = splitmix64(splitmix64(seed ^ (static_cast<uint64_t>(r) * 0x100000001B3ull)) + c);
  P[r * pitch + c] = (static_cast<double>(h >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}

template <bool CPLX>
__global__ void __launch_bounds__(kPT) k_residual_norms(const double* ZW, const double* QW, const double* evals, int64_t m,
                                                        int q, int pitch, double* col_part, int* ticket, double* out) {
  __shared__ double s_g[8][kMaxCols];
  int64_t r0, r1;
  block_rows(m, r0, r1);
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  for (int cc = c; cc < q; cc += 32) {
    const double th = evals[cc];
    double s = 0.0;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      if (CPLX) {
        const double dr = ZW[r * pitch + 2 * cc] - th * QW[r * pitch + 2 * cc];
        const double di = ZW[r * pitch + 2 * cc + 1] - th * QW[r * pitch + 2 * cc + 1];
        s += dr * dr + di * di;
      } else {
        const double dr = ZW[r * pitch + cc] - th * QW[r * pitch + cc];
        s += dr * dr;
      }
    }
    s_g[rl][cc] = s;
  }
  __syncthreads();
  if (threadIdx.x < q) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += s_g[k][threadIdx.x];
    col_part[static_cast<size_t>(blockIdx.x) * kMaxCols + threadIdx.x] = s;
  }
  if (last_block_done(ticket) && threadIdx.x < q) {
    double s = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) s += __ldcg(col_part + static_cast<size_t>(k) * kMaxCols + threadIdx.x);
    out[threadIdx.x] = sqrt(s);
  }
}

__global__ void k_take_columns(const double* In, int64_t m, int ncr_out, int pitch_in, int pitch_out, double* Out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= m * ncr_out) return;
  const int64_t r = idx / ncr_out;
  const int c = static_cast<int>(idx - r * ncr_out);
  Out[r * pitch_out + c] = In[r * pitch_in + c];
}

__global__ void k_sum_panels(const double* In, int nparts, size_t part_stride, int64_t n, double* Out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double s = 0.0;
  for (int r = 0; r < nparts; ++r) s += In[r * part_stride + idx];   // fixed rank order
  Out[idx] = s;
}

// ------------------------------------------------------------------------------------------------
// K3: q x q Hermitian eigen-solver, single block.  One-sided (Hestenes) Jacobi on F = C: warp w owns
// one column pair of the round-robin ordering, lane i holds rows i and i+32 of both columns of F and
// of the accumulated rotations W; the three inner products per pair are warp-shuffle butterflies.
// Replaces the q x q core of LAPACK's svd()/eigen() at R/spEigen.R:76,130,180.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

template <bool CPLX>
__global__ void __launch_bounds__(1024) k_small(const double* gram_part, int nparts, int q, int mode, double* Msmall,
                                                double* evals, int* status) {
  extern __shared__ double sm[];
  constexpr int E = CPLX ? 2 : 1;
  double* F = sm;                      // column-major: F[(c*q + i)*E + e]
  double* W = F + q * q * E;
  double* lam = W + q * q * E;         // [q]
  int* rank = reinterpret_cast<int*>(lam + q);
  __shared__ int s_rot;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;

  // fixed-order sum of the stage-1 Gram partials; C[a][b] sits at part[(a*q+b)*E]
  for (int e = tid; e < q * q; e += nth) {
    const int a = e / q, b = e - a * q;
    double sr = 0.0, si = 0.0;
    for (int k = 0; k < nparts; ++k) {
      const double* src = gram_part + (static_cast<size_t>(k) * q * q + e) * E;
      sr += __ldcg(src);
      if (CPLX) si += __ldcg(src + 1);
    }
    F[(b * q + a) * E] = sr;
    if (CPLX) F[(b * q + a) * E + 1] = si;
    W[(b * q + a) * E] = (a == b) ? 1.0 : 0.0;
    if (CPLX) W[(b * q + a) * E + 1] = 0.0;
  }
  if (tid == 0) s_rot = 0;
  __syncthreads();

  const int qe = q + (q & 1);
  const int npairs = qe / 2;
  const int n1 = qe - 1;
  const double tol = 8.0 * DBL_EPSILON;
  const int i0 = lane, i1 = lane + 32;
  for (int sweep = 0; sweep < 60; ++sweep) {
    for (int s = 0; s < n1; ++s) {
      if (warp < npairs) {
        int pa, pb;
        if (warp == 0) { pa = s; pb = qe - 1; }
        else { pa = (s + warp) % n1; pb = (s - warp + n1) % n1; }
        if (pa > pb) { const int t = pa; pa = pb; pb = t; }
        if (pb < q) {
          double x[2][E], y[2][E];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = h ? i1 : i0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
              x[h][e] = (i < q) ? F[(pa * q + i) * E + e] : 0.0;
              y[h][e] = (i < q) ? F[(pb * q + i) * E + e] : 0.0;
            }
          }
          double al = 0.0, be = 0.0, gr = 0.0, gi = 0.0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (CPLX) {
              al += x[h][0] * x[h][0] + x[h][E - 1] * x[h][E - 1];
              be += y[h][0] * y[h][0] + y[h][E - 1] * y[h][E - 1];
              gr += x[h][0] * y[h][0] + x[h][E - 1] * y[h][E - 1];     // conj(x) * y
              gi += x[h][0] * y[h][E - 1] - x[h][E - 1] * y[h][0];
            } else {
              al += x[h][0] * x[h][0];
              be += y[h][0] * y[h][0];
              gr += x[h][0] * y[h][0];
            }
          }
          al = warp_sum(al); be = warp_sum(be); gr = warp_sum(gr);
          if (CPLX) gi = warp_sum(gi);
          const double g = CPLX ? hypot(gr, gi) : fabs(gr);
          if (al > 0.0 && be > 0.0 && g > tol * sqrt(al) * sqrt(be)) {
            const double er = gr / g, ei = CPLX ? gi / g : 0.0;
            const double zeta = (be - al) / (2.0 * g);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t);
            const double sn = cs * t;
            // x' = cs*x - sn*conj(e)*y ; y' = sn*e*x + cs*y    (applied to F and W columns)
#pragma unroll
            for (int mat = 0; mat < 2; ++mat) {
              double* M = mat ? W : F;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int i = h ? i1 : i0;
                if (i < q) {
                  double xr, xi = 0.0, yr, yi = 0.0;
                  if (mat == 0) { xr = x[h][0]; yr = y[h][0]; if (CPLX) { xi = x[h][E - 1]; yi = y[h][E - 1]; } }
                  else {
                    xr = M[(pa * q + i) * E]; yr = M[(pb * q + i) * E];
                    if (CPLX) { xi = M[(pa * q + i) * E + E - 1]; yi = M[(pb * q + i) * E + E - 1]; }
                  }
                  if (CPLX) {
                    // conj(e)*y = (er - i ei)(yr + i yi) ;  e*x = (er + i ei)(xr + i xi)
                    const double cyr = er * yr + ei * yi, cyi = er * yi - ei * yr;
                    const double exr = er * xr - ei * xi, exi = er * xi + ei * xr;
                    M[(pa * q + i) * E] = cs * xr - sn * cyr;
                    M[(pa * q + i) * E + E - 1] = cs * xi - sn * cyi;
                    M[(pb * q + i) * E] = sn * exr + cs * yr;
                    M[(pb * q + i) * E + E - 1] = sn * exi + cs * yi;
                  } else {
                    M[(pa * q + i) * E] = cs * xr - sn * er * yr;
                    M[(pb * q + i) * E] = sn * er * xr + cs * yr;
                  }
                }
              }
            }
            if (lane == 0) s_rot = 1;
          }
        }
      }
      __syncthreads();
    }
    const int rot = s_rot;
    __syncthreads();
    if (tid == 0) s_rot = 0;
    __syncthreads();
    if (!rot) break;
  }

  // eigenvalues as Rayleigh quotients lam_c = Re(w_c^H f_c)  (f_c = C w_c)
  for (int c = warp; c < q; c += (nth >> 5)) {
    double s = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = h ? i1 : i0;
      if (i < q) {
        s += W[(c * q + i) * E] * F[(c * q + i) * E];
        if (CPLX) s += W[(c * q + i) * E + E - 1] * F[(c * q + i) * E + E - 1];
      }
    }
    s = warp_sum(s);
    if (lane == 0) lam[c] = s;
  }
  __syncthreads();

  if (mode == SMALL_INVSQRT) {
    if (tid == 0) {
      double lmax = 0.0;
      for (int c = 0; c < q; ++c) lmax = fmax(lmax, lam[c]);
      int bad = 0;
      for (int c = 0; c < q; ++c) {
        if (!(lam[c] > lmax * 1e-300) || !isfinite(lam[c])) { bad = 1; lam[c] = 0.0; }
        else lam[c] = 1.0 / sqrt(lam[c]);
      }
      if (bad) atomicOr(status, 1);
    }
    __syncthreads();
    // M[a][b] = sum_c W[a][c] * lam_c * conj(W[b][c])
    for (int e = tid; e < q * q; e += nth) {
      const int a = e / q, b = e - a * q;
      double sr = 0.0, si = 0.0;
      for (int c = 0; c < q; ++c) {
        const double war = W[(c * q + a) * E], wbr = W[(c * q + b) * E];
        if (CPLX) {
          const double wai = W[(c * q + a) * E + E - 1], wbi = W[(c * q + b) * E + E - 1];
          sr += lam[c] * (war * wbr + wai * wbi);
          si += lam[c] * (wai * wbr - war * wbi);
        } else {
          sr += lam[c] * war * wbr;
        }
      }
      Msmall[e * E] = sr;
      if (CPLX) Msmall[e * E + E - 1] = si;
    }
  } else {
    if (tid < q) {
      int r = 0;
      for (int c = 0; c < q; ++c) r += (lam[c] > lam[tid]) || (lam[c] == lam[tid] && c < tid);
      rank[tid] = r;
      evals[r] = lam[tid];
    }
    __syncthreads();
    for (int e = tid; e < q * q; e += nth) {
      const int a = e / q, c = e - a * q;
      const int rc = rank[c];
      Msmall[(a * q + rc) * E] = W[(c * q + a) * E];
      if (CPLX) Msmall[(a * q + rc) * E + E - 1] = W[(c * q + a) * E + E - 1];
    }
  }
}


template <bool CPLX>
__global__ void __launch_bounds__(kPT) k_coldots(const double* V, const double* Y, int64_t m, int q, int pitch,
                                                 double* dots_part) {
  __shared__ double s_g[8][kMaxCols];
  int64_t r0, r1;
  block_rows(m, r0, r1);
  const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
  for (int cc = c; cc < q; cc += 32) {
    double s = 0.0;
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      if (CPLX) s += V[r * pitch + 2 * cc] * Y[r * pitch + 2 * cc] + V[r * pitch + 2 * cc + 1] * Y[r * pitch + 2 * cc + 1];
      else s += V[r * pitch + cc] * Y[r * pitch + cc];
    }
    s_g[rl][cc] = s;
  }
  __syncthreads();
  if (threadIdx.x < q) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += s_g[k][threadIdx.x];
    dots_part[static_cast<size_t>(blockIdx.x) * q + threadIdx.x] = s;
  }
}

__global__ void k_sum_parts(const double* parts, int nparts, int len, double* out) {
  const int i = threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int k = 0; k < nparts; ++k) s += parts[static_cast<size_t>(k) * len + i];
  out[i] = s;
}

size_t produce_smem(const PanelDims& pd, int op) {
  size_t n = 2 * static_cast<size_t>(kTileRows) * pd.pitch;
  if (op == OP_APPLY) n += static_cast<size_t>(pd.q) * pd.q * (pd.cplx ? 2 : 1);
  else n += 3 * kMaxCols;
  return n * sizeof(double);
}

template <int OP>
int launch_produce(const PanelDims& pd, const ProdArgs& pa, cudaStream_t st) {
  const size_t smem = produce_smem(pd, OP);
  if (pd.cplx) {
    static bool set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& set = set_dev[dev_ & 63];
    if (!set) { SPE_CUDA(cudaFuncSetAttribute(k_produce<OP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set = true; }
    k_produce<OP, true><<<kPanelBlocks, kPT, smem, st>>>(pa);
  } else {
    static bool set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& set = set_dev[dev_ & 63];
    if (!set) { SPE_CUDA(cudaFuncSetAttribute(k_produce<OP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set = true; }
    k_produce<OP, false><<<kPanelBlocks, kPT, smem, st>>>(pa);
  }
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

ProdArgs base_args(const PanelDims& pd) {
  ProdArgs pa{};
  pa.m = pd.m; pa.q = pd.q; pa.ncr = pd.ncr; pa.pitch = pd.pitch;
  return pa;
}

inline int nblk(int64_t n, int t = 256) { return static_cast<int>((n + t - 1) / t); }

}  // namespace

int launch_gram(const PanelDims& pd, const double* A, const double* B, double* gram_part, cudaStream_t st) {
  const size_t smem = (A == B ? 1 : 2) * static_cast<size_t>(kTileRows) * pd.pitch * sizeof(double);
  if (pd.cplx) k_gram<true><<<kPanelBlocks, kPT, smem, st>>>(A, B, pd.m, pd.q, pd.pitch, gram_part);
  else k_gram<false><<<kPanelBlocks, kPT, smem, st>>>(A, B, pd.m, pd.q, pd.pitch, gram_part);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_small(const PanelDims& pd, const double* gram_part, int nparts, int mode, double* Msmall, double* evals,
                 int* status, cudaStream_t st) {
  const int q = pd.q;
  const int qe = q + (q & 1);
  int threads = 32 * (qe / 2);
  if (threads < 64) threads = 64;
  if (threads > 1024) threads = 1024;
  const size_t smem = (2 * static_cast<size_t>(q) * q * (pd.cplx ? 2 : 1) + q) * sizeof(double) + q * sizeof(int) + 16;
  if (pd.cplx) {
    static bool set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& set = set_dev[dev_ & 63];
    if (!set) { SPE_CUDA(cudaFuncSetAttribute(k_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set = true; }
    k_small<true><<<1, threads, smem, st>>>(gram_part, nparts, q, mode, Msmall, evals, status);
  } else {
    static bool set_dev[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& set = set_dev[dev_ & 63];
    if (!set) { SPE_CUDA(cudaFuncSetAttribute(k_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set = true; }
    k_small<false><<<1, threads, smem, st>>>(gram_part, nparts, q, mode, Msmall, evals, status);
  }
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_apply(const PanelDims& pd, const double* A, const double* Msmall, double* Out, double* gram_part_out,
                 double* absmin_part_out, cudaStream_t st) {
  ProdArgs pa = base_args(pd);
  pa.in0 = A; pa.Msmall = Msmall; pa.Out = Out; pa.gram_part = gram_part_out; pa.absmin_part = absmin_part_out;
  return launch_produce<OP_APPLY>(pd, pa, st);
}

int launch_reweight(const PanelDims& pd, const double* Y, const double* V, const double* d, const double* rho,
                    const double* wmax, StageConsts sc, double* B, double* gram_part_out, cudaStream_t st) {
  ProdArgs pa = base_args(pd);
  pa.in0 = Y; pa.in1 = V; pa.d = d; pa.rho = rho; pa.wmax = wmax; pa.sc = sc; pa.Out = B; pa.gram_part = gram_part_out;
  return launch_produce<OP_REWEIGHT>(pd, pa, st);
}

int launch_extrapolate(const PanelDims& pd, const double* V, const double* V1, const double* V2, double a, double* Out,
                       double* gram_part_out, cudaStream_t st) {
  ProdArgs pa = base_args(pd);
  pa.in0 = V; pa.in1 = V1; pa.in2 = V2; pa.a = a; pa.Out = Out; pa.gram_part = gram_part_out;
  return launch_produce<OP_EXTRAP>(pd, pa, st);
}

int launch_absmin(const PanelDims& pd, const double* V, double* absmin_part, cudaStream_t st) {
  if (pd.cplx) k_absmin<true><<<kPanelBlocks, kPT, 0, st>>>(V, pd.m, pd.q, pd.pitch, absmin_part);
  else k_absmin<false><<<kPanelBlocks, kPT, 0, st>>>(V, pd.m, pd.q, pd.pitch, absmin_part);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_wmax(const PanelDims& pd, const double* absmin_part, StageConsts sc, double* wmax, cudaStream_t st) {
  k_wmax<<<1, kMaxCols, 0, st>>>(absmin_part, kPanelBlocks, pd.q, sc, wmax);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_squarem_norms(const PanelDims& pd, const double* V, const double* V1, const double* V2,
                         const PanelScratch& sc, double* norms, cudaStream_t st) {
  k_squarem_norms<<<kPanelBlocks, kPT, 0, st>>>(V, V1, V2, pd.m, pd.ncr, pd.pitch, sc.col_part, sc.ticket, norms);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_objective(const PanelDims& pd, const double* V0, StageConsts stc, const double* dots_part, int n_dot_parts,
                     const double* d, const double* rho, const PanelScratch& sc, double* out, cudaStream_t st) {
  if (pd.cplx)
    k_objective<true><<<kPanelBlocks, kPT, 0, st>>>(V0, pd.m, pd.q, pd.pitch, stc, dots_part, n_dot_parts, d, rho,
                                                    sc.col_part, sc.ticket, out);
  else
    k_objective<false><<<kPanelBlocks, kPT, 0, st>>>(V0, pd.m, pd.q, pd.pitch, stc, dots_part, n_dot_parts, d, rho,
                                                     sc.col_part, sc.ticket, out);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_finalize(const PanelDims& pd, const double* V, double thres, const PanelScratch& sc, double* norms_tmp,
                    double* out_colmajor, cudaStream_t st) {
  if (pd.cplx) {
    k_thres_norms<true><<<kPanelBlocks, kPT, 0, st>>>(V, pd.m, pd.q, pd.pitch, thres, sc.col_part, sc.ticket, norms_tmp);
    k_thres_scale_colmajor<true><<<nblk(pd.m * pd.q), 256, 0, st>>>(V, pd.m, pd.q, pd.pitch, thres, norms_tmp, out_colmajor);
  } else {
    k_thres_norms<false><<<kPanelBlocks, kPT, 0, st>>>(V, pd.m, pd.q, pd.pitch, thres, sc.col_part, sc.ticket, norms_tmp);
    k_thres_scale_colmajor<false><<<nblk(pd.m * pd.q), 256, 0, st>>>(V, pd.m, pd.q, pd.pitch, thres, norms_tmp, out_colmajor);
  }
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_panel_to_colmajor(const PanelDims& pd, const double* P, double* out_colmajor, cudaStream_t st) {
  k_panel_to_colmajor<<<nblk(pd.m * pd.q), 256, 0, st>>>(P, pd.m, pd.q, pd.cplx ? 2 : 1, pd.pitch, out_colmajor);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_colmajor_to_panel(const PanelDims& pd, const double* in_colmajor, double* P, cudaStream_t st) {
  k_colmajor_to_panel<<<nblk(pd.m * pd.q), 256, 0, st>>>(in_colmajor, pd.m, pd.q, pd.cplx ? 2 : 1, pd.pitch, P);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_hat(const PanelDims& pd, const double* V, double* Vhat, int conj_form, cudaStream_t st) {
  k_hat<<<nblk(pd.m * pd.q), 256, 0, st>>>(V, pd.m, pd.q, pd.pitch, conj_form, Vhat);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_tr_combine(int64_t n, int q, const double* Pm, double* That, cudaStream_t st) {
  k_tr_combine<<<nblk(n * q), 256, 0, st>>>(Pm, n, q, panel_pitch(2 * q), That);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_random_panel(const PanelDims& pd, uint64_t seed, double* P, cudaStream_t st) {
  k_random_panel<<<nblk(pd.m * pd.ncr), 256, 0, st>>>(pd.m, pd.ncr, pd.pitch, seed, P);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_residual_norms(const PanelDims& pd, const double* ZW, const double* QW, const double* evals,
                          const PanelScratch& sc, double* out, cudaStream_t st) {
  if (pd.cplx) k_residual_norms<true><<<kPanelBlocks, kPT, 0, st>>>(ZW, QW, evals, pd.m, pd.q, pd.pitch, sc.col_part, sc.ticket, out);
  else k_residual_norms<false><<<kPanelBlocks, kPT, 0, st>>>(ZW, QW, evals, pd.m, pd.q, pd.pitch, sc.col_part, sc.ticket, out);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_take_columns(const PanelDims& pd_in, const double* In, const PanelDims& pd_out, double* Out, cudaStream_t st) {
  k_take_columns<<<nblk(pd_out.m * pd_out.ncr), 256, 0, st>>>(In, pd_out.m, pd_out.ncr, pd_in.pitch, pd_out.pitch, Out);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_sum_panels(const PanelDims& pd, const double* In, int nparts, size_t part_stride, double* Out, cudaStream_t st) {
  const int64_t n = padded_rows(pd.m) * pd.pitch;
  k_sum_panels<<<nblk(n), 256, 0, st>>>(In, nparts, part_stride, n, Out);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

int launch_coldots(const PanelDims& pd, const double* V, const double* Y, double* dots_part, cudaStream_t st) {
  if (pd.cplx) k_coldots<true><<<kPanelBlocks, kPT, 0, st>>>(V, Y, pd.m, pd.q, pd.pitch, dots_part);
  else k_coldots<false><<<kPanelBlocks, kPT, 0, st>>>(V, Y, pd.m, pd.q, pd.pitch, dots_part);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_sum_parts(const double* parts, int nparts, int len, double* out, cudaStream_t st) {
  k_sum_parts<<<1, kMaxCols * 2, 0, st>>>(parts, nparts, len, out);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

}  // namespace speigen
