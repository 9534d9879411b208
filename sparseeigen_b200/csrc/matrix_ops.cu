#include "matrix_ops.cuh"
#include "host_pack.h"
#include "../../include/speigen_b200.h"
#include <cfloat>

namespace speigen {
void note_launch(int n);
namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 32 x 32 element tiles, 32 x 8 threads.  E = doubles per element.
struct ScanPeers {
  const double* A[8];   // shard base pointers of every row block (as mapped into this device); [0] = own for 1 GPU
  int64_t rps;          // rows per shard
};

template <int E>
__global__ void __launch_bounds__(256) k_scan_cov(const double* A, int64_t rows, int64_t m, int64_t ld, int64_t lo,
                                                  int full_check, const ScanPeers peers, double* part, int* ticket,
                                                  double* out4) {
  __shared__ double tileT[32][32 * E + 1];
  __shared__ double s_red[5][8];
  __shared__ int s_is_last;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tr = (rows + 31) / 32, tc = (m + 31) / 32;
  const int64_t ntiles = tr * tc;
  double sdiff = 0.0, sabs = 0.0, dmax = -DBL_MAX, nnan = 0.0, dmin = DBL_MAX;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t tj = t / tc, ti = t - tj * tc;
    if (full_check) {
      // transposed tile: global rows ti*32.. (owned by shard g/rps), global columns lo + tj*32..
      for (int r = ty; r < 32; r += 8) {
        const int64_t gg = ti * 32 + r, ii = lo + tj * 32 + tx;
        const bool ok = gg < m && (tj * 32 + tx) < rows;
        const double* src = ok ? peers.A[gg / peers.rps] + (gg % peers.rps) * ld + ii * E : nullptr;
#pragma unroll
        for (int e = 0; e < E; ++e) tileT[r][tx * E + e] = ok ? src[e] : 0.0;
      }
      __syncthreads();
    }
    for (int r = ty; r < 32; r += 8) {
      const int64_t j = tj * 32 + r, i = ti * 32 + tx;
      if (j < rows && i < m) {
        double a[E];
#pragma unroll
        for (int e = 0; e < E; ++e) a[e] = A[j * ld + i * E + e];
        const double ab = (E == 2) ? hypot(a[0], a[E - 1]) : fabs(a[0]);
        if (isnan(ab)) nnan += 1.0;
        sabs += ab;
        if (full_check) {
          // S[i][j] element lives at tileT[tx][r]; compare with conj
          const double br = tileT[tx][r * E], bi = (E == 2) ? -tileT[tx][r * E + E - 1] : 0.0;
          sdiff += (E == 2) ? hypot(a[0] - br, a[E - 1] - bi) : fabs(a[0] - br);
        }
        if (i == lo + j) { dmax = fmax(dmax, a[0]); dmin = fmin(dmin, a[0]); }
      }
    }
    if (full_check) __syncthreads();
  }
  sdiff = warp_sum_d(sdiff); sabs = warp_sum_d(sabs); nnan = warp_sum_d(nnan); dmax = warp_max_d(dmax);
  dmin = -warp_max_d(-dmin);
  if (tx == 0) { s_red[0][ty] = sdiff; s_red[1][ty] = sabs; s_red[2][ty] = dmax; s_red[3][ty] = nnan; s_red[4][ty] = dmin; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = -DBL_MAX, d = 0, mn = DBL_MAX;
    for (int w = 0; w < 8; ++w) { a += s_red[0][w]; b += s_red[1][w]; c = fmax(c, s_red[2][w]); d += s_red[3][w]; mn = fmin(mn, s_red[4][w]); }
    part[blockIdx.x * 8 + 0] = a; part[blockIdx.x * 8 + 1] = b; part[blockIdx.x * 8 + 2] = c; part[blockIdx.x * 8 + 3] = d;
    part[blockIdx.x * 8 + 4] = mn;
    __threadfence();
    const int tk = atomicAdd(ticket, 1);
    s_is_last = (tk == (int)gridDim.x - 1);
    if (s_is_last) *ticket = 0;
  }
  __syncthreads();
  if (s_is_last && threadIdx.x == 0) {
    __threadfence();
    double a = 0, b = 0, c = -DBL_MAX, d = 0, mn = DBL_MAX;
    for (unsigned k = 0; k < gridDim.x; ++k) {
      a += __ldcg(part + k * 8); b += __ldcg(part + k * 8 + 1); c = fmax(c, __ldcg(part + k * 8 + 2)); d += __ldcg(part + k * 8 + 3);
      mn = fmin(mn, __ldcg(part + k * 8 + 4));
    }
    out4[0] = a; out4[1] = b; out4[2] = c; out4[3] = d; out4[4] = mn;      // out4[4]: min Re diag (a PD matrix has a positive diagonal)
  }
}

// Second half of the Hermitian-half upload (host_pack.h): every element whose super-tile did not cross PCIe is rebuilt as the
// conjugate of its transposed partner, which did - read from the shard that owns it (own matrix on one GPU, a peer's over
// NVLink when S is row-sharded inside one process).  Same 32 x 32 staging as the scan above; uploaded elements are never
// written and non-uploaded ones never read, so the kernels of different GPUs may run concurrently.
template <int E>
__global__ void __launch_bounds__(256) k_mirror_cov(double* A, int64_t rows, int64_t m, int64_t ld, int64_t lo, int64_t block,
                                                    const ScanPeers peers) {
  __shared__ double tileT[32][32 * E + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tr = (rows + 31) / 32, tc = (m + 31) / 32;
  const int64_t ntiles = tr * tc;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t tj = t / tc, ti = t - tj * tc;
    // a 32 x 32 tile meets at most 2 x 2 super-tiles: its corners tell whether anything in it has to be rebuilt
    const int64_t rf = lo + tj * 32, rl = lo + min(tj * 32 + 31, rows - 1), cf = ti * 32, cl = min(ti * 32 + 31, m - 1);
    const bool need = !half_tile_uploaded(rf, cf, block) || !half_tile_uploaded(rf, cl, block) ||
                      !half_tile_uploaded(rl, cf, block) || !half_tile_uploaded(rl, cl, block);
    if (!need) continue;
    for (int r = ty; r < 32; r += 8) {
      const int64_t gg = ti * 32 + r, jl = tj * 32 + tx, ii = lo + jl;      // source element (gg, ii) = partner of (ii, gg)
      const bool ok = gg < m && jl < rows && !half_tile_uploaded(ii, gg, block);
      const double* src = ok ? peers.A[gg / peers.rps] + (gg % peers.rps) * ld + ii * E : nullptr;
#pragma unroll
      for (int e = 0; e < E; ++e) tileT[r][tx * E + e] = ok ? src[e] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t j = tj * 32 + r, i = ti * 32 + tx;
      if (j < rows && i < m && !half_tile_uploaded(lo + j, i, block)) {
        A[j * ld + i * E] = tileT[tx][r * E];
        if (E == 2) A[j * ld + i * E + 1] = -tileT[tx][r * E + E - 1];
      }
    }
    __syncthreads();
  }
}

template <int E>
__global__ void __launch_bounds__(256) k_row_sums(const double* A, int64_t nloc, int64_t ld, double* sums) {
  __shared__ double s_red[E][8];
  const double* row = A + static_cast<int64_t>(blockIdx.x) * ld;
  double s[E];
#pragma unroll
  for (int e = 0; e < E; ++e) s[e] = 0.0;
  for (int64_t i = threadIdx.x; i < nloc; i += 256) {
#pragma unroll
    for (int e = 0; e < E; ++e) s[e] += row[i * E + e];
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    s[e] = warp_sum_d(s[e]);
    if ((threadIdx.x & 31) == 0) s_red[e][threadIdx.x >> 5] = s[e];
  }
  __syncthreads();
  if (threadIdx.x < E) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    sums[static_cast<int64_t>(blockIdx.x) * E + threadIdx.x] = t;
  }
}

template <int E>
__global__ void __launch_bounds__(256) k_center_sqnorm(double* A, int64_t nloc, int64_t ld, const double* sums,
                                                       double inv_n, double* sq) {
  __shared__ double s_red[8];
  double* row = A + static_cast<int64_t>(blockIdx.x) * ld;
  double mu[E];
#pragma unroll
  for (int e = 0; e < E; ++e) mu[e] = sums[static_cast<int64_t>(blockIdx.x) * E + e] * inv_n;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < nloc; i += 256) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const double v = row[i * E + e] - mu[e];
      row[i * E + e] = v;
      s += v * v;
    }
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    sq[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k_vec_max(const double* v, int64_t n, double* out2) {
  __shared__ double s_m[8], s_n[8];
  double mx = -DBL_MAX, nn = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const double x = v[i];
    if (isnan(x)) nn += 1.0; else mx = fmax(mx, x);
  }
  mx = warp_max_d(mx); nn = warp_sum_d(nn);
  if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = mx; s_n[threadIdx.x >> 5] = nn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mx = fmax(mx, s_m[w]); nn += s_n[w]; }
    out2[0] = mx; out2[1] = nn;
  }
}
}  // namespace

int launch_scan_cov(const double* A, int64_t rows, int64_t m, int64_t ld, int64_t lo, int cplx, int full_check,
                    const double* const* shard_ptrs, int nshards, int64_t rows_per_shard, double* part, int* ticket,
                    double* out4, cudaStream_t st) {
  ScanPeers peers{};
  for (int i = 0; i < 8; ++i) peers.A[i] = (shard_ptrs && i < nshards) ? shard_ptrs[i] : A;
  peers.rps = (shard_ptrs && nshards > 1) ? rows_per_shard : (m > 0 ? m : 1);
  if (cplx) k_scan_cov<2><<<kScanBlocks, 256, 0, st>>>(A, rows, m, ld, lo, full_check, peers, part, ticket, out4);
  else k_scan_cov<1><<<kScanBlocks, 256, 0, st>>>(A, rows, m, ld, lo, full_check, peers, part, ticket, out4);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_mirror_cov(double* A, int64_t rows, int64_t m, int64_t ld, int64_t lo, int cplx, int64_t block,
                      const double* const* shard_ptrs, int nshards, int64_t rows_per_shard, cudaStream_t st) {
  if (rows <= 0) return SPEIGEN_OK;
  ScanPeers peers{};
  for (int i = 0; i < 8; ++i) peers.A[i] = (shard_ptrs && i < nshards) ? shard_ptrs[i] : A;
  peers.rps = (shard_ptrs && nshards > 1) ? rows_per_shard : (m > 0 ? m : 1);
  if (cplx) k_mirror_cov<2><<<kScanBlocks, 256, 0, st>>>(A, rows, m, ld, lo, block, peers);
  else k_mirror_cov<1><<<kScanBlocks, 256, 0, st>>>(A, rows, m, ld, lo, block, peers);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_row_sums(const double* A, int64_t rows, int64_t nloc, int64_t ld, int cplx, double* sums, cudaStream_t st) {
  if (cplx) k_row_sums<2><<<(unsigned)rows, 256, 0, st>>>(A, nloc, ld, sums);
  else k_row_sums<1><<<(unsigned)rows, 256, 0, st>>>(A, nloc, ld, sums);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_center_sqnorm(double* A, int64_t rows, int64_t nloc, int64_t ld, int cplx, const double* sums, double inv_n,
                         double* sq, cudaStream_t st) {
  if (cplx) k_center_sqnorm<2><<<(unsigned)rows, 256, 0, st>>>(A, nloc, ld, sums, inv_n, sq);
  else k_center_sqnorm<1><<<(unsigned)rows, 256, 0, st>>>(A, nloc, ld, sums, inv_n, sq);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}
int launch_vec_max(const double* v, int64_t n, double* out2, cudaStream_t st) {
  k_vec_max<<<1, 256, 0, st>>>(v, n, out2);
  SPE_CUDA(cudaGetLastError());
  note_launch(1);
  return SPEIGEN_OK;
}

}  // namespace speigen
