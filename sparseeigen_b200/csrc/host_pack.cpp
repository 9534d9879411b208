// Host-only part of the Hermitian-half upload: unit planner and the tile compare of isSymmetric.matrix (R/spEigen.R:75).
// See host_pack.h.  Compiled without CUDA; the AVX2 bodies are selected at run time.
#include "host_pack.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace speigen {

std::vector<HalfUnit> plan_half_upload(int64_t m, int e, const std::vector<std::pair<int64_t, int64_t>>& dev_ranges,
                                       int64_t block, size_t slot_bytes) {
  std::vector<std::vector<HalfUnit>> per_dev(dev_ranges.size());
  const int64_t nb = (m + block - 1) / block;
  for (size_t dv = 0; dv < dev_ranges.size(); ++dv) {
    const int64_t lo = dev_ranges[dv].first, hi = std::min(dev_ranges[dv].second, m);
    if (hi <= lo) continue;
    for (int64_t I = lo / block; I * block < hi; ++I) {
      const int64_t r_begin = std::max(lo, I * block), r_end = std::min(hi, (I + 1) * block);
      for (int64_t J = 0; J < nb; ++J) {
        if (!half_tile_uploaded(I * block, J * block, block)) continue;
        const int64_t c0 = J * block, nc = std::min(block, m - c0);
        const size_t row_bytes = static_cast<size_t>(nc) * e * sizeof(double);
        const int64_t rows_per = std::max<int64_t>(1, static_cast<int64_t>(slot_bytes / row_bytes));
        for (int64_t r = r_begin; r < r_end; r += rows_per) {
          HalfUnit u;
          u.dev = static_cast<int>(dv);
          u.r0 = r;
          u.nr = std::min(rows_per, r_end - r);
          u.c0 = c0;
          u.nc = nc;
          u.diag = (I == J);
          per_dev[dv].push_back(u);
        }
      }
    }
  }
  // interleave the devices so that every PCIe link is busy at once
  std::vector<HalfUnit> order;
  for (size_t i = 0;; ++i) {
    bool any = false;
    for (auto& v : per_dev)
      if (i < v.size()) { order.push_back(v[i]); any = true; }
    if (!any) break;
  }
  return order;
}

bool host_has_avx2() {
  static const bool has = __builtin_cpu_supports("avx2");
  return has;
}

namespace {

constexpr int kSub = 128;   // sub-tile edge: a 128 x 128 partner copy (128 KB real, 256 KB complex) stays in L2

// ---- real --------------------------------------------------------------------------------------------------------------
// a: ni x nj, pitch lda; bt: nj x ni, pitch ldb (the partner as stored: rows j, columns i)
void real_scalar(const double* a, size_t lda, const double* bt, size_t ldb, int ni, int nj, double acc[3]) {
  double d[4] = {0, 0, 0, 0}, sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
  for (int i = 0; i < ni; ++i) {
    const double* ar = a + static_cast<size_t>(i) * lda;
    const double* bc = bt + i;
    for (int j = 0; j < nj; ++j) {
      const double x = ar[j], y = bc[static_cast<size_t>(j) * ldb];
      d[j & 3] += std::fabs(x - y);
      sa[j & 3] += std::fabs(x);
      sb[j & 3] += std::fabs(y);
    }
  }
  acc[0] += (d[0] + d[1]) + (d[2] + d[3]);
  acc[1] += (sa[0] + sa[1]) + (sa[2] + sa[3]);
  acc[2] += (sb[0] + sb[1]) + (sb[2] + sb[3]);
}

__attribute__((target("avx2"))) void real_avx2(const double* a, size_t lda, const double* bt, size_t ldb, int ni, int nj,
                                                double acc[3]) {
  const __m256d absmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7fffffffffffffffLL));
  __m256d d0 = _mm256_setzero_pd(), d1 = d0, sa0 = d0, sa1 = d0, sb0 = d0, sb1 = d0;
  const int ni4 = ni & ~3, nj4 = nj & ~3;
  for (int i0 = 0; i0 < ni4; i0 += 4) {
    const double* a0 = a + static_cast<size_t>(i0) * lda;
    for (int j0 = 0; j0 < nj4; j0 += 4) {
      const double* bb = bt + static_cast<size_t>(j0) * ldb + i0;
      const __m256d r0 = _mm256_loadu_pd(bb), r1 = _mm256_loadu_pd(bb + ldb), r2 = _mm256_loadu_pd(bb + 2 * ldb),
                    r3 = _mm256_loadu_pd(bb + 3 * ldb);
      const __m256d t0 = _mm256_unpacklo_pd(r0, r1), t1 = _mm256_unpackhi_pd(r0, r1), t2 = _mm256_unpacklo_pd(r2, r3),
                    t3 = _mm256_unpackhi_pd(r2, r3);
      const __m256d c0 = _mm256_permute2f128_pd(t0, t2, 0x20), c1 = _mm256_permute2f128_pd(t1, t3, 0x20);
      const __m256d c2 = _mm256_permute2f128_pd(t0, t2, 0x31), c3 = _mm256_permute2f128_pd(t1, t3, 0x31);
      const __m256d x0 = _mm256_loadu_pd(a0 + j0), x1 = _mm256_loadu_pd(a0 + lda + j0), x2 = _mm256_loadu_pd(a0 + 2 * lda + j0),
                    x3 = _mm256_loadu_pd(a0 + 3 * lda + j0);
      d0 = _mm256_add_pd(d0, _mm256_and_pd(absmask, _mm256_sub_pd(x0, c0)));
      d1 = _mm256_add_pd(d1, _mm256_and_pd(absmask, _mm256_sub_pd(x1, c1)));
      d0 = _mm256_add_pd(d0, _mm256_and_pd(absmask, _mm256_sub_pd(x2, c2)));
      d1 = _mm256_add_pd(d1, _mm256_and_pd(absmask, _mm256_sub_pd(x3, c3)));
      sa0 = _mm256_add_pd(sa0, _mm256_add_pd(_mm256_and_pd(absmask, x0), _mm256_and_pd(absmask, x2)));
      sa1 = _mm256_add_pd(sa1, _mm256_add_pd(_mm256_and_pd(absmask, x1), _mm256_and_pd(absmask, x3)));
      sb0 = _mm256_add_pd(sb0, _mm256_add_pd(_mm256_and_pd(absmask, c0), _mm256_and_pd(absmask, c2)));
      sb1 = _mm256_add_pd(sb1, _mm256_add_pd(_mm256_and_pd(absmask, c1), _mm256_and_pd(absmask, c3)));
    }
  }
  double t[4];
  _mm256_storeu_pd(t, _mm256_add_pd(d0, d1));
  acc[0] += (t[0] + t[1]) + (t[2] + t[3]);
  _mm256_storeu_pd(t, _mm256_add_pd(sa0, sa1));
  acc[1] += (t[0] + t[1]) + (t[2] + t[3]);
  _mm256_storeu_pd(t, _mm256_add_pd(sb0, sb1));
  acc[2] += (t[0] + t[1]) + (t[2] + t[3]);
  if (nj4 < nj) real_scalar(a + nj4, lda, bt + static_cast<size_t>(nj4) * ldb, ldb, ni4, nj - nj4, acc);
  if (ni4 < ni) real_scalar(a + static_cast<size_t>(ni4) * lda, lda, bt + ni4, ldb, ni - ni4, nj, acc);
}

// ---- complex (interleaved re, im; pitches in doubles) -------------------------------------------------------------------
void cplx_scalar(const double* a, size_t lda, const double* bt, size_t ldb, int ni, int nj, double acc[3]) {
  double d[2] = {0, 0}, sa[2] = {0, 0}, sb[2] = {0, 0};
  for (int i = 0; i < ni; ++i) {
    const double* ar = a + static_cast<size_t>(i) * lda;
    const double* bc = bt + 2 * i;
    for (int j = 0; j < nj; ++j) {
      const double xr = ar[2 * j], xi = ar[2 * j + 1];
      const double yr = bc[static_cast<size_t>(j) * ldb], yi = bc[static_cast<size_t>(j) * ldb + 1];
      const double dr = xr - yr, di = xi + yi;      // a - conj(b)
      d[j & 1] += std::sqrt(dr * dr + di * di);
      sa[j & 1] += std::sqrt(xr * xr + xi * xi);
      sb[j & 1] += std::sqrt(yr * yr + yi * yi);
    }
  }
  acc[0] += d[0] + d[1];
  acc[1] += sa[0] + sa[1];
  acc[2] += sb[0] + sb[1];
}

__attribute__((target("avx2"))) void cplx_avx2(const double* a, size_t lda, const double* bt, size_t ldb, int ni, int nj,
                                                double acc[3]) {
  const __m256d conj_sign = _mm256_castsi256_pd(_mm256_set_epi64x(static_cast<long long>(0x8000000000000000ULL), 0,
                                                                  static_cast<long long>(0x8000000000000000ULL), 0));
  __m256d d0 = _mm256_setzero_pd(), sa0 = d0, sb0 = d0;
  const int ni2 = ni & ~1, nj2 = nj & ~1;
  for (int i0 = 0; i0 < ni2; i0 += 2) {
    const double* a0 = a + static_cast<size_t>(i0) * lda;
    for (int j0 = 0; j0 < nj2; j0 += 2) {
      const double* bb = bt + static_cast<size_t>(j0) * ldb + 2 * i0;
      const __m256d r0 = _mm256_loadu_pd(bb), r1 = _mm256_loadu_pd(bb + ldb);       // (b[j][i], b[j][i+1]), (b[j+1][i], b[j+1][i+1])
      const __m256d c0 = _mm256_xor_pd(_mm256_permute2f128_pd(r0, r1, 0x20), conj_sign);   // conj(b[j][i]),   conj(b[j+1][i])
      const __m256d c1 = _mm256_xor_pd(_mm256_permute2f128_pd(r0, r1, 0x31), conj_sign);   // conj(b[j][i+1]), conj(b[j+1][i+1])
      const __m256d x0 = _mm256_loadu_pd(a0 + 2 * j0), x1 = _mm256_loadu_pd(a0 + lda + 2 * j0);
      const __m256d e0 = _mm256_sub_pd(x0, c0), e1 = _mm256_sub_pd(x1, c1);
      d0 = _mm256_add_pd(d0, _mm256_sqrt_pd(_mm256_hadd_pd(_mm256_mul_pd(e0, e0), _mm256_mul_pd(e1, e1))));
      sa0 = _mm256_add_pd(sa0, _mm256_sqrt_pd(_mm256_hadd_pd(_mm256_mul_pd(x0, x0), _mm256_mul_pd(x1, x1))));
      sb0 = _mm256_add_pd(sb0, _mm256_sqrt_pd(_mm256_hadd_pd(_mm256_mul_pd(c0, c0), _mm256_mul_pd(c1, c1))));
    }
  }
  double t[4];
  _mm256_storeu_pd(t, d0);
  acc[0] += (t[0] + t[1]) + (t[2] + t[3]);
  _mm256_storeu_pd(t, sa0);
  acc[1] += (t[0] + t[1]) + (t[2] + t[3]);
  _mm256_storeu_pd(t, sb0);
  acc[2] += (t[0] + t[1]) + (t[2] + t[3]);
  if (nj2 < nj) cplx_scalar(a + 2 * nj2, lda, bt + static_cast<size_t>(nj2) * ldb, ldb, ni2, nj - nj2, acc);
  if (ni2 < ni) cplx_scalar(a + static_cast<size_t>(ni2) * lda, lda, bt + 2 * ni2, ldb, ni - ni2, nj, acc);
}

// Streaming (non-temporal) copy of one row segment: the staging slot is written once and next read by the DMA engine, so the
// lines should neither be fetched for ownership nor evict the tiles being compared.  Falls back to memcpy when the segment
// is not 32-byte aligned at the destination.
__attribute__((target("avx2"))) void copy_segment_nt_avx2(double* dst, const double* src, size_t n) {
  size_t k = 0;
  for (; k + 16 <= n; k += 16) {
    const __m256d v0 = _mm256_loadu_pd(src + k), v1 = _mm256_loadu_pd(src + k + 4), v2 = _mm256_loadu_pd(src + k + 8),
                  v3 = _mm256_loadu_pd(src + k + 12);
    _mm256_stream_pd(dst + k, v0);
    _mm256_stream_pd(dst + k + 4, v1);
    _mm256_stream_pd(dst + k + 8, v2);
    _mm256_stream_pd(dst + k + 12, v3);
  }
  for (; k + 4 <= n; k += 4) _mm256_stream_pd(dst + k, _mm256_loadu_pd(src + k));
  for (; k < n; ++k) dst[k] = src[k];
}
inline void copy_segment(double* dst, const double* src, size_t n, bool avx2) {
  if (avx2 && (reinterpret_cast<uintptr_t>(dst) & 31u) == 0) copy_segment_nt_avx2(dst, src, n);
  else std::memcpy(dst, src, n * sizeof(double));
}

}  // namespace

void hermitian_unit_sums_scalar(const double* a, size_t lda, const double* b, size_t ldb, int64_t nr, int64_t nc, int e,
                                double acc[3]) {
  for (int64_t i0 = 0; i0 < nr; i0 += kSub)
    for (int64_t j0 = 0; j0 < nc; j0 += kSub) {
      const int ni = static_cast<int>(std::min<int64_t>(kSub, nr - i0)), nj = static_cast<int>(std::min<int64_t>(kSub, nc - j0));
      const double* as = a + static_cast<size_t>(i0) * lda + j0 * e;
      const double* bs = b + static_cast<size_t>(j0) * ldb + i0 * e;
      if (e == 2) cplx_scalar(as, lda, bs, ldb, ni, nj, acc);
      else real_scalar(as, lda, bs, ldb, ni, nj, acc);
    }
}

void hermitian_unit_sums(const double* a, size_t lda, const double* b, size_t ldb, int64_t nr, int64_t nc, int e,
                         double* scratch, double acc[3]) {
  half_unit_pack_and_sums(a, lda, b, ldb, nr, nc, e, nullptr, true, scratch, acc);
}

void half_unit_pack_and_sums(const double* a, size_t lda, const double* b, size_t ldb, int64_t nr, int64_t nc, int e,
                             double* stage, bool compare, double* scratch, double acc[3]) {
  const bool avx2 = host_has_avx2();
  const size_t lstage = static_cast<size_t>(nc) * e;
  if (stage && !compare) {
    // pack only (the compare is deferred to the background workers): whole row segments, which the hardware prefetchers follow
    // further than the 1 KB pieces of a sub-tile (10-20 % faster on the hosts measured)
    for (int64_t i = 0; i < nr; ++i) copy_segment(stage + static_cast<size_t>(i) * lstage, a + static_cast<size_t>(i) * lda, lstage, avx2);
    _mm_sfence();
    return;
  }
  for (int64_t i0 = 0; i0 < nr; i0 += kSub)
    for (int64_t j0 = 0; j0 < nc; j0 += kSub) {
      const int ni = static_cast<int>(std::min<int64_t>(kSub, nr - i0)), nj = static_cast<int>(std::min<int64_t>(kSub, nc - j0));
      if (stage) {
        // the unit's own sub-tile: DRAM -> cache (kept for the compare below) -> staging slot (streamed)
        for (int i = 0; i < ni; ++i)
          copy_segment(stage + static_cast<size_t>(i0 + i) * lstage + j0 * e, a + static_cast<size_t>(i0 + i) * lda + j0 * e,
                       static_cast<size_t>(nj) * e, avx2);
      }
      if (!compare) continue;
      // The partner sub-tile is nj short row segments a whole matrix row apart: copy it next to the core first (sequential
      // lines per segment, which the prefetchers follow), then walk the dense copy in transposed order.
      const size_t lds = static_cast<size_t>(ni) * e;
      for (int j = 0; j < nj; ++j)
        std::memcpy(scratch + static_cast<size_t>(j) * lds, b + static_cast<size_t>(j0 + j) * ldb + i0 * e, lds * sizeof(double));
      const double* as = a + static_cast<size_t>(i0) * lda + j0 * e;
      if (e == 2) {
        if (avx2) cplx_avx2(as, lda, scratch, lds, ni, nj, acc);
        else cplx_scalar(as, lda, scratch, lds, ni, nj, acc);
      } else {
        if (avx2) real_avx2(as, lda, scratch, lds, ni, nj, acc);
        else real_scalar(as, lda, scratch, lds, ni, nj, acc);
      }
    }
  if (stage) _mm_sfence();      // the streamed lines are globally visible before the caller hands the slot to the DMA engine
}

void emulate_half_upload(const double* S, int64_t m, int e, int world, int64_t block, size_t slot_bytes, int use_simd,
                         double* A_out, double* diff, double* sabs) {
  const int64_t per = (m + world - 1) / world;
  std::vector<std::pair<int64_t, int64_t>> ranges;
  for (int r = 0; r < world; ++r) ranges.emplace_back(std::min(m, per * r), std::min(m, per * (r + 1)));
  const std::vector<HalfUnit> units = plan_half_upload(m, e, ranges, block, slot_bytes);
  std::vector<double> stage, scratch(2 * kSub * kSub);
  double dsum = 0.0, ssum = 0.0;
  const size_t ld = static_cast<size_t>(m) * e;
  for (const HalfUnit& u : units) {
    const size_t w = static_cast<size_t>(u.nc) * e;
    stage.assign(static_cast<size_t>(u.nr) * w + 4, 0.0);
    double* st = stage.data();
    if (use_simd) while (reinterpret_cast<uintptr_t>(st) & 31u) ++st;     // exercise the streaming-store path
    double acc[3] = {0.0, 0.0, 0.0};
    const double* src = S + (u.r0) * ld + u.c0 * e;
    const double* partner = S + static_cast<size_t>(u.c0) * ld + u.r0 * e;
    if (use_simd == 2) {      // the one-shot call's flow: pack now, evaluate later against the caller's matrix
      half_unit_pack_and_sums(src, ld, partner, ld, u.nr, u.nc, e, st, false, scratch.data(), acc);
      half_unit_pack_and_sums(src, ld, partner, ld, u.nr, u.nc, e, nullptr, true, scratch.data(), acc);
    } else if (use_simd) {
      half_unit_pack_and_sums(src, ld, partner, ld, u.nr, u.nc, e, st, true, scratch.data(), acc);
    } else {
      for (int64_t r = 0; r < u.nr; ++r) std::memcpy(st + r * w, src + r * ld, w * sizeof(double));
      hermitian_unit_sums_scalar(st, w, partner, ld, u.nr, u.nc, e, acc);
    }
    for (int64_t r = 0; r < u.nr; ++r) std::memcpy(A_out + (u.r0 + r) * ld + u.c0 * e, st + r * w, w * sizeof(double));
    double d1, s1;
    half_unit_contribution(u, acc, &d1, &s1);
    dsum += d1;
    ssum += s1;
  }
  for (int64_t r = 0; r < m; ++r)
    for (int64_t c = 0; c < m; ++c)
      if (!half_tile_uploaded(r, c, block)) {
        A_out[r * ld + c * e] = A_out[c * ld + r * e];
        if (e == 2) A_out[r * ld + c * e + 1] = -A_out[c * ld + r * e + 1];
      }
  *diff = dsum;
  *sabs = ssum;
}

}  // namespace speigen
