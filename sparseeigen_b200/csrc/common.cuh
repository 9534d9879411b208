// Shared device/host helpers for the speigen_b200 CUDA library (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>

namespace speigen {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void note_launch(int n);   // counts kernel launches (bench.py's gpu_launches)
#define SPE_CUDA(x)                                                                        \
  do {                                                                                     \
    cudaError_t e__ = (x);                                                                 \
    if (e__ != cudaSuccess) {                                                              \
      ::speigen::set_error("CUDA error %s (%s) at %s:%d", cudaGetErrorName(e__),           \
                           cudaGetErrorString(e__), __FILE__, __LINE__);                   \
      return SPEIGEN_ERR_CUDA;                                                             \
    }                                                                                      \
  } while (0)

// ---- panel layout ---------------------------------------------------------------------------
// Every m x q "panel" (V, Y, B, R, U ...) lives row-major in HBM with `ncr` real columns
// (q for real, 2q = interleaved {re,im} for complex) and a row pitch of panel_pitch(ncr)
// doubles.  pitch == 4 (mod 8) makes the DMMA B-fragment shared-memory reads of a staged
// panel tile bank-conflict free (see contract.cu) and keeps rows 32-byte aligned.
__host__ __device__ inline int panel_pitch(int ncr) {
  int p = ncr;
  while ((p & 7) != 4) ++p;
  return p;
}
constexpr int kMaxCols = 64;   // max real columns handled by one contraction launch (q<=64 real, q<=32 complex)
constexpr int kRowPad = 256;   // panel row count is padded to a multiple of this (zero rows)

// ---- PTX wrappers ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// L2 eviction policies for the streamed matrix (read once) and the re-read panel.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// TMA: 2-D tiled tensor load global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// TMA: 1-D bulk copy global -> shared.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// FP64 tensor-core MMA (SASS: DMMA.8x8x4): C(8x8) += A(8x4,row) * B(4x8,col).
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

// ---- reweighting arithmetic (shared by the K1 epilogue and the standalone kernel) ----------
// Follows R/spEigen.R:108-112 (stage constants) and :170-176 (weights, H).
struct StageConsts {
  double p, eps, c1, c2, w0, half_over_c1;
};
#ifdef __CUDACC__
__host__ __device__ inline double weight_of_abs(double a, const StageConsts& sc) {
#ifdef __CUDA_ARCH__
  // rounding sequence pinned (no compiler-chosen FMA contraction): every kernel derives the same weight from the same |V|
  return (a > sc.eps) ? __ddiv_rn(sc.half_over_c1, __fma_rn(a, a, __dmul_rn(sc.p, a))) : sc.w0;   // R/spEigen.R:170-173
#else
  return (a > sc.eps) ? sc.half_over_c1 / (a * a + sc.p * a) : sc.w0;
#endif
}
// One component of B = d*Y - (w(|V|) - wmax) * V * rho (R/spEigen.R:176-177, G - H) with the rounding sequence pinned, so
// that the contraction's fused epilogue and the tail's producer give the same bits whatever the compiler contracts.
__device__ __forceinline__ double reweight_elem(double y, double v, double w_minus_wmax, double d, double rho) {
  return __fma_rn(d, y, -__dmul_rn(__dmul_rn(w_minus_wmax, v), rho));
}
#endif

}  // namespace speigen
