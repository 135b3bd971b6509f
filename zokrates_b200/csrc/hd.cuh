// Host/device portability shims.
//
// The product is the sm_100a build (nvcc, __CUDA_ARCH__ defined inside kernels).  The same
// device functions also compile as plain C++ so that tests/host_emu can step every kernel body
// thread-by-thread on a CPU-only machine (this authoring container has no GPU).  The host
// emulation is a TEST harness: nothing in libzkb200.so ever takes the host path.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_HDN __host__ __device__
#define ZKB_NI __host__ __device__ __noinline__
#else
#define ZKB_HD inline
#define ZKB_HDN
#define ZKB_NI __attribute__((noinline))
#endif

namespace zkb {
namespace ptx {

#if defined(__CUDA_ARCH__)
// Carry-chain primitives.  `asm volatile` keeps the program order of the CC-flag users; ptxas
// turns mad.lo.cc/madc.hi.cc pairs on the same operands into IMAD.WIDE.U32(.X) with predicate
// carries (checked with cuobjdump -sass).
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// lo/hi halves of one 32x32 product issued from ONE asm statement so that both halves name the same
// virtual registers: ptxas then emits a single IMAD.WIDE.U32(.X) with predicate carry-in/out.
__device__ __forceinline__ void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c0), "r"(c1));
}
__device__ __forceinline__ void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c0), "r"(c1));
}
__device__ __forceinline__ void madc_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, %5;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(c0), "r"(c1));
}
#else
// Host emulation of the PTX carry flag (single-threaded test harness only).
static thread_local uint32_t g_cc = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cc; }
// PTX: sub.cc writes the borrow-out to CC.CF; subc computes a - (b + CC.CF).
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cc; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return (uint32_t)(((uint64_t)a * b) >> 32) + c + g_cc; }
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) { uint32_t l = mad_lo_cc(a, b, c0); uint32_t h = madc_hi_cc(a, b, c1); lo = l; hi = h; }
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) { uint32_t l = madc_lo_cc(a, b, c0); uint32_t h = madc_hi_cc(a, b, c1); lo = l; hi = h; }
inline void madc_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t c0, uint32_t c1) { uint32_t l = madc_lo_cc(a, b, c0); uint32_t h = madc_hi(a, b, c1); lo = l; hi = h; }
#endif

}  // namespace ptx
}  // namespace zkb
