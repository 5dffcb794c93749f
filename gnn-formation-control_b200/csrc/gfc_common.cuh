// gfc_common.cuh — shared helpers for libgfc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "gfc.h"

namespace gfc {

// ---- host-side error plumbing (never throws across the C ABI) --------------
void set_error(const char* fmt, ...);
int& launch_counter();

#define GFC_CUDA_TRY(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::gfc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                       __FILE__, __LINE__);                                       \
      return GFC_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define GFC_LAUNCH_CHECK(name)                                                    \
  do {                                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) {                                                      \
      ::gfc::set_error("launch of %s failed: %s (%s:%d)", name,                   \
                       cudaGetErrorString(_e), __FILE__, __LINE__);               \
      return GFC_ERR_CUDA;                                                        \
    }                                                                             \
    ++::gfc::launch_counter();                                                    \
  } while (0)

#define GFC_REQUIRE(cond, code, ...)                                              \
  do {                                                                            \
    if (!(cond)) {                                                                \
      ::gfc::set_error(__VA_ARGS__);                                              \
      return code;                                                                \
    }                                                                             \
  } while (0)

struct DeviceInfo {
  int sm_count;
  int cc_major, cc_minor;
  int smem_optin;  // max dynamic shared memory per block (opt-in)
};
int get_device_info(DeviceInfo* out);

// Largest double s such that sqrt_rn(s) <= R (inclusive) / sqrt_rn(s) < R
// (strict).  Comparing the fp64 squared distance against it reproduces the
// reference's `sqrt(d2) <= R` (scene.py:147-149) / `pdist < R`
// (multirobotsim_dcenlocal.py:306-307) bit for bit without a device sqrt.
double squared_threshold(double radius, bool inclusive);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- device helpers ----------------------------------------------------------
#ifdef __CUDACC__

// fp32 -> tf32 by "round half ulp, truncate": add half a tf32 ulp to the bit pattern and clear
// the 13 low mantissa bits.  Two integer-pipe instructions; cvt.rna.tf32.f32 runs on the
// quarter-rate conversion pipe and was the top stall of every MMA phase (profiles/r1).
__device__ __forceinline__ uint32_t to_tf32(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}

// x = hi + lo with hi the tf32 rounding of x; lo = x - hi is exact in fp32 and carries <= 13
// significant bits, of which the tensor core keeps the top 11 (it ignores the low 13 mantissa
// bits of a tf32 operand).  hi*hi + hi*lo + lo*hi then carries ~2^-21 relative error per term
// (the "3xTF32" scheme).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// D(16x8) += A(16x8, row) * B(8x8, col); fragment layout per PTX ISA:
//   g = lane>>2, t = lane&3
//   a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4);  b0:(k=t,n=g) b1:(k=t+4,n=g)
//   c0:(g,2t) c1:(g,2t+1) c2:(g+8,2t) c3:(g+8,2t+1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4],
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 "
      "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == GFC_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == GFC_ACT_LEAKY_RELU) return v > 0.f ? v : v * slope;
  return v;
}
// d(pre) given the forward OUTPUT yo (sign(yo) == sign(pre) for slope > 0; for
// ReLU yo > 0 <=> pre > 0) — matches torch's leaky_relu/relu backward (x > 0).
__device__ __forceinline__ float act_grad(float dy, float yo, int act, float slope) {
  if (act == GFC_ACT_RELU) return yo > 0.f ? dy : 0.f;
  if (act == GFC_ACT_LEAKY_RELU) return yo > 0.f ? dy : dy * slope;
  return dy;
}

// fp64 squared distance with explicitly rounded ops (no FMA contraction), the
// arithmetic of `(xi-xj)**2 + (yi-yj)**2` on python floats (scene.py:147-148).
__device__ __forceinline__ double sqdist64(float xi, float yi, float xj, float yj) {
  double dx = __dsub_rn((double)xi, (double)xj);
  double dy = __dsub_rn((double)yi, (double)yj);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// 1/sqrt(deg) in fp64 as the reference: np.sqrt(1. / deg), 0 for deg == 0
// (multirobotsim_dcenlocal.py:309-313).
__device__ __forceinline__ double inv_sqrt_deg(int deg) {
  if (deg <= 0) return 0.0;
  return __dsqrt_rn(__ddiv_rn(1.0, (double)deg));
}

#endif  // __CUDACC__

}  // namespace gfc
