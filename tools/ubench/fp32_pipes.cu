// Micro-benchmark: issue cost (cycles per warp-instruction per SM sub-partition) of the fp32 instructions the
// cfg2 producer warps are made of, at the occupancy those kernels run at (16 warps per SM = 4 per scheduler).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float s0, float s1) {
  float a[8], b = s0 + threadIdx.x * 1e-9f, c = s1;
  uint64_t pa[8], pb, pc;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pa[i]) : "f"(a[i]), "f"(a[i] + 1.f));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b), "f"(b));
  asm volatile("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
  unsigned mask = (unsigned)(s0 * 0.f);   // 0: predicates off
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pb), "l"(pc));
      if (MODE == 2) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb));
      if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(b));          // 2 distinct regs
      if (MODE == 6) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; @p add.rn.f32 %0, %0, %1;}" : "+f"(a[i]) : "f"(b), "r"(mask));
      if (MODE == 7) asm volatile("lop3.b32 %0, %0, %1, 0x12345, 0x96;" : "+r"(*(unsigned*)&a[i]) : "r"(__float_as_uint(b)));
      if (MODE == 8) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(a[i]) : "f"(c));   // immediate multiplier
      if (MODE == 9) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(*(unsigned*)&a[i]) : "r"(__float_as_uint(b)));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(pa[i]));
    s += a[i] + lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  k<MODE><<<148, 512>>>(out, cyc, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  k<MODE><<<148, 512>>>(out, cyc, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < 148; ++i) m += h[i];
  m /= 148;
  // 16 warps / 4 schedulers = 4 warps per scheduler, ITERS*8 instructions each
  printf("%-34s %.2f cycles per warp-instruction per scheduler (%s)\n", name, m / (4.0 * ITERS * 8), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("FFMA  (3 distinct regs)", out, cyc);
  run<5>("FFMA  (2 distinct regs)", out, cyc);
  run<8>("FFMA  (immediate multiplier)", out, cyc);
  run<1>("FFMA2 (packed pair)", out, cyc);
  run<2>("FADD", out, cyc);
  run<3>("FADD2 (packed pair)", out, cyc);
  run<4>("FMUL", out, cyc);
  run<6>("FADD predicated off (+SETP)", out, cyc);
  run<7>("LOP3", out, cyc);
  run<9>("PRMT", out, cyc);
  return 0;
}
