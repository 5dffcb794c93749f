// gfc_tc5_n8.cu — tcgen05 / TMEM kernels of the fused filter for the 8-robot shape (cfg2: N=8, G=F=32, K=3):
// forward  y = act(sum_k z_k H_k + b),  z_k = z_{k-1} S        (BatchLSIGF, utils/graphUtils/graphML.py:2342-2366)
// backward V_0 = dY o act'(y), V_k = V_{k-1} S^T;  dX = sum_k V_k H_k^T,  dH_k = V_k^T X,  db = colsum(V_0)
//
// Work decomposition.  A tile is 16 graphs = 128 rows (graph, node).  Warp j of the 16 producer warps owns graph j
// of the tile, lane c owns feature column c: the thread keeps the 8 node values of its column in registers, so a
// diffusion hop is 64 FFMAs against the graph's 8x8 GSO (128-bit shared-memory broadcasts) and never leaves the
// register file.  Everything a warp needs is warp-local — its GSO is built from the 8 positions with shuffles
// (bit-exact radius rule, pair_adjacent) into a private 256-byte slot — so producer warps never meet at a CTA
// barrier: they hand finished operand buffers to the issuer warps through mbarriers and run ahead into the next
// tile while the tensor core contracts the previous one.
//
// Operands.  The thread's 8 values are exactly one 16-byte bf16 chunk of the TRANSPOSED state matrix
//     ZT[c][row]   byte(c, row) = (row / 8) * PV + c * 16 + (row % 8) * 2         (c = k*32 + column, PV = 96 * 16)
// the UMMA canonical no-swizzle layout, read MN-major (M = row) by the tap contraction and K-major (M = c) by the
// dH contraction.  fp32 parity comes from three bf16 planes (successive truncation, x = p0 + p1 + p2 exactly to
// 2^-24) and the six products p0q0 p0q1 p0q2 p1q0 p1q1 p2q0; the three planes of the B operand sit side by side
// along N, so ONE MMA of N = 96 / 64 / 32 columns per A plane produces them into three accumulator column blocks
// that the epilogue adds in fp32 (round to nearest) — A is fetched from shared memory 3 times instead of 6.
// (tf32 operands would need K-major storage on both sides: an MN-major tf32 MMA returns zeros on sm_100a.)
#include "gfc_tile_kernels.cuh"
#include "gfc_tc5.cuh"

namespace gfc {

int g_pdl = 1;   // programmatic dependent launch of the n8 kernels (gfc_set_option key 4)

namespace n8 {

constexpr int N = 8, G = 32, F = 32, K = 3, KF = K * F, KG = K * G;
constexpr int ROWS = 128, GPC = ROWS / N;
constexpr int PV = KF * 16;                    // bytes per 8-row group of ZT / VT
constexpr int V_PLANE = GPC * PV;              // 24 KB
constexpr int PB = 3 * 32 * 16;                // B operands: [8-k chunk][plane][n = 32] x 16 B
constexpr int XB_BYTES = GPC * PB;             // XT, three planes side by side along N: 24 KB
constexpr int HB_BYTES = (KF / 8) * PB;        // taps: 18 KB
constexpr int S_BYTES = GPC * N * N * 4;
constexpr int kProducers = 512;

// packed fp32 pairs (FFMA2 / FADD2 / FMUL2: two lanes per issue slot)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d;
}
__device__ __forceinline__ void upk2(uint64_t d, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(d));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// 8 consecutive values -> one 16-byte chunk in each of three planes `stride` bytes apart.  Planes by successive
// truncation (x = p0 + p1 + p2 to 2^-24 |x|, every subtraction exact), two values per instruction.
__device__ __forceinline__ void store_chunk3(unsigned char* plane0, int stride, const float (&v)[8]) {
  uint32_t a[4], b[4], c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x0 = __float_as_uint(v[2 * i]), x1 = __float_as_uint(v[2 * i + 1]);
    a[i] = __byte_perm(x0, x1, 0x7632);
    const uint64_t r = sub2(pk2(v[2 * i], v[2 * i + 1]),
                            pk2(__uint_as_float(x0 & 0xffff0000u), __uint_as_float(x1 & 0xffff0000u)));
    float r0, r1;
    upk2(r, r0, r1);
    const uint32_t y0 = __float_as_uint(r0), y1 = __float_as_uint(r1);
    b[i] = __byte_perm(y0, y1, 0x7632);
    const uint64_t t = sub2(r, pk2(__uint_as_float(y0 & 0xffff0000u), __uint_as_float(y1 & 0xffff0000u)));
    float t0, t1;
    upk2(t, t0, t1);
    c[i] = __byte_perm(__float_as_uint(t0), __float_as_uint(t1), 0x7632);
  }
  *reinterpret_cast<uint4*>(plane0) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(plane0 + stride) = make_uint4(b[0], b[1], b[2], b[3]);
  *reinterpret_cast<uint4*>(plane0 + 2 * stride) = make_uint4(c[0], c[1], c[2], c[3]);
}

// one diffusion hop of a register column:  v[n] <- sum_m T[n][m] v[m],  T row-major in the warp's slot
// (forward: T = S^T, backward: T = S).  Rows arrive as 128-bit broadcasts; pairs (m, m+1) are multiplied together.
__device__ __forceinline__ void hop8(const float* __restrict__ T, float (&v)[8]) {
  uint64_t vp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) vp[i] = pk2(v[2 * i], v[2 * i + 1]);
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const ulonglong2 ta = *reinterpret_cast<const ulonglong2*>(T + n * 8);
    const ulonglong2 tb = *reinterpret_cast<const ulonglong2*>(T + n * 8 + 4);
    uint64_t acc = mul2(ta.x, vp[0]);
    acc = fma2(ta.y, vp[1], acc);
    acc = fma2(tb.x, vp[2], acc);
    acc = fma2(tb.y, vp[3], acc);
    float lo, hi;
    upk2(acc, lo, hi);
    v[n] = lo + hi;
  }
}

// sqrt(1/deg) in fp64 for deg < 8 == __dsqrt_rn(__ddiv_rn(1, deg)) (both correctly rounded), inv_sqrt_deg of
// gfc_common.cuh without the software division / square root in the hot loop
__device__ __forceinline__ double inv_sqrt_deg8(int deg) {
  unsigned long long b = 0ull;
  switch (deg) {
    case 1: b = 0x3ff0000000000000ull; break;
    case 2: b = 0x3fe6a09e667f3bcdull; break;
    case 3: b = 0x3fe279a74590331cull; break;
    case 4: b = 0x3fe0000000000000ull; break;
    case 5: b = 0x3fdc9f25c5bfedd9ull; break;
    case 6: b = 0x3fda20bd700c2c3eull; break;
    case 7: b = 0x3fd83091e6a7f7e6ull; break;
    default: break;
  }
  return __longlong_as_double((long long)b);
}

// GSO of this warp's graph into its private slot (fp32, row-major; TRANSPOSED: slot[n][m] = S[m][n]).
// POS: from the 8 positions held by lanes 0..7 (the rule is symmetric).  DENSE: s0 / s1 = S[lane], S[lane + 32].
template <int GSRC, bool TRANSPOSED>
__device__ __forceinline__ void warp_gso(float* __restrict__ Sw, const TileArgs& a, bool ok, float2 mypos,
                                         float s0, float s1, int lane) {
  if (GSRC == GSRC_POS) {
    const int m = lane & 7, n0 = lane >> 3;          // entries o = lane (rows 0..3) and lane + 32 (rows 4..7)
    const float mx = __shfl_sync(0xffffffffu, mypos.x, m), my = __shfl_sync(0xffffffffu, mypos.y, m);
    const float ax = __shfl_sync(0xffffffffu, mypos.x, n0), ay = __shfl_sync(0xffffffffu, mypos.y, n0);
    const float bx = __shfl_sync(0xffffffffu, mypos.x, n0 + 4), by = __shfl_sync(0xffffffffu, mypos.y, n0 + 4);
    const bool e0 = ok && (n0 != m) && pair_adjacent(ax, ay, mx, my, a);
    const bool e1 = ok && (n0 + 4 != m) && pair_adjacent(bx, by, mx, my, a);
    float v0 = e0 ? 1.f : 0.f, v1 = e1 ? 1.f : 0.f;
    if (a.norm) {
      const unsigned lo = __ballot_sync(0xffffffffu, e0), hi = __ballot_sync(0xffffffffu, e1);
      const int dm = __popc(((m < 4 ? lo : hi) >> ((m & 3) * 8)) & 0xffu);      // degree of node m (symmetric rule)
      const int d0 = __popc((lo >> (n0 * 8)) & 0xffu), d1 = __popc((hi >> (n0 * 8)) & 0xffu);
      const double im = inv_sqrt_deg8(dm);
      if (e0) v0 = (float)__dmul_rn(inv_sqrt_deg8(d0), im);
      if (e1) v1 = (float)__dmul_rn(inv_sqrt_deg8(d1), im);
    }
    Sw[lane] = v0;
    Sw[lane + 32] = v1;
  } else if (TRANSPOSED) {
    const int m = lane & 7, n0 = lane >> 3;          // s0 = S[n0][m], s1 = S[n0 + 4][m]
    Sw[m * 8 + n0] = s0;
    Sw[m * 8 + n0 + 4] = s1;
  } else {
    Sw[lane] = s0;
    Sw[lane + 32] = s1;
  }
  __syncwarp();
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// taps -> B operand planes [8-k chunk][plane][n] : chunk `cchunk`, column n = lane
__device__ __forceinline__ void store_tap_chunk(unsigned char* Hb, int cchunk, int n, const float (&v)[8]) {
  store_chunk3(Hb + cchunk * PB + n * 16, 32 * 16, v);
}

// all threads of the CTA: producers arrive after their first tile, issuer warps after the one-time setup
template <int NT>
__device__ __forceinline__ void cta_bar1() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// ======================================================================================================
// forward
// ======================================================================================================
struct FwdLayout {
  static constexpr int Z_BYTES = 3 * V_PLANE;                 // one operand buffer (72 KB); two of them
  static constexpr int OFF_Z = 0;
  static constexpr int OFF_H = 2 * Z_BYTES;
  static constexpr int OFF_S = OFF_H + HB_BYTES;
  static constexpr int OFF_BIAS = OFF_S + S_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + F * 4;            // ops_ready[2], done[2], tmem ptr
  static constexpr size_t BYTES = OFF_BAR + 64;
  static constexpr int TMEM_COLS = 256;                       // two accumulators of 96 columns at 0 and 128
};
constexpr int kFwdThreads = kProducers + 32;

template <int GSRC>
__global__ void __launch_bounds__(kFwdThreads, 1)
tc5_n8_fwd_kernel(const TileArgs a) {
  using L = FwdLayout;
  extern __shared__ __align__(16) float smem[];
  unsigned char* sm8 = reinterpret_cast<unsigned char*>(smem);
  unsigned char* Zb = sm8 + L::OFF_Z;
  unsigned char* Hb = sm8 + L::OFF_H;
  float* Ss = reinterpret_cast<float*>(sm8 + L::OFF_S);
  float* bias_s = reinterpret_cast<float*>(sm8 + L::OFF_BIAS);
  uint64_t* ops_ready = reinterpret_cast<uint64_t*>(sm8 + L::OFF_BAR);
  uint64_t* done = ops_ready + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm8 + L::OFF_BAR + 48);
  const TilePlan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = warp, c = lane;
  const bool producer = tid < kProducers;
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);

  // ---- mbarriers (nothing here touches global memory: under PDL it overlaps the previous kernel's tail) -----
  if (tid == 0) {
    tc5::mbar_init(&ops_ready[0], 16);
    tc5::mbar_init(&ops_ready[1], 16);
    tc5::mbar_init(&done[0], 1);
    tc5::mbar_init(&done[1], 1);
    tc5::fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  float xcol[N];
  float2 mypos = make_float2(0.f, 0.f);
  float s0 = 0.f, s1 = 0.f;
  auto request = [&](int tile) {
    const int b0 = tile * GPC;
    const bool ok = j < min(GPC, p.B - b0);
    if (ok) load_x_column<TileCfg<8, 32, 32, 3, 512, 2>>(xcol, a.x, b0, j, c, 1);
    else {
#pragma unroll
      for (int n = 0; n < N; ++n) xcol[n] = 0.f;
    }
    if (GSRC == GSRC_POS) {
      mypos = (ok && lane < N) ? __ldg(reinterpret_cast<const float2*>(a.pos) + (size_t)(b0 + j) * N + lane)
                               : make_float2(0.f, 0.f);
    } else {
      const float* src = a.S + (size_t)(b0 + j) * N * N;
      s0 = ok ? __ldg(src + lane) : 0.f;
      s1 = ok ? __ldg(src + lane + 32) : 0.f;
    }
  };

  if (!producer) {
    // =========================== issuer warp: one-time setup, then the MMA chains ============================
    tc5::tmem_alloc(tmem_ptr, L::TMEM_COLS);
    {  // taps -> B[n = f][k = c] planes (c = k*G + g: the rows of h are already c-contiguous), bias
      const int f = lane;
      float4 u[KG / 8], w[KG / 8];
#pragma unroll
      for (int cc = 0; cc < KG / 8; ++cc) {
        const float4* src = reinterpret_cast<const float4*>(a.h + (size_t)f * KG + cc * 8);
        u[cc] = __ldg(src);
        w[cc] = __ldg(src + 1);
      }
      bias_s[lane] = a.bias ? __ldg(a.bias + lane) : 0.f;
#pragma unroll
      for (int cc = 0; cc < KG / 8; ++cc) {
        const float v[8] = {u[cc].x, u[cc].y, u[cc].z, u[cc].w, w[cc].x, w[cc].y, w[cc].z, w[cc].w};
        store_tap_chunk(Hb, cc, f, v);
      }
    }
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    cta_bar1<kFwdThreads>();
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    if (tc5::elect_one()) {
      constexpr uint32_t kI96 = tc5::idesc_bf16(128, 96, 1, 0), kI64 = tc5::idesc_bf16(128, 64, 1, 0),
                         kI32 = tc5::idesc_bf16(128, 32, 1, 0);
      const uint32_t z_base = tc5::smem_u32(Zb), h_base = tc5::smem_u32(Hb);
      const uint64_t bd = tc5::make_desc(h_base, PB, 128);
      uint32_t itl = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++itl) {
        const int buf = itl & 1;
        tc5::mbar_wait(&ops_ready[buf], (itl >> 1) & 1);
        tc5::fence_after_sync();
        const uint32_t zb = z_base + buf * L::Z_BYTES;
        uint64_t a0 = tc5::make_desc(zb, 128, PV), a1 = tc5::make_desc(zb + V_PLANE, 128, PV),
                 a2 = tc5::make_desc(zb + 2 * V_PLANE, 128, PV);
        uint64_t b = bd;
        const uint32_t acc = tmem + buf * 128;
#pragma unroll 1
        for (int s = 0; s < KG / 16; ++s) {
          tc5::mma_bf16_ss(acc, a0, b, kI96, s == 0 ? 0u : 1u);     // p0 . [q0 | q1 | q2]
          tc5::mma_bf16_ss(acc, a1, b, kI64, 1u);                   // p1 . [q0 | q1]
          tc5::mma_bf16_ss(acc, a2, b, kI32, 1u);                   // p2 .  q0
          a0 += 256 >> 4; a1 += 256 >> 4; a2 += 256 >> 4;
          b += (2 * PB) >> 4;
        }
        tc5::mma_commit(&done[buf]);
      }
    }
    __syncwarp();
  } else {
    // =========================== producers / epilogue ========================================================
    float* Sw = Ss + j * N * N;
    auto produce = [&](int tile, int buf) {
      const int b0 = tile * GPC;
      const bool ok = j < min(GPC, p.B - b0);
      unsigned char* Zt = Zb + buf * L::Z_BYTES + j * PV + c * 16;
      warp_gso<GSRC, true>(Sw, a, ok, mypos, s0, s1, lane);     // slot = S^T: z_{k+1}[n] = sum_m S[m][n] z_k[m]
      float z[N];
#pragma unroll
      for (int n = 0; n < N; ++n) z[n] = xcol[n];
      const int nxt = tile + gridDim.x;
      if (nxt < p.ntiles) request(nxt);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        store_chunk3(Zt + k * G * 16, V_PLANE, z);
        if (k + 1 < K) hop8(Sw, z);
      }
      tc5::fence_proxy_async();
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(&ops_ready[buf]);
    };

    request(blockIdx.x);
    GFC_STAMP(a, 0);
    produce(blockIdx.x, 0);
    GFC_STAMP(a, 1);
    cta_bar1<kFwdThreads>();
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    const int q = warp & 3, cg = warp >> 2;              // TMEM lane quadrant, 8-column group of this warp
    float bb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bb[i] = bias_s[cg * 8 + i];
    uint32_t itl = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++itl) {
      const int buf = itl & 1;
      const int b0 = tile * GPC;
      const int rows_used = min(GPC, p.B - b0) * N;
      {
        const int nxt = tile + gridDim.x;
        if (nxt < p.ntiles) produce(nxt, buf ^ 1);
      }
      if (itl == 0) GFC_STAMP(a, 2);
      tc5::mbar_wait(&done[buf], (itl >> 1) & 1);
      tc5::fence_after_sync();
      uint32_t r0[8], r1[8], r2[8];
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + cg * 8);
      tc5::tmem_ld8u(taddr, r0);
      tc5::tmem_ld8u(taddr + 32, r1);
      tc5::tmem_ld8u(taddr + 64, r2);
      tc5::tmem_ld_wait();
      const int r = q * 32 + lane;
      if (r < rows_used) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = apply_act((__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + __uint_as_float(r2[i]) + bb[i],
                           a.act, a.slope);
        float4* dst = reinterpret_cast<float4*>(a.y + ((size_t)b0 * N + r) * F + cg * 8);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      tc5::fence_before_sync();
      if (itl == 0) GFC_STAMP(a, 3);
    }
  }
  GFC_STAMP_NS(a, 9);
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == kProducers / 32) tc5::tmem_dealloc(*tmem_ptr, L::TMEM_COLS);
}

// ======================================================================================================
// backward
// ======================================================================================================
struct BwdLayout {
  static constexpr int OFF_V = 0;                              // 3 planes; the M = 128 over-read of the dH chain
  static constexpr int OFF_X = 3 * V_PLANE;                    //   (K*F = 96 lanes used) runs into XT
  static constexpr int VX_BYTES = 3 * V_PLANE + XB_BYTES;      // one operand buffer (96 KB); two of them
  static constexpr int OFF_H = 2 * VX_BYTES;
  static constexpr int OFF_S = OFF_H + HB_BYTES;
  static constexpr int OFF_DB = OFF_S + S_BYTES;
  static constexpr int OFF_BAR = OFF_DB + F * 4;               // ops_ready[2], dx_done[2], dh_done[2], tmem ptr
  static constexpr size_t BYTES = OFF_BAR + 64;
  static constexpr int TMEM_COLS = 512;                        // buffer b: dX blocks at b*256, dH blocks at b*256+128
};
constexpr int kBwdThreads = kProducers + 64;                   // + warp 16 (dX chain) + warp 17 (dH chain)

template <int GSRC>
__global__ void __launch_bounds__(kBwdThreads, 1)
tc5_n8_bwd_kernel(const TileArgs a) {
  using L = BwdLayout;
  extern __shared__ __align__(16) float smem[];
  unsigned char* sm8 = reinterpret_cast<unsigned char*>(smem);
  unsigned char* Vb = sm8 + L::OFF_V;
  unsigned char* Hb = sm8 + L::OFF_H;
  float* Ss = reinterpret_cast<float*>(sm8 + L::OFF_S);
  float* dbs = reinterpret_cast<float*>(sm8 + L::OFF_DB);
  uint64_t* ops_ready = reinterpret_cast<uint64_t*>(sm8 + L::OFF_BAR);
  uint64_t* dx_done = ops_ready + 2;
  uint64_t* dh_done = ops_ready + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm8 + L::OFF_BAR + 48);
  const TilePlan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = warp, c = lane;
  const bool producer = tid < kProducers;
  const bool want_dx = a.dX != nullptr, want_dh = a.dHp != nullptr, want_db = a.dbp != nullptr;
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      tc5::mbar_init(&ops_ready[i], 16);
      tc5::mbar_init(&dx_done[i], 1);
      tc5::mbar_init(&dh_done[i], 1);
    }
    tc5::fence_mbar_init();
  }
  if (tid < F) dbs[tid] = 0.f;
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  float xcol[N], dyc[N], yoc[N];
  float2 mypos = make_float2(0.f, 0.f);
  float s0 = 0.f, s1 = 0.f;
  auto request = [&](int tile) {
    const int b0 = tile * GPC;
    const bool ok = j < min(GPC, p.B - b0);
    if (GSRC == GSRC_POS) {
      mypos = (ok && lane < N) ? __ldg(reinterpret_cast<const float2*>(a.pos) + (size_t)(b0 + j) * N + lane)
                               : make_float2(0.f, 0.f);
    } else {
      const float* src = a.S + (size_t)(b0 + j) * N * N;
      s0 = ok ? __ldg(src + lane) : 0.f;
      s1 = ok ? __ldg(src + lane + 32) : 0.f;
    }
    const float* dsrc = a.dY + ((size_t)(b0 + j) * N) * F + c;
    const float* ysrc = (a.act != GFC_ACT_NONE) ? a.yout + ((size_t)(b0 + j) * N) * F + c : nullptr;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      dyc[n] = ok ? __ldg(dsrc + (size_t)n * F) : 0.f;
      yoc[n] = (ok && ysrc) ? __ldg(ysrc + (size_t)n * F) : 1.f;
    }
    if (want_dh) {
      if (ok) load_x_column<TileCfg<8, 32, 32, 3, 512, 2>>(xcol, a.x, b0, j, c, 1);
      else {
#pragma unroll
        for (int n = 0; n < N; ++n) xcol[n] = 0.f;
      }
    }
  };
  const int q = warp & 3, cg = warp >> 2;
  float hacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hacc[i] = 0.f;
  float dbreg = 0.f;

  if (!producer) {
    // =========================== issuers: warp 16 -> dX chain, warp 17 -> dH chain ==========================
    const bool is_dx = warp == kProducers / 32;
    if (is_dx) tc5::tmem_alloc(tmem_ptr, L::TMEM_COLS);
    if (want_dx) {
      // taps -> B[n = g][k = c] planes, c = k*F + f:  B[g][c] = h[f][k*G + g]  (loads coalesced along g);
      // each issuer warp converts half of the twelve 8-c chunks
      constexpr int HALF = KF / 16;
      const int g = lane, cc0 = is_dx ? 0 : HALF;
      float v[HALF][8];
#pragma unroll
      for (int cc = 0; cc < HALF; ++cc) {
        const int cchunk = cc0 + cc;
        const int k = cchunk / (F / 8), f0 = (cchunk % (F / 8)) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[cc][i] = __ldg(a.h + (size_t)(f0 + i) * KG + k * G + g);
      }
#pragma unroll
      for (int cc = 0; cc < HALF; ++cc) store_tap_chunk(Hb, cc0 + cc, g, v[cc]);
    }
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    cta_bar1<kBwdThreads>();
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    if ((is_dx ? want_dx : want_dh) && tc5::elect_one()) {
      const uint32_t v_base = tc5::smem_u32(Vb), h_base = tc5::smem_u32(Hb);
      uint32_t itl = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++itl) {
        const int buf = itl & 1;
        tc5::mbar_wait(&ops_ready[buf], (itl >> 1) & 1);
        tc5::fence_after_sync();
        const uint32_t vb = v_base + buf * L::VX_BYTES;
        if (is_dx) {
          // dX[row][g] = sum_c V[row][c] H[c][g]: A = VT read MN-major (M = row), 16 columns c per MMA
          constexpr uint32_t kI96 = tc5::idesc_bf16(128, 96, 1, 0), kI64 = tc5::idesc_bf16(128, 64, 1, 0),
                             kI32 = tc5::idesc_bf16(128, 32, 1, 0);
          uint64_t a0 = tc5::make_desc(vb, 128, PV), a1 = tc5::make_desc(vb + V_PLANE, 128, PV),
                   a2 = tc5::make_desc(vb + 2 * V_PLANE, 128, PV);
          uint64_t b = tc5::make_desc(h_base, PB, 128);
          const uint32_t acc = tmem + buf * 256;
#pragma unroll 1
          for (int s = 0; s < KF / 16; ++s) {
            tc5::mma_bf16_ss(acc, a0, b, kI96, s == 0 ? 0u : 1u);
            tc5::mma_bf16_ss(acc, a1, b, kI64, 1u);
            tc5::mma_bf16_ss(acc, a2, b, kI32, 1u);
            a0 += 256 >> 4; a1 += 256 >> 4; a2 += 256 >> 4;
            b += (2 * PB) >> 4;
          }
          tc5::mma_commit(&dx_done[buf]);
        } else {
          // dH[c][g] = sum_rows VT[c][row] XT[g][row]: A = VT K-major (M = c), 16 rows (two graphs) per MMA
          constexpr uint32_t kI96 = tc5::idesc_bf16(128, 96, 0, 0), kI64 = tc5::idesc_bf16(128, 64, 0, 0),
                             kI32 = tc5::idesc_bf16(128, 32, 0, 0);
          uint64_t a0 = tc5::make_desc(vb, PV, 128), a1 = tc5::make_desc(vb + V_PLANE, PV, 128),
                   a2 = tc5::make_desc(vb + 2 * V_PLANE, PV, 128);
          uint64_t b = tc5::make_desc(vb + L::OFF_X, PB, 128);
          const uint32_t acc = tmem + buf * 256 + 128;
#pragma unroll 1
          for (int s = 0; s < ROWS / 16; ++s) {
            tc5::mma_bf16_ss(acc, a0, b, kI96, s == 0 ? 0u : 1u);
            tc5::mma_bf16_ss(acc, a1, b, kI64, 1u);
            tc5::mma_bf16_ss(acc, a2, b, kI32, 1u);
            a0 += (2 * PV) >> 4; a1 += (2 * PV) >> 4; a2 += (2 * PV) >> 4;
            b += (2 * PB) >> 4;
          }
          tc5::mma_commit(&dh_done[buf]);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== producers / epilogue ========================================================
    float* Sw = Ss + j * N * N;
    auto produce = [&](int tile, int buf) {
      const int b0 = tile * GPC;
      const bool ok = j < min(GPC, p.B - b0);
      unsigned char* Vt = Vb + buf * L::VX_BYTES + j * PV + c * 16;
      unsigned char* Xt = Vb + buf * L::VX_BYTES + L::OFF_X + j * PB + c * 16;
      warp_gso<GSRC, false>(Sw, a, ok, mypos, s0, s1, lane);    // slot = S: V_{k+1}[n] = sum_m S[n][m] V_k[m]
      float v[N];
#pragma unroll
      for (int n = 0; n < N; ++n) v[n] = act_grad(dyc[n], yoc[n], a.act, a.slope);
      if (want_db) {
#pragma unroll
        for (int n = 0; n < N; ++n) dbreg += v[n];
      }
      if (want_dh) store_chunk3(Xt, 32 * 16, xcol);
      const int nxt = tile + gridDim.x;
      if (nxt < p.ntiles) request(nxt);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        store_chunk3(Vt + k * F * 16, V_PLANE, v);
        if (k + 1 < K) hop8(Sw, v);
      }
      tc5::fence_proxy_async();
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(&ops_ready[buf]);
    };

    request(blockIdx.x);
    GFC_STAMP(a, 0);
    produce(blockIdx.x, 0);
    GFC_STAMP(a, 1);
    cta_bar1<kBwdThreads>();
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    uint32_t itl = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++itl) {
      const int buf = itl & 1;
      const int b0 = tile * GPC;
      const int rows_used = min(GPC, p.B - b0) * N;
      {
        const int nxt = tile + gridDim.x;
        if (nxt < p.ntiles) produce(nxt, buf ^ 1);
      }
      if (itl == 0) GFC_STAMP(a, 2);
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cg * 8);
      // ---- dX straight from TMEM: this thread's row (j', n'), 8 channels g ------------------------------
      if (want_dx) {
        tc5::mbar_wait(&dx_done[buf], (itl >> 1) & 1);
        tc5::fence_after_sync();
        uint32_t r0[8], r1[8], r2[8];
        tc5::tmem_ld8u(taddr, r0);
        tc5::tmem_ld8u(taddr + 32, r1);
        tc5::tmem_ld8u(taddr + 64, r2);
        tc5::tmem_ld_wait();
        const int r = q * 32 + lane;
        if (r < rows_used) {
          const int jj = r / N, nn = r - jj * N;
          float* dst = a.dX + ((size_t)(b0 + jj) * G + cg * 8) * N + nn;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[(size_t)i * N] = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + __uint_as_float(r2[i]);
        }
      }
      if (itl == 0) GFC_STAMP(a, 3);
      // ---- dH tile -> running fp32 sums (lane = (k, f), 8 channels g) -----------------------------------
      if (want_dh) {
        tc5::mbar_wait(&dh_done[buf], (itl >> 1) & 1);
        tc5::fence_after_sync();
        if (q < 3) {
          uint32_t r0[8], r1[8], r2[8];
          tc5::tmem_ld8u(taddr + 128, r0);
          tc5::tmem_ld8u(taddr + 160, r1);
          tc5::tmem_ld8u(taddr + 192, r2);
          tc5::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            hacc[i] += (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + __uint_as_float(r2[i]);
        }
      }
      tc5::fence_before_sync();
      if (itl == 0) GFC_STAMP(a, 4);
    }
    // ---- per-CTA partials: dH[f][k*G + g] from lane (k, f) = q*32 + lane; db through shared memory ------
    if (want_dh && q < 3) {
      const int row = q * 32 + lane;
      const int k = row / F, f = row - k * F;
      float* dst = a.dHp + (size_t)blockIdx.x * F * KG + (size_t)f * KG + k * G + cg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(hacc[4], hacc[5], hacc[6], hacc[7]);
    }
    if (want_db) atomicAdd(dbs + c, dbreg);
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (want_db && tid < F) a.dbp[(size_t)blockIdx.x * F + tid] = dbs[tid];
  GFC_STAMP(a, 6);
  GFC_STAMP_NS(a, 9);
  if (warp == kProducers / 32) tc5::tmem_dealloc(*tmem_ptr, L::TMEM_COLS);
}

template <typename Kern>
static int launch_n8(Kern kern, const TileArgs& a, cudaStream_t st, const char* name) {
  GFC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.p.smem_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.p.grid);
  cfg.blockDim = dim3(a.p.threads);
  cfg.dynamicSmemBytes = a.p.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  GFC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
  GFC_LAUNCH_CHECK(name);
  return GFC_OK;
}

}  // namespace n8

static int n8_grid(int B) {
  DeviceInfo di;
  if (get_device_info(&di)) return 0;
  const int ntiles = ceil_div(B, n8::GPC);
  return ntiles < di.sm_count ? ntiles : di.sm_count;
}

// persistent grid of the backward kernel (one CTA per SM): the caller sizes the partial buffers with it
int tc5_bwd_grid_n8_32_32_3(int B) { return n8_grid(B); }

int tc5_fwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st) {
  TileArgs b = a;
  b.p.gpc = n8::GPC;
  b.p.ntiles = ceil_div(a.p.B, n8::GPC);
  b.p.grid = n8_grid(a.p.B);
  if (b.p.grid <= 0) return GFC_ERR_CUDA;
  b.p.threads = n8::kFwdThreads;
  b.p.smem_bytes = n8::FwdLayout::BYTES;
  if (gsrc == GSRC_POS) return n8::launch_n8(n8::tc5_n8_fwd_kernel<GSRC_POS>, b, st, "tc5_n8_fwd<pos>");
  return n8::launch_n8(n8::tc5_n8_fwd_kernel<GSRC_DENSE>, b, st, "tc5_n8_fwd<dense>");
}

int tc5_bwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st) {
  TileArgs b = a;
  b.p.gpc = n8::GPC;
  b.p.ntiles = ceil_div(a.p.B, n8::GPC);
  b.p.grid = n8_grid(a.p.B);
  if (b.p.grid <= 0) return GFC_ERR_CUDA;
  b.p.threads = n8::kBwdThreads;
  b.p.smem_bytes = n8::BwdLayout::BYTES;
  if (gsrc == GSRC_POS) return n8::launch_n8(n8::tc5_n8_bwd_kernel<GSRC_POS>, b, st, "tc5_n8_bwd<pos>");
  return n8::launch_n8(n8::tc5_n8_bwd_kernel<GSRC_DENSE>, b, st, "tc5_n8_bwd<dense>");
}

}  // namespace gfc
