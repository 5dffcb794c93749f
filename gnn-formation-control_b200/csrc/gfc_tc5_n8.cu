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


// contiguous graph range of a CTA, cut into tiles of up to GPC graphs (the last one may be partial: graphs are
// spread evenly over the CTAs instead of handing out whole tiles, so no SM waits for a neighbour's extra tile)
struct CtaRange {
  int begin, end, ntiles;
  __device__ __forceinline__ CtaRange(int B) {
    begin = (int)(((long long)blockIdx.x * B) / gridDim.x);
    end = (int)(((long long)(blockIdx.x + 1) * B) / gridDim.x);
    ntiles = (end - begin + GPC - 1) / GPC;
  }
};

// ======================================================================================================
// forward
// ======================================================================================================
struct FwdLayout {
  static constexpr int Z_BYTES = 3 * V_PLANE;                 // one operand buffer (72 KB); two of them
  static constexpr int OFF_Z = 0;
  static constexpr int OFF_H = 2 * Z_BYTES;
  static constexpr int OFF_S = OFF_H + HB_BYTES;
  static constexpr int OFF_BAR = OFF_S + S_BYTES;             // ops_ready[2][K], done[2], tmem_ready, tmem ptr
  static constexpr size_t BYTES = OFF_BAR + 96;
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
  uint64_t* ops_ready = reinterpret_cast<uint64_t*>(sm8 + L::OFF_BAR);   // [buffer][tap k]: state k of the tile stored
  uint64_t* done = ops_ready + 2 * K;
  uint64_t* tmem_ready = done + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm8 + L::OFF_BAR + 80);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = warp, c = lane;
  const bool producer = tid < kProducers;
  const CtaRange rg(a.p.B);
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);

  // ---- mbarriers (nothing here touches global memory: under PDL it overlaps the previous kernel's tail) -----
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2 * K; ++i) tc5::mbar_init(&ops_ready[i], 16);
    tc5::mbar_init(&done[0], 1);
    tc5::mbar_init(&done[1], 1);
    tc5::mbar_init(tmem_ready, 1);
    tc5::fence_mbar_init();
  }
  // L2 prefetch of this CTA's first two tiles (no data is consumed, so it may run ahead of the PDL dependency:
  // the DRAM / TLB latency of the first loads overlaps the previous kernel's tail)
  if (producer && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int g0 = rg.begin + t * GPC + j;
      if (g0 < rg.end) tc5::bulk_prefetch_l2(a.x + (size_t)g0 * G * N, G * N * 4);
    }
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (!producer) {
    // =========================== issuer warp: TMEM, then the MMA chains =====================================
    tc5::tmem_alloc(tmem_ptr, L::TMEM_COLS);
    tc5::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc5::mbar_arrive(tmem_ready);
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    if (tc5::elect_one()) {
      constexpr uint32_t kI96 = tc5::idesc_bf16(128, 96, 1, 0), kI64 = tc5::idesc_bf16(128, 64, 1, 0),
                         kI32 = tc5::idesc_bf16(128, 32, 1, 0);
      const uint32_t z_base = tc5::smem_u32(Zb), h_base = tc5::smem_u32(Hb);
      const uint64_t bd = tc5::make_desc(h_base, PB, 128);
      for (int itl = 0; itl < rg.ntiles; ++itl) {
        const int buf = itl & 1;
        const uint32_t zb = z_base + buf * L::Z_BYTES;
        uint64_t a0 = tc5::make_desc(zb, 128, PV), a1 = tc5::make_desc(zb + V_PLANE, 128, PV),
                 a2 = tc5::make_desc(zb + 2 * V_PLANE, 128, PV);
        uint64_t b = bd;
        const uint32_t acc = tmem + buf * 128;
#pragma unroll 1
        for (int s = 0; s < KG / 16; ++s) {
          if ((s & (G / 16 - 1)) == 0) {   // the MMAs of tap k start as soon as z_k is stored (the hops go on)
            tc5::mbar_wait_suspend(&ops_ready[buf * K + s / (G / 16)], (itl >> 1) & 1);   // (k = 0 also orders the
            tc5::fence_after_sync();                                                      //  taps of the first tile)
          }
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
    // running global pointers of this (graph slot, column) thread; they advance by one tile per request
    int gq = rg.begin + j;
    const float* px = a.x + ((size_t)gq * G + c) * N;
    const float2* ppos = reinterpret_cast<const float2*>(a.pos) + (size_t)gq * N + lane;
    const float* ps = a.S + (size_t)gq * a.s_bstride + lane;
    float xcol[N];
    float2 mypos = make_float2(0.f, 0.f);
    float s0 = 0.f, s1 = 0.f;
    bool okr = false;
    auto request = [&]() {
      okr = gq < rg.end;
      if (okr) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(px)), w = __ldg(reinterpret_cast<const float4*>(px) + 1);
        xcol[0] = u.x; xcol[1] = u.y; xcol[2] = u.z; xcol[3] = u.w;
        xcol[4] = w.x; xcol[5] = w.y; xcol[6] = w.z; xcol[7] = w.w;
      } else {
#pragma unroll
        for (int n = 0; n < N; ++n) xcol[n] = 0.f;
      }
      if (GSRC == GSRC_POS) {
        mypos = (okr && lane < N) ? __ldg(ppos) : make_float2(0.f, 0.f);
      } else {
        s0 = okr ? __ldg(ps) : 0.f;
        s1 = okr ? __ldg(ps + 32) : 0.f;
      }
      gq += GPC; px += (size_t)GPC * G * N; ppos += GPC * N; ps += (size_t)GPC * a.s_bstride;
      if (lane == 0 && gq < rg.end) tc5::bulk_prefetch_l2(px - c * N, G * N * 4);   // the tile after, into L2
    };
    auto produce = [&](int t, int buf) {
      const bool ok = okr;
      unsigned char* Zt = Zb + buf * L::Z_BYTES + j * PV + c * 16;
      warp_gso<GSRC, true>(Sw, a, ok, mypos, s0, s1, lane);     // slot = S^T: z_{k+1}[n] = sum_m S[m][n] z_k[m]
      float z[N];
#pragma unroll
      for (int n = 0; n < N; ++n) z[n] = xcol[n];
      if (t + 1 < rg.ntiles) request();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        store_chunk3(Zt + k * G * 16, V_PLANE, z);
        tc5::fence_proxy_async();
        if (k == 0) tc5::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc5::mbar_arrive(&ops_ready[buf * K + k]);
        if (k + 1 < K) hop8(Sw, z);
      }
    };

    // taps -> B[n = f][k = c] planes (c = k*G + g: the rows of h are already c-contiguous); warps 0..11 convert
    // one 8-c chunk each, visible to the issuer through the first ops_ready arrival
    float4 tu, tw;
    if (warp < KG / 8) {
      const float4* src = reinterpret_cast<const float4*>(a.h + (size_t)lane * KG + warp * 8);
      tu = __ldg(src);
      tw = __ldg(src + 1);
    }
    request();
    GFC_STAMP(a, 0);
    const int q = warp & 3, cg = warp >> 2;              // TMEM lane quadrant, 8-column group of this warp
    float bb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bb[i] = a.bias ? __ldg(a.bias + cg * 8 + i) : 0.f;
    if (warp < KG / 8) {
      const float v[8] = {tu.x, tu.y, tu.z, tu.w, tw.x, tw.y, tw.z, tw.w};
      store_tap_chunk(Hb, warp, lane, v);
    }
    produce(0, 0);
    GFC_STAMP(a, 1);
    uint32_t tmem = 0;
    for (int itl = 0; itl < rg.ntiles; ++itl) {
      const int buf = itl & 1;
      const int b0 = rg.begin + itl * GPC;
      const int rows_used = min(GPC, rg.end - b0) * N;
      if (itl + 1 < rg.ntiles) produce(itl + 1, buf ^ 1);
      if (itl == 0) {
        GFC_STAMP(a, 2);
        tc5::mbar_wait_suspend(tmem_ready, 0);
        tc5::fence_after_sync();
        tmem = *tmem_ptr;
      }
      tc5::mbar_wait_suspend(&done[buf], (itl >> 1) & 1);
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

// (Tried at the end of round 2: the activation mask as bits for this shape too — the forward epilogue storing one byte per
// row and column group, the backward reading 32 bytes of mask per graph instead of 1 KB of y, gfc_use_mask.  The
// producers are bound by instruction issue, not by bytes: cfg2 step 19.9 -> 20.2 us, 64 batches per launch 0.581 ->
// 0.599 ms.  Dropped; the wide kernels keep it, where it removes 4.2 GB per cfg3 step.)
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
  static constexpr int OFF_BAR = OFF_DB + F * 4;               // ops_ready[2][K], dx_done[2], dh_done[2], tmem_ready, ptr
  static constexpr size_t BYTES = OFF_BAR + 112;
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
  uint64_t* dx_done = ops_ready + 2 * K;   // ops_ready[buffer][k]: V_k (and, with k = 0, XT) of the tile stored
  uint64_t* dh_done = dx_done + 2;
  uint64_t* tmem_ready = dh_done + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm8 + L::OFF_BAR + 96);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = warp, c = lane;
  const bool producer = tid < kProducers;
  const bool want_dx = a.dX != nullptr, want_dh = a.dHp != nullptr, want_db = a.dbp != nullptr;
  const CtaRange rg(a.p.B);
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2 * K; ++i) tc5::mbar_init(&ops_ready[i], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      tc5::mbar_init(&dx_done[i], 1);
      tc5::mbar_init(&dh_done[i], 1);
    }
    tc5::mbar_init(tmem_ready, 1);
    tc5::fence_mbar_init();
  }
  if (tid < F) dbs[tid] = 0.f;
  // L2 prefetch of this CTA's first two tiles of dY and x (nothing is consumed: may run ahead of the PDL dependency)
  if (producer && lane == 0) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int g0 = rg.begin + t * GPC + j;
      if (g0 < rg.end) {
        tc5::bulk_prefetch_l2(a.dY + (size_t)g0 * N * F, N * F * 4);
        if (want_dh) tc5::bulk_prefetch_l2(a.x + (size_t)g0 * G * N, G * N * 4);
      }
    }
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (!producer) {
    // =========================== issuers: warp 16 -> dX chain, warp 17 -> dH chain ==========================
    const bool is_dx = warp == kProducers / 32;
    if (is_dx) {
      tc5::tmem_alloc(tmem_ptr, L::TMEM_COLS);
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(tmem_ready);
    } else {
      tc5::mbar_wait_suspend(tmem_ready, 0);
    }
    tc5::fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    if ((is_dx ? want_dx : want_dh) && tc5::elect_one()) {
      const uint32_t v_base = tc5::smem_u32(Vb), h_base = tc5::smem_u32(Hb);
      for (int itl = 0; itl < rg.ntiles; ++itl) {
        const int buf = itl & 1;
        const uint32_t vb = v_base + buf * L::VX_BYTES;
        if (!is_dx) {   // the dH chain contracts over rows: it needs every V_k of the tile
#pragma unroll
          for (int k = 0; k < K; ++k) tc5::mbar_wait_suspend(&ops_ready[buf * K + k], (itl >> 1) & 1);
          tc5::fence_after_sync();
        }
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
            if ((s & (F / 16 - 1)) == 0) {   // the MMAs of V_k start as soon as V_k is stored (the hops go on)
              tc5::mbar_wait_suspend(&ops_ready[buf * K + s / (F / 16)], (itl >> 1) & 1);
              tc5::fence_after_sync();
            }
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
    int gq = rg.begin + j;
    const float* px = a.x + ((size_t)gq * G + c) * N;
    const float* pdy = a.dY + (size_t)gq * N * F + c;
    const float* py = a.yout + (size_t)gq * N * F + c;
    const float2* ppos = reinterpret_cast<const float2*>(a.pos) + (size_t)gq * N + lane;
    const float* ps = a.S + (size_t)gq * a.s_bstride + lane;
    const bool has_act = a.act != GFC_ACT_NONE;
    // d(pre) = yo > 0 ? dy : dy * neg  (act_grad of gfc_common.cuh without the per-element branches; yo = 1 when
    // there is no activation)
    const float neg = a.act == GFC_ACT_LEAKY_RELU ? a.slope : (a.act == GFC_ACT_RELU ? 0.f : 1.f);
    float xcol[N], dyc[N], yoc[N];
    float2 mypos = make_float2(0.f, 0.f);
    float s0 = 0.f, s1 = 0.f;
    bool okr = false;
    auto request = [&]() {
      okr = gq < rg.end;
      if (GSRC == GSRC_POS) {
        mypos = (okr && lane < N) ? __ldg(ppos) : make_float2(0.f, 0.f);
      } else {
        s0 = okr ? __ldg(ps) : 0.f;
        s1 = okr ? __ldg(ps + 32) : 0.f;
      }
#pragma unroll
      for (int n = 0; n < N; ++n) {
        dyc[n] = okr ? __ldg(pdy + n * F) : 0.f;
        yoc[n] = (okr && has_act) ? __ldg(py + n * F) : 1.f;
      }
      if (want_dh) {
        if (okr) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(px)), w = __ldg(reinterpret_cast<const float4*>(px) + 1);
          xcol[0] = u.x; xcol[1] = u.y; xcol[2] = u.z; xcol[3] = u.w;
          xcol[4] = w.x; xcol[5] = w.y; xcol[6] = w.z; xcol[7] = w.w;
        } else {
#pragma unroll
          for (int n = 0; n < N; ++n) xcol[n] = 0.f;
        }
      }
      gq += GPC; px += (size_t)GPC * G * N; pdy += (size_t)GPC * N * F; py += (size_t)GPC * N * F;
      ppos += GPC * N; ps += (size_t)GPC * a.s_bstride;
      if (lane == 0 && gq < rg.end) {   // the tile after, into L2
        tc5::bulk_prefetch_l2(pdy - c, N * F * 4);
        if (has_act) tc5::bulk_prefetch_l2(py - c, N * F * 4);
        if (want_dh) tc5::bulk_prefetch_l2(px - c * N, G * N * 4);
      }
    };
    float dbreg = 0.f;
    auto produce = [&](int t, int buf) {
      const bool ok = okr;
      unsigned char* Vt = Vb + buf * L::VX_BYTES + j * PV + c * 16;
      unsigned char* Xt = Vb + buf * L::VX_BYTES + L::OFF_X + j * PB + c * 16;
      warp_gso<GSRC, false>(Sw, a, ok, mypos, s0, s1, lane);    // slot = S: V_{k+1}[n] = sum_m S[n][m] V_k[m]
      float v[N];
#pragma unroll
      for (int n = 0; n < N; ++n) v[n] = yoc[n] > 0.f ? dyc[n] : dyc[n] * neg;
      if (want_db) {
#pragma unroll
        for (int n = 0; n < N; ++n) dbreg += v[n];
      }
      if (want_dh) store_chunk3(Xt, 32 * 16, xcol);
      if (t + 1 < rg.ntiles) request();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        store_chunk3(Vt + k * F * 16, V_PLANE, v);
        tc5::fence_proxy_async();
        if (k == 0) tc5::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc5::mbar_arrive(&ops_ready[buf * K + k]);
        if (k + 1 < K) hop8(Sw, v);
      }
    };

    // taps -> B[n = g][k = c] planes, c = k*F + f:  B[g][c] = h[f][k*G + g]  (loads coalesced along g); warps
    // 0..11 convert one 8-c chunk each, visible to the issuers through the first ops_ready arrival
    float tv[8];
    if (want_dx && warp < KF / 8) {
      const int k = warp / (F / 8), f0 = (warp % (F / 8)) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) tv[i] = __ldg(a.h + (size_t)(f0 + i) * KG + k * G + lane);
    }
    request();
    GFC_STAMP(a, 0);
    if (want_dx && warp < KF / 8) store_tap_chunk(Hb, warp, lane, tv);
    produce(0, 0);
    GFC_STAMP(a, 1);
    const int q = warp & 3, cg = warp >> 2;
    float hacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hacc[i] = 0.f;
    uint32_t tmem = 0;
    for (int itl = 0; itl < rg.ntiles; ++itl) {
      const int buf = itl & 1;
      const int b0 = rg.begin + itl * GPC;
      const int rows_used = min(GPC, rg.end - b0) * N;
      if (itl + 1 < rg.ntiles) produce(itl + 1, buf ^ 1);
      if (itl == 0) {
        GFC_STAMP(a, 2);
        tc5::mbar_wait_suspend(tmem_ready, 0);
        tc5::fence_after_sync();
        tmem = *tmem_ptr;
      }
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + cg * 8);
      // ---- dX straight from TMEM: this thread's row (j', n'), 8 channels g ------------------------------
      if (want_dx) {
        tc5::mbar_wait_suspend(&dx_done[buf], (itl >> 1) & 1);
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
        tc5::mbar_wait_suspend(&dh_done[buf], (itl >> 1) & 1);
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
