// gfc_tc5_wide.cu — warp-specialised tcgen05 / TMEM kernels of the fused graph filter for
// wide feature counts (64..128 channels), GSO rebuilt on chip from positions.
//
// One persistent CTA per SM works on tiles of 128 packed rows r = (graph j, node n).  ALL the
// arithmetic of BatchLSIGF (utils/graphUtils/graphML.py:2342-2366) runs on the 5th-generation
// tensor cores with fp32 accumulators in tensor memory:
//   * the diffusion state W_k [128 rows x CIN channels] lives in shared memory as THREE bf16
//     planes (successive truncation, W = p0 + p1 + p2 to 2^-24) in the UMMA canonical no-swizzle
//     layout, split into two channel slabs that are processed as two independent chains;
//   * the hop  W_{k+1} = P W_k  (graphML.py:2349-2352) is an MMA with the block-diagonal 0/1
//     matrix P of the tile's graphs (exact in bf16) as A operand and the three planes of W_k
//     as MN-major B operands: the result in TMEM is the exact fp32 hop;
//   * the tap contraction  OUT += W_k H_k  (graphML.py:2361-2362) is the 6-term product of the
//     bf16x3 planes (error ~2^-23, fp32-equivalent); the taps stream through a ring of
//     shared-memory stages filled by the TMA engine (cp.async.bulk) from a pre-packed copy;
//   * worker warps read the hop result back (tcgen05.ld), split it into planes for the next
//     tap, and run the epilogue (bias + activation, or the dX transpose) of tile t while the
//     issuing thread already feeds the tensor core with tile t+1.
// MODE 0: forward,  IN = x [B,G,N],  OUT = y [B,N,F]   (CIN = G, COUT = F)
// MODE 2: forward with the input given node-major, IN = x [B,N,G] (a previous layer's output), OUT = y [B,N,F]
// MODE 1: backward dX: V_0 = dY o act'(y), V_k = P V_{k-1}, dX = sum_k V_k H_k^T-contraction
//         (IN = dY [B,N,F], OUT = dX [B,G,N]; CIN = F, COUT = G) — the closed form of the autograd
//         graph of graphML.py:2342-2366 for a symmetric 0/1 GSO.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <type_traits>
#include "gfc_common.cuh"
#include "gfc_tc5.cuh"
#include "gfc_tc5_wide.cuh"
#include <string.h>

namespace gfc {
using tc5::make_desc;
using tc5::store_chunk3;

template <int CIN, int COUT>
struct WideLayout {
  static constexpr int ROWS = 128;
  static constexpr int CS = CIN / 2;                 // channels per slab
  static constexpr int CPT = CS / 2;                 // state columns per worker thread
  static constexpr int NCH = CS / 8;                 // 16-byte chunks (8 bf16) per row and slab
  static constexpr int PW = ROWS * 16;               // bytes between chunks (one chunk column of all rows)
  static constexpr int PLANE = NCH * PW;             // one bf16 plane of a slab
  static constexpr int SLAB = 3 * PLANE;
  static constexpr int STAGE = COUT * 96;            // taps of 16 channels: 3 planes x 2 chunks x COUT x 16 B
  static constexpr int NSTAGE = 6;
  static constexpr int KSTEPS = CS / 16;             // tap MMA k-steps (= ring stages) per phase
  static constexpr int OFF_W = 0;
  static constexpr int OFF_RING = OFF_W + 2 * SLAB;
  static constexpr int OFF_STAGE = (OFF_RING + NSTAGE * STAGE + 1023) / 1024 * 1024;   // 2 TMA store staging buffers [128 rows x 32 cols] fp32, 128B swizzle
  static constexpr int OFF_BIAS = OFF_STAGE + 2 * ROWS * 128;
  static constexpr int OFF_SP = OFF_BIAS + COUT * 4;         // float2 positions of the tile rows
  static constexpr int OFF_BAR = OFF_SP + 2 * ROWS * 8;     // (two position buffers, alternating per tile)
  static constexpr int NBAR = 2 * NSTAGE + 2 + 2 + 1 + 2 + 2;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr int TM_OUT = 0;                   // two output accumulators [128 x COUT]
  static constexpr int TM_HOP = 2 * COUT;            // two hop accumulators   [128 x CS]
  static constexpr int TM_P = 2 * COUT + 2 * CS;     // two buffers of the block-diagonal 0/1 hop matrix P, bf16 [128 lanes x 128 k] = 64 columns each
  static constexpr int TM_USED = 2 * COUT + 2 * CS + 128;
  static constexpr int TM_COLS = TM_USED <= 32 ? 32 : TM_USED <= 64 ? 64 : TM_USED <= 128 ? 128 : TM_USED <= 256 ? 256 : 512;
  static_assert(CIN % 32 == 0 && CIN >= 32 && CIN <= 128, "CIN in {32,64,96,128}");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 128, "COUT multiple of 16, <= 128");
  static_assert(CPT % 8 == 0, "whole chunks per worker thread");
  static_assert(BYTES <= 227 * 1024, "shared memory");
};

constexpr int kWideThreads = 320;   // warp 0: MMA issuer, warp 1: TMA producer + TMEM owner, warps 2..9: workers
constexpr int kWorkerWarps = 8;
int g_wide_flush_every = 2;   // gfc_set_option(GFC_OPT_WIDE_FLUSH_EVERY)
int g_wide_no_prefetch = 0;    // experiment switch

// exact fp64 rule, kept out of line so the (rare) rounding-band case is a real branch and the fp64 /
// conversion instructions are not if-converted into every pair test
static __device__ __noinline__ bool wide_adjacent_exact(float ax, float ay, float bx, float by, double thr) {
  return sqdist64(ax, ay, bx, by) <= thr;
}
__device__ __forceinline__ bool wide_adjacent(float2 a, float2 b, double thr, float thr_lo, float thr_hi) {
  const float dx = a.x - b.x, dy = a.y - b.y;
  const float s = fmaf(dx, dx, dy * dy);
  if (s < thr_lo) return true;
  if (s > thr_hi) return false;
  return wide_adjacent_exact(a.x, a.y, b.x, b.y, thr);
}

// adjacency of tile row `pr` (position `me`) with the 8 tile rows c0..c0+7: fp32 bit patterns 1.0f / 0.
// All 8 squared distances are formed first (8 independent shared-memory loads); the exact fp64 rule only
// runs for the rare pairs inside the fp32 rounding band.
__device__ __forceinline__ void adjacency8(uint32_t (&e)[8], const float2* __restrict__ sp, float2 me, int pr,
                                           int c0, int c_lo, int c_hi, int rows_used, double thr, float thr_lo,
                                           float thr_hi) {
  float sv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 o = sp[c0 + i];
    const float dx = me.x - o.x, dy = me.y - o.y;
    sv[i] = fmaf(dx, dx, dy * dy);
  }
  uint32_t band = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    const bool ok = (c >= c_lo) && (c < c_hi) && (c != pr) && (pr < rows_used) && (c < rows_used);
    e[i] = (ok && sv[i] < thr_lo) ? 0x3f800000u : 0u;
    if (ok && sv[i] >= thr_lo && sv[i] <= thr_hi) band |= 1u << i;
  }
  if (band) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (band & (1u << i)) {
        const float2 o = sp[c0 + i];
        e[i] = wide_adjacent_exact(me.x, me.y, o.x, o.y, thr) ? 0x3f800000u : 0u;
      }
    }
  }
}

// Predicated read-only loads as volatile asm: issued exactly where written (never sunk to the first use)
// and with the predicate inside the statement, so no select / move depends on the value right after it.
__device__ __forceinline__ float ldg_f32(const float* p, bool pred) {
  float v;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\t"
      "@q ld.global.nc.f32 %0, [%1];\n\t}"
      : "=f"(v)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ float4 ldg_f32x4(const float* p, bool pred) {
  float4 v;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
      "@q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "r"((int)pred));
  return v;
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float ldg_cg(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(kWideThreads, 1)
tc5_wide_kernel(const __grid_constant__ WideArgs w, const __grid_constant__ CUtensorMap tmap_out) {
  using L = WideLayout<CIN, COUT>;
  extern __shared__ __align__(128) unsigned char wsmem[];
  unsigned char* Wb = wsmem + L::OFF_W;
  unsigned char* Rb = wsmem + L::OFF_RING;
  float2* sp_all = reinterpret_cast<float2*>(wsmem + L::OFF_SP);
  unsigned char* stage_out = wsmem + L::OFF_STAGE;
  float* sbias = reinterpret_cast<float*>(wsmem + L::OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsmem + L::OFF_BAR);
  uint64_t* h_full = bars;                      // [NSTAGE]
  uint64_t* h_empty = bars + L::NSTAGE;         // [NSTAGE]
  uint64_t* w_ready = bars + 2 * L::NSTAGE;     // [2]
  uint64_t* mma_done = w_ready + 2;             // [2]
  uint64_t* p_ready = mma_done + 2;             // [1]
  uint64_t* out_full = p_ready + 1;             // [2]
  uint64_t* out_free = out_full + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(out_free + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = w.N, K = w.K;
  // a CTA owns a CONTIGUOUS range of tiles: its streams stay inside a few 2 MB pages for many tiles (strided
  // assignment made every CTA touch new pages of every array at every tile: TLB-miss storms at tile boundaries)
  const int tiles_per_cta = (w.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t_begin = min(w.ntiles, (int)blockIdx.x * tiles_per_cta);
  const int t_end = min(w.ntiles, t_begin + tiles_per_cta);

  // ---- one-time setup -------------------------------------------------------------------------
  for (int i = tid; i < COUT; i += kWideThreads) sbias[i] = (MODE != 1 && w.bias) ? __ldg(w.bias + i) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < L::NSTAGE; ++i) { tc5::mbar_init(&h_full[i], 1); tc5::mbar_init(&h_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc5::mbar_init(&w_ready[i], kWorkerWarps);
      tc5::mbar_init(&mma_done[i], 1);
      tc5::mbar_init(&out_full[i], 1);
      tc5::mbar_init(&out_free[i], kWorkerWarps);
    }
    tc5::mbar_init(p_ready, kWorkerWarps);
    tc5::fence_mbar_init();
  }
  if (warp == 1) tc5::tmem_alloc(tmem_ptr, L::TM_COLS);
  tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  // NOTE on code size: every role's per-tile code is executed once per ~15 us while the other roles run
  // their own loops, so the instruction caches only hold it if it is small.  Loops are deliberately kept
  // rolled (one copy of the plane-split / pair-test / MMA-issue code each) unless a register array forces
  // unrolling; barrier parities are bit masks so that they can be indexed at run time.
  if (warp == 0) {
    // =========================== MMA issuer (one elected thread) ===============================
    if (tc5::elect_one()) {
      constexpr uint32_t kIdescTap = tc5::idesc_bf16(128, COUT, 0, 0);
      constexpr uint32_t kIdescHop = tc5::idesc_bf16(128, L::CS, 0, 1);
      const uint32_t w_addr = tc5::smem_u32(Wb), r_addr = tc5::smem_u32(Rb);
      uint32_t par_wr = 0, par_of = 0, par_pr = 0, par_hf = 0;
      int st = 0;
      int nstamp = 0;
      const bool dbg = w.dbg != nullptr && blockIdx.x == 0;
#define GFC_WSTAMP(tag) do { if (dbg && nstamp < 1000) { w.dbg[2 * nstamp] = clock64(); w.dbg[2 * nstamp + 1] = (tag); ++nstamp; } } while (0)
      const int hop_ksteps = (w.gpc * N + 15) >> 4;
      int it = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        const int ob = it & 1;
        if (it >= 2) { tc5::mbar_wait(&out_free[ob], (par_of >> ob) & 1); par_of ^= 1u << ob; }
        tc5::mbar_wait(p_ready, par_pr); par_pr ^= 1;
        const uint32_t d_out = tmem + L::TM_OUT + ob * COUT;
#pragma unroll 1
        for (int ph = 0; ph < 2 * K; ++ph) {
          const int k = ph >> 1, s = ph & 1;
          GFC_WSTAMP(100 + s);
          tc5::mbar_wait(&w_ready[s], (par_wr >> s) & 1); par_wr ^= 1u << s;
          tc5::fence_after_sync();
          GFC_WSTAMP(110 + s);
          const uint32_t ws = w_addr + s * L::SLAB;
          if (k + 1 < K) {
            // hop: D_hop[s] = P * W_k[slab s]   (A = P from tensor memory, B = state planes MN-major)
            const uint32_t d_hop = tmem + L::TM_HOP + s * L::CS;
            uint32_t acc = 0;
#pragma unroll 1
            for (int j = 0; j < hop_ksteps; ++j) {
              const uint32_t pa = tmem + L::TM_P + ob * 64 + j * 8;   // 16 source rows = 8 columns of bf16 pairs
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                tc5::mma_bf16_ts(d_hop, pa, make_desc(ws + pl * L::PLANE + j * 256, 128, L::PW), kIdescHop, acc);
                acc = 1;
              }
            }
          }
          // taps: D_out += W_k[slab s] * H_k[slab s]   (6-term bf16x3 product)
#pragma unroll 1
          for (int i = 0; i < L::KSTEPS; ++i) {
            if (i == 0) GFC_WSTAMP(120 + s);
            tc5::mbar_wait(&h_full[st], par_hf);
            tc5::fence_after_sync();
            GFC_WSTAMP(130 + i);
            const uint32_t hs = r_addr + st * L::STAGE;
            const uint32_t wa = ws + i * 2 * L::PW;
            const uint64_t a0 = make_desc(wa, L::PW, 128);
            const uint64_t a1 = make_desc(wa + L::PLANE, L::PW, 128);
            const uint64_t a2 = make_desc(wa + 2 * L::PLANE, L::PW, 128);
            const uint64_t b0 = make_desc(hs, COUT * 16, 128);
            const uint64_t b1 = make_desc(hs + COUT * 32, COUT * 16, 128);
            const uint64_t b2 = make_desc(hs + COUT * 64, COUT * 16, 128);
            const uint32_t first = (ph == 0 && i == 0) ? 0u : 1u;
            tc5::mma_bf16_ss(d_out, a0, b0, kIdescTap, first);
            tc5::mma_bf16_ss(d_out, a0, b1, kIdescTap, 1u);
            tc5::mma_bf16_ss(d_out, a1, b0, kIdescTap, 1u);
            tc5::mma_bf16_ss(d_out, a1, b1, kIdescTap, 1u);
            tc5::mma_bf16_ss(d_out, a0, b2, kIdescTap, 1u);
            tc5::mma_bf16_ss(d_out, a2, b0, kIdescTap, 1u);
            tc5::mma_commit(&h_empty[st]);
            if (++st == L::NSTAGE) { st = 0; par_hf ^= 1; }
          }
          tc5::mma_commit(&mma_done[s]);
          GFC_WSTAMP(140 + s);
        }
        tc5::mma_commit(&out_full[ob]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== tap producer (TMA bulk copies) ===============================
    if (tc5::elect_one()) {
      int st = 0;
      uint32_t par_he = 0;
      bool primed = false;   // the first NSTAGE fills need no wait
      int filled = 0;
      const int stages_per_tile = K * (CIN / 16);
      for (int tile = t_begin; tile < t_end; ++tile) {
#pragma unroll 1
        for (int u = 0; u < stages_per_tile; ++u) {
          if (primed) { tc5::mbar_wait(&h_empty[st], par_he); }
          tc5::mbar_arrive_expect_tx(&h_full[st], L::STAGE);
          tc5::bulk_g2s(Rb + st * L::STAGE, reinterpret_cast<const unsigned char*>(w.hpack) + (size_t)u * L::STAGE,
                        L::STAGE, &h_full[st]);
          if (++st == L::NSTAGE) { st = 0; if (primed) par_he ^= 1; }
          if (!primed && ++filled == L::NSTAGE) primed = true;
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== workers ======================================================
    const int wt = tid - 64;                 // 0..255
    const int q = warp & 3;                  // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;        // which half of a slab's columns
    const int r = q * 32 + lane;             // tile row owned by this thread (= its TMEM lane)
    const int jr = r / N, nr = r - jr * N;   // (graph, node) of the row
    const uint32_t tm_lane = tmem + ((uint32_t)(q * 32) << 16);
    float xin[L::CPT];
    float2 mypos = make_float2(0.f, 0.f);
    uint32_t par_md = 0, par_ofl = 0;
    int nstamp = 0;
    const bool dbg = w.dbg != nullptr && blockIdx.x == 0 && tid == 64;
#define GFC_KSTAMP(tag) do { if (dbg && nstamp < 1000) { w.dbg[2048 + 2 * nstamp] = clock64(); w.dbg[2048 + 2 * nstamp + 1] = (tag); ++nstamp; } } while (0)

    // Operands of the next tile.  Early in the current tile one thread asks the TMA engine to pull the
    // (contiguous) input tile into L2; the register loads then happen slab by slab right before the slab is
    // written (short live ranges, L2-hit latency hidden behind the P build / the wait for the slab).
    auto prefetch_tile = [&](int tile) {
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      const uint32_t bytes = (uint32_t)gcount * (uint32_t)N * CIN * 4u;   // multiple of 16 (CIN % 32 == 0)
      const size_t off = (size_t)b0 * N * CIN;
      tc5::bulk_prefetch_l2(w.in + off, bytes);
      if (MODE == 1 && w.act != GFC_ACT_NONE) tc5::bulk_prefetch_l2(w.yout + off, bytes);
    };
    auto load_pos = [&](int tile) {
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      if (wt < 128)   // rows 0..127 in thread order wt: position of row wt
        mypos = (wt < gcount * N) ? __ldg(reinterpret_cast<const float2*>(w.pos) + (size_t)b0 * N + wt)
                                  : make_float2(0.f, 0.f);
    };
    auto load_slab = [&](int tile, int s) {
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      const bool valid = r < gcount * N;
      const int c0 = s * L::CS + half * L::CPT;
      if (MODE == 0) {
        const float* src = w.in + ((size_t)(b0 + jr) * CIN + c0) * N + nr;
#pragma unroll
        for (int i = 0; i < L::CPT; ++i) xin[i] = ldg_f32(src + (size_t)i * N, valid);
      } else {
        // dY / y are row-major [rows x CIN]: a warp instruction reads whole 16-byte pieces of RPI consecutive rows
        // (4 full lines) instead of 32 scattered ones; the halves of a bf16 chunk meet by a lane-pair shuffle
        // at store time.  Warp ww owns rows 16 ww .. 16 ww + 15, lane = (row offset, piece).
        constexpr int PPR = L::CS / 4, RPI = 32 / PPR;
        const int ww = warp - 2;
        const int rows_used = gcount * N;
#pragma unroll
        for (int i = 0; i < L::CPT / 4; ++i) {
          const int row = 16 * ww + RPI * i + lane / PPR;
          const bool ok = row < rows_used;
          const size_t off = ((size_t)b0 * N + row) * CIN + s * L::CS + 4 * (lane % PPR);
          float4 v = ldg_f32x4(w.in + off, ok);
          if (MODE == 1 && w.act != GFC_ACT_NONE) {
            const float4 yo = ldg_f32x4(w.yout + off, ok);
            v.x = act_grad(v.x, yo.x, w.act, w.slope);
            v.y = act_grad(v.y, yo.y, w.act, w.slope);
            v.z = act_grad(v.z, yo.z, w.act, w.slope);
            v.w = act_grad(v.w, yo.w, w.act, w.slope);
            // hand dY o act'(y) to the dH kernel: it then streams ONE tensor and never waits on the mask
            if (w.d_out && ok) *reinterpret_cast<float4*>(w.d_out + off) = v;
          }
          xin[4 * i] = v.x; xin[4 * i + 1] = v.y; xin[4 * i + 2] = v.z; xin[4 * i + 3] = v.w;
        }
      }
    };
    // registers of load_slab -> the three bf16 planes of W_0[slab s]
    auto store_slab = [&](int s) {
      if (MODE == 0) {
        unsigned char* base = Wb + s * L::SLAB + (half * (L::CPT / 8)) * L::PW + r * 16;
#pragma unroll
        for (int c = 0; c < L::CPT / 8; ++c) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = xin[c * 8 + i];
          store_chunk3(base + c * L::PW, L::PLANE, v);
        }
      } else {
        constexpr int PPR = L::CS / 4, RPI = 32 / PPR;
        const int ww = warp - 2;
        const int pi = lane % PPR;
        const bool odd = pi & 1;
#pragma unroll
        for (int i = 0; i < L::CPT / 4; i += 2) {
          // even lane keeps its piece of row(i) and receives the partner's; odd lane does the same for row(i+1)
          float snd[4], rcv[4], own[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            snd[e] = odd ? xin[4 * i + e] : xin[4 * (i + 1) + e];
            own[e] = odd ? xin[4 * (i + 1) + e] : xin[4 * i + e];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) rcv[e] = __shfl_xor_sync(0xffffffffu, snd[e], 1);
          float v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) { v[e] = odd ? rcv[e] : own[e]; v[4 + e] = odd ? own[e] : rcv[e]; }
          const int row = 16 * ww + RPI * (i + (odd ? 1 : 0)) + lane / PPR;
          store_chunk3(Wb + s * L::SLAB + (pi >> 1) * L::PW + row * 16, L::PLANE, v);
        }
      }
    };
    auto publish = [&](uint64_t* bar) {   // this warp's shared-memory writes -> tensor core, then arrive
      tc5::fence_proxy_async();
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(bar);
    };

    // P[r][c] = 1 iff rows r and c belong to the same graph and are adjacent (symmetric rule).  This thread owns
    // TMEM lane r; the two warps of a quadrant interleave 4-chunk groups of source rows.  Chunks t0..t1-1 of 8.
    auto build_p_part = [&](int tile, int pbuf, int t0, int t1) {
      const float2* sp = sp_all + pbuf * L::ROWS;
      const int gcount = min(w.gpc, w.B - tile * w.gpc);
      const int rows_used = gcount * N;
      const int c_lo = jr * N, c_hi = c_lo + N;
      const float2 me = sp[r];
      const bool row_ok = r < w.gpc * N;
#pragma unroll 1
      for (int t = t0; t < t1; ++t) {
        const int qc = (t >> 2) * 8 + half * 4 + (t & 3);   // chunk of 8 source rows = 4 TMEM columns
        uint32_t e[8];
        if (row_ok && qc * 8 < c_hi && qc * 8 + 8 > c_lo) {
          adjacency8(e, sp, me, r, qc * 8, c_lo, c_hi, rows_used, w.thr, w.thr_lo, w.thr_hi);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = 0u;
        }
        tc5::tmem_st4(tm_lane + L::TM_P + pbuf * 64 + qc * 4, tc5::pack_bf16_hi(e[0], e[1]), tc5::pack_bf16_hi(e[2], e[3]),
                      tc5::pack_bf16_hi(e[4], e[5]), tc5::pack_bf16_hi(e[6], e[7]));
      }
    };
    auto publish_p = [&]() {
      tc5::tmem_st_wait();
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(p_ready);
    };

    // The loop runs once per tile plus one leading iteration (it == -1) that only prepares the first tile's
    // operands, so that the per-tile code exists exactly once.  P of the next tile is double-buffered in TMEM and
    // built in two halves right after the first two write-backs of the live tile (where the workers have slack).
    int tile = t_begin;               // tile whose MMAs run during this iteration (none when it == -1)
    int it = -1;
    int next = t_begin;
    if (next < t_end) load_pos(next);
    while (true) {
      const bool live = it >= 0;
      const bool has_next = next < t_end;
      if (!live && !has_next) break;
      const int pbuf = (it + 1) & 1;
      if (has_next) {
        if (wt < 128) sp_all[pbuf * L::ROWS + wt] = mypos;   // positions of `next` (loaded one tile ahead)
        worker_bar();
        if (next + 1 < t_end) load_pos(next + 1);
        if (wt == 0 && !w.no_prefetch) {   // L2 prefetch runs two tiles ahead
          if (!live) prefetch_tile(next);
          if (next + 1 < t_end) prefetch_tile(next + 1);
        }
        if (!live || K == 1) {
          if (K > 1) build_p_part(next, pbuf, 0, 8);
          publish_p();
        }
      }
      // ---- write-backs of the K-1 hops ------------------------------------------------------------------
      const int nwb = live ? 2 * (K - 1) : 0;
      bool slab0_loaded = false;
#pragma unroll 1
      for (int ph = 0; ph < nwb; ++ph) {
        // next tile's slab 0: requested two write-backs ahead of its use (memory latency behind the last taps)
        if (has_next && !slab0_loaded && ph + 2 >= nwb) { load_slab(next, 0); slab0_loaded = true; }
        {
          const int s = ph & 1;
          GFC_KSTAMP(200 + s);
          tc5::mbar_wait(&mma_done[s], (par_md >> s) & 1); par_md ^= 1u << s;
          tc5::fence_after_sync();
          GFC_KSTAMP(210 + s);
          // hop result (exact fp32) -> three bf16 planes of W_{k+1}[slab s], 8 columns at a time
          const uint32_t taddr = tm_lane + L::TM_HOP + s * L::CS + half * L::CPT;
          unsigned char* base = Wb + s * L::SLAB + (half * (L::CPT / 8)) * L::PW + r * 16;
#pragma unroll 1
          for (int c = 0; c < L::CPT / 8; ++c) {
            uint32_t v[8];
            tc5::tmem_ld8u(taddr + c * 8, v);
            tc5::tmem_ld_wait();
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]);
            store_chunk3(base + c * L::PW, L::PLANE, f);
          }
          GFC_KSTAMP(220 + s);
          publish(&w_ready[s]);
          GFC_KSTAMP(230 + s);
        }
        if (has_next && ph < 2) {   // K > 1 here
          build_p_part(next, pbuf, 4 * ph, 4 * ph + 4);
          if (ph == 1) publish_p();
        }
      }
      // ---- the live tile's slabs become free one by one: next tile's state W_0 ---------------------------
      GFC_KSTAMP(240);
      if (has_next && !slab0_loaded) load_slab(next, 0);
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        if (has_next && s == 1) load_slab(next, 1);   // in flight while the last tap of slab 1 finishes
        if (live) { tc5::mbar_wait(&mma_done[s], (par_md >> s) & 1); par_md ^= 1u << s; }
        GFC_KSTAMP(250 + s);
        if (has_next) {
          store_slab(s);
          publish(&w_ready[s]);
        }
        GFC_KSTAMP(260 + s);
      }
      // ---- epilogue of the live tile (the issuer is already working on the next one) -------------------
      if (live) {
        const int b0 = tile * w.gpc;
        const int rows_used = min(w.gpc, w.B - b0) * N;
        const int ob = it & 1;
        tc5::mbar_wait(&out_full[ob], (par_ofl >> ob) & 1); par_ofl ^= 1u << ob;
        tc5::fence_after_sync();
        GFC_KSTAMP(271);
        if constexpr (MODE != 1) {
          // y tile through swizzled staging buffers and the TMA store engine: full 128-byte lines
#pragma unroll 1
          for (int pc = 0; pc < COUT / 32; ++pc) {
            const int col = pc * 32 + half * 16;
            uint32_t v[16];
            tc5::tmem_ld16(tm_lane + L::TM_OUT + ob * COUT + col, v);
            tc5::tmem_ld_wait();
            unsigned char* sbuf = stage_out + (pc & 1) * (L::ROWS * 128);
            unsigned char* srow = sbuf + r * 128;
            // the TMA store that last used this buffer (two pieces ago) has finished reading it
            if (wt == 0) tc5::tma_store_wait_read1();
            worker_bar();
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bb = *reinterpret_cast<const float4*>(sbias + col + i4 * 4);
              float4 o;
              o.x = apply_act(__uint_as_float(v[i4 * 4 + 0]) + bb.x, w.act, w.slope);
              o.y = apply_act(__uint_as_float(v[i4 * 4 + 1]) + bb.y, w.act, w.slope);
              o.z = apply_act(__uint_as_float(v[i4 * 4 + 2]) + bb.z, w.act, w.slope);
              o.w = apply_act(__uint_as_float(v[i4 * 4 + 3]) + bb.w, w.act, w.slope);
              const int cc = half * 4 + i4;
              *reinterpret_cast<float4*>(srow + ((cc ^ (r & 7)) << 4)) = o;
            }
            tc5::fence_proxy_async();
            worker_bar();
            if (wt == 0) {
              tc5::tma_store_2d(&tmap_out, sbuf, pc * 32, b0 * N);
              tc5::tma_store_commit();
            }
          }
        } else {
          // dX[(b0 + j), g, n]: lanes = consecutive nodes n, one coalesced store per channel
          const bool valid = r < rows_used;
#pragma unroll 1
          for (int cb = 0; cb < COUT / 2; cb += 16) {
            const int col = half * (COUT / 2) + cb;
            uint32_t v[16];
            tc5::tmem_ld16(tm_lane + L::TM_OUT + ob * COUT + col, v);
            tc5::tmem_ld_wait();
            if (valid) {
              float* dst = w.out + ((size_t)(b0 + jr) * COUT + col) * N + nr;
#pragma unroll
              for (int i = 0; i < 16; ++i) dst[(size_t)i * N] = __uint_as_float(v[i]);
            }
          }
        }
        tc5::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc5::mbar_arrive(&out_free[ob]);
        GFC_KSTAMP(270);
      }
      if (!has_next) break;
      tile = next;
      next += 1;
      ++it;
    }
  }
  if (tid == 64) tc5::tma_store_wait_all();
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc5::tmem_dealloc(tmem, L::TM_COLS);
}

// =====================================================================================================
// backward dH / db:  dH[f][k*G+g] = sum_rows V_k[r][f] X[r][g],  V_0 = dY o act'(y),  V_k = P V_{k-1}
// (the same gradient as sum_rows D[r][f] Z_k[r][g] with Z_k = P^k X, because P is symmetric; it needs no
// recomputation of the diffusion states).  A CTA owns the feature slice fh (FH columns of f) for the
// whole kernel and keeps its K x [G x FH] accumulators in tensor memory across its tiles:
//   * X^T (lanes = g, columns = tile rows, bf16x3 planes) is the A operand, stored into TMEM by the workers
//     straight from x's native [B,G,N] layout (tcgen05.st);
//   * V_k lives in shared memory as bf16x3 planes, in a ring of three buffers (tap k of a tile uses buffer
//     (base + k) % 3, so the next tile's V_0 can be written while the last taps still run); it is the MN-major
//     B operand of both the dH product and the hop;
//   * P (block-diagonal 0/1, exact in bf16) is double-buffered in shared memory (K-major A operand of the hop).
// One chain per tile: hop(k) -> [write-back(k) by the workers || dH(k) on the tensor core] -> hop(k+1) ...
// =====================================================================================================
template <int G, int F, int FH>
struct DhLayout {
  static constexpr int ROWS = 128;
  static constexpr int NFH = F / FH;
  static constexpr int CPT = FH / 2;                 // V columns per worker thread in a write-back
  static constexpr int NCH = FH / 8;
  static constexpr int PW = ROWS * 16;
  static constexpr int PLANE = NCH * PW;
  static constexpr int VBUF = 3 * PLANE;
  static constexpr int P_BYTES = (ROWS / 8) * PW;
  static constexpr int OFF_V = 0;
  static constexpr int OFF_P = OFF_V + 3 * VBUF;
  static constexpr int OFF_SP = OFF_P + 2 * P_BYTES;
  static constexpr int OFF_DB = OFF_SP + 2 * ROWS * 8;
  static constexpr int OFF_BAR = OFF_DB + FH * 4;
  static constexpr int NBAR = 6;
  static constexpr int BYTES_MIN = OFF_BAR + NBAR * 8 + 16;
  static constexpr int BYTES = BYTES_MIN < 120 * 1024 ? 120 * 1024 : BYTES_MIN;   // one CTA per SM (TMEM is taken whole)
  static constexpr int TM_X = 0;                     // 3 planes x 64 columns (128 rows, two per column)
  static constexpr int TM_ACC = 192;                 // acc(k) at TM_ACC + k*FH, hop result at TM_ACC + K*FH
  static_assert(FH % 32 == 0 && F % FH == 0 && CPT % 8 == 0, "feature slice");
  static_assert(G == 128 || G == 64 || G == 32, "G");
  static_assert(BYTES <= 227 * 1024, "shared memory");
};

template <int G, int F, int FH>
__global__ void __launch_bounds__(kWideThreads, 1)
tc5_wide_dh_kernel(const __grid_constant__ WideDhArgs w) {
  using L = DhLayout<G, F, FH>;
  extern __shared__ __align__(128) unsigned char wsmem[];
  unsigned char* Vb = wsmem + L::OFF_V;
  unsigned char* Pb = wsmem + L::OFF_P;
  float2* sp_all = reinterpret_cast<float2*>(wsmem + L::OFF_SP);
  float* dbs = reinterpret_cast<float*>(wsmem + L::OFF_DB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsmem + L::OFF_BAR);
  uint64_t* v_ready = bars;          // workers -> issuer: the next V buffer is written
  uint64_t* hop_done = bars + 1;     // issuer (commit) -> workers: hop result in TMEM, older MMAs complete
  uint64_t* item_done = bars + 2;    // issuer (commit) -> workers: every MMA of the tile complete
  uint64_t* p_ready = bars + 3;
  uint64_t* x_ready = bars + 4;
  // V_0 of the next tile has its own barrier: a fast warp may publish it right after its last write-back, and
  // on a shared barrier that second arrival could complete the write-back phase before a slow warp arrived
  uint64_t* v0_ready = bars + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = w.N, K = w.K;
  const int fh = blockIdx.x % L::NFH, part = blockIdx.x / L::NFH, nparts = gridDim.x / L::NFH;
  const int TM_HOP = L::TM_ACC + K * FH;
  // contiguous tile range per CTA group (see tc5_wide_kernel)
  const int tiles_per_part = (w.ntiles + nparts - 1) / nparts;
  const int t_begin = min(w.ntiles, part * tiles_per_part);
  const int t_end = min(w.ntiles, t_begin + tiles_per_part);

  for (int i = tid; i < 2 * L::P_BYTES / 16; i += kWideThreads) reinterpret_cast<uint4*>(Pb)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < FH; i += kWideThreads) dbs[i] = 0.f;
  if (tid == 0) {
    tc5::mbar_init(v_ready, kWorkerWarps);
    tc5::mbar_init(hop_done, 1);
    tc5::mbar_init(item_done, 1);
    tc5::mbar_init(p_ready, kWorkerWarps);
    tc5::mbar_init(x_ready, kWorkerWarps);
    tc5::mbar_init(v0_ready, kWorkerWarps);
    tc5::fence_mbar_init();
  }
  if (warp == 1) tc5::tmem_alloc(tmem_ptr, 512);
  tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  if (warp == 0) {
    // =========================== MMA issuer (one elected thread) ===============================
    if (tc5::elect_one()) {
      constexpr uint32_t kIdesc = tc5::idesc_bf16(128, FH, 0, 1);   // B = V planes, MN-major
      const uint32_t v_addr = tc5::smem_u32(Vb), p_addr = tc5::smem_u32(Pb);
      uint32_t par_vr = 0, par_v0 = 0, par_pr = 0, par_xr = 0;
      const int ksteps = (w.gpc * N + 15) >> 4;
      int it = 0, vbase = 0;
      int nstamp = 0;
      const bool dbg = w.dbg != nullptr && blockIdx.x == 0;
#define GFC_DSTAMP(tag) do { if (dbg && nstamp < 1000) { w.dbg[2 * nstamp] = clock64(); w.dbg[2 * nstamp + 1] = (tag); ++nstamp; } } while (0)
      for (int tile = t_begin; tile < t_end; ++tile, ++it) {
        GFC_DSTAMP(300);
        tc5::mbar_wait(p_ready, par_pr); par_pr ^= 1;
        GFC_DSTAMP(301);
        const bool fresh = (it % w.flush_every) == 0;   // the workers drained the accumulators before x_ready
        const uint32_t pa = p_addr + (it & 1) * L::P_BYTES;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
          GFC_DSTAMP(310);
          if (k == 0) { tc5::mbar_wait(v0_ready, par_v0); par_v0 ^= 1; }
          else { tc5::mbar_wait(v_ready, par_vr); par_vr ^= 1; }
          tc5::fence_after_sync();
          GFC_DSTAMP(311);
          const uint32_t vs = v_addr + ((vbase + k) % 3) * L::VBUF;
          if (k + 1 < K) {
            // hop: D_hop = P * V_k
            uint32_t acc = 0;
#pragma unroll 1
            for (int j = 0; j < ksteps; ++j) {
              const uint64_t da = make_desc(pa + j * 2 * L::PW, L::PW, 128);
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                tc5::mma_bf16_ss(tmem + TM_HOP, da, make_desc(vs + pl * L::PLANE + j * 256, 128, L::PW), kIdesc, acc);
                acc = 1;
              }
            }
            tc5::mma_commit(hop_done);
            GFC_DSTAMP(312);
          }
          if (k == 0) { tc5::mbar_wait(x_ready, par_xr); par_xr ^= 1; tc5::fence_after_sync(); GFC_DSTAMP(313); }
          // dH: acc(k)[g][f] += X^T[g][rows] * V_k[rows][f]   (6-term bf16x3 product, A from tensor memory)
          const uint32_t d_acc = tmem + L::TM_ACC + k * FH;
#pragma unroll 1
          for (int j = 0; j < ksteps; ++j) {
            const uint32_t xa = tmem + L::TM_X + j * 8;
            const uint64_t b0 = make_desc(vs + j * 256, 128, L::PW);
            const uint64_t b1 = make_desc(vs + L::PLANE + j * 256, 128, L::PW);
            const uint64_t b2 = make_desc(vs + 2 * L::PLANE + j * 256, 128, L::PW);
            const uint32_t first = (fresh && j == 0) ? 0u : 1u;
            tc5::mma_bf16_ts(d_acc, xa, b0, kIdesc, first);
            tc5::mma_bf16_ts(d_acc, xa, b1, kIdesc, 1u);
            tc5::mma_bf16_ts(d_acc, xa + 64, b0, kIdesc, 1u);
            tc5::mma_bf16_ts(d_acc, xa + 64, b1, kIdesc, 1u);
            tc5::mma_bf16_ts(d_acc, xa, b2, kIdesc, 1u);
            tc5::mma_bf16_ts(d_acc, xa + 128, b0, kIdesc, 1u);
          }
          GFC_DSTAMP(314);
        }
        tc5::mma_commit(item_done);
        vbase = (vbase + K) % 3;
      }
    }
    __syncwarp();
  } else if (warp >= 2) {
    // =========================== workers ======================================================
    const int wt = tid - 64;
    const int ww = warp - 2;
    const int q = warp & 3;
    const int hf = ww >> 2;
    const int r = q * 32 + lane;               // tile row (write-back) and feature lane g (X^T, drain)
    const int jr = r / N;
    const bool g_ok = r < G;
    const uint32_t tm_lane = tmem + ((uint32_t)(q * 32) << 16);
    constexpr int PPR = FH / 4, RPI = 32 / PPR;   // V_0 loads: 16-byte pieces per row, rows per warp instruction
    float xt[64];                              // x[g = r][rows 64 hf .. 64 hf + 63] of the next tile
    float xin[FH / 2];                         // V_0 pieces of the next tile (short-lived)
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};     // column sums of V_0 (columns 4*(lane % PPR) .. +3)
    float2 mypos = make_float2(0.f, 0.f);
    uint32_t par_hd = 0, par_id = 0;
    int nstamp = 0;
    const bool dbg = w.dbg != nullptr && blockIdx.x == 0 && tid == 64;
#define GFC_ESTAMP(tag) do { if (dbg && nstamp < 1000) { w.dbg[2048 + 2 * nstamp] = clock64(); w.dbg[2048 + 2 * nstamp + 1] = (tag); ++nstamp; } } while (0)

    auto publish = [&](uint64_t* bar) {
      tc5::fence_proxy_async();
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(bar);
    };
    auto load_pos = [&](int tile) {
      const int b0 = tile * w.gpc;
      const int rows_used = min(w.gpc, w.B - b0) * N;
      if (wt < 128)
        mypos = (wt < rows_used) ? __ldg(reinterpret_cast<const float2*>(w.pos) + (size_t)b0 * N + wt) : make_float2(0.f, 0.f);
    };
    auto prefetch_tile = [&](int tile) {
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      if (w.dpre) {
        tc5::bulk_prefetch_l2(w.dpre + (size_t)b0 * N * F, (uint32_t)gcount * N * F * 4u);
      } else {
        tc5::bulk_prefetch_l2(w.dY + (size_t)b0 * N * F, (uint32_t)gcount * N * F * 4u);
        if (w.act != GFC_ACT_NONE) tc5::bulk_prefetch_l2(w.yout + (size_t)b0 * N * F, (uint32_t)gcount * N * F * 4u);
      }
      if (fh == 0) tc5::bulk_prefetch_l2(w.x + (size_t)b0 * G * N, (uint32_t)gcount * G * N * 4u);
    };
    // X^T operand: x[(b0 + j), g, n] -> TMEM lane g, column (tile row / 2).  The 16x256b store shape lets a
    // thread own 4 consecutive tile rows (one float4 of x) of the feature lanes  16 h + t/4  and  16 h + t/4 + 8:
    // a warp load instruction then touches 8 lines instead of 32.  xt[16 P + 8 e + 4 u + i]: part P = 2 h + e
    // (e selects rho in {2e, 2e+1}), u = rho & 1 ... see store_xt for the register order of the store.
    auto load_xt_part = [&](int tile, auto part_c) {
      constexpr int PART = decltype(part_c)::value;
      constexpr int h16 = PART >> 1, e = PART & 1;
      const int b0 = tile * w.gpc;
      const int rows_used = min(w.gpc, w.B - b0) * N;
#pragma unroll
      for (int u = 0; u < 2; ++u) {          // rho = 2 e + u
#pragma unroll
        for (int gs = 0; gs < 2; ++gs) {     // feature lane t/4 (+ 8)
          const int g = q * 32 + 16 * h16 + 8 * gs + (lane >> 2);
          const int rr = hf * 64 + 16 * (2 * e + u) + 4 * (lane & 3);
          float* dst = &xt[16 * PART + 8 * u + 4 * gs];
          if ((N & 3) == 0) {
            const int j = rr / N, n = rr - j * N;
            const float4 v = ldg_f32x4(w.x + ((size_t)(b0 + j) * G + g) * N + n, g < G && rr < rows_used);
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = (rr + i) / N, n = (rr + i) - j * N;
              dst[i] = ldg_f32(w.x + ((size_t)(b0 + j) * G + g) * N + n, g < G && rr + i < rows_used);
            }
          }
        }
      }
    };
    auto load_xt_slot = [&](int tile, int slot) {
      if (slot == 0) load_xt_part(tile, std::integral_constant<int, 0>{});
      else if (slot == 1) load_xt_part(tile, std::integral_constant<int, 1>{});
      else if (slot == 2) load_xt_part(tile, std::integral_constant<int, 2>{});
      else load_xt_part(tile, std::integral_constant<int, 3>{});
    };
    auto store_xt = [&]() {
      if (q * 32 < G) {   // warp-uniform: this quadrant holds real feature lanes
#pragma unroll
        for (int h16 = 0; h16 < 2; ++h16) {
          uint32_t p0[16], p1[16], p2[16];
#pragma unroll
          for (int rho = 0; rho < 4; ++rho) {
#pragma unroll
            for (int gs = 0; gs < 2; ++gs) {
              const float* src = &xt[16 * (2 * h16 + (rho >> 1)) + 8 * (rho & 1) + 4 * gs];
              uint32_t a0, a1, a2, b0, b1, b2, c0, c1, c2, d0, d1, d2;
              tc5::split_bf16x3(src[0], a0, a1, a2);
              tc5::split_bf16x3(src[1], b0, b1, b2);
              tc5::split_bf16x3(src[2], c0, c1, c2);
              tc5::split_bf16x3(src[3], d0, d1, d2);
              p0[4 * rho + 2 * gs] = tc5::pack_bf16_hi(a0, b0); p0[4 * rho + 2 * gs + 1] = tc5::pack_bf16_hi(c0, d0);
              p1[4 * rho + 2 * gs] = tc5::pack_bf16_hi(a1, b1); p1[4 * rho + 2 * gs + 1] = tc5::pack_bf16_hi(c1, d1);
              p2[4 * rho + 2 * gs] = tc5::pack_bf16_hi(a2, b2); p2[4 * rho + 2 * gs + 1] = tc5::pack_bf16_hi(c2, d2);
            }
          }
          const uint32_t ta = tmem + ((uint32_t)(q * 32 + 16 * h16) << 16) + L::TM_X + hf * 32;
          tc5::tmem_st_16x256b_x4(ta, p0);
          tc5::tmem_st_16x256b_x4(ta + 64, p1);
          tc5::tmem_st_16x256b_x4(ta + 128, p2);
        }
        tc5::tmem_st_wait();
      }
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(x_ready);
    };
    // V_0 = dY o act'(y) of this CTA's feature slice: coalesced pieces -> registers (+ db), then planes
    auto load_v0 = [&](int tile) {
      const int b0 = tile * w.gpc;
      const int rows_used = min(w.gpc, w.B - b0) * N;
#pragma unroll
      for (int i = 0; i < FH / 8; ++i) {
        const int row = 16 * ww + RPI * i + lane / PPR;
        const bool ok = row < rows_used;
        const size_t off = ((size_t)b0 * N + row) * F + fh * FH + 4 * (lane % PPR);
        float4 v;
        if (w.dpre) {
          v = ldg_f32x4(w.dpre + off, ok);      // not consumed before store_v0: the load latency is never exposed
        } else {
          v = ldg_f32x4(w.dY + off, ok);
          if (w.act != GFC_ACT_NONE) {
            const float4 yo = ldg_f32x4(w.yout + off, ok);
            v.x = act_grad(v.x, yo.x, w.act, w.slope);
            v.y = act_grad(v.y, yo.y, w.act, w.slope);
            v.z = act_grad(v.z, yo.z, w.act, w.slope);
            v.w = act_grad(v.w, yo.w, w.act, w.slope);
          }
        }
        xin[4 * i] = v.x; xin[4 * i + 1] = v.y; xin[4 * i + 2] = v.z; xin[4 * i + 3] = v.w;
      }
    };
    auto store_v0 = [&](unsigned char* vbuf) {
      const int pi = lane % PPR;
      const bool odd = pi & 1;
#pragma unroll
      for (int i = 0; i < FH / 8; ++i) {
        dbacc[0] += xin[4 * i]; dbacc[1] += xin[4 * i + 1]; dbacc[2] += xin[4 * i + 2]; dbacc[3] += xin[4 * i + 3];
      }
#pragma unroll
      for (int i = 0; i < FH / 8; i += 2) {
        float snd[4], rcv[4], own[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          snd[e] = odd ? xin[4 * i + e] : xin[4 * (i + 1) + e];
          own[e] = odd ? xin[4 * (i + 1) + e] : xin[4 * i + e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) rcv[e] = __shfl_xor_sync(0xffffffffu, snd[e], 1);
        float v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[e] = odd ? rcv[e] : own[e]; v[4 + e] = odd ? own[e] : rcv[e]; }
        const int row = 16 * ww + RPI * (i + (odd ? 1 : 0)) + lane / PPR;
        store_chunk3(vbuf + (pi >> 1) * L::PW + row * 16, L::PLANE, v);
      }
    };
    // P[r][c] (smem, K-major A operand): chunks t0..t1-1 of the 8 this thread owns
    auto build_p_part = [&](int tile, int pbuf, int t0, int t1) {
      const float2* sp = sp_all + pbuf * L::ROWS;
      unsigned char* pb = Pb + pbuf * L::P_BYTES;
      const int rows_used = min(w.gpc, w.B - tile * w.gpc) * N;
      const int c_lo = jr * N, c_hi = c_lo + N;
      const float2 me = sp[r];
      if (r < w.gpc * N) {
#pragma unroll 1
        for (int t = t0; t < t1; ++t) {
          const int qc = 2 * t + hf;
          if (qc * 8 < c_hi && qc * 8 + 8 > c_lo) {
            uint32_t e[8];
            adjacency8(e, sp, me, r, qc * 8, c_lo, c_hi, rows_used, w.thr, w.thr_lo, w.thr_hi);
            *reinterpret_cast<uint4*>(pb + qc * L::PW + r * 16) =
                make_uint4(tc5::pack_bf16_hi(e[0], e[1]), tc5::pack_bf16_hi(e[2], e[3]), tc5::pack_bf16_hi(e[4], e[5]),
                           tc5::pack_bf16_hi(e[6], e[7]));
          }
        }
      }
    };
    // Drain the TMEM accumulators into this CTA group's partial buffer (zeroed by the host, L2-resident, every
    // address owned by exactly one thread of one CTA) with fire-and-forget reductions: no read latency, and the
    // per-address order is this thread's program order, so the result is deterministic.  The tensor core's fp32
    // accumulation truncates, so the error grows with the number of MMAs chained into one accumulator;
    // draining every `flush_every` tiles bounds it.
    float* dst = w.dHp + (size_t)part * F * K * G;
    auto flush = [&]() {
      tc5::fence_after_sync();
      if (q * 32 < G) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
#pragma unroll 1
          for (int cb = 0; cb < FH / 2; cb += 8) {
            const int col = hf * (FH / 2) + cb;
            uint32_t v[8];
            tc5::tmem_ld8u(tm_lane + L::TM_ACC + k * FH + col, v);
            tc5::tmem_ld_wait();
            if (g_ok) {
              float* d0 = dst + (size_t)(fh * FH + col) * K * G + k * G + r;
#pragma unroll
              for (int i = 0; i < 8; ++i) red_add_f32(d0 + (size_t)i * K * G, __uint_as_float(v[i]));
            }
          }
        }
      }
      tc5::fence_before_sync();
    };

    // One leading iteration (it == -1) prepares the first tile; afterwards iteration `it` serves tile `tile`
    // (write-backs) and prepares `next`.
    int tile = t_begin, next = t_begin, it = -1, vbase = 0, n_items = 0;
    if (next < t_end) load_pos(next);
    while (true) {
      const bool live = it >= 0;
      const bool has_next = next < t_end;
      if (!live && !has_next) break;
      const int pbuf = (it + 1) & 1;
      int xt_slot = 0;
      bool p_done = false;
      if (has_next) {
        if (wt < 128) sp_all[pbuf * L::ROWS + wt] = mypos;
        worker_bar();
        if (next + 1 < t_end) load_pos(next + 1);
        if (wt == 0 && !w.no_prefetch) {   // L2 prefetch runs two tiles ahead
          if (!live) prefetch_tile(next);
          if (next + 1 < t_end) prefetch_tile(next + 1);
        }
      }
      // ---- write-backs of the K-1 hops; the next tile's x columns and P are prepared in the gaps ----------
      const int nwb = live ? K - 1 : 0;
      bool v0_loaded = false;
#pragma unroll 1
      for (int k = 0; k < nwb; ++k) {
        // next tile's V_0 pieces: requested two write-backs ahead of their use so that the (long) memory latency
        // overlaps the remaining taps of the live tile
        if (has_next && !v0_loaded && k + 2 >= nwb) { load_v0(next); v0_loaded = true; }
        GFC_ESTAMP(400);
        tc5::mbar_wait(hop_done, par_hd); par_hd ^= 1;
        tc5::fence_after_sync();
        GFC_ESTAMP(401);
        unsigned char* base = Vb + ((vbase + k + 1) % 3) * L::VBUF + (hf * (L::CPT / 8)) * L::PW + r * 16;
        const uint32_t taddr = tm_lane + TM_HOP + hf * L::CPT;
#pragma unroll 1
        for (int c = 0; c < L::CPT / 8; ++c) {
          uint32_t v[8];
          tc5::tmem_ld8u(taddr + c * 8, v);
          tc5::tmem_ld_wait();
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[i]);
          store_chunk3(base + c * L::PW, L::PLANE, f);
        }
        publish(v_ready);
        GFC_ESTAMP(402);
        if (has_next) {
          if (K > 1 && k < 2) {
            build_p_part(next, pbuf, 4 * k, 4 * k + 4);
            if (k == 1 || nwb == 1) {
              if (nwb == 1) build_p_part(next, pbuf, 4, 8);
              publish(p_ready); p_done = true;
            }
          }
          if (xt_slot < 2) load_xt_slot(next, xt_slot++);   // half of X^T early; the rest once V_0's registers are free
        }
        GFC_ESTAMP(403);
      }
      if (has_next) {
        if (!p_done) { if (K > 1) build_p_part(next, pbuf, 0, 8); publish(p_ready); }
        // V_0 of the next tile goes into the ring buffer after the live tile's last one (free: its last reader
        // was tap K-3 of the live tile, which completed before the hop result of tap K-2 was signalled)
        GFC_ESTAMP(404);
        if (!v0_loaded) load_v0(next);
        GFC_ESTAMP(405);
        store_v0(Vb + ((vbase + (live ? K : 0)) % 3) * L::VBUF);
        publish(v0_ready);
        GFC_ESTAMP(406);
        while (xt_slot < 4) load_xt_slot(next, xt_slot++);   // latency hides behind the last tap's MMAs / the drain
      }
      if (live) {
        tc5::mbar_wait(item_done, par_id); par_id ^= 1;   // every dH product of the live tile has completed
        GFC_ESTAMP(407);
        ++n_items;
        if (!has_next || (n_items % w.flush_every) == 0) flush();
        GFC_ESTAMP(408);
        vbase = (vbase + K) % 3;
      }
      if (has_next) store_xt();
      GFC_ESTAMP(409);
      if (!has_next) break;
      tile = next;
      next += 1;
      ++it;
    }
    if (w.dbp) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(dbs + 4 * (lane % PPR) + i, dbacc[i]);
      worker_bar();
      if (wt < FH) w.dbp[(size_t)part * F + fh * FH + wt] = dbs[wt];
    }
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc5::tmem_dealloc(tmem, 512);
}

// ---- tap packing: bf16x3 planes in ring-stage order ----------------------------------------------
// stage u = k * (CIN/16) + c16 holds channels [16 c16, 16 c16 + 16) of tap k:
//   byte(plane, cc, n, e) = plane*COUT*32 + cc*COUT*16 + n*16 + e*2   with channel = 16 c16 + 8 cc + e
// MODE 0: B[n = f][channel = g] = h[f][k*G + g];  MODE 1: B[n = g][channel = f] = h[f][k*G + g]
__global__ void __launch_bounds__(256)
wide_pack_taps_kernel(const float* __restrict__ h, int G, int F, int K, int mode, uint16_t* __restrict__ out) {
  const int CIN = mode == 0 ? G : F, COUT = mode == 0 ? F : G;
  const int total = K * CIN * COUT;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int k = idx / (CIN * COUT), rem = idx - k * CIN * COUT;
    const int ch = rem / COUT, n = rem - ch * COUT;
    const int f = mode == 0 ? n : ch, g = mode == 0 ? ch : n;
    const float v = h[(size_t)f * K * G + k * G + g];
    uint32_t p0, p1, p2;
    tc5::split_bf16x3(v, p0, p1, p2);
    const int c16 = ch >> 4, cc = (ch >> 3) & 1, e = ch & 7;
    const size_t stage = (size_t)(k * (CIN / 16) + c16) * (COUT * 48);   // in uint16 units (96 B * COUT / 2)
    const size_t o = stage + (size_t)cc * COUT * 8 + (size_t)n * 8 + e;
    out[o] = (uint16_t)(p0 >> 16);
    out[o + (size_t)COUT * 16] = (uint16_t)(p1 >> 16);
    out[o + (size_t)COUT * 32] = (uint16_t)(p2 >> 16);
  }
}

int launch_wide_pack(const float* h, int G, int F, int K, int mode, uint16_t* out, cudaStream_t st) {
  const int total = K * G * F;
  int grid = ceil_div(total, 256);
  if (grid > 592) grid = 592;
  wide_pack_taps_kernel<<<grid, 256, 0, st>>>(h, G, F, K, mode, out);
  GFC_LAUNCH_CHECK("wide_pack_taps_kernel");
  return GFC_OK;
}

bool wide_supported(int N, int G, int F, int K, int mode) {
  const int CIN = mode == 0 ? G : F, COUT = mode == 0 ? F : G;
  if (N < 1 || N > 128 || K < 1 || K > 16) return false;
  return (CIN == 128 || CIN == 64) && (COUT == 128 || COUT == 64);
}

size_t wide_pack_bytes(int G, int F, int K) { return align_up((size_t)K * G * F * 6, 256); }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
static int encode_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                          uint32_t box_outer) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GFC_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    GFC_REQUIRE(p && qres == cudaDriverEntryPointSuccess, GFC_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GFC_REQUIRE(r == CUDA_SUCCESS, GFC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return GFC_OK;
}

template <int CIN, int COUT, int MODE>
static int launch_wide_t(const WideArgs& a0, cudaStream_t st) {
  using L = WideLayout<CIN, COUT>;
  WideArgs a = a0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  a.tma_out = 0;
  if (MODE != 1) {   // y viewed as [B*N rows, COUT cols]; one box = [gpc*N rows x 32 cols]
    int rc = encode_tmap_2d(&tmap, a.out, COUT, (uint64_t)a.B * a.N, 32, (uint32_t)(a.gpc * a.N));
    if (rc) return rc;
    a.tma_out = 1;
  }
  auto kern = tc5_wide_kernel<CIN, COUT, MODE>;
  GFC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const int grid = a.ntiles < di.sm_count ? a.ntiles : di.sm_count;
  kern<<<grid, kWideThreads, L::BYTES, st>>>(a, tmap);
  GFC_LAUNCH_CHECK(MODE == 0 ? "tc5_wide_kernel<fwd>" : MODE == 1 ? "tc5_wide_kernel<dX>" : "tc5_wide_kernel<fwd,node-major in>");
  return GFC_OK;
}

int launch_wide(const WideArgs& a0, int G, int F, int mode, cudaStream_t st) {
  WideArgs a = a0;
  a.no_prefetch = g_wide_no_prefetch;
  a.gpc = 128 / a.N;
  if (a.gpc > a.B) a.gpc = a.B;
  a.ntiles = ceil_div(a.B, a.gpc);
  const int CIN = mode != 1 ? G : F, COUT = mode != 1 ? F : G;
#define GFC_WIDE_CASE(ci, co)                                                   \
  if (CIN == ci && COUT == co)                                                  \
    return mode == 0 ? launch_wide_t<ci, co, 0>(a, st)                          \
         : mode == 1 ? launch_wide_t<ci, co, 1>(a, st) : launch_wide_t<ci, co, 2>(a, st);
  GFC_WIDE_CASE(128, 128)
  GFC_WIDE_CASE(64, 64)
  GFC_WIDE_CASE(128, 64)
  GFC_WIDE_CASE(64, 128)
#undef GFC_WIDE_CASE
  set_error("launch_wide: unsupported channel counts %d -> %d", CIN, COUT);
  return GFC_ERR_UNSUPPORTED;
}


template <int G, int F, int FH>
static int launch_wide_dh_t(const WideDhArgs& a0, int* nparts_out, cudaStream_t st) {
  using L = DhLayout<G, F, FH>;
  WideDhArgs a = a0;
  auto kern = tc5_wide_dh_kernel<G, F, FH>;
  GFC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
  kern<<<a.nparts * L::NFH, kWideThreads, L::BYTES, st>>>(a);
  GFC_LAUNCH_CHECK("tc5_wide_dh_kernel");
  if (nparts_out) *nparts_out = a.nparts;
  return GFC_OK;
}

// feature slice: the K+1 [128 x FH] fp32 regions (K accumulators + the hop result) share 320 TMEM columns
static int dh_slice(int F, int K) {
  if ((K + 1) * 64 <= 320 && F % 64 == 0) return 64;
  if ((K + 1) * 32 <= 320 && F % 32 == 0) return 32;
  return 0;
}

bool wide_dh_supported(int N, int G, int F, int K) {
  if (N < 1 || N > 128 || K < 1) return false;
  if (!(G == 128 || G == 64)) return false;
  if (!(F == 128 || F == 64)) return false;
  return dh_slice(F, K) != 0;
}

int wide_dh_nparts(int B, int N, int F, int K) {
  DeviceInfo di;
  if (get_device_info(&di)) return 0;
  const int fhs = dh_slice(F, K);
  if (!fhs) return 0;
  const int nfh = F / fhs;
  int gpc = 128 / N; if (gpc > B) gpc = B;
  const int ntiles = ceil_div(B, gpc);
  int np = di.sm_count / nfh;
  if (np > ntiles) np = ntiles;
  if (np < 1) np = 1;
  return np;
}

int launch_wide_dh(const WideDhArgs& a0, int G, int F, cudaStream_t st) {
  WideDhArgs a = a0;
  a.gpc = 128 / a.N;
  if (a.gpc > a.B) a.gpc = a.B;
  a.ntiles = ceil_div(a.B, a.gpc);
  a.nparts = wide_dh_nparts(a.B, a.N, F, a.K);
  if (a.flush_every <= 0) a.flush_every = g_wide_flush_every;
  a.no_prefetch = g_wide_no_prefetch;
  const int fhs = dh_slice(F, a.K);
#define GFC_DH_CASE(g, f, s) if (G == g && F == f && fhs == s) return launch_wide_dh_t<g, f, s>(a, nullptr, st);
  GFC_DH_CASE(128, 128, 64)
  GFC_DH_CASE(128, 128, 32)
  GFC_DH_CASE(64, 64, 64)
  GFC_DH_CASE(64, 64, 32)
  GFC_DH_CASE(128, 64, 64)
  GFC_DH_CASE(128, 64, 32)
  GFC_DH_CASE(64, 128, 64)
  GFC_DH_CASE(64, 128, 32)
#undef GFC_DH_CASE
  set_error("launch_wide_dh: unsupported shape G=%d F=%d K=%d", G, F, a.K);
  return GFC_ERR_UNSUPPORTED;
}

}  // namespace gfc
