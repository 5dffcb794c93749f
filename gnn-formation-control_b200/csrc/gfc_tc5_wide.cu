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
// MODE 1: backward dX: V_0 = dY o act'(y), V_k = P V_{k-1}, dX = sum_k V_k H_k^T-contraction
//         (IN = dY [B,N,F], OUT = dX [B,G,N]; CIN = F, COUT = G) — the closed form of the autograd
//         graph of graphML.py:2342-2366 for a symmetric 0/1 GSO.
#include "gfc_common.cuh"
#include "gfc_tc5.cuh"
#include "gfc_tc5_wide.cuh"

namespace gfc {

template <int CIN, int COUT>
struct WideLayout {
  static constexpr int ROWS = 128;
  static constexpr int CS = CIN / 2;                 // channels per slab
  static constexpr int CPT = CS / 2;                 // state columns per worker thread
  static constexpr int NCH = CS / 8;                 // 16-byte chunks (8 bf16) per row and slab
  static constexpr int PW = ROWS * 16;               // bytes between chunks (one chunk column of all rows)
  static constexpr int PLANE = NCH * PW;             // one bf16 plane of a slab
  static constexpr int SLAB = 3 * PLANE;
  static constexpr int P_BYTES = (ROWS / 8) * PW;    // block-diagonal hop matrix, bf16 [128 x 128]
  static constexpr int STAGE = COUT * 96;            // taps of 16 channels: 3 planes x 2 chunks x COUT x 16 B
  static constexpr int NSTAGE = 6;
  static constexpr int KSTEPS = CS / 16;             // tap MMA k-steps (= ring stages) per phase
  static constexpr int OFF_W = 0;
  static constexpr int OFF_P = OFF_W + 2 * SLAB;
  static constexpr int OFF_RING = OFF_P + P_BYTES;
  static constexpr int OFF_SP = OFF_RING + NSTAGE * STAGE;   // float2 positions of the tile rows
  static constexpr int OFF_BAR = OFF_SP + ROWS * 8;
  static constexpr int NBAR = 2 * NSTAGE + 2 + 2 + 1 + 2 + 2;
  static constexpr int BYTES = OFF_BAR + NBAR * 8 + 16;
  static constexpr int TM_OUT = 0;                   // two output accumulators [128 x COUT]
  static constexpr int TM_HOP = 2 * COUT;            // two hop accumulators   [128 x CS]
  static constexpr int TM_USED = 2 * COUT + 2 * CS;
  static constexpr int TM_COLS = TM_USED <= 32 ? 32 : TM_USED <= 64 ? 64 : TM_USED <= 128 ? 128 : TM_USED <= 256 ? 256 : 512;
  static_assert(CIN % 32 == 0 && CIN >= 32 && CIN <= 128, "CIN in {32,64,96,128}");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 128, "COUT multiple of 16, <= 128");
  static_assert(CPT % 8 == 0, "whole chunks per worker thread");
  static_assert(BYTES <= 227 * 1024, "shared memory");
};

constexpr int kWideThreads = 320;   // warp 0: MMA issuer, warp 1: TMA producer + TMEM owner, warps 2..9: workers
constexpr int kWorkerWarps = 8;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (((lbo >> 4) & 0x3fffu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3fffu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ bool wide_adjacent(float2 a, float2 b, const WideArgs& w) {
  const float dx = a.x - b.x, dy = a.y - b.y;
  const float s = fmaf(dx, dx, dy * dy);
  if (s < w.thr_lo) return true;
  if (s > w.thr_hi) return false;
  return sqdist64(a.x, a.y, b.x, b.y) <= w.thr;
}

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// 8 consecutive channels of one row -> one 16-byte chunk in each of the three planes
__device__ __forceinline__ void store_chunk3(unsigned char* plane0, int plane_bytes, const float (&v)[8]) {
  uint32_t a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tc5::split_bf16x3(v[i], a[i], b[i], c[i]);
  *reinterpret_cast<uint4*>(plane0) = make_uint4(tc5::pack_bf16_hi(a[0], a[1]), tc5::pack_bf16_hi(a[2], a[3]),
                                                 tc5::pack_bf16_hi(a[4], a[5]), tc5::pack_bf16_hi(a[6], a[7]));
  *reinterpret_cast<uint4*>(plane0 + plane_bytes) =
      make_uint4(tc5::pack_bf16_hi(b[0], b[1]), tc5::pack_bf16_hi(b[2], b[3]), tc5::pack_bf16_hi(b[4], b[5]),
                 tc5::pack_bf16_hi(b[6], b[7]));
  *reinterpret_cast<uint4*>(plane0 + 2 * plane_bytes) =
      make_uint4(tc5::pack_bf16_hi(c[0], c[1]), tc5::pack_bf16_hi(c[2], c[3]), tc5::pack_bf16_hi(c[4], c[5]),
                 tc5::pack_bf16_hi(c[6], c[7]));
}

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(kWideThreads, 1)
tc5_wide_kernel(const WideArgs w) {
  using L = WideLayout<CIN, COUT>;
  extern __shared__ __align__(128) unsigned char wsmem[];
  unsigned char* Wb = wsmem + L::OFF_W;
  unsigned char* Pb = wsmem + L::OFF_P;
  unsigned char* Rb = wsmem + L::OFF_RING;
  float2* sp = reinterpret_cast<float2*>(wsmem + L::OFF_SP);
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

  // ---- one-time setup -------------------------------------------------------------------------
  for (int i = tid; i < L::P_BYTES / 16; i += kWideThreads) reinterpret_cast<uint4*>(Pb)[i] = make_uint4(0, 0, 0, 0);
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

  if (warp == 0) {
    // =========================== MMA issuer (one thread) ======================================
    if (lane == 0) {
      constexpr uint32_t kIdescTap = tc5::idesc_bf16(128, COUT, 0, 0);
      constexpr uint32_t kIdescHop = tc5::idesc_bf16(128, L::CS, 0, 1);
      const uint32_t w_addr = tc5::smem_u32(Wb), p_addr = tc5::smem_u32(Pb), r_addr = tc5::smem_u32(Rb);
      uint32_t par_wr[2] = {0, 0}, par_of[2] = {0, 0}, par_pr = 0, par_hf = 0;
      int st = 0;
      const int hop_ksteps = (w.gpc * N + 15) >> 4;
      int it = 0;
      for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x, ++it) {
        const int ob = it & 1;
        if (it >= 2) { tc5::mbar_wait(&out_free[ob], par_of[ob]); par_of[ob] ^= 1; }
        tc5::mbar_wait(p_ready, par_pr); par_pr ^= 1;
        const uint32_t d_out = tmem + L::TM_OUT + ob * COUT;
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            tc5::mbar_wait(&w_ready[s], par_wr[s]); par_wr[s] ^= 1;
            tc5::fence_after_sync();
            const uint32_t ws = w_addr + s * L::SLAB;
            if (k + 1 < K) {
              // hop: D_hop[s] = P * W_k[slab s]   (A = P K-major, B = state planes MN-major)
              const uint32_t d_hop = tmem + L::TM_HOP + s * L::CS;
              uint32_t acc = 0;
              for (int j = 0; j < hop_ksteps; ++j) {
                const uint64_t da = make_desc(p_addr + j * 2 * L::PW, L::PW, 128);
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                  const uint64_t db = make_desc(ws + pl * L::PLANE + j * 256, 128, L::PW);
                  tc5::mma_bf16_ss(d_hop, da, db, kIdescHop, acc);
                  acc = 1;
                }
              }
            }
            // taps: D_out += W_k[slab s] * H_k[slab s]   (6-term bf16x3 product)
#pragma unroll 1
            for (int i = 0; i < L::KSTEPS; ++i) {
              tc5::mbar_wait(&h_full[st], par_hf);
              tc5::fence_after_sync();
              const uint32_t hs = r_addr + st * L::STAGE;
              const uint32_t wa = ws + i * 2 * L::PW;
              const uint64_t a0 = make_desc(wa, L::PW, 128);
              const uint64_t a1 = make_desc(wa + L::PLANE, L::PW, 128);
              const uint64_t a2 = make_desc(wa + 2 * L::PLANE, L::PW, 128);
              const uint64_t b0 = make_desc(hs, COUT * 16, 128);
              const uint64_t b1 = make_desc(hs + COUT * 32, COUT * 16, 128);
              const uint64_t b2 = make_desc(hs + COUT * 64, COUT * 16, 128);
              const uint32_t first = (k == 0 && s == 0 && i == 0) ? 0u : 1u;
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
          }
        }
        tc5::mma_commit(&out_full[ob]);
      }
    }
  } else if (warp == 1) {
    // =========================== tap producer (TMA bulk copies) ===============================
    if (lane == 0) {
      int st = 0;
      uint32_t par_he = 0;
      bool primed = false;   // the first NSTAGE fills need no wait
      int filled = 0;
      const int stages_per_tile = K * (CIN / 16);
      for (int tile = blockIdx.x; tile < w.ntiles; tile += gridDim.x) {
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
  } else {
    // =========================== workers ======================================================
    const int wt = tid - 64;                 // 0..255
    const int q = warp & 3;                  // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;        // which half of a slab's columns
    const int r = q * 32 + lane;             // tile row owned by this thread
    const int jr = r / N, nr = r - jr * N;   // (graph, node) of the row
    float xin[2][L::CPT];
    float2 mypos = make_float2(0.f, 0.f);
    uint32_t par_md[2] = {0, 0}, par_ofl[2] = {0, 0};

    auto load_inputs = [&](int tile) {
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      const bool valid = r < gcount * N;
      if (wt < 128) {   // rows 0..127 in thread order wt: position of row wt
        const int rr = wt;
        mypos = (rr < gcount * N) ? __ldg(reinterpret_cast<const float2*>(w.pos) + (size_t)b0 * N + rr)
                                  : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int c0 = s * L::CS + half * L::CPT;
        if (MODE == 0) {
          const float* src = w.in + ((size_t)(b0 + jr) * CIN + c0) * N + nr;
#pragma unroll
          for (int i = 0; i < L::CPT; ++i) xin[s][i] = valid ? __ldg(src + (size_t)i * N) : 0.f;
        } else {
          const size_t off = ((size_t)b0 * N + r) * CIN + c0;
#pragma unroll
          for (int i4 = 0; i4 < L::CPT / 4; ++i4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
              v = __ldg(reinterpret_cast<const float4*>(w.in + off) + i4);
              if (w.act != GFC_ACT_NONE) {
                const float4 yo = __ldg(reinterpret_cast<const float4*>(w.yout + off) + i4);
                v.x = act_grad(v.x, yo.x, w.act, w.slope);
                v.y = act_grad(v.y, yo.y, w.act, w.slope);
                v.z = act_grad(v.z, yo.z, w.act, w.slope);
                v.w = act_grad(v.w, yo.w, w.act, w.slope);
              }
            }
            xin[s][4 * i4] = v.x; xin[s][4 * i4 + 1] = v.y; xin[s][4 * i4 + 2] = v.z; xin[s][4 * i4 + 3] = v.w;
          }
        }
      }
    };
    auto store_w0 = [&](int s) {
      unsigned char* base = Wb + s * L::SLAB + (half * (L::CPT / 8)) * L::PW + r * 16;
#pragma unroll
      for (int c = 0; c < L::CPT / 8; ++c) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = xin[s][c * 8 + i];
        store_chunk3(base + c * L::PW, L::PLANE, v);
      }
    };
    auto publish = [&](uint64_t* bar) {   // this warp's shared-memory writes -> tensor core, then arrive
      tc5::fence_proxy_async();
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(bar);
    };
    auto build_p = [&](int tile) {
      // P[r][c] = 1 iff rows r and c belong to the same graph and are adjacent (symmetric rule)
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      const int rows_used = gcount * N;
      const int pr = wt & 127, hsel = wt >> 7;
      if (pr < w.gpc * N) {
        const int pj = pr / N;
        const int c_lo = pj * N, c_hi = c_lo + N;        // block columns [c_lo, c_hi)
        const float2 me = sp[pr];
        for (int qc = (c_lo >> 3); qc <= ((c_hi - 1) >> 3); ++qc) {
          if ((qc & 1) != hsel) continue;
          uint32_t e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = qc * 8 + i;
            bool on = false;
            if (c >= c_lo && c < c_hi && c != pr && pr < rows_used && c < rows_used) on = wide_adjacent(me, sp[c], w);
            e[i] = on ? 0x3f800000u : 0u;
          }
          *reinterpret_cast<uint4*>(Pb + qc * L::PW + pr * 16) =
              make_uint4(tc5::pack_bf16_hi(e[0], e[1]), tc5::pack_bf16_hi(e[2], e[3]), tc5::pack_bf16_hi(e[4], e[5]),
                         tc5::pack_bf16_hi(e[6], e[7]));
        }
      }
    };

    // ---- prologue: first tile's operands -----------------------------------------------------
    int tile = blockIdx.x;
    if (tile < w.ntiles) {
      load_inputs(tile);
      if (wt < 128) sp[wt] = mypos;
      worker_bar();
      if (K > 1) build_p(tile);
      publish(p_ready);
      store_w0(0); publish(&w_ready[0]);
      store_w0(1); publish(&w_ready[1]);
    }
    int it = 0;
    for (; tile < w.ntiles; tile += gridDim.x, ++it) {
      const int next = tile + gridDim.x;
      const bool has_next = next < w.ntiles;
      const int b0 = tile * w.gpc;
      const int gcount = min(w.gpc, w.B - b0);
      const int rows_used = gcount * N;
      if (has_next) load_inputs(next);       // in flight during this tile's tensor-core phases
      for (int k = 0; k + 1 < K; ++k) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          tc5::mbar_wait(&mma_done[s], par_md[s]); par_md[s] ^= 1;
          tc5::fence_after_sync();
          // hop result (exact fp32) -> three bf16 planes of W_{k+1}[slab s]
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + L::TM_HOP + s * L::CS + half * L::CPT;
          unsigned char* base = Wb + s * L::SLAB + (half * (L::CPT / 8)) * L::PW + r * 16;
          if constexpr (L::CPT == 32) {
            uint32_t v[32];
            tc5::tmem_ld32(taddr, v);
            tc5::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[c * 8 + i]);
              store_chunk3(base + c * L::PW, L::PLANE, f);
            }
          } else {
            uint32_t v[16];
            tc5::tmem_ld16(taddr, v);
            tc5::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[c * 8 + i]);
              store_chunk3(base + c * L::PW, L::PLANE, f);
            }
          }
          publish(&w_ready[s]);
        }
      }
      // ---- tail: all hops of this tile are done -> next tile's P; slabs free one by one --------
      if (has_next) {
        if (wt < 128) sp[wt] = mypos;
        worker_bar();
        if (K > 1) build_p(next);
        publish(p_ready);
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        tc5::mbar_wait(&mma_done[s], par_md[s]); par_md[s] ^= 1;
        if (has_next) { store_w0(s); publish(&w_ready[s]); }
      }
      // ---- epilogue of this tile (the issuer is already working on the next one) ---------------
      const int ob = it & 1;
      tc5::mbar_wait(&out_full[ob], par_ofl[ob]); par_ofl[ob] ^= 1;
      tc5::fence_after_sync();
      const bool valid = r < rows_used;
#pragma unroll 1
      for (int cb = 0; cb < COUT / 2; cb += 16) {
        const int col = half * (COUT / 2) + cb;
        uint32_t v[16];
        tc5::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + L::TM_OUT + ob * COUT + col, v);
        tc5::tmem_ld_wait();
        if (valid) {
          if (MODE == 0) {
            float* dst = w.out + ((size_t)b0 * N + r) * COUT + col;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              float o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float bb = w.bias ? __ldg(w.bias + col + i4 * 4 + i) : 0.f;
                o[i] = apply_act(__uint_as_float(v[i4 * 4 + i]) + bb, w.act, w.slope);
              }
              reinterpret_cast<float4*>(dst)[i4] = make_float4(o[0], o[1], o[2], o[3]);
            }
          } else {
            float* dst = w.out + ((size_t)(b0 + jr) * COUT + col) * N + nr;
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[(size_t)i * N] = __uint_as_float(v[i]);
          }
        }
      }
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(&out_free[ob]);
    }
  }
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc5::tmem_dealloc(tmem, L::TM_COLS);
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

template <int CIN, int COUT, int MODE>
static int launch_wide_t(const WideArgs& a, cudaStream_t st) {
  using L = WideLayout<CIN, COUT>;
  auto kern = tc5_wide_kernel<CIN, COUT, MODE>;
  GFC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  const int grid = a.ntiles < di.sm_count ? a.ntiles : di.sm_count;
  kern<<<grid, kWideThreads, L::BYTES, st>>>(a);
  GFC_LAUNCH_CHECK(MODE == 0 ? "tc5_wide_kernel<fwd>" : "tc5_wide_kernel<dX>");
  return GFC_OK;
}

int launch_wide(const WideArgs& a0, int G, int F, int mode, cudaStream_t st) {
  WideArgs a = a0;
  a.gpc = 128 / a.N;
  if (a.gpc > a.B) a.gpc = a.B;
  a.ntiles = ceil_div(a.B, a.gpc);
  const int CIN = mode == 0 ? G : F, COUT = mode == 0 ? F : G;
#define GFC_WIDE_CASE(ci, co)                                                   \
  if (CIN == ci && COUT == co)                                                  \
    return mode == 0 ? launch_wide_t<ci, co, 0>(a, st) : launch_wide_t<ci, co, 1>(a, st);
  GFC_WIDE_CASE(128, 128)
  GFC_WIDE_CASE(64, 64)
  GFC_WIDE_CASE(128, 64)
  GFC_WIDE_CASE(64, 128)
#undef GFC_WIDE_CASE
  set_error("launch_wide: unsupported channel counts %d -> %d", CIN, COUT);
  return GFC_ERR_UNSUPPORTED;
}

}  // namespace gfc
