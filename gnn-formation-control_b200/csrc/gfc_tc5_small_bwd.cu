// gfc_tc5_small_bwd.cu — tcgen05 / TMEM version of the fused filter BACKWARD for small static shapes
// (cfg2: N=8, G=F=32, K=3).  It uses the closed form of the autograd graph of BatchLSIGF
// (utils/graphUtils/graphML.py:2342-2366) that needs no recomputation of the diffusion states:
//      V_0 = dY o act'(y),   V_k = V_{k-1} . S^T   (per graph, row-vector convention, any S)
//      dX  = sum_k V_k H_k          ->  ONE MMA chain  [128 rows x K*F] . [K*F x G]
//      dH  = [V_0|V_1|V_2]^T X      ->  ONE MMA chain  [K*F (lanes) x 128 rows] . [128 rows x G]
//      db  = column sums of V_0
// Every (graph, column) thread keeps its column of V in registers across the hops (FP32 pipes, S as
// shared-memory broadcasts) and writes the hi/lo TF32 split of every V_k and of its x column into UMMA
// operand panels; one elected thread issues the two 3xTF32 chains; dX leaves straight from TMEM, the dH
// tile is drained into fp32 registers after every tile (the tensor core's accumulation truncates; the
// running sums over a CTA's tiles are ordinary round-to-nearest adds) and written once per CTA.
#include "gfc_tile_kernels.cuh"
#include "gfc_tc5.cuh"

namespace gfc {

// Operand layout.  A (graph j, column c) thread owns the 8 consecutive tile rows j*8..j*8+7 of its column, i.e.
// exactly one 16-byte bf16 chunk of the TRANSPOSED matrices
//     VT[c][row]   (c = k*F + f, all taps side by side)      XT[g][row]
// stored as  byte(c, row) = (row / 8) * PANEL + c * 16 + (row % 8) * 2  — the UMMA canonical no-swizzle layout.
// The same bytes are
//   * a K-major A operand with M = c,   K = row   (dH  = VT . X : SBO = 128 B between 8-c groups, LBO = PANEL)
//   * an MN-major A operand with M = row, K = c   (dX  = V . H  : SBO = PANEL between 8-row groups, LBO = 128 B)
// (tf32 operands must be K-major — an MN-major tf32 MMA silently produces zeros on sm_100a — so this kernel
// computes in bf16x3: three planes by successive truncation, 6-term products, error ~2^-23 like the wide path.)
template <typename CFG>
struct Tc5BwdLayout {
  static constexpr int N = CFG::sN, G = CFG::sG, F = CFG::sF, K = CFG::sK, KF = K * F, KG = K * G;
  static_assert(N == 8, "one 16-byte chunk = the 8 nodes of a graph");
  static_assert(G == 32 && F == 32, "one (graph, column) thread per warp lane");
  static_assert(KF <= 128 && KF % 16 == 0, "the dH tile has K*F <= 128 lanes");
  static constexpr int ROWS = 128;
  static constexpr int GPC = ROWS / N;                      // = number of 8-row groups
  static constexpr int PV = KF * 16;                        // bytes per 8-row group of VT (one chunk per column)
  static constexpr int PX = G * 16;
  static constexpr int PH = G * 16;                         // taps: B[n = g][k = c], chunks of 8 c
  static constexpr int V_PLANE = GPC * PV;                  // 24 KB
  static constexpr int X_PLANE = GPC * PX;                  // 8 KB
  static constexpr int H_PLANE = (KF / 8) * PH;             // 6 KB
  static constexpr int OFF_V = 0;                           // 3 planes; the M = 128 over-read of plane 2 runs into XT
  static constexpr int OFF_X = OFF_V + 3 * V_PLANE;
  static constexpr int VX_BYTES = 3 * V_PLANE + 3 * X_PLANE; // one operand buffer; two of them (tile t's MMAs || tile t+1's hops)
  static constexpr int OFF_H = 2 * VX_BYTES;
  static constexpr int OFF_S = OFF_H + 3 * H_PLANE;         // fp32 [GPC][N][N]
  static constexpr int OFF_ISD = OFF_S + GPC * N * N * 4;   // doubles [GPC*N]
  static constexpr int OFF_DB = OFF_ISD + GPC * N * 8;
  static constexpr int OFF_SP = OFF_DB + F * 4;             // float2 positions of the tile rows
  static constexpr int OFF_BAR = OFF_SP + ROWS * 8;         // 5 mbarriers + tmem ptr
  static constexpr size_t BYTES = OFF_BAR + 64;
  static constexpr int TMEM_COLS = 64;                      // dX accumulator [128 x G] + dH accumulator [KF x G]
};

// GSO tile from positions staged in shared memory (binary modes and sym-norm; barriers taken by the whole CTA)
template <typename CFG>
__device__ __forceinline__ void gso_tile_from_smem(float* __restrict__ Ss, double* __restrict__ isd,
                                                   const float2* __restrict__ sp, const TileArgs& a, int gcount) {
  constexpr int N = CFG::sN;
  const int nn = gcount * N, total = nn * N;
  for (int o = threadIdx.x; o < total; o += CFG::kThreads) {
    const int r = o / N, n2 = o - r * N;
    const int jj = r / N, m = r - jj * N;
    if (m < n2) {
      const int qq = jj * N + n2;
      const float2 pi = sp[r], pj = sp[qq];
      const float v = pair_adjacent(pi.x, pi.y, pj.x, pj.y, a) ? 1.f : 0.f;
      Ss[o] = v;
      Ss[(size_t)qq * N + m] = v;
    } else if (m == n2) {
      Ss[o] = 0.f;
    }
  }
  if (a.norm) {
    asm volatile("bar.sync 1, 512;" ::: "memory");
    for (int r = threadIdx.x; r < nn; r += CFG::kThreads) {
      float deg = 0.f;
#pragma unroll
      for (int m = 0; m < N; ++m) deg += Ss[(size_t)r * N + m];
      isd[r] = inv_sqrt_deg((int)deg);
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    for (int o = threadIdx.x; o < total; o += CFG::kThreads) {
      const int r = o / N, n2 = o - r * N;
      const int jj = r / N;
      if (Ss[o] != 0.f) Ss[o] = (float)__dmul_rn(isd[r], isd[jj * N + n2]);
    }
  }
}

constexpr int kBwdProducers = 512;             // 16 warps: one (graph, column) thread each, also the epilogue
constexpr int kBwdThreads = kBwdProducers + 64;  // + warp 16 (dX chain issuer) + warp 17 (dH chain issuer)
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <typename CFG, int GSRC>
__global__ void __launch_bounds__(kBwdThreads, 1)
tc5_bwd_kernel(const TileArgs a) {
  using L = Tc5BwdLayout<CFG>;
  constexpr int N = L::N, G = L::G, F = L::F, K = L::K, KF = L::KF, KG = L::KG;
  extern __shared__ __align__(16) float smem[];
  unsigned char* sm8 = reinterpret_cast<unsigned char*>(smem);
  unsigned char* Vb = sm8 + L::OFF_V;
  unsigned char* Xb = sm8 + L::OFF_X;
  unsigned char* Hb = sm8 + L::OFF_H;
  float* Ss = reinterpret_cast<float*>(sm8 + L::OFF_S);
  double* isd = reinterpret_cast<double*>(sm8 + L::OFF_ISD);
  float* dbs = reinterpret_cast<float*>(sm8 + L::OFF_DB);
  float2* sp = reinterpret_cast<float2*>(sm8 + L::OFF_SP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm8 + L::OFF_BAR);
  uint64_t* ops_ready = bars + 2;      // [2] producers -> issuers: operand buffer written (16 warp arrivals)
  uint64_t* acc_free = bars + 4;       // producers -> issuers: the TMEM accumulators have been drained
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sm8 + L::OFF_BAR + 48);
  const TilePlan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = warp, c = lane;                       // (graph of the tile, column f == g)
  const bool want_dx = a.dX != nullptr, want_dh = a.dHp != nullptr, want_db = a.dbp != nullptr;
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);

  // ---- operands of the first tile are requested before the one-time setup -------------------------
  float xcol[N], dyc[N], yoc[N];
  float2 mypos = make_float2(0.f, 0.f);
  auto request = [&](int tile) {
    const int b0 = tile * L::GPC;
    const bool ok = j < min(L::GPC, p.B - b0);
    if (GSRC == GSRC_POS && tid < L::ROWS)
      mypos = (tid < min(L::GPC, p.B - b0) * N) ? __ldg(reinterpret_cast<const float2*>(a.pos) + (size_t)b0 * N + tid)
                                                : make_float2(0.f, 0.f);
    if (want_dh) {
      if (ok) load_x_column<CFG>(xcol, a.x, b0, j, c, a.vec_ok);
      else {
#pragma unroll
        for (int n = 0; n < N; ++n) xcol[n] = 0.f;
      }
    }
    const float* dsrc = a.dY + ((size_t)(b0 + j) * N) * F + c;
    const float* ysrc = (a.act != GFC_ACT_NONE) ? a.yout + ((size_t)(b0 + j) * N) * F + c : nullptr;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      dyc[n] = ok ? __ldg(dsrc + (size_t)n * F) : 0.f;
      yoc[n] = (ok && ysrc) ? __ldg(ysrc + (size_t)n * F) : 1.f;
    }
  };
  const bool producer = tid < kBwdProducers;
  if (producer && (int)blockIdx.x < p.ntiles) request(blockIdx.x);

  // ---- one-time setup: TMEM, mbarriers, taps -> bf16x3 planes B[n = g][k = c] ----------------------
  if (warp == 0) tc5::tmem_alloc(tmem_ptr, L::TMEM_COLS);
  if (tid == 32) {
    tc5::mbar_init(&bars[0], 1);
    tc5::mbar_init(&bars[1], 1);
    tc5::mbar_init(&ops_ready[0], 16);
    tc5::mbar_init(&ops_ready[1], 16);
    tc5::mbar_init(acc_free, 16);
    tc5::fence_mbar_init();
  }
  if (want_dx) {
    for (int idx = tid; idx < F * KG; idx += kBwdThreads) {
      const int f = idx / KG, cc = idx - f * KG;
      const int k = cc / G, g = cc - k * G;            // B[n = g][c = k*F + f] = h[f][k*G + g]
      uint32_t p0, p1, p2;
      tc5::split_bf16x3(__ldg(a.h + idx), p0, p1, p2);
      const int cidx = k * F + f;
      unsigned short* dst = reinterpret_cast<unsigned short*>(Hb + (cidx >> 3) * L::PH + g * 16 + (cidx & 7) * 2);
      dst[0] = (unsigned short)(p0 >> 16);
      dst[L::H_PLANE / 2] = (unsigned short)(p1 >> 16);
      dst[L::H_PLANE] = (unsigned short)(p2 >> 16);
    }
  }
  if (tid < F) dbs[tid] = 0.f;
  tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  constexpr uint32_t kIdescDx = tc5::idesc_bf16(128, G, 1, 0);   // A = VT read MN-major (M = row), B = taps K-major
  constexpr uint32_t kIdescDh = tc5::idesc_bf16(128, G, 0, 0);   // A = VT K-major (M = c), B = XT K-major
  const uint32_t v_base = tc5::smem_u32(Vb), x_base = tc5::smem_u32(Xb), h_base = tc5::smem_u32(Hb);
  const uint64_t dxa0 = tc5::make_desc(v_base, 128, L::PV), dxa1 = tc5::make_desc(v_base + L::V_PLANE, 128, L::PV),
                 dxa2 = tc5::make_desc(v_base + 2 * L::V_PLANE, 128, L::PV);
  const uint64_t dxb0 = tc5::make_desc(h_base, L::PH, 128), dxb1 = tc5::make_desc(h_base + L::H_PLANE, L::PH, 128),
                 dxb2 = tc5::make_desc(h_base + 2 * L::H_PLANE, L::PH, 128);
  const uint64_t dha0 = tc5::make_desc(v_base, L::PV, 128), dha1 = tc5::make_desc(v_base + L::V_PLANE, L::PV, 128),
                 dha2 = tc5::make_desc(v_base + 2 * L::V_PLANE, L::PV, 128);
  const uint64_t dhb0 = tc5::make_desc(x_base, L::PX, 128), dhb1 = tc5::make_desc(x_base + L::X_PLANE, L::PX, 128),
                 dhb2 = tc5::make_desc(x_base + 2 * L::X_PLANE, L::PX, 128);
  const int q = warp & 3, cg = warp >> 2;              // TMEM lane quadrant, 8-column group of this warp
  float hacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hacc[i] = 0.f;
  float dbreg = 0.f;
  uint32_t it = 0;
  GFC_STAMP(a, 0);

  // produce(tile, buf): GSO tile, V_0 / x chunks, the K-1 hops in registers, all operand planes of buffer `buf`;
  // then requests the registers of the tile after it.  Runs on the CUDA cores while the tensor core works on
  // the previous tile's buffer.
  auto produce = [&](int tile, int buf) {
    const int b0 = tile * L::GPC;
    const int gcount = min(L::GPC, p.B - b0);
    unsigned char* Vt = Vb + buf * L::VX_BYTES;
    unsigned char* Xt = Xb + buf * L::VX_BYTES;
    if (GSRC == GSRC_POS) {
      // positions were requested a tile ahead: stage them, then the pair tests run out of shared memory
      if (tid < L::ROWS) sp[tid] = mypos;
      producer_bar();
      gso_tile_from_smem<CFG>(Ss, isd, sp, a, gcount);
    } else {
      load_gso_tile<CFG, GSRC>(Ss, nullptr, isd, a, b0, gcount);
    }
    float v[N];
#pragma unroll
    for (int n = 0; n < N; ++n) v[n] = act_grad(dyc[n], yoc[n], a.act, a.slope);
    if (want_db) {
#pragma unroll
      for (int n = 0; n < N; ++n) dbreg += v[n];
    }
    if (want_dh) tc5::store_chunk3(Xt + j * L::PX + c * 16, L::X_PLANE, xcol);
    const int nxt = tile + gridDim.x;
    if (nxt < p.ntiles) request(nxt);
    producer_bar();                                     // Ss visible
    const float* Sj = Ss + (size_t)j * N * N;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      tc5::store_chunk3(Vt + j * L::PV + (k * F + c) * 16, L::V_PLANE, v);
      if (k + 1 < K) {   // V_{k+1}[n] = sum_m S[n][m] V_k[m]   (row n of S_j as 128-bit broadcasts)
        float vn[N];
#pragma unroll
        for (int n = 0; n < N; ++n) {
          float s = 0.f;
#pragma unroll
          for (int m4 = 0; m4 < N / 4; ++m4) {
            const float4 sv = *reinterpret_cast<const float4*>(Sj + n * N + 4 * m4);
            s = fmaf(sv.x, v[4 * m4], s); s = fmaf(sv.y, v[4 * m4 + 1], s);
            s = fmaf(sv.z, v[4 * m4 + 2], s); s = fmaf(sv.w, v[4 * m4 + 3], s);
          }
          vn[n] = s;
        }
#pragma unroll
        for (int n = 0; n < N; ++n) v[n] = vn[n];
      }
    }
    // publish the buffer to the issuers (and order this warp's reads of Ss / sp before the next produce)
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncwarp();
    if (lane == 0) tc5::mbar_arrive(&ops_ready[buf]);
    producer_bar();
  };

  if (!producer) {
    // =========================== issuers: warp 16 -> dX chain, warp 17 -> dH chain ==========================
    // (different accumulators; the tensor pipe takes the instructions of both elected threads as they arrive;
    //  descriptors are built once and advanced by adding to their 14-bit address field)
    const bool is_dx = warp == 16;
    if ((is_dx ? want_dx : want_dh) && tc5::elect_one()) {
      uint32_t par_ops = 0, par_acc = 0, itl = 0;
      int buf = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++itl, buf ^= 1) {
        tc5::mbar_wait(&ops_ready[buf], (par_ops >> buf) & 1); par_ops ^= 1u << buf;
        if (itl > 0) { tc5::mbar_wait(acc_free, par_acc); par_acc ^= 1; }
        tc5::fence_after_sync();
        const uint64_t bo = (uint64_t)((buf * L::VX_BYTES) >> 4);
        if (is_dx) {
          // dX[row][g] = sum_c V[row][c] H[c][g]: 16 columns c per MMA (two 8-c groups, 128 B apart)
          uint64_t a0 = dxa0 + bo, a1 = dxa1 + bo, a2 = dxa2 + bo, b0d = dxb0, b1d = dxb1, b2d = dxb2;
#pragma unroll 1
          for (int s = 0; s < KF / 16; ++s) {
            tc5::mma_bf16_ss(tmem, a0, b0d, kIdescDx, s == 0 ? 0u : 1u);
            tc5::mma_bf16_ss(tmem, a0, b1d, kIdescDx, 1u);
            tc5::mma_bf16_ss(tmem, a1, b0d, kIdescDx, 1u);
            tc5::mma_bf16_ss(tmem, a1, b1d, kIdescDx, 1u);
            tc5::mma_bf16_ss(tmem, a0, b2d, kIdescDx, 1u);
            tc5::mma_bf16_ss(tmem, a2, b0d, kIdescDx, 1u);
            a0 += 256 >> 4; a1 += 256 >> 4; a2 += 256 >> 4;
            b0d += (2 * L::PH) >> 4; b1d += (2 * L::PH) >> 4; b2d += (2 * L::PH) >> 4;
          }
          tc5::mma_commit(&bars[0]);
        } else {
          // dH[c][g] = sum_rows VT[c][row] XT[g][row]: 16 rows (two graphs) per MMA
          uint64_t a0 = dha0 + bo, a1 = dha1 + bo, a2 = dha2 + bo, b0d = dhb0 + bo, b1d = dhb1 + bo, b2d = dhb2 + bo;
#pragma unroll 1
          for (int s = 0; s < L::ROWS / 16; ++s) {
            tc5::mma_bf16_ss(tmem + G, a0, b0d, kIdescDh, s == 0 ? 0u : 1u);
            tc5::mma_bf16_ss(tmem + G, a0, b1d, kIdescDh, 1u);
            tc5::mma_bf16_ss(tmem + G, a1, b0d, kIdescDh, 1u);
            tc5::mma_bf16_ss(tmem + G, a1, b1d, kIdescDh, 1u);
            tc5::mma_bf16_ss(tmem + G, a0, b2d, kIdescDh, 1u);
            tc5::mma_bf16_ss(tmem + G, a2, b0d, kIdescDh, 1u);
            a0 += (2 * L::PV) >> 4; a1 += (2 * L::PV) >> 4; a2 += (2 * L::PV) >> 4;
            b0d += (2 * L::PX) >> 4; b1d += (2 * L::PX) >> 4; b2d += (2 * L::PX) >> 4;
          }
          tc5::mma_commit(&bars[1]);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== producers / epilogue ========================================================
    int buf = 0;
    if ((int)blockIdx.x < p.ntiles) produce(blockIdx.x, 0);
    if ((int)blockIdx.x < p.ntiles) GFC_STAMP(a, 1);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it, buf ^= 1) {
      const int b0 = tile * L::GPC;
      const int rows_used = min(L::GPC, p.B - b0) * N;
      // ---- the next tile's operands are produced while the tensor core works on this one ----------------
      {
        const int nxt = tile + gridDim.x;
        if (nxt < p.ntiles) produce(nxt, buf ^ 1);
      }
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 2);
      // ---- dX straight from TMEM: this thread's row (j', n'), 8 channels g ------------------------------
      if (want_dx) {
        tc5::mbar_wait(&bars[0], it & 1);
        tc5::fence_after_sync();
        uint32_t r8[8];
        tc5::tmem_ld8u(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 8), r8);
        tc5::tmem_ld_wait();
        const int r = q * 32 + lane;
        if (r < rows_used) {
          const int jj = r / N, nn = r - jj * N;
          float* dst = a.dX + ((size_t)(b0 + jj) * G + cg * 8) * N + nn;
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[(size_t)i * N] = __uint_as_float(r8[i]);
        }
      }
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 3);
      // ---- dH tile -> running fp32 sums (lane = (k, f), 8 channels g) -----------------------------------
      if (want_dh) {
        tc5::mbar_wait(&bars[1], it & 1);
        tc5::fence_after_sync();
        uint32_t r8[8];
        tc5::tmem_ld8u(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(G + cg * 8), r8);
        tc5::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) hacc[i] += __uint_as_float(r8[i]);
      }
      // the accumulators may be overwritten by the next tile's chains
      tc5::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc5::mbar_arrive(acc_free);
      if (tile == (int)blockIdx.x) GFC_STAMP(a, 4);
    }
  }
  tc5::fence_before_sync();
  __syncthreads();

  // ---- per-CTA partials: dH[f][k*G + g] from lane (k, f) = q*32 + lane, db[f] ---------------------------
  if (want_dh && producer) {
    const int row = q * 32 + lane;
    if (row < KF) {
      const int k = row / F, f = row - k * F;
      float* dst = a.dHp + (size_t)blockIdx.x * F * KG + (size_t)f * KG + k * G + cg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(hacc[0], hacc[1], hacc[2], hacc[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(hacc[4], hacc[5], hacc[6], hacc[7]);
    }
  }
  if (want_db) {
    if (producer) atomicAdd(dbs + c, dbreg);
    __syncthreads();
    if (tid < F) a.dbp[(size_t)blockIdx.x * F + tid] = dbs[tid];
  }
  GFC_STAMP(a, 6);
  GFC_STAMP_NS(a, 9);
  tc5::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc5::tmem_dealloc(tmem, L::TMEM_COLS);
}

using Cfg_tc5b_n8 = TileCfg<8, 32, 32, 3, 512, 2>;

// persistent grid of this kernel (one CTA per SM): the caller sizes the partial buffers with it
int tc5_bwd_grid_n8_32_32_3(int B) {
  DeviceInfo di;
  if (get_device_info(&di)) return 0;
  const int ntiles = ceil_div(B, Tc5BwdLayout<Cfg_tc5b_n8>::GPC);
  return ntiles < di.sm_count ? ntiles : di.sm_count;
}

int tc5_bwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st) {
  using L = Tc5BwdLayout<Cfg_tc5b_n8>;
  TileArgs b = a;
  b.p.gpc = L::GPC;
  b.p.ntiles = ceil_div(a.p.B, L::GPC);
  b.p.grid = tc5_bwd_grid_n8_32_32_3(a.p.B);
  b.p.threads = kBwdThreads;
  b.p.smem_bytes = L::BYTES;
  if (gsrc == GSRC_POS) return launch_tile_kernel(tc5_bwd_kernel<Cfg_tc5b_n8, GSRC_POS>, b, st, "tc5_bwd<pos>");
  return launch_tile_kernel(tc5_bwd_kernel<Cfg_tc5b_n8, GSRC_DENSE>, b, st, "tc5_bwd<dense>");
}

}  // namespace gfc
