// gfc_tc5_small.cu — tcgen05 / TMEM version of the fused filter FORWARD for small static
// shapes (cfg2: N=8, G=F=32, K=3).  Same math as tile_fwd_kernel (gfc_tile_kernels.cuh);
// the tap contraction runs on the 5th-generation tensor cores:
//   * every (graph, feature) thread keeps its diffusion column in registers; per tap k it
//     writes the hi/lo TF32 split of z_k straight into a UMMA canonical K-major panel buffer
//     (double-buffered over taps), then runs the next hop in registers while
//   * ONE elected thread issues the 3xTF32 tcgen05.mma chain of that tap
//     (D[128 x F] in TMEM += Z_k[128 x G] * H_k^T), committing to an mbarrier;
//   * the epilogue reads the accumulator with tcgen05.ld, adds bias, applies the activation
//     and stores node-major y — 32-byte sectors, fully written.
// Reference semantics: BatchLSIGF utils/graphUtils/graphML.py:2273-2367.
#include "gfc_tile_kernels.cuh"
#include "gfc_tc5.cuh"

namespace gfc {

template <typename CFG>
struct Tc5FwdLayout {
  static constexpr int N = CFG::sN, G = CFG::sG, F = CFG::sF, K = CFG::sK, KG = K * G;
  static constexpr int ROWS = 128;                          // UMMA M
  static constexpr int GPC = ROWS / N;                      // graphs per tile
  static constexpr int PA = tc5::panel_floats(ROWS);        // A panel stride (floats)
  static constexpr int PB = tc5::panel_floats(F);           // B panel stride (floats)
  static constexpr int A_TERM = (G / 4) * PA;               // one tap, one term (hi or lo)
  static constexpr int OFF_A = 0;                           // [2 buffers][2 terms][G/4 panels][PA]
  static constexpr int OFF_B = OFF_A + 4 * A_TERM;          // [2 terms][KG/4 panels][PB]
  static constexpr int B_TERM = (KG / 4) * PB;
  static constexpr int OFF_S = OFF_B + 2 * B_TERM;          // [GPC][N][N]
  static constexpr int OFF_ISD = OFF_S + GPC * N * N;       // doubles [GPC*N]
  static constexpr int OFF_BAR = OFF_ISD + 2 * GPC * N;     // 2 mbarriers (uint64) + tmem ptr
  static constexpr int TOTAL = OFF_BAR + 8;
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
  static constexpr int TMEM_COLS = F < 32 ? 32 : F;         // power of two >= 32 (F in {32,64,128,256})
};

template <typename CFG, int GSRC>
__global__ void __launch_bounds__(CFG::kThreads, 2)
tc5_fwd_kernel(const TileArgs a) {
  using L = Tc5FwdLayout<CFG>;
  constexpr int N = L::N, G = L::G, F = L::F, K = L::K;
  extern __shared__ __align__(16) float smem[];
  float* As = smem + L::OFF_A;
  float* Bs = smem + L::OFF_B;
  float* Ss = smem + L::OFF_S;
  double* isd = reinterpret_cast<double*>(smem + L::OFF_ISD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_BAR + 4);
  const TilePlan& p = a.p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool single = a.single_pass != 0;
  GFC_STAMP(a, 7);
  GFC_STAMP_NS(a, 8);

  // the first tile's x column is requested before the one-time setup so its DRAM latency
  // overlaps the TMEM allocation and the tap packing
  const int j = tid / G, gcol = tid - j * G;
  float z[N];
  {
    const int b0 = blockIdx.x * L::GPC;
    const int gcount = min(L::GPC, p.B - b0);
    if ((int)blockIdx.x < p.ntiles && tid < gcount * G) load_x_column<CFG>(z, a.x, b0, j, gcol, a.vec_ok);
  }
  // ---- one-time setup: TMEM, mbarriers, taps -> B operand panels (hi / lo) ----------------
  if (warp == 0) tc5::tmem_alloc(tmem_ptr, L::TMEM_COLS);
  if (tid == 32) {
    tc5::mbar_init(&bars[0], 1);
    tc5::mbar_init(&bars[1], 1);
    tc5::fence_mbar_init();
  }
  for (int idx = tid; idx < F * L::KG; idx += CFG::kThreads) {
    const int f = idx / L::KG, c = idx - f * L::KG;      // B[n = f][k = c] = h[f][c]
    uint32_t hi, lo;
    split_tf32(__ldg(a.h + idx), hi, lo);
    const int off = tc5::panel_off(f, c, L::PB);
    Bs[off] = __uint_as_float(hi);
    Bs[L::B_TERM + off] = __uint_as_float(lo);
  }
  tc5::fence_proxy_async();
  tc5::fence_before_sync();
  __syncthreads();
  tc5::fence_after_sync();
  const uint32_t tmem_d = *tmem_ptr;
  constexpr uint32_t kIdesc = tc5::idesc_tf32(L::ROWS, F, 0, 0);
  const uint32_t a_base = tc5::smem_u32(As), b_base = tc5::smem_u32(Bs);
  uint32_t nwait0 = 0, nwait1 = 0;  // completed waits per mbarrier (phase parity bookkeeping)
  GFC_STAMP(a, 0);

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int b0 = tile * L::GPC;
    const int gcount = min(L::GPC, p.B - b0);
    const int rows_used = gcount * N;
    const bool has_col = tid < gcount * G;
    if (has_col) {
      if (tile != (int)blockIdx.x) load_x_column<CFG>(z, a.x, b0, j, gcol, a.vec_ok);
    } else {
#pragma unroll
      for (int n = 0; n < N; ++n) z[n] = 0.f;
    }
    if (GSRC == GSRC_POS) gso_tile_from_global<CFG>(Ss, isd, a, b0, gcount);
    else load_gso_tile<CFG, GSRC>(Ss, nullptr, isd, a, b0, gcount);
    // (the barrier that publishes Ss is the first per-tap barrier below)
    if (tile == (int)blockIdx.x) GFC_STAMP(a, 10);

#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int buf = k & 1;
      if (k >= 2) {  // the MMAs of tap k-2 must have finished reading this buffer
        if (buf == 0) { tc5::mbar_wait(&bars[0], nwait0 & 1); ++nwait0; }
        else { tc5::mbar_wait(&bars[1], nwait1 & 1); ++nwait1; }
        if (k == 2 && tile == (int)blockIdx.x) GFC_STAMP(a, 15);
      }
      if (tid < L::GPC * G) {  // rows of absent graphs (tail tile) are written as zeros
        float* Ahi = As + (size_t)(buf * 2) * L::A_TERM;
        float* Alo = Ahi + L::A_TERM;
#pragma unroll
        for (int n = 0; n < N; ++n) {
          uint32_t hi, lo;
          split_tf32(z[n], hi, lo);
          const int off = tc5::panel_off(j * N + n, gcol, L::PA);
          Ahi[off] = __uint_as_float(hi);
          Alo[off] = __uint_as_float(lo);
        }
      }
      if (k == 0 && tile == (int)blockIdx.x) GFC_STAMP(a, 11);
      tc5::fence_proxy_async();
      if (k == 0 && tile == (int)blockIdx.x) GFC_STAMP(a, 12);
      __syncthreads();
      if (k == 0 && tile == (int)blockIdx.x) GFC_STAMP(a, 1);
      if (k == 1 && tile == (int)blockIdx.x) GFC_STAMP(a, 14);
      if (warp == 0 && tc5::elect_one()) {
        tc5::fence_after_sync();
        const uint32_t a_hi = a_base + (uint32_t)((buf * 2) * L::A_TERM * 4);
        const uint32_t a_lo = a_hi + (uint32_t)(L::A_TERM * 4);
        const uint32_t b_hi = b_base + (uint32_t)((k * (G / 4)) * L::PB * 4);
        const uint32_t b_lo = b_hi + (uint32_t)(L::B_TERM * 4);
#pragma unroll
        for (int s = 0; s < G / 8; ++s) {
          const uint32_t ao = (uint32_t)(s * 2 * L::PA * 4), bo = (uint32_t)(s * 2 * L::PB * 4);
          const uint64_t dah = tc5::smem_desc(a_hi + ao, L::PA * 4, 128);
          const uint64_t dal = tc5::smem_desc(a_lo + ao, L::PA * 4, 128);
          const uint64_t dbh = tc5::smem_desc(b_hi + bo, L::PB * 4, 128);
          const uint64_t dbl = tc5::smem_desc(b_lo + bo, L::PB * 4, 128);
          const uint32_t first = (k == 0 && s == 0) ? 0u : 1u;
          if (!single) {
            tc5::mma_tf32_ss(tmem_d, dal, dbh, kIdesc, first);
            tc5::mma_tf32_ss(tmem_d, dah, dbl, kIdesc, 1u);
            tc5::mma_tf32_ss(tmem_d, dah, dbh, kIdesc, 1u);
          } else {
            tc5::mma_tf32_ss(tmem_d, dah, dbh, kIdesc, first);
          }
        }
        tc5::mma_commit(&bars[buf]);
        if (k == 0 && tile == (int)blockIdx.x) GFC_STAMP(a, 13);
      }
      if (k + 1 < K && has_col) {  // next hop in registers while the tensor core works
        const float* Sj = Ss + (size_t)j * N * N;
        float zn[N];
#pragma unroll
        for (int n = 0; n < N; ++n) zn[n] = 0.f;
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const float zm = z[m];
          if ((N & 3) == 0) {  // 128-bit broadcast loads of row m of S_j
#pragma unroll
            for (int q = 0; q < N / 4; ++q) {
              const float4 sv = *reinterpret_cast<const float4*>(Sj + m * N + 4 * q);
              zn[4 * q] = fmaf(sv.x, zm, zn[4 * q]);
              zn[4 * q + 1] = fmaf(sv.y, zm, zn[4 * q + 1]);
              zn[4 * q + 2] = fmaf(sv.z, zm, zn[4 * q + 2]);
              zn[4 * q + 3] = fmaf(sv.w, zm, zn[4 * q + 3]);
            }
          } else {
#pragma unroll
            for (int n = 0; n < N; ++n) zn[n] = fmaf(Sj[m * N + n], zm, zn[n]);
          }
        }
#pragma unroll
        for (int n = 0; n < N; ++n) z[n] = zn[n];
      }
    }
    if (tile == (int)blockIdx.x) GFC_STAMP(a, 2);
    // ---- wait for the outstanding commits (in order), then the epilogue from TMEM ---------
    if (K >= 2) {
      if (((K - 2) & 1) == 0) { tc5::mbar_wait(&bars[0], nwait0 & 1); ++nwait0; }
      else { tc5::mbar_wait(&bars[1], nwait1 & 1); ++nwait1; }
    }
    if (((K - 1) & 1) == 0) { tc5::mbar_wait(&bars[0], nwait0 & 1); ++nwait0; }
    else { tc5::mbar_wait(&bars[1], nwait1 & 1); ++nwait1; }
    tc5::fence_after_sync();
    {
      const int q = warp & 3, cg0 = warp >> 2;           // TMEM lane quadrant, first column group
      const int r = q * 32 + lane;
      for (int cg = cg0; cg < F / 8; cg += CFG::kWarps / 4) {
        float v[8];
        tc5::tmem_ld8(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 8), v);
        if (r < rows_used) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float bb = a.bias ? __ldg(a.bias + cg * 8 + i) : 0.f;
            v[i] = apply_act(v[i] + bb, a.act, a.slope);
          }
          float4* dst = reinterpret_cast<float4*>(a.y + ((size_t)b0 * N + r) * F + cg * 8);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
    tc5::fence_before_sync();
    __syncthreads();  // accumulator, operand buffers and Ss are reused by the next tile
    if (tile == (int)blockIdx.x) GFC_STAMP(a, 3);
  }
  GFC_STAMP_NS(a, 9);
  if (warp == 0) tc5::tmem_dealloc(tmem_d, L::TMEM_COLS);
}

using Cfg_tc5_n8 = TileCfg<8, 32, 32, 3, 512, 2>;

int tc5_fwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st) {
  using L = Tc5FwdLayout<Cfg_tc5_n8>;
  TileArgs b = a;
  b.p.gpc = L::GPC;
  b.p.ntiles = ceil_div(a.p.B, L::GPC);
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  int per_sm = (int)((size_t)di.smem_optin / (L::BYTES + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  b.p.grid = di.sm_count * per_sm < b.p.ntiles ? di.sm_count * per_sm : b.p.ntiles;
  b.p.threads = Cfg_tc5_n8::kThreads;
  b.p.smem_bytes = L::BYTES;
  if (gsrc == GSRC_POS) return launch_tile_kernel(tc5_fwd_kernel<Cfg_tc5_n8, GSRC_POS>, b, st, "tc5_fwd<pos>");
  return launch_tile_kernel(tc5_fwd_kernel<Cfg_tc5_n8, GSRC_DENSE>, b, st, "tc5_fwd<dense>");
}

}  // namespace gfc
