// gfc_api.cu — the extern "C" surface declared in include/gfc.h: validation,
// dispatch between the fused tile kernels (path A) and the workspace pipeline
// (path B, incl. the CSR variant), second-stage gradient reductions.
#include "gfc_tile.cuh"
#include "gfc_generic.cuh"
#include "gfc_tc5_wide.cuh"
#include "gfc_dp.cuh"
#include <string.h>
#include <math.h>

namespace gfc {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static int g_skip_grad_reduce = 0;
static thread_local float* g_next_stats = nullptr;   // gfc_use_stats: applies to the next filter call of this thread
static thread_local uint32_t* g_next_mask = nullptr; // gfc_use_mask: likewise
static thread_local size_t g_next_mask_bytes = 0;
static thread_local int g_mask_filled = 0;           // gfc_mask_filled: did the last forward call of this thread write its mask
static long long* g_dbg_clk = nullptr;  // device buffer [>= grid][16], see gfc_set_debug_clock_buffer

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int& launch_counter() { return g_launches; }

int get_device_info(DeviceInfo* out) {
  static thread_local int cached_dev = -1;
  static thread_local DeviceInfo cached;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    // no device visible (CPU-only planning / unit tests): assume one B200
    (void)cudaGetLastError();
    out->sm_count = 148; out->cc_major = 10; out->cc_minor = 0; out->smem_optin = 232448;
    return GFC_OK;
  }
  if (dev != cached_dev) {
    GFC_CUDA_TRY(cudaDeviceGetAttribute(&cached.sm_count, cudaDevAttrMultiProcessorCount, dev));
    GFC_CUDA_TRY(cudaDeviceGetAttribute(&cached.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    GFC_CUDA_TRY(cudaDeviceGetAttribute(&cached.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    GFC_CUDA_TRY(cudaDeviceGetAttribute(&cached.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cached_dev = dev;
  }
  *out = cached;
  return GFC_OK;
}

int gso_mode_threshold(int mode, double radius, double* thr, bool* norm);  // gfc_gso.cu

static int g_csr_fused = 1;
static int g_wide_fwd_mask = 1;        // gfc_set_option(GFC_OPT_WIDE_FWD_MASK)
static int g_wide_mask_handover = 1;   // gfc_set_option(GFC_OPT_WIDE_MASK_HANDOVER)
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// fp32 screening band around the fp64 threshold: the fp32 squared distance differs from the
// fp64 one by < 2^-22 relative, the band is 4e-6 wide on each side; inside it the kernels
// fall back to the exact fp64 rule.
static void set_thresholds(TileArgs& a, double thr, bool norm) {
  a.thr = thr;
  a.norm = norm ? 1 : 0;
  if (!(thr >= 0)) { a.thr_lo = -1.f; a.thr_hi = -1.f; return; }
  a.thr_lo = nextafterf((float)(thr * (1.0 - 4e-6)), -INFINITY);
  a.thr_hi = nextafterf((float)(thr * (1.0 + 4e-6)), INFINITY);
}

struct GsoSrc {
  int kind;  // GSRC_DENSE / GSRC_POS
  const float* S;
  const float* pos;
  double radius;
  int mode;
  int binary = 0;   // dense: the caller vouches that every entry is 0 or 1 (GFC_PREC_FLAG_BINARY_GSO)
  int shared = 0;   // dense: S is [E,N,N], one GSO for the whole batch (GFC_PREC_FLAG_SHARED_GSO)
};
static thread_local int g_last_path = 0;   // gfc_last_path(): kernel family of this thread's last filter call

// ---- path B workspace plan -----------------------------------------------------
struct GenericPlan {
  int nsplit;          // split of the dH reduction
  int nchunks, rows_per_chunk;  // db partial sums
  size_t ws_s, ws_z, ws_d, ws_dhp, ws_dbp, ws_bytes;
  int fused_bwd;       // CSR: the whole backward of a graph runs in one CTA (gfc_csr_fused.cu)
  size_t ws_fdh, ws_fdb;   //   ... its per-graph partials [B][F*C] and [B][F]
  int rows_ok;         // the tap contractions run on the tensor-core tile kernels ("rows" plan, VAR_ROWS)
  TilePlan rows;       //   ... with this plan; its partial buffers / packed taps live at ws_rows
  size_t ws_rows;
};

// Tap contractions of the workspace pipeline on the tensor cores: the states Zw are a row-major [rows x C] matrix,
// i.e. `rows` one-node graphs with C input features and one tap, so the fused tile kernels apply unchanged
// (forward: Y = act(Zw H^T + b); backward: D = dY o act', db, dH += D^T Zw, U = D H written over Zw in place).
static bool rows_plan(long long rows, long long C, int F, int backward, TilePlan* p) {
  if (rows <= 0 || rows > 0x7fffffffLL || C > 8192) return false;
  if (!plan_tile((int)rows, 1, (int)C, F, 1, backward, GSRC_DENSE, p)) return false;
  return p->variant == VAR_ROWS;
}

static void plan_generic(int B, int N, int G, int F, int K, int E, int backward, int need_s, GenericPlan* g) {
  const long long rows = (long long)B * N;
  const long long C = (long long)E * K * G;
  size_t off = 0;
  g->ws_s = off; if (need_s) off += align_up((size_t)rows * N * sizeof(float), 256);
  g->ws_z = off; off += align_up((size_t)rows * C * sizeof(float), 256);
  g->ws_d = off;
  g->nsplit = 1; g->nchunks = 1; g->rows_per_chunk = (int)(rows > 0 ? rows : 1);
  g->ws_dhp = g->ws_dbp = off;
  if (backward) {
    off += align_up((size_t)rows * F * sizeof(float), 256);
    const long long blocks = (long long)ceil_div(F, 32) * ((C + 31) / 32);
    long long ns = (148LL * 4 + blocks - 1) / blocks;
    long long maxns = (rows + 255) / 256;
    if (ns > maxns) ns = maxns;
    if (ns > 128) ns = 128;
    if (ns < 1) ns = 1;
    g->nsplit = (int)ns;
    g->ws_dhp = off; off += align_up((size_t)g->nsplit * F * C * sizeof(float), 256);
    long long nc = (rows + 511) / 512;
    if (nc > 1024) nc = 1024;
    if (nc < 1) nc = 1;
    g->nchunks = (int)nc;
    g->rows_per_chunk = (int)((rows + nc - 1) / nc);
    g->nchunks = (int)((rows + g->rows_per_chunk - 1) / g->rows_per_chunk);
    if (g->nchunks < 1) g->nchunks = 1;
    g->ws_dbp = off; off += align_up((size_t)g->nchunks * F * sizeof(float), 256);
  }
  g->fused_bwd = (backward && E == 1 && !need_s && csr_bwd_fused_supported(N, G, F, K, nullptr)) ? 1 : 0;
  g->ws_fdh = g->ws_fdb = off;
  if (g->fused_bwd) {
    g->ws_fdh = off; off += align_up((size_t)B * F * C * sizeof(float), 256);
    g->ws_fdb = off; off += align_up((size_t)B * F * sizeof(float), 256);
  }
  g->rows_ok = rows_plan(rows, C, F, backward, &g->rows) ? 1 : 0;
  g->ws_rows = off;
  if (g->rows_ok) off += align_up(g->rows.ws_bytes, 256);
  g->ws_bytes = off;
}

static int check_common(const char* fn, int B, int N, int G, int F, int K, int E, int act, int prec, float slope = 0.f) {
  // the backward recovers the activation mask from the sign of the forward OUTPUT (act_grad): valid for slope >= 0 only
  GFC_REQUIRE(act != GFC_ACT_LEAKY_RELU || slope >= 0.f, GFC_ERR_BAD_ARG, "%s: negative_slope %g < 0 is not supported",
              fn, (double)slope);
  GFC_REQUIRE(B >= 0 && N > 0 && G > 0 && F > 0 && K > 0 && E > 0, GFC_ERR_BAD_ARG,
              "%s: bad shape B=%d N=%d G=%d F=%d K=%d E=%d", fn, B, N, G, F, K, E);
  GFC_REQUIRE(act >= GFC_ACT_NONE && act <= GFC_ACT_LEAKY_RELU, GFC_ERR_BAD_ARG, "%s: bad activation %d", fn, act);
  GFC_REQUIRE(prec == GFC_PREC_FP32_3XTF32 || prec == GFC_PREC_TF32 || prec == GFC_PREC_F16, GFC_ERR_BAD_ARG,
              "%s: bad precision %d", fn, prec);
  return GFC_OK;
}

static int need_ws(const char* fn, const void* ws, size_t have, size_t need) {
  if (need == 0) return GFC_OK;
  GFC_REQUIRE(ws != nullptr && have >= need, GFC_ERR_WORKSPACE, "%s: workspace %zu B < required %zu B", fn, have, need);
  GFC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, GFC_ERR_WORKSPACE, "%s: workspace must be 256-byte aligned", fn);
  return GFC_OK;
}

static int rows_fwd(const GenericPlan& g, char* wsb, const float* Zw, const float* h, const float* bias, float* y,
                    int act, float slope, int prec, cudaStream_t st) {
  const TilePlan& p = g.rows;
  TileArgs a{};
  a.S = Zw;   // one dummy weight per row (K = 1: no hop ever reads it)
  a.x = Zw; a.h = h; a.bias = bias; a.y = y;
  a.act = act; a.slope = slope; a.single_pass = (prec != GFC_PREC_FP32_3XTF32);
  a.vec_ok = aligned16(Zw) && aligned16(y);
  a.p = p;
  if (!p.h_smem) {
    float4* hp = reinterpret_cast<float4*>(wsb + g.ws_rows + p.ws_hpack);
    int rc = launch_pack_taps(h, p.F, p.KG, 0, hp, st);
    if (rc) return rc;
    a.hpack = hp;
  }
  return launch_tile_fwd(a, GSRC_DENSE, st);
}

// Zw holds the states Z on entry (read only when dH is wanted) and U = D H on return (when want_u)
static int rows_bwd(const GenericPlan& g, char* wsb, float* Zw, const float* h, const float* yout, const float* dY,
                    bool want_u, float* dH, float* db, int act, float slope, int prec, const DpCtx* dp,
                    cudaStream_t st) {
  const TilePlan& p = g.rows;
  const size_t nH = (size_t)p.F * p.KG;
  TileArgs a{};
  a.S = Zw;
  a.x = Zw; a.h = h; a.yout = yout; a.dY = dY;
  a.dX = want_u ? Zw : nullptr;
  a.dHp = dH ? reinterpret_cast<float*>(wsb + g.ws_rows + p.ws_dhp) : nullptr;
  a.dbp = db ? reinterpret_cast<float*>(wsb + g.ws_rows + p.ws_dbp) : nullptr;
  a.act = act; a.slope = slope; a.single_pass = (prec != GFC_PREC_FP32_3XTF32);
  a.vec_ok = aligned16(Zw) && aligned16(dY) && (!yout || aligned16(yout));
  a.p = p;
  int rc;
  if (!p.h_smem && want_u) {
    float4* hp = reinterpret_cast<float4*>(wsb + g.ws_rows + p.ws_hpack);
    rc = launch_pack_taps(h, p.F, p.KG, 1, hp, st);
    if (rc) return rc;
    a.hpack = hp;
  }
  if (dH && !p.acc_regs) GFC_CUDA_TRY(cudaMemsetAsync(a.dHp, 0, (size_t)p.nparts * nH * sizeof(float), st));
  rc = launch_tile_bwd(a, GSRC_DENSE, st);
  if (rc) return rc;
  if (!dH && !db) return GFC_OK;
  if (dp) return launch_reduce_allreduce(a.dHp, p.nparts, (int)nH, a.dbp, p.grid, p.F, dH, *dp, st);
  return launch_reduce_parts(dH ? a.dHp : nullptr, p.nparts, (int)nH, dH, db ? a.dbp : nullptr, p.grid, p.F, db, st);
}

// ---- tcgen05 wide path eligibility -------------------------------------------------
// positions (binary or sym-norm rule), G, F in {64,128}, N <= 128, fp16 headroom c (K-1) <= 21
static bool use_wide(const GsoSrc& gs, int N, int G, int F, int K, int mode, int prec) {
  (void)prec;
  if (g_disable_tcgen05) return false;
  if (gs.kind == GSRC_POS) return wide_supported(N, G, F, K, mode);
  // dense 0/1 GSO (the reference's own addGSO call): P is exact in fp16; a node may have N neighbours (self loop)
  return gs.binary && wide_supported(N + 1, G, F, K, mode) && N <= 128;
}
static int wide_cshift_of(const GsoSrc& gs, int N, bool norm) { return wide_cshift(gs.kind == GSRC_POS ? N : N + 1, norm); }
// fp16 planes per operand: 2 (hi/lo, fp32-equivalent) unless a looser mode was asked for AND the single-plane kernels
// are instantiated for the shape (128 -> 128, dH feature slice 64: K <= 5); elsewhere the looser modes simply get the
// two-plane kernels (more accurate than requested, never less)
static int wide_planes(int prec, int G, int F, int K) {
  if (prec == GFC_PREC_FP32_3XTF32) return 2;
  return (G == 128 && F == 128 && (K + 1) * 64 <= 384) ? 1 : 2;
}
// workspace behind the tile plan's own: [packed taps | amax (256 B) | dH partials | db partials | dY o act'(y)]
struct WideWs {
  size_t pack, amax, dhp, dbp, dpre, bytes;
};
static WideWs wide_ws(int B, int N, int G, int F, int K, int backward) {
  WideWs o{};
  size_t off = 0;
  if (wide_supported(N, G, F, K, 0) || wide_supported(N, G, F, K, 1)) {
    o.pack = off; off += wide_pack_bytes(G, F, K);
    o.amax = off; off += 256;
    o.dhp = o.dbp = o.dpre = off;
    if (backward && wide_dh_supported(N, G, F, K)) {
      const size_t np = (size_t)wide_dh_nparts(B, N, F, K);
      o.dhp = off; off += align_up(np * F * K * G * sizeof(float), 256);
      o.dbp = off; off += align_up(np * F * sizeof(float), 256);
      // hand-over from the dX kernel to the dH kernel: the activation mask as bits, 2 KB per tile (default), or
      // dY o act'(y) itself (GFC_OPT_WIDE_MASK_HANDOVER = 0)
      int gpc = 128 / (N > 0 ? N : 1); if (gpc > B) gpc = B; if (gpc < 1) gpc = 1;
      const size_t ntiles = ((size_t)B + gpc - 1) / gpc;
      const size_t full = (size_t)B * N * F * sizeof(float), bits = ntiles * 2048;
      o.dpre = off; off += align_up(full > bits ? full : bits, 256);
    }
  }
  o.bytes = off;
  return o;
}
static size_t wide_ws_extra(int B, int N, int G, int F, int K, int backward) { return wide_ws(B, N, G, F, K, backward).bytes; }
static void fill_wide_graph(WideGraph& g, const GsoSrc& gs, const TileArgs& a, bool norm, int N, int transpose) {
  g = WideGraph{};
  g.pos = gs.pos; g.thr = a.thr; g.thr_lo = a.thr_lo; g.thr_hi = a.thr_hi; g.norm = norm ? 1 : 0;
  if (gs.kind == GSRC_DENSE) { g.pos = nullptr; g.S = gs.S; g.s_bstride = gs.shared ? 0 : (long long)N * N; g.s_transpose = transpose; g.norm = 0; }
}

// ---- forward ------------------------------------------------------------------
static int filter_fwd_impl(const char* fn, const GsoSrc& gs, const float* x, const float* h, const float* bias,
                           float* y, int B, int N, int G, int F, int K, int E, int act, float slope, int prec,
                           void* ws, size_t ws_bytes, cudaStream_t st) {
  launch_counter() = 0;
  float* stats = g_next_stats;   // one-shot
  g_next_stats = nullptr;
  uint32_t* fmask = g_next_mask;
  const size_t fmask_bytes = g_next_mask_bytes;
  g_next_mask = nullptr; g_next_mask_bytes = 0; g_mask_filled = 0;
  int rc = check_common(fn, B, N, G, F, K, E, act, prec, slope);
  if (rc) return rc;
  if (stats) GFC_CUDA_TRY(cudaMemsetAsync(stats, 0, 4 * sizeof(float), st));   // stats[3] stays 0: "not filled"
  if (B == 0) return GFC_OK;
  GFC_REQUIRE(x && h && y, GFC_ERR_BAD_ARG, "%s: NULL tensor pointer", fn);
  GFC_REQUIRE(gs.kind == GSRC_DENSE ? gs.S != nullptr : gs.pos != nullptr, GFC_ERR_BAD_ARG, "%s: NULL graph source", fn);
  double thr = 0; bool norm = false;
  if (gs.kind == GSRC_POS) {
    GFC_REQUIRE(E == 1, GFC_ERR_UNSUPPORTED, "%s: position-built GSO has E = 1", fn);
    rc = gso_mode_threshold(gs.mode, gs.radius, &thr, &norm);
    if (rc) return rc;
  }
  TilePlan p;
  g_last_path = 2;
  if (E == 1 && plan_tile(B, N, G, F, K, 0, gs.kind, &p)) {
    g_last_path = 1;
    rc = need_ws(fn, ws, ws_bytes, p.ws_bytes + wide_ws_extra(B, N, G, F, K, 0));
    if (rc) return rc;
    TileArgs a{};
    a.S = gs.S; a.pos = gs.pos; set_thresholds(a, thr, norm);
    a.s_bstride = gs.shared ? 0 : (long long)N * N;
    a.x = x; a.h = h; a.bias = bias; a.y = y;
    a.act = act; a.slope = slope; a.single_pass = (prec != GFC_PREC_FP32_3XTF32);
    a.vec_ok = aligned16(x) && aligned16(y) && (gs.kind == GSRC_POS || aligned16(gs.S));
    a.p = p;
    a.dbg_clk = g_dbg_clk;
    if (use_wide(gs, N, G, F, K, 0, prec) && a.vec_ok && aligned16(gs.pos) && aligned16(h)) {
      // tcgen05 / TMEM path: fp16 hi/lo planes, hops and taps on the tensor cores
      const WideWs wws = wide_ws(B, N, G, F, K, 0);
      unsigned char* hp = reinterpret_cast<unsigned char*>(ws) + p.ws_bytes + wws.pack;
      const int cs = wide_cshift_of(gs, N, norm), np = wide_planes(prec, G, F, K);
      rc = launch_wide_pack(h, G, F, K, 0, cs, np, hp, st);
      if (rc) return rc;
      WideArgs wa{};
      fill_wide_graph(wa.g, gs, a, norm, N, 1);   // forward hop: z_{k+1} = z_k S  ->  P = S^T
      g_last_path = 3;
      wa.in = x; wa.hpack = hp; wa.bias = bias; wa.out = y; wa.B = B; wa.N = N; wa.K = K; wa.cshift = cs;
      wa.act = act; wa.slope = slope; wa.dbg = g_dbg_clk;
      wa.amax = stats;   // stats[0] = max |x|: a by-product of the per-tile operand scales
      if (fmask && g_wide_fwd_mask && act != GFC_ACT_NONE && fmask_bytes >= wide_mask_bytes(B, N)) {
        wa.fmask_out = fmask;   // signs of the outputs, 2 KB per tile: the backward call of this batch then never reads y
        g_mask_filled = 1;
      }
      wa.mark_stats = stats ? 1 : 0;   // stats[3] = 1 ("filled") is written by the kernel itself
      return launch_wide(wa, G, F, 0, np, st);
    }
    if (!p.h_smem) {
      float4* hp = reinterpret_cast<float4*>(static_cast<char*>(ws) + p.ws_hpack);
      rc = launch_pack_taps(h, F, p.KG, 0, hp, st);
      if (rc) return rc;
      a.hpack = hp;
    }
    return launch_tile_fwd(a, gs.kind, st);
  }
  // path B
  GenericPlan g;
  plan_generic(B, N, G, F, K, E, 0, gs.kind == GSRC_POS, &g);
  rc = need_ws(fn, ws, ws_bytes, g.ws_bytes);
  if (rc) return rc;
  char* wsb = static_cast<char*>(ws);
  const float* S = gs.S;
  const int s_shared = (gs.kind == GSRC_DENSE && gs.shared) ? 1 : 0;
  if (gs.kind == GSRC_POS) {
    float* Sw = reinterpret_cast<float*>(wsb + g.ws_s);
    int lc = launch_counter();
    rc = gfc_gso_build(gs.pos, B, N, gs.radius, gs.mode, nullptr, Sw, st);
    launch_counter() += lc;
    if (rc) return rc;
    S = Sw;
  }
  float* Zw = reinterpret_cast<float*>(wsb + g.ws_z);
  const long long C = (long long)E * K * G;
  rc = launch_xpose_in(x, Zw, B, N, G, E, K, st);
  if (rc) return rc;
  for (int k = 1; k < K; ++k) {
    rc = launch_hop_dense(Zw, S, B, N, G, E, K, k - 1, k, 0, s_shared, st);
    if (rc) return rc;
  }
  if (g.rows_ok) return rows_fwd(g, wsb, Zw, h, bias, y, act, slope, prec, st);
  return launch_sgemm(Zw, C, 1, h, 1, C, y, F, 0, (long long)B * N, F, C, 1, bias, act, slope, st);
}

// ---- backward -----------------------------------------------------------------
// dp != nullptr: dH / db are the two halves of one flat bucket (db == dH + F*E*K*G) and the second-stage
// reduction is fused with the all-reduce over peer memory (gfc_dp.cu)
static int filter_bwd_impl(const char* fn, const GsoSrc& gs, const float* x, const float* h, const float* yout,
                           const float* dY, float* dX, float* dH, float* db,
                           int B, int N, int G, int F, int K, int E, int act, float slope, int prec,
                           void* ws, size_t ws_bytes, cudaStream_t st, const DpCtx* dp = nullptr) {
  launch_counter() = 0;
  const float* stats = g_next_stats;   // one-shot
  g_next_stats = nullptr;
  const uint32_t* fmask = g_next_mask;   // filled by the forward call of this batch (the caller vouches: gfc_mask_filled)
  if (fmask && g_next_mask_bytes < wide_mask_bytes(B, N)) fmask = nullptr;
  g_next_mask = nullptr; g_next_mask_bytes = 0;
  int rc = check_common(fn, B, N, G, F, K, E, act, prec, slope);
  if (rc) return rc;
  const size_t nH = (size_t)F * E * K * G;
  if (B == 0) {
    if (dH) GFC_CUDA_TRY(cudaMemsetAsync(dH, 0, nH * sizeof(float), st));
    if (db) GFC_CUDA_TRY(cudaMemsetAsync(db, 0, (size_t)F * sizeof(float), st));
    if (dp) return launch_reduce_allreduce(dH, 1, (int)nH, db, 1, F, dH, *dp, st);
    return GFC_OK;
  }
  GFC_REQUIRE(h && dY, GFC_ERR_BAD_ARG, "%s: NULL tensor pointer", fn);
  GFC_REQUIRE(!dH || x, GFC_ERR_BAD_ARG, "%s: x is required for dH", fn);
  GFC_REQUIRE(!dp || (dH && db && db == dH + nH), GFC_ERR_BAD_ARG, "%s: the dp bucket must be contiguous [dH | db]", fn);
  GFC_REQUIRE(act == GFC_ACT_NONE || yout, GFC_ERR_BAD_ARG, "%s: y_out is required with a fused activation", fn);
  GFC_REQUIRE(gs.kind == GSRC_DENSE ? gs.S != nullptr : gs.pos != nullptr, GFC_ERR_BAD_ARG, "%s: NULL graph source", fn);
  if (!dX && !dH && !db) return GFC_OK;
  double thr = 0; bool norm = false;
  if (gs.kind == GSRC_POS) {
    GFC_REQUIRE(E == 1, GFC_ERR_UNSUPPORTED, "%s: position-built GSO has E = 1", fn);
    rc = gso_mode_threshold(gs.mode, gs.radius, &thr, &norm);
    if (rc) return rc;
  }
  TilePlan p;
  g_last_path = 2;
  if (E == 1 && plan_tile(B, N, G, F, K, 1, gs.kind, &p)) {
    g_last_path = 1;
    rc = need_ws(fn, ws, ws_bytes, p.ws_bytes + wide_ws_extra(B, N, G, F, K, 1));
    if (rc) return rc;
    char* wsb = static_cast<char*>(ws);
    TileArgs a{};
    a.S = gs.S; a.pos = gs.pos; set_thresholds(a, thr, norm);
    a.s_bstride = gs.shared ? 0 : (long long)N * N;
    a.x = x; a.h = h; a.yout = yout; a.dY = dY; a.dX = dX;
    a.dHp = dH ? reinterpret_cast<float*>(wsb + p.ws_dhp) : nullptr;
    a.dbp = db ? reinterpret_cast<float*>(wsb + p.ws_dbp) : nullptr;
    a.act = act; a.slope = slope; a.single_pass = (prec != GFC_PREC_FP32_3XTF32);
    a.vec_ok = aligned16(dY) && (!x || aligned16(x)) && (!yout || aligned16(yout)) &&
               (gs.kind == GSRC_POS || aligned16(gs.S));
    a.p = p;
    a.dbg_clk = g_dbg_clk;
    if (use_wide(gs, N, G, F, K, 1, prec) && a.vec_ok && aligned16(gs.pos) && aligned16(h) && (!dX || aligned16(dX)) &&
        (!dH || wide_dh_supported(N, G, F, K))) {
      // tcgen05 path: dX kernel (V_k = P^k (dY o act'), dX = sum_k V_k H_k) and dH / db kernel (accumulators in
      // tensor memory across all tiles of a CTA)
      const WideWs wws = wide_ws(B, N, G, F, K, 1);
      char* wb = wsb + p.ws_bytes;
      const int cs = wide_cshift_of(gs, N, norm), np = wide_planes(prec, G, F, K);
      g_last_path = 3;
      float* amax = reinterpret_cast<float*>(wb + wws.amax);
      float* wide_dpre = nullptr;
      uint32_t* wide_vmask = nullptr;
      const bool want_grads = dH != nullptr;   // db alone is served below by the tile kernels
      if (!want_grads && db && !dX) goto tile_path;
      if (want_grads) GFC_CUDA_TRY(cudaMemsetAsync(amax, 0, 2 * sizeof(float), st));
      if (dX) {
        unsigned char* hp = reinterpret_cast<unsigned char*>(wb + wws.pack);
        rc = launch_wide_pack(h, G, F, K, 1, cs, np, hp, st);
        if (rc) return rc;
        WideArgs wa{};
        fill_wide_graph(wa.g, gs, a, norm, N, 0);   // backward hop: V_{k+1} = S V_k  ->  P = S
        wa.in = dY; wa.yout = (act != GFC_ACT_NONE) ? yout : nullptr; wa.hpack = hp; wa.out = dX;
        wa.B = B; wa.N = N; wa.K = K; wa.cshift = cs; wa.act = act; wa.slope = slope; wa.dbg = g_dbg_clk;
        if (act != GFC_ACT_NONE && fmask) wa.fmask = fmask;   // the forward kernel's mask: y is not read
        if (want_grads) {
          wa.amax = amax;   // max |dY o act'(y)| of the batch: by-product of the tile scales
          // hand-over to the dH kernel: the activation mask as bits (default; 1/32 of the bytes) or dY o act'(y) itself
          if (act != GFC_ACT_NONE && fmask) {}   // the dH kernel reads the forward kernel's mask as well
          else if (act != GFC_ACT_NONE && g_wide_mask_handover) { wide_vmask = reinterpret_cast<uint32_t*>(wb + wws.dpre); wa.vmask = wide_vmask; }
          else if (act != GFC_ACT_NONE) { wide_dpre = reinterpret_cast<float*>(wb + wws.dpre); wa.d_out = wide_dpre; }
        }
        rc = launch_wide(wa, G, F, 1, np, st);
        if (rc) return rc;
        a.dX = nullptr;
        if (!dH && !db) return GFC_OK;
      }
      if (want_grads) {
        const int npart = wide_dh_nparts(B, N, F, K);
        float* dhp = reinterpret_cast<float*>(wb + wws.dhp);
        float* dbp = reinterpret_cast<float*>(wb + wws.dbp);
        // launch-wide operand scales: max |x| always, max |dY| (x max(1, slope)) when the dX kernel did not run
        const float vbound = (act == GFC_ACT_LEAKY_RELU && slope > 1.f) ? slope : 1.f;
        rc = launch_wide_absmax(x, (size_t)B * G * N, dX ? nullptr : dY, (size_t)B * N * F, vbound, amax, stats, st);
        if (rc) return rc;
        WideDhArgs da{};
        fill_wide_graph(da.g, gs, a, norm, N, 0);
        da.x = x; da.dY = dY; da.yout = (act != GFC_ACT_NONE) ? yout : nullptr; da.dpre = wide_dpre; da.vmask = wide_vmask; da.amax = amax;
        if (act != GFC_ACT_NONE && fmask) da.fmask = fmask;
        da.dHp = dhp; da.dbp = db ? dbp : nullptr;
        da.B = B; da.N = N; da.K = K; da.cshift = cs; da.act = act; da.slope = slope; da.dbg = g_dbg_clk;
        GFC_CUDA_TRY(cudaMemsetAsync(dhp, 0, (size_t)npart * nH * sizeof(float), st));   // partials are accumulated with red.add
        rc = launch_wide_dh(da, G, F, np, st);
        if (rc) return rc;
        if (g_skip_grad_reduce) return GFC_OK;
        if (dp) return launch_reduce_allreduce(dhp, npart, (int)nH, dbp, npart, F, dH, *dp, st);
        return launch_reduce_parts(dhp, npart, (int)nH, dH, db ? dbp : nullptr, npart, F, db, st);
      }
    }
  tile_path:
    if (!p.h_smem && a.dX) {
      float4* hp = reinterpret_cast<float4*>(wsb + p.ws_hpack);
      rc = launch_pack_taps(h, F, p.KG, 1, hp, st);
      if (rc) return rc;
      a.hpack = hp;
    }
    if (dH && !p.acc_regs)
      GFC_CUDA_TRY(cudaMemsetAsync(a.dHp, 0, (size_t)p.nparts * nH * sizeof(float), st));
    rc = launch_tile_bwd(a, gs.kind, st);
    if (rc) return rc;
    if (g_skip_grad_reduce) return GFC_OK;  // profiling aid, see gfc_set_option
    if (dp) return launch_reduce_allreduce(a.dHp, p.nparts, (int)nH, a.dbp, p.grid, F, dH, *dp, st);
    return launch_reduce_parts(dH ? a.dHp : nullptr, p.nparts, (int)nH, dH,
                               db ? a.dbp : nullptr, p.grid, F, db, st);
  }
  // path B
  GenericPlan g;
  plan_generic(B, N, G, F, K, E, 1, gs.kind == GSRC_POS, &g);
  rc = need_ws(fn, ws, ws_bytes, g.ws_bytes);
  if (rc) return rc;
  char* wsb = static_cast<char*>(ws);
  const float* S = gs.S;
  const int s_shared = (gs.kind == GSRC_DENSE && gs.shared) ? 1 : 0;
  if (gs.kind == GSRC_POS) {
    float* Sw = reinterpret_cast<float*>(wsb + g.ws_s);
    int lc = launch_counter();
    rc = gfc_gso_build(gs.pos, B, N, gs.radius, gs.mode, nullptr, Sw, st);
    launch_counter() += lc;
    if (rc) return rc;
    S = Sw;
  }
  float* Zw = reinterpret_cast<float*>(wsb + g.ws_z);
  float* Dw = reinterpret_cast<float*>(wsb + g.ws_d);
  const long long rows = (long long)B * N;
  const long long C = (long long)E * K * G;
  if (g.rows_ok) {
    // tensor-core pipeline: states -> one fused kernel (D, db, dH, U in place) -> Horner hops
    if (dH) {
      rc = launch_xpose_in(x, Zw, B, N, G, E, K, st);
      if (rc) return rc;
      for (int k = 1; k < K; ++k) {
        rc = launch_hop_dense(Zw, S, B, N, G, E, K, k - 1, k, 0, s_shared, st);
        if (rc) return rc;
      }
    }
    rc = rows_bwd(g, wsb, Zw, h, (act != GFC_ACT_NONE) ? yout : nullptr, dY, dX != nullptr, dH, db, act, slope, prec,
                  dp, st);
    if (rc) return rc;
    if (dX) {
      for (int k = K - 2; k >= 0; --k) {
        rc = launch_hop_dense(Zw, S, B, N, G, E, K, k + 1, k, 1, s_shared, st);
        if (rc) return rc;
      }
      rc = launch_xpose_out(Zw, dX, B, N, G, E, K, st);
      if (rc) return rc;
    }
    return GFC_OK;
  }
  rc = launch_dpre(dY, yout, Dw, rows * F, act, slope, st);
  if (rc) return rc;
  if (db) {
    float* part = reinterpret_cast<float*>(wsb + g.ws_dbp);
    rc = launch_colsum(Dw, rows, F, g.rows_per_chunk, g.nchunks, part, st);
    if (rc) return rc;
    rc = launch_reduce_parts(part, g.nchunks, F, db, nullptr, 0, 0, nullptr, st);
    if (rc) return rc;
  }
  if (dH) {
    rc = launch_xpose_in(x, Zw, B, N, G, E, K, st);
    if (rc) return rc;
    for (int k = 1; k < K; ++k) {
      rc = launch_hop_dense(Zw, S, B, N, G, E, K, k - 1, k, 0, s_shared, st);
      if (rc) return rc;
    }
    float* part = reinterpret_cast<float*>(wsb + g.ws_dhp);
    // dH[f][c] = sum_r D[r][f] Zw[r][c]
    rc = launch_sgemm(Dw, 1, F, Zw, C, 1, part, C, (long long)nH, F, (int)C, rows, g.nsplit, nullptr,
                      GFC_ACT_NONE, 0.f, st);
    if (rc) return rc;
    rc = launch_reduce_parts(part, g.nsplit, (int)nH, dH, nullptr, 0, 0, nullptr, st);
    if (rc) return rc;
  }
  if (dX) {
    // U[r][c] = sum_f D[r][f] h[f][c]
    rc = launch_sgemm(Dw, F, 1, h, C, 1, Zw, C, 0, rows, (int)C, F, 1, nullptr, GFC_ACT_NONE, 0.f, st);
    if (rc) return rc;
    for (int k = K - 2; k >= 0; --k) {
      rc = launch_hop_dense(Zw, S, B, N, G, E, K, k + 1, k, 1, s_shared, st);
      if (rc) return rc;
    }
    rc = launch_xpose_out(Zw, dX, B, N, G, E, K, st);
    if (rc) return rc;
  }
  if (dp) return launch_reduce_allreduce(dH, 1, (int)nH, db, 1, F, dH, *dp, st);
  return GFC_OK;
}

}  // namespace gfc

using namespace gfc;

extern "C" int gfc_version(void) { return GFC_VERSION; }
extern "C" const char* gfc_last_error(void) { return g_err; }
extern "C" int gfc_last_launch_count(void) { return g_launches; }
extern "C" int gfc_set_debug_clock_buffer(void* device_i64, size_t bytes) {
  g_dbg_clk = (bytes >= 16 * sizeof(long long) * 1184) ? static_cast<long long*>(device_i64) : nullptr;
  return (device_i64 && !g_dbg_clk) ? GFC_ERR_BAD_ARG : GFC_OK;
}
extern "C" int gfc_set_option(int key, int value) {
  if (key == GFC_OPT_SKIP_GRAD_REDUCE) { g_skip_grad_reduce = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_DISABLE_TCGEN05) { g_disable_tcgen05 = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_PDL) { g_pdl = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_WIDE_NO_PREFETCH) { g_wide_no_prefetch = value; return GFC_OK; }
  if (key == GFC_OPT_DP_TIMEOUT_MS) { g_dp_timeout_ms = value > 0 ? value : 10000; return GFC_OK; }
  if (key == GFC_OPT_CSR_STAGE_IDX) { g_csr_stage_idx = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_WIDE_FWD_MASK) { g_wide_fwd_mask = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_WIDE_MASK_HANDOVER) { g_wide_mask_handover = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_CSR_FUSED) { g_csr_fused = value ? 1 : 0; return GFC_OK; }
  if (key == GFC_OPT_WIDE_FLUSH_EVERY) { g_wide_flush_every = value > 0 ? value : 3; return GFC_OK; }
  set_error("gfc_set_option: unknown key %d", key);
  return GFC_ERR_BAD_ARG;
}

extern "C" int gfc_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes) {
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc) return rc;
  if (sm_count) *sm_count = di.sm_count;
  if (cc_major) *cc_major = di.cc_major;
  if (cc_minor) *cc_minor = di.cc_minor;
  if (smem_optin_bytes) *smem_optin_bytes = di.smem_optin;
  return GFC_OK;
}

extern "C" int gfc_filter_path(int B, int N, int G, int F, int K, int E, int backward) {
  if (B <= 0 || N <= 0 || G <= 0 || F <= 0 || K <= 0 || E <= 0) return 0;
  TilePlan p;
  if (E == 1 && plan_tile(B, N, G, F, K, backward, GSRC_DENSE, &p)) return 1;
  return 2;
}

extern "C" int gfc_tile_plan_info(int B, int N, int G, int F, int K, int backward, int from_positions, int* out) {
  if (!out) return GFC_ERR_BAD_ARG;
  TilePlan p;
  plan_tile(B, N, G, F, K, backward, from_positions ? GSRC_POS : GSRC_DENSE, &p);
  out[0] = p.ok; out[1] = p.gpc; out[2] = p.rows; out[3] = p.rpad; out[4] = p.ntiles; out[5] = p.grid;
  out[6] = (int)p.smem_bytes; out[7] = p.h_smem; out[8] = p.acc_regs; out[9] = p.nb_dh; out[10] = p.nparts;
  out[11] = p.ldz;
  return GFC_OK;
}

extern "C" size_t gfc_filter_workspace_bytes(int B, int N, int G, int F, int K, int E, int backward) {
  if (B <= 0 || N <= 0 || G <= 0 || F <= 0 || K <= 0 || E <= 0) return 0;
  size_t need = 0;
  TilePlan p;
  bool all_tile = (E == 1);
  for (int src = 0; src < 2 && E == 1; ++src) {
    if (plan_tile(B, N, G, F, K, backward, src, &p)) {
      const size_t nb = p.ws_bytes + wide_ws_extra(B, N, G, F, K, backward);
      if (nb > need) need = nb;
    }
    else all_tile = false;
  }
  if (!all_tile) {
    GenericPlan g;
    plan_generic(B, N, G, F, K, E, backward, E == 1, &g);
    if (g.ws_bytes > need) need = g.ws_bytes;
  }
  return need;
}

extern "C" int gfc_filter_fwd(const float* x, const float* S, const float* h, const float* bias, float* y,
                              int B, int N, int G, int F, int K, int E, int act, float slope, int precision,
                              void* workspace, size_t workspace_bytes, void* stream) {
  GsoSrc gs{GSRC_DENSE, S, nullptr, 0.0, 0};
  gs.binary = (precision & GFC_PREC_FLAG_BINARY_GSO) != 0 && E == 1;
  gs.shared = (precision & GFC_PREC_FLAG_SHARED_GSO) != 0;
  precision &= ~(GFC_PREC_FLAG_BINARY_GSO | GFC_PREC_FLAG_SHARED_GSO);
  return filter_fwd_impl("gfc_filter_fwd", gs, x, h, bias, y, B, N, G, F, K, E, act, slope, precision,
                         workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gfc_filter_fwd_pos(const float* x, const float* pos, double radius, int mode, const float* h,
                                  const float* bias, float* y, int B, int N, int G, int F, int K,
                                  int act, float slope, int precision,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  GsoSrc gs{GSRC_POS, nullptr, pos, radius, mode};
  return filter_fwd_impl("gfc_filter_fwd_pos", gs, x, h, bias, y, B, N, G, F, K, 1, act, slope, precision,
                         workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gfc_filter_fwd_pos_nm(const float* x_nm, const float* pos, double radius, int mode, const float* h,
                                     const float* bias, float* y, int B, int N, int G, int F, int K,
                                     int act, float slope, int precision,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "gfc_filter_fwd_pos_nm";
  cudaStream_t st = (cudaStream_t)stream;
  launch_counter() = 0;
  g_next_stats = nullptr;   // the node-major entry has no backward partner: a pending gfc_use_stats is dropped
  g_next_mask = nullptr; g_next_mask_bytes = 0; g_mask_filled = 0;
  int rc = check_common(fn, B, N, G, F, K, 1, act, precision, slope);
  if (rc) return rc;
  if (B == 0) return GFC_OK;
  GFC_REQUIRE(x_nm && pos && h && y, GFC_ERR_BAD_ARG, "%s: NULL pointer", fn);
  double thr = 0; bool norm = false;
  rc = gso_mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  GsoSrc gs{GSRC_POS, nullptr, pos, radius, mode};
  GFC_REQUIRE(use_wide(gs, N, G, F, K, 0, precision) && aligned16(x_nm) && aligned16(y) && aligned16(pos) && aligned16(h),
              GFC_ERR_UNSUPPORTED, "%s: node-major input needs the tcgen05 wide path (positions, G, F in {64,128}, "
              "N <= 128, 16-byte aligned tensors); transpose to [B,G,N] and call gfc_filter_fwd_pos", fn);
  TilePlan p;
  GFC_REQUIRE(plan_tile(B, N, G, F, K, 0, GSRC_POS, &p), GFC_ERR_UNSUPPORTED, "%s: shape not covered", fn);
  rc = need_ws(fn, workspace, workspace_bytes, p.ws_bytes + wide_ws_extra(B, N, G, F, K, 0));
  if (rc) return rc;
  TileArgs a{};
  set_thresholds(a, thr, norm);
  const WideWs wws = wide_ws(B, N, G, F, K, 0);
  unsigned char* hp = reinterpret_cast<unsigned char*>(workspace) + p.ws_bytes + wws.pack;
  const int cs = wide_cshift(N, norm), np = wide_planes(precision, G, F, K);
  rc = launch_wide_pack(h, G, F, K, 0, cs, np, hp, st);
  if (rc) return rc;
  WideArgs wa{};
  fill_wide_graph(wa.g, gs, a, norm, N, 1);
  g_last_path = 3;
  wa.in = x_nm; wa.hpack = hp; wa.bias = bias; wa.out = y; wa.B = B; wa.N = N; wa.K = K; wa.cshift = cs;
  wa.act = act; wa.slope = slope;
  return launch_wide(wa, G, F, 2, np, st);
}

extern "C" int gfc_filter_bwd(const float* x, const float* S, const float* h, const float* y_out,
                              const float* dY, float* dX, float* dH, float* db,
                              int B, int N, int G, int F, int K, int E, int act, float slope, int precision,
                              void* workspace, size_t workspace_bytes, void* stream) {
  GsoSrc gs{GSRC_DENSE, S, nullptr, 0.0, 0};
  gs.binary = (precision & GFC_PREC_FLAG_BINARY_GSO) != 0 && E == 1;
  gs.shared = (precision & GFC_PREC_FLAG_SHARED_GSO) != 0;
  precision &= ~(GFC_PREC_FLAG_BINARY_GSO | GFC_PREC_FLAG_SHARED_GSO);
  return filter_bwd_impl("gfc_filter_bwd", gs, x, h, y_out, dY, dX, dH, db, B, N, G, F, K, E, act, slope,
                         precision, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int gfc_filter_bwd_pos(const float* x, const float* pos, double radius, int mode, const float* h,
                                  const float* y_out, const float* dY, float* dX, float* dH, float* db,
                                  int B, int N, int G, int F, int K, int act, float slope, int precision,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  GsoSrc gs{GSRC_POS, nullptr, pos, radius, mode};
  return filter_bwd_impl("gfc_filter_bwd_pos", gs, x, h, y_out, dY, dX, dH, db, B, N, G, F, K, 1, act, slope,
                         precision, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ---- (e) data-parallel variants: gradients leave through the fused reduce + all-reduce kernel --------
extern "C" int gfc_filter_bwd_pos_dp(const float* x, const float* pos, double radius, int mode, const float* h,
                                     const float* y_out, const float* dY, float* dX, float* grads,
                                     int B, int N, int G, int F, int K, int act, float slope, int precision,
                                     void* workspace, size_t workspace_bytes,
                                     void* const* peer_buf, void* const* peer_sig, int rank, int world, float scale,
                                     void* stream) {
  GsoSrc gs{GSRC_POS, nullptr, pos, radius, mode};
  DpCtx dp{peer_buf, peer_sig, rank, world, scale};
  if (!grads) { set_error("gfc_filter_bwd_pos_dp: NULL gradient bucket"); return GFC_ERR_BAD_ARG; }
  return filter_bwd_impl("gfc_filter_bwd_pos_dp", gs, x, h, y_out, dY, dX, grads, grads + (size_t)F * K * G,
                         B, N, G, F, K, 1, act, slope, precision, workspace, workspace_bytes, (cudaStream_t)stream, &dp);
}

extern "C" int gfc_filter_bwd_dp(const float* x, const float* S, const float* h, const float* y_out,
                                 const float* dY, float* dX, float* grads,
                                 int B, int N, int G, int F, int K, int E, int act, float slope, int precision,
                                 void* workspace, size_t workspace_bytes,
                                 void* const* peer_buf, void* const* peer_sig, int rank, int world, float scale,
                                 void* stream) {
  GsoSrc gs{GSRC_DENSE, S, nullptr, 0.0, 0};
  gs.binary = (precision & GFC_PREC_FLAG_BINARY_GSO) != 0 && E == 1;
  gs.shared = (precision & GFC_PREC_FLAG_SHARED_GSO) != 0;
  precision &= ~(GFC_PREC_FLAG_BINARY_GSO | GFC_PREC_FLAG_SHARED_GSO);
  DpCtx dp{peer_buf, peer_sig, rank, world, scale};
  if (!grads) { set_error("gfc_filter_bwd_dp: NULL gradient bucket"); return GFC_ERR_BAD_ARG; }
  return filter_bwd_impl("gfc_filter_bwd_dp", gs, x, h, y_out, dY, dX, grads, grads + (size_t)F * E * K * G,
                         B, N, G, F, K, E, act, slope, precision, workspace, workspace_bytes, (cudaStream_t)stream, &dp);
}

/* all-reduce of an arbitrary flat fp32 bucket through the same kernel (in place allowed) */
extern "C" int gfc_dp_allreduce(const float* in, float* out, int n, void* const* peer_buf, void* const* peer_sig,
                                int rank, int world, float scale, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(in && out && n > 0, GFC_ERR_BAD_ARG, "gfc_dp_allreduce: bad arguments");
  DpCtx dp{peer_buf, peer_sig, rank, world, scale};
  return launch_reduce_allreduce(in, 1, n, nullptr, 0, 0, out, dp, (cudaStream_t)stream);
}

// ---- (d) CSR variant: workspace pipeline with SpMM hops -------------------------
extern "C" size_t gfc_filter_csr_workspace_bytes(int B, int N, int G, int F, int K, int backward) {
  if (B <= 0 || N <= 0 || G <= 0 || F <= 0 || K <= 0) return 0;
  GenericPlan g;
  plan_generic(B, N, G, F, K, 1, backward, 0, &g);
  return g.ws_bytes;
}

extern "C" int gfc_filter_csr_fwd(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                  int64_t nnz_stride, const float* h, const float* bias, float* y,
                                  int B, int N, int G, int F, int K, int act, float slope, int precision,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "gfc_filter_csr_fwd";
  launch_counter() = 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = check_common(fn, B, N, G, F, K, 1, act, precision, slope);
  if (rc) return rc;
  if (B == 0) return GFC_OK;
  GFC_REQUIRE(x && rowptr && colidx && h && y, GFC_ERR_BAD_ARG, "%s: NULL pointer", fn);
  GFC_REQUIRE((G & 3) == 0, GFC_ERR_UNSUPPORTED, "%s: G=%d must be a multiple of 4", fn, G);
  GenericPlan g;
  plan_generic(B, N, G, F, K, 1, 0, 0, &g);
  rc = need_ws(fn, workspace, workspace_bytes, g.ws_bytes);
  if (rc) return rc;
  if (g_csr_fused && csr_fwd_fused_supported(N, G, F, K, nullptr) && aligned16(y))
    return launch_csr_fwd_fused(x, rowptr, colidx, vals, nnz_stride, h, bias, y, B, N, G, F, K, act, slope,
                                precision != GFC_PREC_FP32_3XTF32, st);
  float* Zw = reinterpret_cast<float*>(static_cast<char*>(workspace) + g.ws_z);
  const long long C = (long long)K * G;
  rc = launch_hops_csr(Zw, rowptr, colidx, vals, nnz_stride, B, N, G, K, 0, x, nullptr, st);
  if (rc) return rc;
  if (g.rows_ok) return rows_fwd(g, static_cast<char*>(workspace), Zw, h, bias, y, act, slope, precision, st);
  return launch_sgemm(Zw, C, 1, h, 1, C, y, F, 0, (long long)B * N, F, C, 1, bias, act, slope, st);
}

extern "C" int gfc_filter_csr_bwd(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                  const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                                  int64_t nnz_stride, const float* h, const float* y_out, const float* dY,
                                  float* dX, float* dH, float* db, int B, int N, int G, int F, int K,
                                  int act, float slope, int precision,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  const char* fn = "gfc_filter_csr_bwd";
  launch_counter() = 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = check_common(fn, B, N, G, F, K, 1, act, precision, slope);
  if (rc) return rc;
  const size_t nH = (size_t)F * K * G;
  if (B == 0) {
    if (dH) GFC_CUDA_TRY(cudaMemsetAsync(dH, 0, nH * sizeof(float), st));
    if (db) GFC_CUDA_TRY(cudaMemsetAsync(db, 0, (size_t)F * sizeof(float), st));
    return GFC_OK;
  }
  GFC_REQUIRE(rowptr && colidx && rowptr_t && colidx_t && h && dY, GFC_ERR_BAD_ARG, "%s: NULL pointer", fn);
  GFC_REQUIRE(!dH || x, GFC_ERR_BAD_ARG, "%s: x is required for dH", fn);
  GFC_REQUIRE(act == GFC_ACT_NONE || y_out, GFC_ERR_BAD_ARG, "%s: y_out is required with a fused activation", fn);
  GFC_REQUIRE((G & 3) == 0, GFC_ERR_UNSUPPORTED, "%s: G=%d must be a multiple of 4", fn, G);
  if (!dX && !dH && !db) return GFC_OK;
  GenericPlan g;
  plan_generic(B, N, G, F, K, 1, 1, 0, &g);
  rc = need_ws(fn, workspace, workspace_bytes, g.ws_bytes);
  if (rc) return rc;
  char* wsb = static_cast<char*>(workspace);
  float* Zw = reinterpret_cast<float*>(wsb + g.ws_z);
  float* Dw = reinterpret_cast<float*>(wsb + g.ws_d);
  const long long rows = (long long)B * N;
  const long long C = (long long)K * G;
  if (g_csr_fused && g.fused_bwd && aligned16(dY) && (act == GFC_ACT_NONE || aligned16(y_out))) {
    // one CTA per graph: V_0 = dY o act'(y), transposed-list hops in shared memory, dX / dH / db from the same state
    float* dhp = reinterpret_cast<float*>(wsb + g.ws_fdh);
    float* dbp = reinterpret_cast<float*>(wsb + g.ws_fdb);
    rc = launch_csr_bwd_fused(x, rowptr_t, colidx_t, vals_t, nnz_stride, h, (act != GFC_ACT_NONE) ? y_out : nullptr, dY,
                              dX, dH ? dhp : nullptr, db ? dbp : nullptr, B, N, G, F, K, act, slope,
                              precision != GFC_PREC_FP32_3XTF32, st);
    if (rc) return rc;
    if (!dH && !db) return GFC_OK;
    return launch_reduce_parts(dH ? dhp : nullptr, B, (int)nH, dH, db ? dbp : nullptr, B, F, db, st);
  }
  if (g.rows_ok) {
    if (dH) {
      rc = launch_hops_csr(Zw, rowptr, colidx, vals, nnz_stride, B, N, G, K, 0, x, nullptr, st);
  if (rc) return rc;
    }
    rc = rows_bwd(g, wsb, Zw, h, (act != GFC_ACT_NONE) ? y_out : nullptr, dY, dX != nullptr, dH, db, act, slope,
                  precision, nullptr, st);
    if (rc) return rc;
    if (dX) {
      rc = launch_hops_csr(Zw, rowptr_t, colidx_t, vals_t, nnz_stride, B, N, G, K, 1, nullptr, dX, st);
      if (rc) return rc;
    }
    return GFC_OK;
  }
  rc = launch_dpre(dY, y_out, Dw, rows * F, act, slope, st);
  if (rc) return rc;
  if (db) {
    float* part = reinterpret_cast<float*>(wsb + g.ws_dbp);
    rc = launch_colsum(Dw, rows, F, g.rows_per_chunk, g.nchunks, part, st);
    if (rc) return rc;
    rc = launch_reduce_parts(part, g.nchunks, F, db, nullptr, 0, 0, nullptr, st);
    if (rc) return rc;
  }
  if (dH) {
    rc = launch_hops_csr(Zw, rowptr, colidx, vals, nnz_stride, B, N, G, K, 0, x, nullptr, st);
  if (rc) return rc;
    float* part = reinterpret_cast<float*>(wsb + g.ws_dhp);
    rc = launch_sgemm(Dw, 1, F, Zw, C, 1, part, C, (long long)nH, F, (int)C, rows, g.nsplit, nullptr,
                      GFC_ACT_NONE, 0.f, st);
    if (rc) return rc;
    rc = launch_reduce_parts(part, g.nsplit, (int)nH, dH, nullptr, 0, 0, nullptr, st);
    if (rc) return rc;
  }
  if (dX) {
    rc = launch_sgemm(Dw, F, 1, h, C, 1, Zw, C, 0, rows, (int)C, F, 1, nullptr, GFC_ACT_NONE, 0.f, st);
    if (rc) return rc;
    rc = launch_hops_csr(Zw, rowptr_t, colidx_t, vals_t, nnz_stride, B, N, G, K, 1, nullptr, dX, st);
    if (rc) return rc;
  }
  return GFC_OK;
}

extern "C" int gfc_use_stats(float* stats) {
  g_next_stats = stats;
  return GFC_OK;
}

extern "C" int gfc_use_mask(void* mask, size_t bytes) {
  GFC_REQUIRE(!mask || (reinterpret_cast<uintptr_t>(mask) & 127) == 0, GFC_ERR_BAD_ARG, "gfc_use_mask: the buffer must be 128-byte aligned");
  g_next_mask = static_cast<uint32_t*>(mask);
  g_next_mask_bytes = mask ? bytes : 0;
  return GFC_OK;
}
extern "C" int gfc_mask_filled(void) { return g_mask_filled; }
extern "C" size_t gfc_filter_mask_bytes(int B, int N, int G, int F, int K) {
  (void)K;
  if (B < 1 || !(G == 64 || G == 128) || !(F == 64 || F == 128)) return 0;   // shapes of the tcgen05 wide kernels
  return wide_mask_bytes(B, N);
}

extern "C" int gfc_last_path(void) { return g_last_path; }
