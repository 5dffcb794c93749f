// gfc_tile.cu — host side of path A: tile plan, variant dispatch, tap packing and the
// deterministic second-stage gradient reduction.  Device code: gfc_tile_kernels.cuh.
#include "gfc_tile_kernels.cuh"

namespace gfc {

int g_disable_tcgen05 = 0;

__global__ void __launch_bounds__(256)
pack_taps_kernel(const float* __restrict__ h, int F, int KG, int for_bwd, float4* __restrict__ out) {
  const int total = (KG * F) >> 1;  // float4 elements
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x)
    out[p] = pack_one(h, F, KG, for_bwd, p);
}

// out[i] = sum_p parts[p][i] in a fixed order (deterministic).  A CTA reduces 32 outputs: 32 partial-classes x 32
// lanes — every thread issues all of its (<= 8) loads before the first add, so the whole reduction is one L2
// round trip for up to 256 partial buffers — then a fixed-order combine through shared memory.  Two independent
// segments (dH and db) share one launch.
constexpr int kReduceClasses = 32;
__global__ void __launch_bounds__(32 * kReduceClasses)
reduce_parts_kernel(const float* __restrict__ pa, int npa, int na, float* __restrict__ oa, int blocks_a,
                    const float* __restrict__ pb, int npb, int nb, float* __restrict__ ob) {
  __shared__ float red[kReduceClasses][33];
  // programmatic dependent launch: wait for the kernel that wrote the partials; let the next kernel's CTAs queue
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const float* parts; int nparts, n; float* out; int blk;
  if ((int)blockIdx.x < blocks_a) { parts = pa; nparts = npa; n = na; out = oa; blk = blockIdx.x; }
  else { parts = pb; nparts = npb; n = nb; out = ob; blk = blockIdx.x - blocks_a; }
  const int lane = threadIdx.x & 31, pg = threadIdx.x >> 5;
  const int i = blk * 32 + lane;
  float s = 0.f;
  if (i < n) {
    for (int p0 = pg; p0 < nparts; p0 += 8 * kReduceClasses) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int p = p0 + u * kReduceClasses;
        v[u] = p < nparts ? parts[(size_t)p * n + i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
  }
  red[pg][lane] = s;
  __syncthreads();
  if (pg == 0 && i < n) {
    float r = red[0][lane];
#pragma unroll
    for (int k = 1; k < kReduceClasses; ++k) r += red[k][lane];
    out[i] = r;
  }
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
static int tile_smem_floats(TilePlan* p, int gsrc, int with_h) {
  int off = 0;
  p->off_z = off; off += p->rpad * p->ldz;
  p->off_s = off; off += (p->gpc * p->N * p->N + 3) & ~3;
  p->off_d = off; if (p->backward) off += p->rpad * p->ldd + ((p->F + 3) & ~3);
  p->off_pos = off; if (gsrc == GSRC_POS) off += (p->gpc * p->N * 2 + 3) & ~3;
  p->off_isd = off; if (gsrc == GSRC_POS) off += (p->gpc * p->N * 2 + 3) & ~3;  // doubles
  p->off_nbr = off;
  p->use_lists = (p->N > 16 && p->N <= 255) ? 1 : 0;
  if (p->use_lists) off += (((p->rows * p->N + p->rows + 3) >> 2) + 3 & ~3) * (p->backward ? 2 : 1);
  p->off_h = off; if (with_h) off += p->KG * p->F * 2;
  return off;
}

int plan_tile(int B, int N, int G, int F, int K, int backward, int gsrc, TilePlan* p) {
  *p = TilePlan{};
  p->B = B; p->N = N; p->G = G; p->F = F; p->K = K; p->KG = K * G; p->backward = backward;
  if (B <= 0 || N <= 0 || G <= 0 || F <= 0 || K <= 0) return 0;
  if ((G & 7) || (F & 7)) return 0;
  if (backward && (F & 15)) return 0;
  if ((long long)K * G > 8192 || N > 4096) return 0;
  // kernel variant
  p->variant = VAR_GENERIC; p->threads = 256; p->nb = 4;
  if (N == 8 && G == 32 && F == 32 && K == 3) { p->variant = VAR_N8_32_32_3; p->threads = 512; p->nb = 2; }
  else if (N == 64 && G == 128 && F == 128 && K == 4) p->variant = VAR_N64_128_128_4;
  else if (G == 128 && F == 128 && K == 3) p->variant = VAR_128_128_3;
  if (N == 1 && K == 1) { p->variant = VAR_ROWS; p->threads = 384; p->nb = 4; }
  if (F > p->threads) return 0;
  DeviceInfo di;
  if (get_device_info(&di)) return 0;
  const long long limit_floats = (di.smem_optin - 1024) / 4;
  p->ldz = p->KG + (backward ? 8 : 4);
  p->ldd = F + 4;
  int gpc = 128 / N;
  // row contraction: one 16-row MMA task per warp and tile (a CTA of this register footprint owns its SM, so the
  // tile must keep all of its warps busy between the two barriers)
  if (p->variant == VAR_ROWS) gpc = 16 * (p->threads / 32);
  if (gpc < 1) gpc = 1;
  if (gpc > B) gpc = B;
  {  // small batches: spread the graphs over the SMs instead of filling 128-row tiles
    int spread = ceil_div(B, di.sm_count);
    int gmin = 16 / N;  // keep at least one full MMA m-tile of rows
    if (gmin < 1) gmin = 1;
    if (spread < gmin) spread = gmin;
    if (spread < gpc) gpc = spread;
  }
  // prefer a footprint that lets two CTAs share an SM, as long as a tile keeps >= 64 rows
  const long long half_floats = limit_floats / 2 - 256;
  int chosen = 0, chosen_h = 0;
  // the row contraction re-reads the packed taps for every 16-row task: insist on taps in shared memory first
  for (int need_h = (p->variant == VAR_ROWS ? 1 : 0); need_h >= 0 && !chosen; --need_h) {
    for (int pass = (p->variant == VAR_ROWS ? 1 : 0); pass < 2 && !chosen; ++pass) {
      const long long lim = pass == 0 ? half_floats : limit_floats;
      for (int gcur = gpc; gcur >= 1; gcur = (gcur > 1 ? (gcur + 1) / 2 : 0)) {
        p->gpc = gcur; p->rows = gcur * N; p->rpad = (p->rows + 15) & ~15;
        if (pass == 0 && p->rows < 64 && gcur != gpc) break;
        if (tile_smem_floats(p, gsrc, 1) <= lim) { chosen = gcur; chosen_h = 1; break; }
        if (!need_h && tile_smem_floats(p, gsrc, 0) <= lim) { chosen = gcur; chosen_h = 0; break; }
        if (gcur == 1) break;
      }
    }
  }
  if (!chosen) return 0;
  p->gpc = chosen; p->rows = chosen * N; p->rpad = (p->rows + 15) & ~15;
  p->h_smem = chosen_h;
  p->smem_bytes = (size_t)tile_smem_floats(p, gsrc, chosen_h) * sizeof(float);
  p->ntiles = ceil_div(B, p->gpc);
  // dH accumulation mode: registers when the [F x KG] output fits one task per warp
  p->nb_dh = p->nb; p->acc_regs = 0;
  if (backward) {
    const int MTd = F >> 4, NTd = p->KG >> 3;
    for (int nb = 1; nb <= p->nb; ++nb) {
      if (MTd * ceil_div(NTd, nb) <= p->threads / 32) { p->nb_dh = nb; p->acc_regs = 1; break; }
    }
  }
  int per_sm = (int)((size_t)di.smem_optin / (p->smem_bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  const int cap = p->threads >= 512 ? 2 : 4;
  if (per_sm > cap) per_sm = cap;
  p->grid = di.sm_count * per_sm;
  if (p->grid > p->ntiles) p->grid = p->ntiles;
  if (backward && p->variant == VAR_N8_32_32_3 && !g_disable_tcgen05) {
    // the tcgen05 backward of this shape is persistent with one CTA per SM and keeps dH in registers
    const int g5 = tc5_bwd_grid_n8_32_32_3(B);
    if (g5 > 0) { p->grid = g5; p->acc_regs = 1; }
  }
  p->nparts = p->acc_regs ? p->grid : (p->grid < 8 ? p->grid : 8);
  // workspace
  size_t off = 0;
  p->ws_hpack = off; if (!p->h_smem) off += align_up((size_t)p->KG * F * 2 * sizeof(float), 256);
  p->ws_dhp = off; if (backward) off += align_up((size_t)p->nparts * F * p->KG * sizeof(float), 256);
  p->ws_dbp = off; if (backward) off += align_up((size_t)p->grid * F * sizeof(float), 256);
  p->ws_bytes = off;
  p->ok = 1;
  return 1;
}

int launch_pack_taps(const float* h, int F, int KG, int for_bwd, float4* out, cudaStream_t st) {
  const int total = (KG * F) >> 1;
  int grid = ceil_div(total, 256);
  if (grid > 1024) grid = 1024;
  pack_taps_kernel<<<grid, 256, 0, st>>>(h, F, KG, for_bwd, out);
  GFC_LAUNCH_CHECK("pack_taps_kernel");
  return GFC_OK;
}

int launch_reduce_parts(const float* parts_a, int nparts_a, int n_a, float* out_a,
                        const float* parts_b, int nparts_b, int n_b, float* out_b, cudaStream_t st) {
  const int blocks_a = parts_a ? ceil_div(n_a, 32) : 0;
  const int blocks_b = parts_b ? ceil_div(n_b, 32) : 0;
  if (blocks_a + blocks_b == 0) return GFC_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks_a + blocks_b);
  cfg.blockDim = dim3(32 * kReduceClasses);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  GFC_CUDA_TRY(cudaLaunchKernelEx(&cfg, reduce_parts_kernel, parts_a, nparts_a, n_a, out_a, blocks_a,
                                  parts_b, nparts_b, n_b, out_b));
  GFC_LAUNCH_CHECK("reduce_parts_kernel");
  return GFC_OK;
}

int launch_tile_fwd(const TileArgs& a, int gsrc, cudaStream_t st) {
  switch (a.p.variant) {
    case VAR_N8_32_32_3:
      if (!g_disable_tcgen05 && a.vec_ok && !a.single_pass) return tc5_fwd_n8_32_32_3(a, gsrc, st);
      return tile_fwd_n8_32_32_3(a, gsrc, st);
    case VAR_128_128_3: return tile_fwd_128_128_3(a, gsrc, st);
    case VAR_N64_128_128_4: return tile_fwd_n64_128_128_4(a, gsrc, st);
    case VAR_ROWS: return tile_fwd_rows(a, gsrc, st);
    default: return tile_fwd_generic(a, gsrc, st);
  }
}

int launch_tile_bwd(const TileArgs& a, int gsrc, cudaStream_t st) {
  switch (a.p.variant) {
    case VAR_N8_32_32_3:
      if (!g_disable_tcgen05 && a.vec_ok && a.p.acc_regs) return tc5_bwd_n8_32_32_3(a, gsrc, st);
      return tile_bwd_n8_32_32_3(a, gsrc, st);
    case VAR_128_128_3: return tile_bwd_128_128_3(a, gsrc, st);
    case VAR_N64_128_128_4: return tile_bwd_n64_128_128_4(a, gsrc, st);
    case VAR_ROWS: return tile_bwd_rows(a, gsrc, st);
    default: return tile_bwd_generic(a, gsrc, st);
  }
}

}  // namespace gfc
