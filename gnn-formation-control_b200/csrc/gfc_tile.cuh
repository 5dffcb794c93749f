// gfc_tile.cuh — plan/arguments of the fused shared-memory tile kernels (path A).
#pragma once
#include "gfc_common.cuh"

namespace gfc {

enum { GSRC_DENSE = 0, GSRC_POS = 1 };

// Kernel variants: compile-time shapes for the BASELINE configs, runtime shapes otherwise.
enum {
  VAR_GENERIC = 0,     // runtime N,G,F,K; 256 threads, 4 n-tiles per warp task
  VAR_N8_32_32_3 = 1,  // cfg2: N=8, G=F=32, K=3; 512 threads, 2 n-tiles per warp task
  VAR_128_128_3 = 2,   // the reference policy's layer (suhaas_model.py:33-42): G=F=128, K=3, any N
  VAR_N64_128_128_4 = 3,  // cfg3
  VAR_ROWS = 4         // N = 1, K = 1: plain row contraction [rows x G] . [G x F] (the workspace pipeline's tap GEMMs); 384 threads
};

// Host-computed plan: how a [B graphs] x [N nodes] batch is cut into tiles of
// `gpc` graphs whose K diffusion states live in shared memory.
struct TilePlan {
  int ok;          // 0 -> shape not covered by path A
  int variant;
  int threads;     // CTA size
  int nb;          // n-tiles (8 columns each) per warp task
  int B, N, G, F, K, KG;
  int backward;
  int gpc;         // graphs per tile
  int rows;        // gpc * N
  int rpad;        // rows rounded up to 16 (MMA m)
  int ldz;         // row stride of the diffusion states Z[r][k*G+g] (floats)
  int ldd;         // row stride of dpre tile D[r][f] (floats, backward)
  int ntiles;      // ceil(B / gpc)
  int h_smem;      // packed taps resident in shared memory
  int acc_regs;    // backward: dH accumulates in registers across tiles
  int nb_dh;       // n-tiles per dH warp task
  int grid;        // persistent CTAs
  int nparts;      // dH partial buffers
  // shared-memory carve-up, float offsets
  int off_z, off_s, off_d, off_h, off_pos, off_isd, off_nbr;
  int use_lists;   // compact neighbour lists for the hops (16 < N <= 255)
  size_t smem_bytes;
  // workspace carve-up, byte offsets
  size_t ws_hpack, ws_dhp, ws_dbp, ws_bytes;
};

int plan_tile(int B, int N, int G, int F, int K, int backward, int gsrc, TilePlan* p);

struct TileArgs {
  // graph source
  const float* S;      // [B,N,N] dense (E = 1)
  long long s_bstride; // floats between consecutive graphs of S: N*N, or 0 = one GSO shared by the whole batch
  const float* pos;    // [B,N,2]
  double thr;          // squared-distance threshold (GSRC_POS), fp64 rule
  float thr_lo, thr_hi;  // fp32 band: s < thr_lo surely inside, s > thr_hi surely outside
  int norm;            // sym-norm weights (GSRC_POS)
  // tensors
  const float* x;      // [B,G,N]
  const float* h;      // [F,KG]
  const float4* hpack; // packed+split B fragments in global (when !h_smem)
  const float* bias;   // [F] or null
  float* y;            // [B,N,F]
  const float* yout;   // [B,N,F] forward output (activation mask), bwd
  const float* dY;     // [B,N,F]
  float* dX;           // [B,G,N] or null
  float* dHp;          // partial dH buffers [nparts][F*KG] or null
  float* dbp;          // partial db [grid][F] or null
  int act;
  float slope;
  int single_pass;     // GFC_PREC_TF32
  int vec_ok;          // all global bases 16-byte aligned
  long long* dbg_clk;  // optional [grid][16] phase time stamps (gfc_set_debug_clock_buffer), else null
  TilePlan p;
};

int launch_tile_fwd(const TileArgs& a, int gsrc, cudaStream_t st);
int launch_tile_bwd(const TileArgs& a, int gsrc, cudaStream_t st);
int launch_pack_taps(const float* h, int F, int KG, int for_bwd, float4* out, cudaStream_t st);
// out_a[i] = sum_p parts_a[p][i] (i < n_a) and, in the same launch, out_b likewise (n_b may be 0)
int launch_reduce_parts(const float* parts_a, int nparts_a, int n_a, float* out_a,
                        const float* parts_b, int nparts_b, int n_b, float* out_b, cudaStream_t st);

// per-variant launchers (one translation unit each, see gfc_tile_inst_*.cu)
typedef int (*tile_launch_fn)(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_fwd_generic(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_bwd_generic(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_fwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_bwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_fwd_128_128_3(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_bwd_128_128_3(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_fwd_n64_128_128_4(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_bwd_n64_128_128_4(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_fwd_rows(const TileArgs& a, int gsrc, cudaStream_t st);
int tile_bwd_rows(const TileArgs& a, int gsrc, cudaStream_t st);
// tcgen05 / TMEM kernels of the cfg2 shape (gfc_tc5_n8.cu)
int tc5_fwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st);
int tc5_bwd_n8_32_32_3(const TileArgs& a, int gsrc, cudaStream_t st);
int tc5_bwd_grid_n8_32_32_3(int B);
extern int g_pdl;              // gfc_set_option(GFC_OPT_PDL), gfc_tc5_n8.cu
extern int g_disable_tcgen05;  // gfc_set_option(GFC_OPT_DISABLE_TCGEN05)

}  // namespace gfc
