// gfc_tile.cuh — plan/arguments of the fused shared-memory tile kernels (path A).
#pragma once
#include "gfc_common.cuh"

namespace gfc {

constexpr int kTileThreads = 256;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kNB = 4;  // n-tiles (of 8 columns) a warp task covers at most

enum { GSRC_DENSE = 0, GSRC_POS = 1 };

// Host-computed plan: how a [B graphs] x [N nodes] batch is cut into tiles of
// `gpc` graphs whose K diffusion states live in shared memory.
struct TilePlan {
  int ok;          // 0 -> shape not covered by path A
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
  int off_z, off_s, off_d, off_h, off_pos, off_isd;
  size_t smem_bytes;
  // workspace carve-up, byte offsets
  size_t ws_hpack, ws_dhp, ws_dbp, ws_bytes;
};

int plan_tile(int B, int N, int G, int F, int K, int backward, int gsrc, TilePlan* p);

struct TileArgs {
  // graph source
  const float* S;      // [B,N,N] dense (E = 1)
  const float* pos;    // [B,N,2]
  double thr;          // squared-distance threshold (GSRC_POS)
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
  TilePlan p;
};

int launch_tile_fwd(const TileArgs& a, int gsrc, cudaStream_t st);
int launch_tile_bwd(const TileArgs& a, int gsrc, cudaStream_t st);
int launch_pack_taps(const float* h, int F, int KG, int for_bwd, float4* out, cudaStream_t st);
int launch_reduce_parts(const float* parts, int nparts, int n, float* out, cudaStream_t st);

}  // namespace gfc
