// gfc_tc5_wide.cuh — host interface of the warp-specialised tcgen05 kernels (gfc_tc5_wide.cu).
#pragma once
#include "gfc_common.cuh"

namespace gfc {

// Graph source of a wide kernel: positions (the GSO is rebuilt on chip, bit-exact radius rule) or a dense
// per-graph GSO S [B,N,N] read from HBM (the reference's own addGSO call, graphML.py:2449-2456).
struct WideGraph {
  const float* pos;        // [B,N,2] or null
  double thr;              // squared-distance threshold, fp64 rule (gfc_gso.cu)
  float thr_lo, thr_hi;    // fp32 screening band
  int norm;                // positions: D^-1/2 A D^-1/2 (multirobotsim_dcenlocal.py:306-315) applied as row scalings
  const float* S;          // dense [B,N,N] fp32 or null (then `pos`)
  long long s_bstride;     // floats between consecutive graphs of S (0: one GSO shared by the whole batch)
  int s_transpose;         // dense: the hop uses S^T (forward: z_{k+1} = z_k S) or S (backward)
  const float* s_bound;    // dense: device float >= max row/column abs sum of any S_b (growth bound per hop)
};

struct WideArgs {
  WideGraph g;
  const float* in;         // MODE 0: x [B,G,N];  MODE 1: dY [B,N,F];  MODE 2: x [B,N,G]
  const float* yout;       // MODE 1: forward output [B,N,F] (activation mask) or null
  const unsigned char* hpack;  // packed taps: 256-byte header {1/t_h, ...} + fp16 planes in ring-stage order
  const float* bias;       // MODE 0: [F] or null
  float* out;              // MODE 0: y [B,N,F];  MODE 1: dX [B,G,N]
  float* d_out;            // MODE 1, optional: dY o act'(y) [B,N,F] written for the dH kernel (else null)
  uint32_t* vmask;         // MODE 1, optional: activation mask bits (y > 0) for the dH kernel, one bit per element of
                           //   [B*N rows][F]: word row * (F/32) + f/32, bit f%32 — 1/32 of the bytes of d_out (else null)
  uint32_t* fmask_out;     // MODE 0 / 2, optional: activation mask bits (y > 0) of the output for the backward kernels, 2 KB per
                           //   tile (gfc_use_mask; layout: wide_fmask_word in gfc_tc5_wide.cu) — the backward then never reads y
  const uint32_t* fmask;   // MODE 1, optional: that mask; yout is not read, d_out / vmask are not needed
  float* amax;             // optional device float[2]: running max |x| (MODE 0/2 -> [0]) / max |dY o act'| (MODE 1 -> [1])
  int mark_stats;          // MODE 0: amax is a gfc_use_stats buffer (float[4]); the kernel also sets amax[3] = 1 ("filled") —
                           //   read by later kernels of the stream only, so no separate marking launch is needed
  int B, N, K;
  int cshift;              // per-hop headroom bits: W_k is carried as W_k 2^(-cshift k), the taps as H_k 2^(+cshift k)
  int gpc, ntiles;         // filled by launch_wide
  int act;
  float slope;
  int no_prefetch;         // experiment: skip the L2 bulk prefetch
  long long* dbg;          // optional clock stamps of CTA 0 (issuer at [0..), worker warp 2 at [4096..)), else null
};

struct WideDhArgs {
  WideGraph g;
  const float* x;          // [B,G,N]
  const float* dY;         // [B,N,F]
  const float* yout;       // [B,N,F] forward output (activation mask) or null
  const float* dpre;       // optional [B,N,F]: dY o act'(y) already formed by the dX kernel (then dY / yout are not read)
  const uint32_t* vmask;   // optional: the mask bits written by the dX kernel (WideArgs::vmask); dY is read, yout is not
  const uint32_t* fmask;   // optional: the mask bits written by the FORWARD kernel (WideArgs::fmask_out); dY is read, yout is not
  const float* amax;       // device float[2]: {max |x|, max |dY o act'|} over the WHOLE batch (launch-wide operand scales)
  float* dHp;              // [nparts][F*K*G] per-CTA-group partial gradients (zeroed by the caller, accumulated with red.add)
  float* dbp;              // [nparts][F] or null
  int B, N, K;
  int cshift;
  int gpc, ntiles, nparts; // filled by launch_wide_dh
  int flush_every;         // tiles chained into the TMEM accumulators between drains (0 = default)
  int no_prefetch;
  int act;
  float slope;
  long long* dbg;          // dH: optional clock stamps of CTA 0 (issuer at [0..), worker warp 2 at [4096..)), else null
};
extern int g_wide_flush_every;
extern int g_wide_no_prefetch;

// per-hop headroom (bits) of the fp16 operand scaling for an N-node 0/1 GSO (0 for the row-normalised form), and
// whether (N, K) keeps the guaranteed fp32-class accuracy (cshift * (K-1) <= 21)
int wide_cshift(int N, int norm);
bool wide_dh_supported(int N, int G, int F, int K);
int wide_dh_nparts(int B, int N, int F, int K);   // number of partial buffers launch_wide_dh writes
int launch_wide_dh(const WideDhArgs& a, int G, int F, int planes, cudaStream_t st);

// mode 0 = forward, 1 = backward dX, 2 = forward with node-major input x [B,N,G] (launch_wide only; pack / support as mode 0)
bool wide_supported(int N, int G, int F, int K, int mode);
size_t wide_pack_bytes(int G, int F, int K);
// taps -> fp16 planes (hi [, lo]) in ring-stage order, scaled by t_h 2^(cshift k); header[0] = 1 / t_h
int launch_wide_pack(const float* h, int G, int F, int K, int mode, int cshift, int planes, unsigned char* out,
                     cudaStream_t st);
int launch_wide(const WideArgs& a, int G, int F, int mode, int planes, cudaStream_t st);
// amax[0] = max |a| (n_a floats), amax[1] = max |b| * bscale (n_b floats); either pointer may be null (slot left as is)
// stats (optional, device float[4] of the forward call, gfc_use_stats): when stats[3] != 0 the pass over `a` is skipped
int launch_wide_absmax(const float* a, size_t n_a, const float* b, size_t n_b, float bscale, float* amax,
                       const float* stats, cudaStream_t st);
int launch_stats_mark(float* stats, cudaStream_t st);   // stats[3] = 1
size_t wide_mask_bytes(int B, int N);   // bytes of the forward / dX activation-mask hand-over: 2 KB per 128-row tile

}  // namespace gfc
