// gfc_tc5_wide.cuh — host interface of the warp-specialised tcgen05 kernels (gfc_tc5_wide.cu).
#pragma once
#include "gfc_common.cuh"

namespace gfc {

struct WideArgs {
  const float* pos;        // [B,N,2]
  double thr;              // squared-distance threshold, fp64 rule (gfc_gso.cu)
  float thr_lo, thr_hi;    // fp32 screening band
  const float* in;         // MODE 0: x [B,G,N];  MODE 1: dY [B,N,F];  MODE 2: x [B,N,G]
  const float* yout;       // MODE 1: forward output [B,N,F] (activation mask) or null
  const uint16_t* hpack;   // taps as bf16x3 planes in ring-stage order (wide_pack_taps_kernel)
  const float* bias;       // MODE 0: [F] or null
  float* out;              // MODE 0: y [B,N,F];  MODE 1: dX [B,G,N]
  float* d_out;            // MODE 1, optional: dY o act'(y) [B,N,F] written for the dH kernel (else null)
  int B, N, K;
  int gpc, ntiles;         // filled by launch_wide
  int act;
  float slope;
  int no_prefetch;         // experiment: skip the L2 bulk prefetch
  int tma_out;             // filled by launch_wide: the y tile leaves through the TMA store engine
  long long* dbg;          // optional clock stamps of CTA 0 (issuer at [0..), worker warp 2 at [2048..)), else null
};

struct WideDhArgs {
  const float* pos;
  double thr;
  float thr_lo, thr_hi;
  const float* x;          // [B,G,N]
  const float* dY;         // [B,N,F]
  const float* yout;       // [B,N,F] forward output (activation mask) or null
  const float* dpre;       // optional [B,N,F]: dY o act'(y) already formed by the dX kernel (then dY / yout are not read)
  float* dHp;              // [nparts][F*K*G] per-CTA-group partial gradients (every element written)
  float* dbp;              // [nparts][F] or null
  int B, N, K;
  int gpc, ntiles, nparts; // filled by launch_wide_dh
  int flush_every;         // tiles chained into the TMEM accumulators between drains (0 = default)
  int no_prefetch;         // experiment: skip the L2 bulk prefetch
  long long* dbg;          // optional clock stamps of CTA 0
  int act;
  float slope;
};
extern int g_wide_flush_every;
extern int g_wide_no_prefetch;
bool wide_dh_supported(int N, int G, int F, int K);
int wide_dh_nparts(int B, int N, int F, int K);   // number of partial buffers launch_wide_dh writes
int launch_wide_dh(const WideDhArgs& a, int G, int F, cudaStream_t st);

// mode 0 = forward, 1 = backward dX, 2 = forward with node-major input x [B,N,G] (launch_wide only; pack / support as mode 0)
bool wide_supported(int N, int G, int F, int K, int mode);
size_t wide_pack_bytes(int G, int F, int K);
int launch_wide_pack(const float* h, int G, int F, int K, int mode, uint16_t* out, cudaStream_t st);
int launch_wide(const WideArgs& a, int G, int F, int mode, cudaStream_t st);

}  // namespace gfc
