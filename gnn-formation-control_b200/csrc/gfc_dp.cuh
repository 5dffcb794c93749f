// gfc_dp.cuh — fused gradient reduction + one-shot all-reduce over peer memory (gfc_dp.cu).
#pragma once
#include "gfc_common.cuh"

#define GFC_DP_MAX_WORLD 16
#define GFC_DP_MAX_BLOCKS 128

namespace gfc {

struct DpCtx {
  void* const* peer_buf;   // [world] device pointers: every rank's exchange buffer (gfc_dp_exchange_bytes, zeroed once)
  void* const* peer_sig;   // [world] device pointers: every rank's signal buffer  (gfc_dp_signal_bytes, zeroed once)
  int rank, world;
  float scale;             // applied to the summed bucket (1/world for a mean)
};

extern int g_dp_timeout_ms;   // bounded wait of the exchange poll (default 10 s)
int dp_blocks(int n);
// out[0..na) = scale * sum_ranks sum_p pa[p][i];  out[na..na+nb) likewise from pb
int launch_reduce_allreduce(const float* pa, int npa, int na, const float* pb, int npb, int nb, float* out,
                            const DpCtx& dp, cudaStream_t st);

}  // namespace gfc
