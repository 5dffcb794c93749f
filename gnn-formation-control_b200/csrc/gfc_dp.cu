// gfc_dp.cu — data-parallel gradient exchange: ONE kernel that finishes the gradient of a rank (fixed-order sum of
// the per-CTA partials the backward kernels left behind) and all-reduces it over NVLink peer memory.
//
//   phase 1  each CTA owns a slice of the flat [dH | db] bucket: it sums the slice over this rank's partial buffers
//            and stores every result, packed with the epoch number into ONE 64-bit word {value, epoch}, straight
//            into slot `rank` of EVERY peer's exchange buffer (P2P stores through NVSwitch);
//   phase 2  the CTA polls its own slice of the `world` slots until every word carries the current epoch and sums
//            the values in rank order (so every rank obtains bit-identical gradients), scales, writes the bucket.
// The flag travels inside the data word (a naturally aligned 64-bit store is single-copy atomic), so there is no
// release fence / acknowledgement round trip on the sender and no separate flag hop: the exchange costs one one-way
// NVLink latency.  2 x (world-1) x n words cross the links per rank; there is no grid-wide synchronisation (slices
// are independent) and no host involvement.  The exchange buffers are double buffered by epoch parity: a rank can
// only start epoch e+2 after every peer has finished reading epoch e, because it needs the peers' epoch e+1 words to
// complete epoch e+1 first.
// The reference has no distributed code (single .to('cuda'), suhaas_agent.py:19); this replaces the
// reduce_parts_kernel + ncclAllReduce pair of the plain DP path.
#include "gfc_common.cuh"
#include "gfc_dp.cuh"
#include "gfc_tile.cuh"   // g_pdl

namespace gfc {

__device__ __forceinline__ void st_word(unsigned long long* p, float v, unsigned int epoch) {
  const unsigned long long w = ((unsigned long long)epoch << 32) | __float_as_uint(v);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct DpPeers {
  unsigned long long* buf[GFC_DP_MAX_WORLD];   // exchange buffer of every rank: {value, epoch} words [2][world][n]
  unsigned int* sig[GFC_DP_MAX_WORLD];         // per-rank state: epochs [nblocks] + one sticky timeout flag (local use only)
};

constexpr int kDpClasses = 32;   // partial-classes per output: 32 lanes x 32 classes = 1024 threads per CTA

__global__ void __launch_bounds__(32 * kDpClasses)
reduce_allreduce_kernel(const float* pa, int npa, int na, const float* pb, int npb, int nb,
                        float* out, const DpPeers peers, int rank, int world, float scale,
                        unsigned long long timeout_ns) {
  // pa / pb / out carry no __restrict__: gfc_dp_allreduce is documented "in place allowed" (out == pa)
  __shared__ unsigned int s_epoch;
  __shared__ float red[kDpClasses][33];
  // programmatic dependent launch: this grid may be scheduled while the backward kernel drains
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int n = na + nb;
  const int nblocks = gridDim.x, c = blockIdx.x;
  const int per = (((n + nblocks - 1) / nblocks) + 31) & ~31;    // slice of this CTA: whole 32-float groups
  const int lo = c * per, hi = min(n, lo + per);
  unsigned int* epochs = peers.sig[rank];
  if (threadIdx.x == 0) { s_epoch = epochs[c] + 1; epochs[c] = s_epoch; }
  __syncthreads();
  const unsigned int e = s_epoch;
  const int par = e & 1;
  // ---- phase 1: local fixed-order reduction of the slice, pushed to every rank ---------------------
  // 32 outputs at a time: 32 partial-classes x 32 lanes; every thread issues all of its loads (<= 8 per pass)
  // before the first add, so a slice of up to 256 partial buffers costs ONE L2 round trip; then a fixed-order
  // combine through shared memory — the same order on every launch and every rank: deterministic gradients
  const int lane = threadIdx.x & 31, pg = threadIdx.x >> 5;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    float s = 0.f;
    if (i < hi) {
      const float* parts; int np, nn, ii;
      if (i < na) { parts = pa; np = npa; nn = na; ii = i; } else { parts = pb; np = npb; nn = nb; ii = i - na; }
      for (int p0 = pg; p0 < np; p0 += 8 * kDpClasses) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = p0 + u * kDpClasses;
          v[u] = p < np ? parts[(size_t)p * nn + ii] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
    }
    red[pg][lane] = s;
    __syncthreads();
    if (pg < world && i < hi) {      // warp r pushes the finished 32 values to rank r (P2P stores through NVSwitch)
      float t = red[0][lane];
#pragma unroll
      for (int k = 1; k < kDpClasses; ++k) t += red[k][lane];
      st_word(peers.buf[pg] + ((size_t)par * world + rank) * n + i, t, e);
    }
    __syncthreads();
  }
  // ---- phase 2: poll every rank's words of the slice, sum in rank order ---------------------------------
  const unsigned long long* mine = peers.buf[rank] + (size_t)par * world * n;
  for (int base = lo; base < hi; base += 32) {
    const int i = base + lane;
    if (pg < world && i < hi) {      // warp r waits for rank r's 32 words
      const unsigned long long* src = mine + (size_t)pg * n + i;
      unsigned long long w = ld_word(src);
      if ((unsigned int)(w >> 32) != e) {
        // bounded wait: a peer that never arrives (crashed rank, ranks that disagreed on the transport) must become
        // an error, not a hung GPU.  On expiry the element is NaN and the sticky flag behind the epochs is raised
        // (read back by gfc_dp_status); the kernel always terminates.
        const unsigned long long t0 = globaltimer_ns();
        while (true) {
          __nanosleep(20);
          w = ld_word(src);
          if ((unsigned int)(w >> 32) == e) break;
          if (globaltimer_ns() - t0 > timeout_ns) {
            atomicExch(epochs + nblocks, 1u + (unsigned int)pg);   // 1 + the rank that was missing
            w = 0x7fc00000ull;                                     // quiet NaN
            break;
          }
        }
      }
      red[pg][lane] = __uint_as_float((unsigned int)w);
    }
    __syncthreads();
    if (pg == 0 && i < hi) {
      float t = red[0][lane];
      for (int r = 1; r < world; ++r) t += red[r][lane];
      out[i] = t * scale;
    }
    __syncthreads();
  }
}

int g_dp_timeout_ms = 10000;   // gfc_set_option(GFC_OPT_DP_TIMEOUT_MS)

int dp_blocks(int n) {
  int nb = ceil_div(n, 32);
  if (nb > GFC_DP_MAX_BLOCKS) nb = GFC_DP_MAX_BLOCKS;
  if (nb < 1) nb = 1;
  return nb;
}

int launch_reduce_allreduce(const float* pa, int npa, int na, const float* pb, int npb, int nb, float* out,
                            const DpCtx& dp, cudaStream_t st) {
  GFC_REQUIRE(dp.world >= 1 && dp.world <= GFC_DP_MAX_WORLD && dp.rank >= 0 && dp.rank < dp.world, GFC_ERR_BAD_ARG,
              "dp: bad rank/world %d/%d", dp.rank, dp.world);
  GFC_REQUIRE(dp.peer_buf && dp.peer_sig && out, GFC_ERR_BAD_ARG, "dp: NULL exchange pointers");
  DpPeers peers;
  for (int r = 0; r < dp.world; ++r) {
    peers.buf[r] = static_cast<unsigned long long*>(dp.peer_buf[r]);
    peers.sig[r] = static_cast<unsigned int*>(dp.peer_sig[r]);
    GFC_REQUIRE(peers.buf[r] && peers.sig[r], GFC_ERR_BAD_ARG, "dp: NULL peer pointer for rank %d", r);
  }
  const int n = na + nb;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(dp_blocks(n));
  cfg.blockDim = dim3(32 * kDpClasses);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  GFC_CUDA_TRY(cudaLaunchKernelEx(&cfg, reduce_allreduce_kernel, pa, npa, na, pb, npb, nb, out, peers, dp.rank,
                                  dp.world, dp.scale, (unsigned long long)g_dp_timeout_ms * 1000000ull));
  GFC_LAUNCH_CHECK("reduce_allreduce_kernel");
  return GFC_OK;
}

}  // namespace gfc

using namespace gfc;

extern "C" size_t gfc_dp_exchange_bytes(int n, int world) {
  if (n <= 0 || world <= 0 || world > GFC_DP_MAX_WORLD) return 0;
  return align_up((size_t)2 * world * n * sizeof(unsigned long long), 256);
}
extern "C" size_t gfc_dp_signal_bytes(int n, int world) {
  if (n <= 0 || world <= 0 || world > GFC_DP_MAX_WORLD) return 0;
  (void)world;
  return align_up((size_t)(dp_blocks(n) + 1) * sizeof(unsigned int), 256);   // epochs + the sticky timeout flag
}

// Synchronising status query of the exchange: copies this rank's sticky timeout flag back (stream-ordered, then waits
// for the stream).  GFC_OK, or GFC_ERR_TIMEOUT when some launch gave up on a peer (then *missing_rank names it and
// the affected bucket elements are NaN).  `n` is the bucket size the signal buffer was sized for.
extern "C" int gfc_dp_status(const void* my_sig, int n, int* missing_rank, void* stream) {
  GFC_REQUIRE(my_sig && n > 0, GFC_ERR_BAD_ARG, "gfc_dp_status: bad arguments");
  unsigned int flag = 0;
  cudaStream_t st = (cudaStream_t)stream;
  GFC_CUDA_TRY(cudaMemcpyAsync(&flag, static_cast<const unsigned int*>(my_sig) + dp_blocks(n), sizeof(flag),
                               cudaMemcpyDeviceToHost, st));
  GFC_CUDA_TRY(cudaStreamSynchronize(st));
  if (missing_rank) *missing_rank = flag ? (int)flag - 1 : -1;
  if (flag) { set_error("gfc_dp_status: the peer exchange timed out waiting for rank %d", (int)flag - 1); return GFC_ERR_TIMEOUT; }
  return GFC_OK;
}
