/*
 * gfc.h — C ABI of libgfc.so: the B200 (sm_100a) graph-filter hot path of
 * soosiey/gnn-formation-control.
 *
 * The reference has no native seam around this path: the seam is the Python
 * nn.Module surface (utils/graphUtils/graphML.py:2369-2488).  Its only FFI
 * precedent is ctypes -> remoteApi.so with int32 status returns (sim.py:21,
 * simConst.py simx_return_ok = 0); this ABI mirrors that convention:
 *   - plain pointers and sizes, no torch types;
 *   - every entry point returns int (0 = GFC_OK); gfc_last_error() gives text;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library
 *     keeps no pointer after return, never synchronises, never touches the
 *     default stream: all work is enqueued on the cudaStream_t passed as
 *     `stream` (void*; pass torch.cuda.current_stream().cuda_stream).
 *
 * Layouts (all device pointers, fp32 unless stated, contiguous row-major):
 *   pos   [B, N, 2]      robot xy positions                (scene.py:147-148)
 *   adj   [B, N, N] u8   1-hop mask, adj[b,i,j] in {0,1}   (scene.py:140-154)
 *   S     [B, E, N, N]   graph shift operator              (graphML.py:2449-2456)
 *   x     [B, G, N]      node signals, feature-major       (graphML.py:2458)
 *   h     [F, E, K, G]   filter taps  (nn.Parameter weight, graphML.py:2434)
 *   bias  [F]            (nn.Parameter bias [F,1],          graphML.py:2436)
 *   y     [B, N, F]      NODE-major memory.  The reference returns a [B,F,N]
 *                        view with strides (N*F, 1, F) over exactly this
 *                        memory (graphML.py:2361-2362: matmul -> permute).
 *   dY    [B, N, F]      upstream gradient in the same memory layout
 *   dX    [B, G, N]
 *   dH    [F, E, K, G],  db [F]
 * Row-vector convention of the reference: z_k = z_{k-1} . S  (graphML.py:2350),
 * y[b,f,n] = sum_{e,k,g} h[f,e,k,g] z[b,e,k,g,n] + bias[f]   (graphML.py:2361-2366).
 */
#ifndef GFC_H_
#define GFC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFC_VERSION 100 /* major*100 + minor */

/* status codes (0 = ok, like simx_return_ok) */
enum {
  GFC_OK = 0,
  GFC_ERR_BAD_ARG = 1,     /* null pointer / non-positive size / bad enum      */
  GFC_ERR_UNSUPPORTED = 2, /* shape outside what the kernels cover             */
  GFC_ERR_WORKSPACE = 3,   /* workspace missing or too small                   */
  GFC_ERR_CUDA = 4,        /* a CUDA call failed; text in gfc_last_error()     */
  GFC_ERR_TIMEOUT = 5      /* gfc_dp_status: the peer exchange gave up on a rank */
};

/* position -> GSO rule */
enum {
  GFC_GSO_BINARY_LE = 0,   /* scene.py:147-152: a = (i!=j) && sqrt(dx^2+dy^2) <= R   */
  GFC_GSO_SYM_NORM_LT = 1, /* multirobotsim_dcenlocal.py:306-315: W=(d<R), D^-1/2 W D^-1/2 */
  GFC_GSO_BINARY_LT = 2    /* the unnormalised W of mode 1                            */
};

/* activation fused behind the filter (suhaas_model.py:120, decentralplanner.py:221) */
enum { GFC_ACT_NONE = 0, GFC_ACT_RELU = 1, GFC_ACT_LEAKY_RELU = 2 };

/* arithmetic of the tap contraction (the diffusion hops are always fp32 FMA)  */
enum {
  GFC_PREC_FP32_3XTF32 = 0, /* tensor cores, hi/lo split, fp32-equivalent (<=1e-5) */
  GFC_PREC_TF32 = 1,        /* single pass tf32, looser bound (~1e-3), opt-in       */
  GFC_PREC_F16 = 2,         /* single fp16 plane on the tcgen05 wide path (1 MMA per product, stated bound 2e-3);
                               shapes outside that path run as GFC_PREC_TF32         */
  /* flag, OR-ed into `precision` of the DENSE entry points (gfc_filter_fwd / _bwd / _bwd_dp): the caller vouches that
     every entry of S is exactly 0 or 1 (the adjacency Scene.readADjMatrix produces, scene.py:140-154, possibly
     asymmetric, possibly with a diagonal).  The hop matrix is then exact in fp16 and wide shapes (G, F in {64,128},
     N <= 127) run on the tcgen05 kernels like the position-built GSOs; without the flag a dense S runs on the
     mma.sync tile kernels (any weights).  gnnfc.GraphFilterBatch.addGSO checks the entries itself.                  */
  GFC_PREC_FLAG_BINARY_GSO = 0x100,
  /* flag, same entry points: S is [E,N,N] — ONE GSO shared by all B graphs of the batch (the reference's same-GSO layer
     GraphFilter / LSIGF, graphML.py:1111, :48) — instead of [B,E,N,N]; no per-graph copies are made or read            */
  GFC_PREC_FLAG_SHARED_GSO = 0x200
};

int gfc_version(void);
/* thread-local text of the last non-zero status returned on this thread */
const char* gfc_last_error(void);
/* SM count, compute capability and opt-in shared memory of the current device */
int gfc_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes);

/* ---- (a) position -> GSO -------------------------------------------------- *
 * Replaces Scene.readADjMatrix (scene.py:140-154) and
 * multiRobotSim.computeAdjacencyMatrix_fixedCommRadius
 * (utils/multirobotsim_dcenlocal.py:291-317).  The comparison is carried out in
 * fp64 on the device so the mask is bit-identical to the reference's python /
 * numpy float64 arithmetic.  adj_out and/or S_out may be NULL.                */
int gfc_gso_build(const float* pos, int B, int N, double radius, int mode,
                  uint8_t* adj_out, float* S_out, void* stream);

/* ---- (b) filter forward --------------------------------------------------- *
 * Replaces GraphFilterBatch.forward -> BatchLSIGF (graphML.py:2458-2477,
 * 2273-2367) plus the activation that follows it in the policy.  Workspace:
 * gfc_filter_workspace_bytes(); may be NULL when that returns 0.              */
size_t gfc_filter_workspace_bytes(int B, int N, int G, int F, int K, int E, int backward);

int gfc_filter_fwd(const float* x, const float* S, const float* h, const float* bias,
                   float* y, int B, int N, int G, int F, int K, int E,
                   int act, float slope, int precision,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Same filter with the GSO rebuilt on chip from positions (E = 1): S never
 * exists in HBM.  Replaces robot.py:654 + suhaas_model.py:149-159,182-185.    */
int gfc_filter_fwd_pos(const float* x, const float* pos, double radius, int mode,
                       const float* h, const float* bias, float* y,
                       int B, int N, int G, int F, int K,
                       int act, float slope, int precision,
                       void* workspace, size_t workspace_bytes, void* stream);
/* gfc_filter_fwd_pos with the input given NODE-major, x_nm [B, N, G] — the memory layout of a previous layer's
 * output y — so stacked layers (suhaas_model.py:112-121 with more than one entry in nGraphFilterTaps) chain
 * without the transposing copy of the reference's [B,F,N] view.  Available where the tcgen05 wide path applies
 * (binary GSO rule, G and F in {64,128}, N <= 128, 16-byte aligned tensors); GFC_ERR_UNSUPPORTED otherwise —
 * transpose and call gfc_filter_fwd_pos.  Same workspace as gfc_filter_workspace_bytes(..., backward = 0). */
int gfc_filter_fwd_pos_nm(const float* x_nm, const float* pos, double radius, int mode,
                          const float* h, const float* bias, float* y,
                          int B, int N, int G, int F, int K,
                          int act, float slope, int precision,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- (c) filter backward (replaces autograd through graphML.py:2342-2366) -- *
 * y_out is the forward OUTPUT (after the activation); it is only read when
 * act != GFC_ACT_NONE (to recover the activation mask) and may be NULL
 * otherwise.  dX / dH / db may each be NULL to skip that gradient.  dH and db
 * are overwritten (not accumulated) and are reduced deterministically.        */
int gfc_filter_bwd(const float* x, const float* S, const float* h, const float* y_out,
                   const float* dY, float* dX, float* dH, float* db,
                   int B, int N, int G, int F, int K, int E,
                   int act, float slope, int precision,
                   void* workspace, size_t workspace_bytes, void* stream);

int gfc_filter_bwd_pos(const float* x, const float* pos, double radius, int mode,
                       const float* h, const float* y_out, const float* dY,
                       float* dX, float* dH, float* db,
                       int B, int N, int G, int F, int K,
                       int act, float slope, int precision,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Operand statistics handed from the forward call of a batch to its backward call.
 * `stats` = device float[4], caller-owned.  gfc_use_stats(stats) applies to the NEXT gfc_filter_fwd* / gfc_filter_bwd*
 * call of the calling thread only (thread-local, consumed by that call; NULL cancels):
 *   - a forward call zeroes stats and, on the tcgen05 path, leaves stats[0] = max |x| (a by-product of its per-tile
 *     operand scales) and stats[3] = 1 ("filled");
 *   - a backward call of the same batch reads stats[3] ON THE DEVICE: if filled, its launch-wide operand scale of x
 *     comes from stats[0] and the extra pass over x (2.1 GB at 65536 x 128 x 64) is skipped; if not (the forward took
 *     another kernel family) it computes the maximum itself.  Results are identical either way.                       */
int gfc_use_stats(float* stats);

/* Activation mask handed from the forward call of a batch to its backward call (tcgen05 wide kernels, fused
 * LeakyReLU / ReLU).  `mask` = caller-owned device buffer, 128-byte aligned, of at least
 * gfc_filter_mask_bytes(B, N, G, F, K) bytes (2 KB per 128-row tile; 0 = the shape has no such kernel).
 * gfc_use_mask(mask, bytes) applies to the NEXT gfc_filter_fwd* / gfc_filter_bwd* call of the calling thread only
 * (thread-local, consumed by that call; NULL cancels):
 *   - a forward call on the tcgen05 wide path writes one bit per output element (y > 0) into it; gfc_mask_filled()
 *     then returns 1 for the calling thread (0 if the call took another kernel family or had no fused activation);
 *   - a backward call of the SAME batch given a filled mask reads it instead of y_out (1/32 of the bytes: the dX and
 *     dH kernels then stream dY only).  Pass a mask to the backward ONLY if gfc_mask_filled() was 1 right after the
 *     forward call that received it.  Results are bit-identical to the y_out path (same predicate y > 0).          */
int gfc_use_mask(void* mask, size_t bytes);
int gfc_mask_filled(void);
size_t gfc_filter_mask_bytes(int B, int N, int G, int F, int K);

/* ---- (d) CSR variant for large sparse swarms ------------------------------ *
 * gfc_csr_count : deg[b,n] = in-degree under the radius rule (int32 [B,N]).
 * gfc_csr_fill  : given rowptr [B, N+1] (exclusive scan of deg per graph, int32,
 *                 offsets relative to the graph's own segment base b*nnz_stride)
 *                 writes colidx (ascending per row) and, for SYM_NORM, vals.
 * The position-built GSOs are symmetric, so one CSR serves both S and S^T.   */
int gfc_csr_count(const float* pos, int B, int N, double radius, int mode,
                  int32_t* deg, void* stream);
/* rowptr[b, 0..N] = exclusive scan of deg[b, :] (rowptr[b, N] = nnz of graph b) */
int gfc_csr_scan(const int32_t* deg, int B, int N, int32_t* rowptr, void* stream);
int gfc_csr_fill(const float* pos, int B, int N, double radius, int mode,
                 const int32_t* rowptr, int64_t nnz_stride,
                 int32_t* colidx, float* vals, void* stream);

/* gfc_csr_build : the three steps above in ONE launch (one CTA per graph, cell list over a uniform grid of edge
 *                 >= R instead of the O(N^2) pair walk; same bit-exact rule, same ascending columns).  nnz_stride is a
 *                 per-graph CAPACITY chosen by the caller — no host round trip to size colidx, so the call can be
 *                 captured in a CUDA graph: a graph that needs more sets *overflow = max(*overflow, needed nnz)
 *                 (int32 device word, zeroed by the caller; may be NULL) and drops the surplus edges.  colidx NULL:
 *                 sizing pass, only rowptr is written.  GFC_ERR_UNSUPPORTED when one graph's positions + lists do not
 *                 fit shared memory (N above ~8000): use count/scan/fill.                                             */
int gfc_csr_build(const float* pos, int B, int N, double radius, int mode,
                  int32_t* rowptr, int64_t nnz_stride, int32_t* colidx, float* vals,
                  int32_t* overflow, void* stream);

/* CSR of S^T ("gather lists": row n holds the m with S[m,n] != 0, value S[m,n]).
 * vals may be NULL (all ones).  csr_t_* is the CSR of S itself, needed by the
 * backward; pass the same arrays when S is symmetric.                         */
size_t gfc_filter_csr_workspace_bytes(int B, int N, int G, int F, int K, int backward);

int gfc_filter_csr_fwd(const float* x, const int32_t* rowptr, const int32_t* colidx,
                       const float* vals, int64_t nnz_stride,
                       const float* h, const float* bias, float* y,
                       int B, int N, int G, int F, int K,
                       int act, float slope, int precision,
                       void* workspace, size_t workspace_bytes, void* stream);

int gfc_filter_csr_bwd(const float* x, const int32_t* rowptr, const int32_t* colidx,
                       const float* vals,
                       const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                       int64_t nnz_stride,
                       const float* h, const float* y_out, const float* dY,
                       float* dX, float* dH, float* db,
                       int B, int N, int G, int F, int K,
                       int act, float slope, int precision,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- (e) data-parallel training: fused gradient reduction + all-reduce ------ *
 * The reference is single-GPU (suhaas_agent.py:19); data parallelism shards the
 * batch of graphs over ranks and sums the flat gradient bucket [dH | db].  These
 * entry points replace "second-stage reduction kernel + ncclAllReduce" by ONE
 * kernel that reduces this rank's per-CTA partials and exchanges the result over
 * NVLink peer memory (one-shot: every rank stores its slice, as 64-bit words
 * {value, epoch}, into every peer's exchange buffer, polls its own buffer until all
 * peers' words carry the epoch, sums in rank order -> every rank holds bit-identical
 * gradients).
 *   peer_buf[r] / peer_sig[r]: device pointers, valid on THIS device, to rank r's
 *   exchange / signal buffer (peer-mapped memory, e.g. torch symmetric memory or
 *   cudaIpc handles), each of gfc_dp_exchange_bytes / gfc_dp_signal_bytes for the
 *   bucket size n = F*E*K*G + F, zero-initialised once before the first call.
 *   grads: the bucket [dH (F*E*K*G) | db (F)], written on every rank.
 *   Every rank of the group must make the same sequence of calls.              */
size_t gfc_dp_exchange_bytes(int n, int world);
size_t gfc_dp_signal_bytes(int n, int world);
int gfc_filter_bwd_pos_dp(const float* x, const float* pos, double radius, int mode,
                          const float* h, const float* y_out, const float* dY,
                          float* dX, float* grads,
                          int B, int N, int G, int F, int K,
                          int act, float slope, int precision,
                          void* workspace, size_t workspace_bytes,
                          void* const* peer_buf, void* const* peer_sig,
                          int rank, int world, float scale, void* stream);
int gfc_filter_bwd_dp(const float* x, const float* S, const float* h, const float* y_out,
                      const float* dY, float* dX, float* grads,
                      int B, int N, int G, int F, int K, int E,
                      int act, float slope, int precision,
                      void* workspace, size_t workspace_bytes,
                      void* const* peer_buf, void* const* peer_sig,
                      int rank, int world, float scale, void* stream);
/* all-reduce (sum * scale) of any flat fp32 bucket through the same kernel; in == out allowed */
int gfc_dp_allreduce(const float* in, float* out, int n,
                     void* const* peer_buf, void* const* peer_sig,
                     int rank, int world, float scale, void* stream);

/* The exchange never hangs the device: a rank that waits longer than GFC_OPT_DP_TIMEOUT_MS (default 10 000 ms)
 * for a peer's words writes NaN into the affected bucket elements, raises a sticky flag in its signal buffer and
 * returns.  gfc_dp_status copies that flag back (it SYNCHRONISES `stream`): GFC_OK, or GFC_ERR_TIMEOUT with
 * *missing_rank = the rank that did not arrive.  my_sig = this rank's own signal buffer, n = the bucket size.   */
int gfc_dp_status(const void* my_sig, int n, int* missing_rank, void* stream);

/* ---- introspection used by bench.py / tests -------------------------------- *
 * Which kernel family a shape dispatches to: 1 = fused shared-memory tile
 * kernel, 2 = workspace pipeline (dense hops), 0 = unsupported.               */
int gfc_filter_path(int B, int N, int G, int F, int K, int E, int backward);
/* Tile plan of path A for a shape: out[12] = {ok, graphs_per_tile, rows, rows_padded,
 * ntiles, grid, smem_bytes, taps_in_smem, dH_in_registers, nb_dh, nparts, ldz}.  */
int gfc_tile_plan_info(int B, int N, int G, int F, int K, int backward, int from_positions, int* out);
/* Process-wide options.  GFC_OPT_SKIP_GRAD_REDUCE = 1 makes gfc_filter_bwd* leave the
 * per-CTA dH / db partials unreduced (dH / db are then NOT written): a profiling aid that
 * lets bench.py time the dominant backward kernel alone, back to back.  Default 0.      */
enum {
  GFC_OPT_SKIP_GRAD_REDUCE = 1,
  GFC_OPT_DISABLE_TCGEN05 = 2, /* 1: use the mma.sync tile kernels where a tcgen05 kernel exists (A/B comparison) */
  GFC_OPT_WIDE_FLUSH_EVERY = 3, /* tiles chained into the TMEM dH accumulators between drains (default 3) */
  GFC_OPT_PDL = 4, /* 1 (default): the cfg2-shape kernels and the gradient reduction use programmatic dependent launch */
  GFC_OPT_CSR_FUSED = 5, /* 1 (default): one-CTA-per-graph fused CSR forward / backward kernels; 0: workspace pipeline (A/B) */
  GFC_OPT_WIDE_NO_PREFETCH = 6, /* 1: tcgen05 wide kernels skip the L2 bulk prefetch of the next tiles (experiment; default 0) */
  GFC_OPT_DP_TIMEOUT_MS = 7, /* bound of the peer-exchange poll in milliseconds (default 10000), see gfc_dp_status */
  GFC_OPT_CSR_STAGE_IDX = 10, /* 1 (default): the fused CSR kernels stage a graph's neighbour lists in shared memory as 16-bit numbers; 0: read from global (A/B) */
  GFC_OPT_WIDE_FWD_MASK = 9, /* 1 (default): a forward call given a buffer with gfc_use_mask fills it; 0: it never does (A/B) */
  GFC_OPT_WIDE_MASK_HANDOVER = 8 /* 1 (default): the tcgen05 dX kernel hands the activation mask to the dH kernel as bits (1/32 of
                                    the bytes); 0: it writes dY o act'(y) [B,N,F] to the workspace as in earlier builds (A/B) */
};
int gfc_set_option(int key, int value);
/* kernel family of the calling thread's last gfc_filter_fwd* / gfc_filter_bwd* call: 1 = fused tile kernels (mma.sync, or the
 * tcgen05 kernels of the 8-node shape), 2 = workspace pipeline, 3 = warp-specialised tcgen05 / TMEM wide kernels */
int gfc_last_path(void);
/* Debug aid: a device buffer of >= 1184*16 int64 in which the fused tile kernels stamp the SM
 * clock at their phase boundaries (first tile of every CTA).  NULL switches it off (default). */
int gfc_set_debug_clock_buffer(void* device_i64, size_t bytes);
/* number of kernel launches the last call on this thread enqueued */
int gfc_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GFC_H_ */
