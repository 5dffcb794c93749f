// gfc_generic.cuh — launchers of the workspace pipeline (path B).
#pragma once
#include "gfc_common.cuh"

namespace gfc {

int launch_xpose_in(const float* x, float* Zw, int B, int N, int G, int E, int K, cudaStream_t st);
int launch_xpose_out(const float* Uw, float* dX, int B, int N, int G, int E, int K, cudaStream_t st);
int launch_hop_dense(float* W, const float* S, int B, int N, int G, int E, int K, int ksrc, int kdst,
                     int transposed, int s_shared, cudaStream_t st);
int launch_hop_csr(float* W, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                   long long nnz_stride, int B, int N, int G, int K, int ksrc, int kdst, int accum,
                   cudaStream_t st);
// all K-1 hops of the CSR filter (horner = 0: slot k = hop(slot k-1); 1: slot k += hop(slot k+1), transposed lists).
// xin != null (forward): the chain starts from x [B,G,N] (slot 0 is filled from it); dxout != null (Horner): the
// result leaves as dX [B,G,N] — the transposing passes are part of the chain.
int launch_hops_csr(float* W, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                    long long nnz_stride, int B, int N, int G, int K, int horner, const float* xin, float* dxout,
                    cudaStream_t st);
// whole CSR forward of a graph in one CTA (gfc_csr_fused.cu); GFC_ERR_UNSUPPORTED when the state does not fit
bool csr_fwd_fused_supported(int N, int G, int F, int K, size_t* smem_bytes);
int launch_csr_fwd_fused(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         long long nnz_stride, const float* h, const float* bias, float* y, int B, int N, int G,
                         int F, int K, int act, float slope, int single, cudaStream_t st);
// whole CSR backward of a graph in one CTA; partial dH [B][F*K*G] and db [B][F] for reduce_parts_kernel
bool csr_bwd_fused_supported(int N, int G, int F, int K, size_t* smem_bytes);
int launch_csr_bwd_fused(const float* x, const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                         long long nnz_stride, const float* h, const float* yout, const float* dY, float* dX,
                         float* dHp, float* dbp, int B, int N, int G, int F, int K, int act, float slope, int single,
                         cudaStream_t st);
int launch_dpre(const float* dY, const float* yout, float* D, long long n, int act, float slope, cudaStream_t st);
int launch_colsum(const float* D, long long rows, int F, int rows_per_chunk, int nchunks, float* part,
                  cudaStream_t st);
// C[m][n] = act(sum_k A(m,k) B(k,n) + bias[n]);  A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs].
// nsplit > 1 writes nsplit partial matrices at C + z*c_zstride (bias/act must be off).
int launch_sgemm(const float* A, long long a_rs, long long a_cs, const float* Bm, long long b_rs, long long b_cs,
                 float* C, long long ldc, long long c_zstride, long long M, int Nn, long long Kd,
                 int nsplit, const float* bias, int act, float slope, cudaStream_t st);

extern int g_csr_stage_idx;   // gfc_set_option(GFC_OPT_CSR_STAGE_IDX): fused CSR kernels keep the lists in shared memory as 16-bit numbers

}  // namespace gfc
