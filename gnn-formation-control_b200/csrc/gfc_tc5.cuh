// gfc_tc5.cuh — Blackwell tensor-core plumbing used by the tcgen05 tile kernels:
// TMEM allocation, UMMA shared-memory / instruction descriptors, tcgen05.mma / commit /
// ld wrappers, mbarrier helpers.  sm_100a only (inline PTX; no CUTLASS dependency).
//
// Operand layout used throughout ("panel layout"): a [rows x cols] fp32 operand is stored
// as cols/4 panels; panel q holds, for every row, the 16-byte chunk of columns 4q..4q+3:
//     byte_offset(row, col) = (col / 4) * panel_bytes + row * 16 + (col % 4) * 4
// This is the UMMA canonical no-swizzle ("interleave") layout made of 8-row x 16-byte core
// matrices.  The same bytes serve as
//   * a K-major operand with MN = row, K = col   (SBO = 128 B between 8-row groups,
//                                                 LBO = panel_bytes between 16-byte K chunks)
//   * an MN-major operand with MN = col, K = row (SBO = panel_bytes, LBO = 128 B)
// because a core matrix is 8 x 16 B either way.  panel_bytes = rows*16 + 16 (the 16-byte pad
// staggers the panels over the shared-memory banks for the CUDA-core writers).
#pragma once
#include <cuda_fp16.h>
#include "gfc_common.cuh"

namespace gfc {
namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// one elected lane of a converged warp (lets the compiler keep the issuing thread's descriptor math
// in uniform registers: UTCHMMA takes its operands from the uniform register file)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMEM -------------------------------------------------------------------------------
// one full warp; writes the base address (lane<<16 | column) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core's reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// same wait, but the warp is suspended by the hardware until the phase completes (or `hint_ns` elapse) instead
// of re-issuing try_wait every few cycles: a spinning warp takes issue slots from the warps it is waiting for
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 1000000u) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(hint_ns)
        : "memory");
  } while (!done);
}

// ---- descriptors --------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;  // version
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                          // c_format  = F32
         | (2u << 7)                        // a_format  = TF32
         | (2u << 10)                       // b_format  = TF32
         | ((uint32_t)a_mn_major << 15)     // a_major   (0 = K, 1 = MN)
         | ((uint32_t)b_mn_major << 16)     // b_major
         | ((uint32_t)(N >> 3) << 17)       // n_dim
         | ((uint32_t)(M >> 4) << 24);      // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -> registers: this thread's lane (row), 8 consecutive fp32 columns ---------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- bf16 operands (kind::f16), fp32 accumulate ---------------------------------------------
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                          // c_format  = F32
         | (1u << 7)                        // a_format  = BF16
         | (1u << 10)                       // b_format  = BF16
         | ((uint32_t)a_mn_major << 15)     // a_major   (0 = K, 1 = MN)
         | ((uint32_t)b_mn_major << 16)     // b_major
         | ((uint32_t)(N >> 3) << 17)       // n_dim
         | ((uint32_t)(M >> 4) << 24);      // m_dim
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand from tensor memory (lanes = M rows, 32-bit columns = two consecutive K elements each)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: this thread's lane, 32 / 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// 16 TMEM lanes x 32 columns in the matrix-fragment distribution: thread t holds, for rho = 0..3,
//   r[4 rho + 0..1] = (lane t/4,     columns 8 rho + 2 (t%4) + {0,1})
//   r[4 rho + 2..3] = (lane t/4 + 8, same columns)
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier: arrive / transaction bytes / bulk async copy (global -> shared) ----------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// contiguous bulk copy, completion counted in bytes on the mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMA tensor store (shared -> global through a CUtensorMap), bulk async-group completion ------
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* src_smem, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(src_smem)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk stores issued so far have finished READING shared memory (the staging buffer is reusable)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ask the TMA engine to pull a contiguous global region into L2 (no destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// ---- TMEM -> registers: this thread's lane (row), 32 consecutive fp32 columns ----------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 -> three bf16 planes by successive truncation: x = p0 + p1 + p2 (+ < 2^-24 |x|), every
// subtraction exact.  Results are fp32 bit patterns whose upper 16 bits are the bf16 values.
__device__ __forceinline__ void split_bf16x3(float x, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  p0 = __float_as_uint(x) & 0xffff0000u;
  const float r1 = x - __uint_as_float(p0);
  p1 = __float_as_uint(r1) & 0xffff0000u;
  const float r2 = r1 - __uint_as_float(p1);
  p2 = __float_as_uint(r2);
}
// two fp32 bit patterns -> one word of two bf16 (upper halves); `a` at the lower address
__device__ __forceinline__ uint32_t pack_bf16_hi(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }

// ---- fp16 operands (kind::f16 with a_format = b_format = F16), fp32 accumulate ------------------------
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                          // c_format  = F32
         | (0u << 7)                        // a_format  = F16
         | (0u << 10)                       // b_format  = F16
         | ((uint32_t)a_mn_major << 15)     // a_major   (0 = K, 1 = MN)
         | ((uint32_t)b_mn_major << 16)     // b_major
         | ((uint32_t)(N >> 3) << 17)       // n_dim
         | ((uint32_t)(M >> 4) << 24);      // m_dim
}
// Two (already scaled) fp32 values -> one word of two fp16: `a` in the low half (lower address).  hi = RN(x),
// lo = RN(x - hi): x = hi + lo to 2^-22 |x| while lo is a normal fp16 (|x| >= 2^-3), to 2^-25 absolute below
// (fp16 subnormals are exact multiples of 2^-24 and run at full rate on the tensor core).
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// 8 consecutive channels of one row (already scaled) -> one 16-byte chunk in each of the NP fp16 planes
template <int NP>
__device__ __forceinline__ void store_chunk_f16(unsigned char* plane0, int plane_bytes, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (NP == 2) split_f16x2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
    else hi[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
  }
  *reinterpret_cast<uint4*>(plane0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (NP == 2) *reinterpret_cast<uint4*>(plane0 + plane_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// power-of-two scale that brings a tensor of maximum magnitude m (bit pattern mbits, m >= 0) just below
// 2^top:  m * s in [2^(top-1), 2^top).  Returns s; *inv = 1/s.  m = 0 (or tiny / huge) clamps to a finite s.
__device__ __forceinline__ float pow2_scale(uint32_t mbits, int top, float* inv) {
  int e = (int)((mbits >> 23) & 0xffu);            // m in [2^(e-127), 2^(e-126))
  int se = 127 + top - (e - 126);                  // biased exponent of s
  if (e == 0) se = 127;                            // zero / denormal input: leave it alone
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  *inv = __uint_as_float((uint32_t)(254 - se) << 23);
  return __uint_as_float((uint32_t)se << 23);
}

// shared-memory matrix descriptor (no swizzle) from byte offsets
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (((lbo >> 4) & 0x3fffu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3fffu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}


// 8 consecutive channels of one row -> one 16-byte chunk in each of the three planes
__device__ __forceinline__ void store_chunk3(unsigned char* plane0, int plane_bytes, const float (&v)[8]) {
  uint32_t a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16x3(v[i], a[i], b[i], c[i]);
  *reinterpret_cast<uint4*>(plane0) = make_uint4(pack_bf16_hi(a[0], a[1]), pack_bf16_hi(a[2], a[3]),
                                                 pack_bf16_hi(a[4], a[5]), pack_bf16_hi(a[6], a[7]));
  *reinterpret_cast<uint4*>(plane0 + plane_bytes) =
      make_uint4(pack_bf16_hi(b[0], b[1]), pack_bf16_hi(b[2], b[3]), pack_bf16_hi(b[4], b[5]),
                 pack_bf16_hi(b[6], b[7]));
  *reinterpret_cast<uint4*>(plane0 + 2 * plane_bytes) =
      make_uint4(pack_bf16_hi(c[0], c[1]), pack_bf16_hi(c[2], c[3]), pack_bf16_hi(c[4], c[5]),
                 pack_bf16_hi(c[6], c[7]));
}


// panel layout helpers (see header comment); offsets in floats
__host__ __device__ constexpr int panel_floats(int rows) { return rows * 4 + 4; }
__device__ __forceinline__ int panel_off(int row, int col, int pfloats) {
  return (col >> 2) * pfloats + row * 4 + (col & 3);
}

}  // namespace tc5
}  // namespace gfc
