// gfc_gso.cu — kernel (a): robot positions -> graph shift operator, plus the CSR
// builder used by the large-swarm path.
//
// Reference behaviour reproduced (bit-exact mask):
//   Scene.readADjMatrix                      scene.py:140-154   (d <= R, zero diag, {0,1})
//   computeAdjacencyMatrix_fixedCommRadius   utils/multirobotsim_dcenlocal.py:291-317
//                                            (d < R, zero diag, D^-1/2 W D^-1/2, deg==0 guard)
// The reference evaluates the distance in float64; so does this kernel (6 fp64
// ops per pair), comparing the squared distance with a host-computed threshold
// that is equivalent to comparing the correctly rounded sqrt with R.
#include "gfc_common.cuh"
#include <math.h>

namespace gfc {

double squared_threshold(double R, bool inclusive) {
  if (isnan(R)) return -1.0;
  if (inclusive) {
    if (R < 0) return -1.0;
    if (isinf(R)) return INFINITY;
    double s = R * R;
    if (isinf(s)) return 1.7976931348623157e308;
    while (sqrt(s) <= R) s = nextafter(s, INFINITY);
    while (s >= 0 && sqrt(s) > R) s = nextafter(s, -INFINITY);
    return s;
  }
  if (R <= 0) return -1.0;
  if (isinf(R)) return 1.7976931348623157e308;
  double s = R * R;
  if (isinf(s)) return 1.7976931348623157e308;
  while (sqrt(s) < R) s = nextafter(s, INFINITY);
  while (s > 0 && !(sqrt(s) < R)) s = nextafter(s, -INFINITY);
  if (!(sqrt(s) < R)) return -1.0;
  return s;
}

// One CTA handles `gpc` consecutive graphs.  smem: [isd: gpc*N doubles (NORM only)]
// [positions: gpc*N*2 floats].
template <bool NORM>
__global__ void __launch_bounds__(256)
gso_build_kernel(const float* __restrict__ pos, int B, int N, double thr, int gpc,
                 uint8_t* __restrict__ adj, float* __restrict__ S) {
  extern __shared__ double smem_d[];
  double* isd = smem_d;
  float* sp = reinterpret_cast<float*>(smem_d + (NORM ? (size_t)gpc * N : 0));
  const int tid = threadIdx.x, nt = blockDim.x;
  const int b0 = blockIdx.x * gpc;
  const int gcount = min(gpc, B - b0);
  const int nn = gcount * N;
  const float* gp = pos + (size_t)b0 * N * 2;
  for (int i = tid; i < nn * 2; i += nt) sp[i] = gp[i];
  __syncthreads();
  if (NORM) {
    for (int r = tid; r < nn; r += nt) {
      const int j = r / N, i = r - j * N;
      const float xi = sp[2 * r], yi = sp[2 * r + 1];
      int deg = 0;
      for (int m = 0; m < N; ++m) {
        const int q = j * N + m;
        deg += (m != i) && (sqdist64(xi, yi, sp[2 * q], sp[2 * q + 1]) <= thr);
      }
      isd[r] = inv_sqrt_deg(deg);
    }
    __syncthreads();
  }
  const size_t base = (size_t)b0 * N * N;
  const int total = nn * N;
  for (int o = tid; o < total; o += nt) {
    const int r = o / N, n2 = o - r * N;
    const int j = r / N, i = r - j * N;
    const int q = j * N + n2;
    const bool a = (i != n2) && (sqdist64(sp[2 * r], sp[2 * r + 1], sp[2 * q], sp[2 * q + 1]) <= thr);
    if (adj) adj[base + o] = a ? 1 : 0;
    if (S) {
      float v = a ? 1.f : 0.f;
      if (NORM) v = a ? (float)__dmul_rn(isd[r], isd[q]) : 0.f;
      S[base + o] = v;
    }
  }
}

// CSR build, pass 1 and 3 share one loop: a CTA owns 256 rows of ONE graph and walks over all nodes of that graph
// staged in shared memory (128-bit broadcasts), deciding each pair with the fp32 screen of the fused kernels
// (s < thr_lo: inside, s > thr_hi: outside) and the exact fp64 rule of (a) only inside the rounding band — the same
// bit-exact decisions as sqdist64(...) <= thr everywhere, at a third of the instruction count.
static __device__ __noinline__ bool csr_pair_exact(float xi, float yi, float xj, float yj, double thr) {
  return sqdist64(xi, yi, xj, yj) <= thr;
}
constexpr int kCsrChunk = 2048;   // nodes staged per pass (16 KB)

template <bool FILL, bool NORM>
__global__ void __launch_bounds__(256)
csr_rows_kernel(const float* __restrict__ pos, int N, double thr, float thr_lo, float thr_hi,
                int32_t* __restrict__ deg, const int32_t* __restrict__ rowptr, long long nnz_stride,
                int32_t* __restrict__ colidx, float* __restrict__ vals) {
  __shared__ float2 sp[kCsrChunk];
  const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
  const bool valid = i < N;
  const float2* gp = reinterpret_cast<const float2*>(pos) + (size_t)b * N;
  const float2 pi = valid ? __ldg(gp + i) : make_float2(0.f, 0.f);
  const int32_t* rp = FILL ? rowptr + (size_t)b * (N + 1) : nullptr;
  int d = 0;
  long long w = 0;
  double isd_i = 0.0;
  if (FILL && valid) {
    w = (long long)b * nnz_stride + rp[i];
    if (NORM) isd_i = inv_sqrt_deg(rp[i + 1] - rp[i]);
  }
  for (int m0 = 0; m0 < N; m0 += kCsrChunk) {
    const int cnt = min(kCsrChunk, N - m0);
    __syncthreads();
    for (int q = threadIdx.x; q < cnt; q += 256) sp[q] = __ldg(gp + m0 + q);
    __syncthreads();
    if (!valid) continue;
#pragma unroll 4
    for (int q = 0; q < cnt; ++q) {
      const float2 pj = sp[q];
      const float dx = pi.x - pj.x, dy = pi.y - pj.y;
      const float sq = fmaf(dx, dx, dy * dy);
      bool e = sq < thr_lo;
      if (!e && !(sq > thr_hi)) e = csr_pair_exact(pi.x, pi.y, pj.x, pj.y, thr);   // rounding band: exact rule
      e = e && (m0 + q != i);
      if (FILL) {
        if (e) {
          const int m = m0 + q;
          colidx[w] = m;
          if (vals) vals[w] = NORM ? (float)__dmul_rn(inv_sqrt_deg(rp[m + 1] - rp[m]), isd_i) : 1.f;
          ++w;
        }
      } else {
        d += e ? 1 : 0;
      }
    }
  }
  if (!FILL && valid) deg[(size_t)b * N + i] = d;
}

// rowptr[b, 0..N] = exclusive scan of deg[b, :]; one CTA per graph.
__global__ void __launch_bounds__(256)
csr_scan_kernel(const int32_t* __restrict__ deg, int N, int32_t* __restrict__ rowptr) {
  __shared__ int part[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int chunk = (N + 255) / 256;
  const int lo = min(N, tid * chunk), hi = min(N, lo + chunk);
  const int32_t* d = deg + (size_t)b * N;
  int32_t* rp = rowptr + (size_t)b * (N + 1);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += d[i];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { int v = part[i]; part[i] = run; run += v; }
    rp[N] = run;
  }
  __syncthreads();
  int run = part[tid];
  for (int i = lo; i < hi; ++i) { rp[i] = run; run += d[i]; }
}

}  // namespace gfc

using namespace gfc;

// fp32 screening band around the squared-distance threshold (same band as set_thresholds in gfc_api.cu: an fp32
// squared distance differs from the fp64 one by < 2^-22 relative; 4e-6 on each side)
static void screen_band(double thr, float* lo, float* hi) {
  if (!(thr >= 0)) { *lo = -1.f; *hi = -1.f; return; }
  *lo = nextafterf((float)(thr * (1.0 - 4e-6)), -INFINITY);
  *hi = nextafterf((float)(thr * (1.0 + 4e-6)), INFINITY);
}

static int mode_threshold(int mode, double radius, double* thr, bool* norm) {
  switch (mode) {
    case GFC_GSO_BINARY_LE: *thr = squared_threshold(radius, true); *norm = false; return GFC_OK;
    case GFC_GSO_SYM_NORM_LT: *thr = squared_threshold(radius, false); *norm = true; return GFC_OK;
    case GFC_GSO_BINARY_LT: *thr = squared_threshold(radius, false); *norm = false; return GFC_OK;
    default: set_error("unknown GSO mode %d", mode); return GFC_ERR_BAD_ARG;
  }
}

// exported for the fused kernels in gfc_tile.cu
namespace gfc {
int gso_mode_threshold(int mode, double radius, double* thr, bool* norm) {
  return mode_threshold(mode, radius, thr, norm);
}
}

extern "C" int gfc_gso_build(const float* pos, int B, int N, double radius, int mode,
                             uint8_t* adj_out, float* S_out, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0, GFC_ERR_BAD_ARG, "gfc_gso_build: negative size B=%d N=%d", B, N);
  if (B == 0 || N == 0) return GFC_OK;
  GFC_REQUIRE(pos != nullptr, GFC_ERR_BAD_ARG, "gfc_gso_build: pos is NULL");
  GFC_REQUIRE(adj_out || S_out, GFC_ERR_BAD_ARG, "gfc_gso_build: both outputs NULL");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  GFC_REQUIRE((long long)N * N <= 0x7fffffffLL / 2, GFC_ERR_UNSUPPORTED,
              "gfc_gso_build: N=%d too large for the dense builder (use gfc_csr_*)", N);
  long long per = (long long)N * N;
  int gpc = (int)(2048 / per);
  if (gpc < 1) gpc = 1;
  if (gpc > 64) gpc = 64;
  if (gpc > B) gpc = B;
  size_t smem = (size_t)gpc * N * 2 * sizeof(float) + (norm ? (size_t)gpc * N * sizeof(double) : 0);
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc) return rc;
  GFC_REQUIRE(smem <= (size_t)di.smem_optin, GFC_ERR_UNSUPPORTED,
              "gfc_gso_build: N=%d needs %zu B shared memory (> %d)", N, smem, di.smem_optin);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = ceil_div(B, gpc);
  if (norm) {
    if (smem > 48 * 1024)
      GFC_CUDA_TRY(cudaFuncSetAttribute(gso_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gso_build_kernel<true><<<grid, 256, smem, st>>>(pos, B, N, thr, gpc, adj_out, S_out);
  } else {
    if (smem > 48 * 1024)
      GFC_CUDA_TRY(cudaFuncSetAttribute(gso_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gso_build_kernel<false><<<grid, 256, smem, st>>>(pos, B, N, thr, gpc, adj_out, S_out);
  }
  GFC_LAUNCH_CHECK("gso_build_kernel");
  return GFC_OK;
}

extern "C" int gfc_csr_count(const float* pos, int B, int N, double radius, int mode,
                             int32_t* deg, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0, GFC_ERR_BAD_ARG, "gfc_csr_count: negative size");
  if (B == 0 || N == 0) return GFC_OK;
  GFC_REQUIRE(pos && deg, GFC_ERR_BAD_ARG, "gfc_csr_count: NULL pointer");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  float lo, hi;
  screen_band(thr, &lo, &hi);
  for (int b0 = 0; b0 < B; b0 += 65535) {   // gridDim.y limit: 65535 graphs per launch
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    csr_rows_kernel<false, false><<<dim3((unsigned)((N + 255) / 256), (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(
        pos + (size_t)b0 * N * 2, N, thr, lo, hi, deg + (size_t)b0 * N, nullptr, 0, nullptr, nullptr);
    GFC_LAUNCH_CHECK("csr_rows_kernel<count>");
  }
  return GFC_OK;
}

extern "C" int gfc_csr_scan(const int32_t* deg, int B, int N, int32_t* rowptr, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0, GFC_ERR_BAD_ARG, "gfc_csr_scan: negative size");
  if (B == 0) return GFC_OK;
  GFC_REQUIRE(deg && rowptr, GFC_ERR_BAD_ARG, "gfc_csr_scan: NULL pointer");
  csr_scan_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(deg, N, rowptr);
  GFC_LAUNCH_CHECK("csr_scan_kernel");
  return GFC_OK;
}

extern "C" int gfc_csr_fill(const float* pos, int B, int N, double radius, int mode,
                            const int32_t* rowptr, int64_t nnz_stride,
                            int32_t* colidx, float* vals, void* stream) {
  launch_counter() = 0;
  GFC_REQUIRE(B >= 0 && N >= 0 && nnz_stride >= 0, GFC_ERR_BAD_ARG, "gfc_csr_fill: negative size");
  if (B == 0 || N == 0) return GFC_OK;
  GFC_REQUIRE(pos && rowptr && colidx, GFC_ERR_BAD_ARG, "gfc_csr_fill: NULL pointer");
  double thr; bool norm;
  int rc = mode_threshold(mode, radius, &thr, &norm);
  if (rc) return rc;
  float lo, hi;
  screen_band(thr, &lo, &hi);
  for (int b0 = 0; b0 < B; b0 += 65535) {   // gridDim.y limit: 65535 graphs per launch
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    const dim3 grid((unsigned)((N + 255) / 256), (unsigned)nb);
    const float* p0 = pos + (size_t)b0 * N * 2;
    const int32_t* rp0 = rowptr + (size_t)b0 * (N + 1);
    int32_t* ci0 = colidx + (size_t)b0 * nnz_stride;
    float* v0 = vals ? vals + (size_t)b0 * nnz_stride : nullptr;
    if (norm)
      csr_rows_kernel<true, true><<<grid, 256, 0, (cudaStream_t)stream>>>(p0, N, thr, lo, hi, nullptr, rp0, nnz_stride,
                                                                           ci0, v0);
    else
      csr_rows_kernel<true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(p0, N, thr, lo, hi, nullptr, rp0, nnz_stride,
                                                                            ci0, v0);
    GFC_LAUNCH_CHECK("csr_rows_kernel<fill>");
  }
  return GFC_OK;
}
