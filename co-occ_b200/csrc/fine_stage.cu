// fine_stage.cu -- CUDA launchers of the OccHead fine / cascade stage (bodies in fine_stage.cuh).
// The arithmetic of every body is verified on the CPU (tests/test_fine_emul.py runs the same source through
// tests/emul/fine_emul.cpp against torch's grid_sample / group_norm and the pinned oracle) and on a B200 against the
// reference fixture (tests/test_gpu_fine.py).  GroupNorm additionally has CUDA-only kernels below (same formulas):
// the generic one-item-per-thread bodies serialise the per-channel dgamma / dbeta atomics of 10^5 point rows.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/coocc_b200.h"
#include "fine_stage.cuh"

namespace coocc {
namespace fine {

template <typename P, void (*Body)(const P&, long long)>
__global__ void __launch_bounds__(256) items_kernel(const P p, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) Body(p, i);
}

template <typename P, void (*Body)(const P&, long long)>
static int run(const P& p, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  long long b = (n + 255) / 256;
  if (b > 148LL * 32) b = 148LL * 32;
  items_kernel<P, Body><<<(unsigned)b, 256, 0, st>>>(p, n);
  return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
}

// ---- GroupNorm, point rows (span == 1): one thread per (row, group), CPG channels of the group in registers -----
// forward: statistics + apply in one pass.  256 threads = 256 / G row lanes x G groups.
template <int CPG>
__global__ void __launch_bounds__(256) gn_rows_fwd_kernel(const GroupNormP p) {
  const int G = p.G, RL = 256 / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  if (rl >= RL) return;
  float ga[CPG], be[CPG];
#pragma unroll
  for (int c = 0; c < CPG; ++c) { ga[c] = p.gamma[g * CPG + c]; be[c] = p.beta[g * CPG + c]; }
  for (long long r = (long long)blockIdx.x * RL + rl; r < p.rows; r += (long long)gridDim.x * RL) {
    const float* row = p.x + r * p.ldx + g * CPG;
    float v[CPG];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) { v[c] = row[c]; sum += v[c]; }
    const float mean = sum / CPG;
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) var += (v[c] - mean) * (v[c] - mean);
    const float rstd = rsqrtf(var / CPG + p.eps);
    p.stats[(r * G + g) * 2 + 0] = mean;
    p.stats[(r * G + g) * 2 + 1] = rstd;
    float* out = p.y + r * p.ldy + g * CPG;
#pragma unroll
    for (int c = 0; c < CPG; ++c) {
      float o = (v[c] - mean) * rstd * ga[c] + be[c];
      if (p.relu && o < 0.f) o = 0.f;
      out[c] = o;
    }
  }
}

// backward: group sums, dx and the dgamma / dbeta partials in one pass; the partials are kept in registers over all
// rows of the thread, reduced over the block's row lanes in shared memory, one global atomic per channel and block.
template <int CPG>
__global__ void __launch_bounds__(256) gn_rows_bwd_kernel(const GroupNormP p) {
  __shared__ float s_dg[128], s_db[128];              // C = G * CPG <= 128 (checked by the launcher)
  const int G = p.G, RL = 256 / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  for (int i = threadIdx.x; i < G * CPG; i += 256) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  float ga[CPG], be[CPG], dg[CPG], db[CPG];
#pragma unroll
  for (int c = 0; c < CPG; ++c) { ga[c] = p.gamma[g * CPG + c]; be[c] = p.beta[g * CPG + c]; dg[c] = 0.f; db[c] = 0.f; }
  const float inv_n = 1.f / CPG;
  if (rl < RL) {
    for (long long r = (long long)blockIdx.x * RL + rl; r < p.rows; r += (long long)gridDim.x * RL) {
      const float mean = p.stats[(r * G + g) * 2 + 0], rstd = p.stats[(r * G + g) * 2 + 1];
      const float* row = p.x + r * p.ldx + g * CPG;
      const float* drow = p.dy + r * p.lddy + g * CPG;
      float xh[CPG], d[CPG];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        xh[c] = (row[c] - mean) * rstd;
        d[c] = drow[c];
        if (p.relu && !(xh[c] * ga[c] + be[c] > 0.f)) d[c] = 0.f;
        dg[c] += d[c] * xh[c];
        db[c] += d[c];
        s1 += d[c] * ga[c];
        s2 += d[c] * ga[c] * xh[c];
      }
      const float m1 = s1 * inv_n, m2 = s2 * inv_n;
      float* out = p.dx + r * p.lddx + g * CPG;
#pragma unroll
      for (int c = 0; c < CPG; ++c) out[c] = rstd * (d[c] * ga[c] - m1 - xh[c] * m2);
    }
#pragma unroll
    for (int c = 0; c < CPG; ++c) { atomicAdd(&s_dg[g * CPG + c], dg[c]); atomicAdd(&s_db[g * CPG + c], db[c]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * CPG; i += 256) { atomicAdd(p.dgamma + i, s_dg[i]); atomicAdd(p.dbeta + i, s_db[i]); }
}

// ---- GroupNorm, feature maps (span > 1): one block per (sample, group) ----------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(256) gn_map_stats_kernel(const GroupNormP p) {
  __shared__ double sh[8];
  const long long s = blockIdx.x / p.G;
  const int g = blockIdx.x % p.G, cpg = p.C / p.G;
  double sum = 0.0, sq = 0.0;
  for (long long i = threadIdx.x; i < (long long)p.span * cpg; i += 256) {
    const float v = p.x[(s * p.span + i / cpg) * p.ldx + g * cpg + (int)(i % cpg)];
    sum += v; sq += (double)v * v;
  }
  sum = block_sum(sum, sh);
  sq = block_sum(sq, sh);
  if (threadIdx.x == 0) {
    const double n = (double)p.span * cpg, mean = sum / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    p.stats[blockIdx.x * 2 + 0] = (float)mean;
    p.stats[blockIdx.x * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
  }
}

__global__ void __launch_bounds__(256) gn_map_bwd_sums_kernel(const GroupNormP p) {
  __shared__ double sh[8];
  const long long s = blockIdx.x / p.G;
  const int g = blockIdx.x % p.G, cpg = p.C / p.G;
  const float mean = p.stats[blockIdx.x * 2 + 0], rstd = p.stats[blockIdx.x * 2 + 1];
  double s1 = 0.0, s2 = 0.0;
  for (int c = 0; c < cpg; ++c) {
    const int ch = g * cpg + c;
    const float gam = p.gamma[ch], bet = p.beta[ch];
    double dg = 0.0, db = 0.0;
    for (long long r = s * p.span + threadIdx.x; r < (s + 1) * p.span; r += 256) {
      const float xh = (p.x[r * p.ldx + ch] - mean) * rstd;
      float d = p.dy[r * p.lddy + ch];
      if (p.relu && !(xh * gam + bet > 0.f)) d = 0.f;
      dg += (double)d * xh;
      db += d;
    }
    dg = block_sum(dg, sh);
    db = block_sum(db, sh);
    if (threadIdx.x == 0) { atomicAdd(p.dgamma + ch, (float)dg); atomicAdd(p.dbeta + ch, (float)db); }
    s1 += db * gam;
    s2 += dg * gam;
  }
  if (threadIdx.x == 0) {
    p.sums[blockIdx.x * 2 + 0] = (float)s1;
    p.sums[blockIdx.x * 2 + 1] = (float)s2;
  }
}

static int rows_grid(long long rows, int G) {
  const int RL = 256 / G;
  long long b = (rows + RL - 1) / RL;
  if (b > 148LL * 8) b = 148LL * 8;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace fine
}  // namespace coocc

using namespace coocc::fine;

extern "C" int coocc_fine_sample3d_fwd(const float* feats, long long ld, int X, int Y, int Z, int C, const int* coords,
                                       int M, int SX, int SY, int SZ, float* out, long long ldo, void* stream) {
  if (!feats || !coords || !out || C < 4 || (C & 3) || (ld & 3) || (ldo & 3) || M < 0 || SX < 2 || SY < 2 || SZ < 2)
    return COOCC_ERR_ARG;
  Sample3dP p{};
  p.feats = feats; p.ld = ld; p.X = X; p.Y = Y; p.Z = Z; p.C = C; p.coords = coords; p.M = M;
  p.SX = SX; p.SY = SY; p.SZ = SZ; p.out = out; p.ldo = ldo;
  return run<Sample3dP, sample3d_fwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample3d_fwd_bf16(const void* feats, long long ld, int X, int Y, int Z, int C, const int* coords,
                                            int M, int SX, int SY, int SZ, float* out, long long ldo, void* stream) {
  if (!feats || !coords || !out || C < 4 || (C & 3) || (ld & 3) || (ldo & 3) || M < 0 || SX < 2 || SY < 2 || SZ < 2)
    return COOCC_ERR_ARG;
  Sample3dP p{};
  p.feats_bf16 = reinterpret_cast<const unsigned short*>(feats); p.ld = ld; p.X = X; p.Y = Y; p.Z = Z; p.C = C;
  p.coords = coords; p.M = M; p.SX = SX; p.SY = SY; p.SZ = SZ; p.out = out; p.ldo = ldo;
  return run<Sample3dP, sample3d_fwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample3d_bwd(const float* gout, long long ldg, int X, int Y, int Z, int C, const int* coords,
                                       int M, int SX, int SY, int SZ, float* dfeats, long long ldd, void* stream) {
  if (!gout || !coords || !dfeats || C < 4 || (C & 3) || (ldg & 3) || (ldd & 3) || M < 0 || SX < 2 || SY < 2 || SZ < 2)
    return COOCC_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(dfeats) & 15) return COOCC_ERR_ALIGN;      // 16-byte vector reductions
  Sample3dP p{};
  p.X = X; p.Y = Y; p.Z = Z; p.C = C; p.coords = coords; p.M = M; p.SX = SX; p.SY = SY; p.SZ = SZ;
  p.gout = gout; p.ldg = ldg; p.dfeats = dfeats; p.ldd = ldd;
  return run<Sample3dP, sample3d_bwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_project(const int* coords, int M, int ncam, const float* vs3, const float* lo3,
                                  const float* inv_bda, const float* cam27, float W_img, float H_img, float* uv,
                                  unsigned char* mask, void* stream) {
  if (!coords || !vs3 || !lo3 || !inv_bda || !cam27 || !uv || !mask || M < 0 || ncam < 1) return COOCC_ERR_ARG;
  ProjectP p{};
  p.coords = coords; p.M = M; p.ncam = ncam;
  for (int a = 0; a < 3; ++a) { p.vs[a] = vs3[a]; p.lo[a] = lo3[a]; }      // host float[3]
  p.inv_bda = inv_bda; p.cam = cam27; p.W_img = W_img; p.H_img = H_img; p.uv = uv; p.mask = mask;
  return run<ProjectP, project_item>(p, (long long)M * ncam, (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample2d_fwd(const float* img, long long ld, int ncam, int H, int W, int C, const float* uv,
                                       const unsigned char* mask, int M, float* out, long long ldo, void* stream) {
  if (!img || !uv || !mask || !out || C < 4 || (C & 3) || (ld & 3) || (ldo & 3) || M < 0) return COOCC_ERR_ARG;
  Sample2dP p{};
  p.img = img; p.ld = ld; p.ncam = ncam; p.H = H; p.W = W; p.C = C; p.uv = uv; p.mask = mask; p.M = M;
  p.out = out; p.ldo = ldo;
  return run<Sample2dP, sample2d_fwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_fine_sample2d_bwd(const float* gout, long long ldg, int ncam, int H, int W, int C, const float* uv,
                                       const unsigned char* mask, int M, float* dimg, long long ldd, void* stream) {
  if (!gout || !uv || !mask || !dimg || C < 4 || (C & 3) || (ldg & 3) || (ldd & 3) || M < 0) return COOCC_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(dimg) & 15) return COOCC_ERR_ALIGN;        // 16-byte vector reductions
  Sample2dP p{};
  p.ncam = ncam; p.H = H; p.W = W; p.C = C; p.uv = uv; p.mask = mask; p.M = M;
  p.gout = gout; p.ldg = ldg; p.dimg = dimg; p.ldd = ldd;
  return run<Sample2dP, sample2d_bwd_item>(p, (long long)M * (C >> 2), (cudaStream_t)stream);
}

extern "C" int coocc_groupnorm_fwd(const float* x, long long ldx, long long rows, int C, int G, int span,
                                   const float* gamma, const float* beta, float eps, int relu, float* stats, float* y,
                                   long long ldy, void* stream) {
  if (!x || !gamma || !beta || !stats || !y || rows < 0 || C < 1 || G < 1 || C % G || span < 1 || rows % span)
    return COOCC_ERR_ARG;
  GroupNormP p{};
  p.x = x; p.ldx = ldx; p.rows = rows; p.C = C; p.G = G; p.span = span; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.relu = relu; p.stats = stats; p.y = y; p.ldy = ldy;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) return 0;
  const int cpg = C / G;
  if (span == 1 && 256 % G == 0 && (cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8)) {
    if (cpg == 1) gn_rows_fwd_kernel<1><<<rows_grid(rows, G), 256, 0, st>>>(p);
    else if (cpg == 2) gn_rows_fwd_kernel<2><<<rows_grid(rows, G), 256, 0, st>>>(p);
    else if (cpg == 4) gn_rows_fwd_kernel<4><<<rows_grid(rows, G), 256, 0, st>>>(p);
    else gn_rows_fwd_kernel<8><<<rows_grid(rows, G), 256, 0, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
  }
  int rc = 0;
  if (span > 1) {
    gn_map_stats_kernel<<<(unsigned)(rows / span * G), 256, 0, st>>>(p);
    if (cudaGetLastError() != cudaSuccess) rc = COOCC_ERR_CUDA;
  } else {
    rc = run<GroupNormP, gn_stats_item>(p, rows / span * G, st);
  }
  if (!rc) rc = run<GroupNormP, gn_apply_item>(p, rows * G, st);
  return rc;
}

extern "C" int coocc_groupnorm_bwd(const float* x, long long ldx, long long rows, int C, int G, int span,
                                   const float* gamma, const float* beta, int relu, const float* stats, const float* dy,
                                   long long lddy, float* sums, float* dx, long long lddx, float* dgamma, float* dbeta,
                                   void* stream) {
  if (!x || !gamma || !beta || !stats || !dy || !sums || !dx || !dgamma || !dbeta || rows < 0 || C < 1 || G < 1 || C % G ||
      span < 1 || rows % span)
    return COOCC_ERR_ARG;
  GroupNormP p{};
  p.x = x; p.ldx = ldx; p.rows = rows; p.C = C; p.G = G; p.span = span; p.gamma = gamma; p.beta = beta; p.relu = relu;
  p.stats = const_cast<float*>(stats); p.dy = dy; p.lddy = lddy; p.sums = sums; p.dx = dx; p.lddx = lddx;
  p.dgamma = dgamma; p.dbeta = dbeta;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) return 0;
  const int cpg = C / G;
  if (span == 1 && 256 % G == 0 && (cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8) && C <= 128) {
    if (cpg == 1) gn_rows_bwd_kernel<1><<<rows_grid(rows, G), 256, 0, st>>>(p);
    else if (cpg == 2) gn_rows_bwd_kernel<2><<<rows_grid(rows, G), 256, 0, st>>>(p);
    else if (cpg == 4) gn_rows_bwd_kernel<4><<<rows_grid(rows, G), 256, 0, st>>>(p);
    else gn_rows_bwd_kernel<8><<<rows_grid(rows, G), 256, 0, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : COOCC_ERR_CUDA;
  }
  int rc = 0;
  if (span > 1) {
    gn_map_bwd_sums_kernel<<<(unsigned)(rows / span * G), 256, 0, st>>>(p);
    if (cudaGetLastError() != cudaSuccess) rc = COOCC_ERR_CUDA;
  } else {
    rc = run<GroupNormP, gn_bwd_sums_item>(p, rows / span * G, st);
  }
  if (!rc) rc = run<GroupNormP, gn_bwd_apply_item>(p, rows * G, st);
  return rc;
}
